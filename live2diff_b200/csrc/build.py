"""Build libl2d_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the library is a
plain C-ABI shared object loaded with ctypes, see include/l2d_b200.h).

    python live2diff_b200/csrc/build.py [--force]

Output: live2diff_b200/libl2d_b200.so (git-ignored; travels to the GPU box with the snapshot).
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libl2d_b200.so")
OBJ = os.path.join(HERE, "_build")
SOURCES = ["api.cu", "kv_attn.cu", "kv_attn_mma.cu", "kv_warmup.cu", "norms.cu", "pointwise.cu", "gemm_tcgen05.cu", "flash_attn.cu", "engine.cu", "stream_state.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def _stamp(path):
    h = hashlib.sha1()
    for f in sorted(os.listdir(HERE)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(HERE, f), "rb").read())
    h.update(open(os.path.join(PKG, "..", "include", "l2d_b200.h"), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [NVCC, *FLAGS, "-c", os.path.join(HERE, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _stamp(HERE)
    if not force and os.path.exists(OUT) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        if verbose:
            print(f"[l2d build] up to date: {OUT}")
        return OUT
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(_compile, srcs))
    cmd = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    open(stamp_file, "w").write(stamp)
    if verbose:
        print(f"[l2d build] built {OUT} from {len(srcs)} sources")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
