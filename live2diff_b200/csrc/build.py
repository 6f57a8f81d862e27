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
SOURCES = sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))
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


MARK = b"L2D_BUILD_HASH="


def built_hash(path=OUT):
    """Source hash the library at `path` was compiled from: the marker string api.cu embeds (read from the
    file, not through dlopen, so that a rebuild inside this process is not shadowed by an already-loaded copy)."""
    try:
        blob = open(path, "rb").read()
    except OSError:
        return None
    i = blob.find(MARK)
    return blob[i + len(MARK): i + len(MARK) + 40].decode("ascii", "replace") if i >= 0 else None


def source_hash():
    return _stamp(HERE)


def _headers_hash():
    h = hashlib.sha1()
    for f in sorted(os.listdir(HERE)):
        if f.endswith((".cuh", ".h")):
            h.update(open(os.path.join(HERE, f), "rb").read())
    h.update(open(os.path.join(PKG, "..", "include", "l2d_b200.h"), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h


def _compile(src, stamp):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    # per-object reuse (scratch files under _build/, untracked): same source + headers + flags -> same object
    h = _headers_hash()
    h.update(open(os.path.join(HERE, src), "rb").read())
    if src == "api.cu":
        h.update(stamp.encode())
    key, keyfile = h.hexdigest(), obj + ".key"
    if os.path.exists(obj) and os.path.exists(keyfile) and open(keyfile).read() == key:
        return obj
    cmd = [NVCC, *FLAGS, *([f'-DL2D_BUILD_HASH_STR="{stamp}"'] if src == "api.cu" else []), "-c", os.path.join(HERE, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    open(keyfile, "w").write(key)
    return obj


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    stamp = _stamp(HERE)
    # the hash lives INSIDE the .so (a sidecar stamp can be checked out without the library it vouches for)
    if not force and built_hash(OUT) == stamp:
        if verbose:
            print(f"[l2d build] up to date: {OUT}")
        return OUT
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(lambda f: _compile(f, stamp), srcs))
    cmd = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if built_hash(OUT) != stamp:
        raise RuntimeError("built library does not carry the expected source hash")
    if verbose:
        print(f"[l2d build] built {OUT} from {len(srcs)} sources")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
