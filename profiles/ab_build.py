"""Build a second copy of the library from a git revision of csrc/ for same-box A/B runs (boxes differ by several
percent, so two builds are only comparable inside one gpurun call):

    python profiles/ab_build.py <git-rev> <tag> [path=replacement ...]  ->  profiles/bin/libl2d_<tag>.so   (git-ignored, travels)
    L2D_LIB_OVERRIDE=profiles/bin/libl2d_<tag>.so python bench.py ...
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from live2diff_b200.csrc import build as B  # noqa: E402

rev, tag = sys.argv[1], sys.argv[2]
replace = dict(a.split("=", 1) for a in sys.argv[3:])          # repo-relative path = file to use instead of the revision's
tmp = tempfile.mkdtemp()
os.makedirs(os.path.join(tmp, "live2diff_b200", "csrc"))
os.makedirs(os.path.join(tmp, "include"))
files = subprocess.run(["git", "ls-tree", "-r", "--name-only", rev, "live2diff_b200/csrc", "include"], cwd=ROOT, capture_output=True,
                       text=True, check=True).stdout.split()
for f in files:
    data = subprocess.run(["git", "show", f"{rev}:{f}"], cwd=ROOT, capture_output=True, check=True).stdout
    if f in replace:
        data = open(replace[f], "rb").read()
    with open(os.path.join(tmp, f), "wb") as fh:
        fh.write(data)
src = os.path.join(tmp, "live2diff_b200", "csrc")
objs = []
for f in sorted(os.listdir(src)):
    if f.endswith(".cu"):
        o = os.path.join(tmp, f[:-3] + ".o")
        extra = [f'-DL2D_BUILD_HASH_STR="ab-{tag}-{rev[:10]}"'] if f == "api.cu" else []
        subprocess.run([B.NVCC, *B.FLAGS, *extra, "-c", os.path.join(src, f), "-o", o], check=True)
        objs.append(o)
out_dir = os.path.join(ROOT, "profiles", "bin")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, f"libl2d_{tag}.so")
subprocess.run([B.NVCC, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"], check=True)
shutil.rmtree(tmp)
print("built", out)
