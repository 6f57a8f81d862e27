"""Per-CTA timeline of the tcgen05 GEMM (clock64 stamps via l2d_gemm_set_debug): where does a CTA spend its time?"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
lib = _lib.lib()
for (m, n, k, act) in [(8192, 960, 320, 0), (8192, 2560, 320, 2), (8192, 320, 1280, 0), (8192, 320, 320, 0)]:
    a = torch.randn(m, k, device=dev).half()
    w = (torch.randn(n, k, device=dev) / math.sqrt(k)).half()
    b = torch.randn(n, device=dev).half()
    res = torch.randn(m, n // (2 if act == 2 else 1), device=dev).half()
    if act == 2:
        w, b = ops.geglu_interleave(w, b, ops.gemm_tile_n(m, n, k))
    for _ in range(2):
        ops.gemm(a, w, bias=b, act=act, residual=None if act == 2 else res)
    tile = ops.gemm_tile_n(m, n, k)
    ctas = math.ceil(n / tile) * math.ceil(m / 128)
    tl = torch.zeros(ctas, 8, dtype=torch.int64, device=dev)
    lib.l2d_gemm_set_debug(tl.data_ptr())
    ops.gemm(a, w, bias=b, act=act, residual=None if act == 2 else res)
    torch.cuda.synchronize()
    lib.l2d_gemm_set_debug(0)
    t = tl.cpu().double()
    t = t[t[:, 0] > 0]           # persistent kernel: one row per CTA (first tile of each CTA)
    d = lambda i, j: float((t[:, j] - t[:, i]).mean())
    span = float(t[:, 6].max() - t[:, 0].min())
    print(f"{m}x{n}x{k} act={act} tile_n={tile} tiles={ctas} ctas={t.shape[0]}: setup {d(0,1):.0f}  first-stage {d(1,2):.0f}  mainloop-issue {d(2,3):.0f}  "
          f"accum-ready-after-last-issue {d(3,4):.0f}  epilogue {d(4,5):.0f}  join {d(5,6):.0f}  total/CTA {d(0,6):.0f} cycles; "
          f"kernel span {span:.0f} cycles")
