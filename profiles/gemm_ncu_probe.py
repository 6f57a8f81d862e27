"""Three GEMM launches for an `ncu --set full` capture (qkv-like, GEGLU-like, long-K)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200 import ops  # noqa: E402

dev = "cuda:0"
for (m, n, k, act) in [(8192, 960, 320, 0), (8192, 2560, 320, 2), (8192, 320, 1280, 0)]:
    a = torch.randn(m, k, device=dev).half()
    w = (torch.randn(n, k, device=dev) / math.sqrt(k)).half()
    b = torch.randn(n, device=dev).half()
    if act == 2:
        w, b = ops.geglu_interleave(w, b, ops.gemm_tile_n(m, n, k))
    for _ in range(2):
        ops.gemm(a, w, bias=b, act=act)
    torch.cuda.synchronize()
