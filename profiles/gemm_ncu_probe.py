"""GEMM launches for an `ncu --set full --import-source on` capture: the level-0 C x C linear with bias + residual
(epilogue-bound: 128 tiles of 128 x 160, 5 k-blocks) and the fused q/k/v projection (no residual), one launch each
between cudaProfilerStart/Stop after warm-up."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200 import ops  # noqa: E402

dev = "cuda:0"
cases = []
for (m, n, k, use_res) in ((8192, 320, 320, True), (8192, 960, 320, False)):
    a = torch.randn(m, k, device=dev).half()
    w = (torch.randn(n, k, device=dev) / math.sqrt(k)).half()
    b = torch.randn(n, device=dev).half()
    res = torch.randn(m, n, device=dev).half() if use_res else None
    cases.append((a, w, b, res))
for a, w, b, res in cases:
    for _ in range(3):
        ops.gemm(a, w, bias=b, residual=res)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for a, w, b, res in cases:
    ops.gemm(a, w, bias=b, residual=res)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled", len(cases), "GEMM launches")
