"""GEMM launches for an `ncu --set full` capture: plain qkv-like (no bias/residual), then with bias+residual."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200 import ops  # noqa: E402

dev = "cuda:0"
m, n, k = 8192, 960, 320
a = torch.randn(m, k, device=dev).half()
w = (torch.randn(n, k, device=dev) / math.sqrt(k)).half()
b = torch.randn(n, device=dev).half()
res = torch.randn(m, n, device=dev).half()
for _ in range(3):
    ops.gemm(a, w)
torch.cuda.synchronize()
for _ in range(2):
    ops.gemm(a, w, bias=b, residual=res)
torch.cuda.synchronize()
