"""K1 tensor-core kernel only (all instantiations), for `compute-sanitizer --tool racecheck`."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200 import ops  # noqa: E402

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g).half()
for L, c, hw in ((16, 320, 32), (16, 640, 8), (16, 1280, 8), (16, 128, 8), (32, 320, 16), (32, 640, 8), (32, 1280, 4)):
    n = 2
    cache = r(n, 2, hw, L, c)
    mask = torch.zeros(n, L, device=dev).half()
    pi = torch.arange(L, device=dev).repeat(n, 1)
    up = torch.tensor([L - 1, L // 2], device=dev)
    ops.kv_attn(r(n, hw, c), r(n, hw, c), r(n, hw, c), cache, r(L, c), r(L, c), r(L, c), mask, pi, up, 8)
torch.cuda.synchronize()
print("sanitize_k1: ran")
