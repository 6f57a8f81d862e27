"""One launch of the tcgen05 attention kernel at the level-0 self-attention shape (and one at level 1) between
cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full --import-source on -o gpurun_out/flash_full ...`."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200 import ops  # noqa: E402

dev = "cuda:0"
cases = []
for b, heads, sq, hd in ((2, 8, 4096, 40), (2, 8, 1024, 80)):
    c = heads * hd
    qkv = torch.randn(b * sq, 3 * c, device=dev).half()
    cases.append((qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], b, heads, sq, sq, hd))
for cs in cases:
    for _ in range(2):
        ops.attention(*cs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for cs in cases:
    ops.attention(*cs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled", len(cases), "launches")
