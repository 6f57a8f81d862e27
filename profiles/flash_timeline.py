"""Per-tile clock64 timeline of CTA (0,0,0) of the tcgen05 attention kernel at the level-0 self-attention shape
(l2d_flash_set_debug): where do the softmax warps and the MMA thread spend a key tile?  Run with L2D_FLASH_PINGPONG=0/1."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
lib = _lib.lib()
b, heads, sq, hd = 2, 8, 4096, 40
c = heads * hd
qkv = torch.randn(b * sq, 3 * c, device=dev).half()
q, k, v = qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:]
for _ in range(3):
    ops.attention(q, k, v, b, heads, sq, sq, hd)
tl = torch.zeros(1088, dtype=torch.int64, device=dev)
lib.l2d_flash_set_debug(tl.data_ptr())
ops.attention(q, k, v, b, heads, sq, sq, hd)
torch.cuda.synchronize()
lib.l2d_flash_set_debug(0)
t = tl.cpu()
sm = t[:1024].view(8, 16, 8).double()
mma = t[1024:].view(2, 16, 2).double()
base = float(sm[0, 2, 0])
names = ["wait S", "TMEM->regs", "max/lazy/PV wait", "token wait", "exp + P stores", "hand-over", "(loop)"]
print(f"pingpong={os.environ.get('L2D_FLASH_PINGPONG', '1')}  (cycles; CTA 0; tiles 2..14 averaged)")
for w in (0, 4, 1, 5):
    d = [float((sm[w, 2:15, i + 1] - sm[w, 2:15, i]).mean()) for i in range(6)]
    period = float((sm[w, 3:16, 0] - sm[w, 2:15, 0]).mean())
    print(f"softmax warp {w} (query tile {w // 4}): period {period:7.0f} | " + "  ".join(f"{n} {x:6.0f}" for n, x in zip(names, d)))
print("absolute times of tiles 4..7 relative to warp 0's tile-2 start:")
for j in range(4, 8):
    for w in (0, 4):
        print(f"  tile {j} warp {w}: " + " ".join(f"{float(sm[w, j, i]) - base:7.0f}" for i in range(7)),
              f"| MMA: S_{j} issued {float(mma[w // 4, j, 0]) - base:7.0f}  PV_{j} issued {float(mma[w // 4, j, 1]) - base:7.0f}")
