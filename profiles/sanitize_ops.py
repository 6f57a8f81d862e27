"""Small invocations of every hand-written kernel family for compute-sanitizer (memcheck / racecheck), SURVEY.md §5:

    compute-sanitizer --tool memcheck python profiles/sanitize_ops.py
    compute-sanitizer --tool racecheck python profiles/sanitize_ops.py
Covers: tcgen05 GEMM (plain, split-K + finish, implicit conv3x3, GEGLU), tcgen05 flash attention (hd 40 / 80, 77 keys),
legacy flash attention, K1 tensor-core kernel (L = 16 and L = 32) and scalar kernel, GroupNorm / LayerNorm, TAESD
encode / decode, and three frames of the tiny streaming UNet through the device-resident stream (PDL on, CUDA graph).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200 import ops  # noqa: E402
from live2diff_b200.device_stream import B200DeviceStream  # noqa: E402
from live2diff_b200.taesd import B200TinyVAE, random_taesd_state_dict  # noqa: E402
from live2diff_b200.unet_step import B200UNetStep  # noqa: E402
from live2diff_b200.weights import UNetDims, random_state_dict  # noqa: E402

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g).half()

# GEMM family
a, w = r(256, 320), r(320, 320) * 0.05
ops.gemm(a, w, bias=r(320), residual=r(256, 320))
ops.gemm(r(128, 11520), r(1280, 11520) * 0.01, bias=r(1280))                      # split-K + finish
x = r(2, 16, 16, 64)
ops.conv3x3(x.reshape(-1, 64), 2, 16, 16, r(64, 9 * 64) * 0.05, bias=r(64))       # implicit conv
# attention
for hd, skv in ((40, 256), (80, 128), (40, 77), (160, 64)):
    c = 8 * hd
    ops.attention(r(2 * 128, c), r(2 * skv, c), r(2 * skv, c), 2, 8, 128, skv, hd)
ops.attention(r(2 * 256, 320), r(2 * 384, 320), r(2 * 384, 320), 2, 8, 256, 384, 40)   # two query tiles per CTA
# K1
for L, c, hw in ((16, 320, 32), (16, 1280, 8), (32, 320, 16), (32, 640, 8), (4, 64, 16), (32, 1280, 4)):
    n = 2
    cache = r(n, 2, hw, L, c)
    mask = torch.zeros(n, L, device=dev).half()
    pi = torch.arange(L, device=dev).repeat(n, 1)
    up = torch.tensor([L - 1, L // 2], device=dev)
    ops.kv_attn(r(n, hw, c), r(n, hw, c), r(n, hw, c), cache, r(L, c), r(L, c), r(L, c), mask, pi, up, 8)
# norms
ops.layernorm(r(64, 320), r(320), r(320))
ops.groupnorm(r(2 * 16 * 16, 320), r(320), r(320), 2, 16, 16, 32, 1e-5, silu=True)                              # one-kernel path
ops.groupnorm(r(2 * 8 * 8, 128), r(192), r(192), 2, 8, 8, 32, 1e-5, silu=True, x2=r(2 * 8 * 8, 64))             # concat sources
ops.groupnorm(r(2 * 8 * 8, 64), r(64), r(64), 2, 8, 8, 32, 1e-5, silu=True, im2col=True, stride=2)              # two-kernel path
# TAESD
vae = B200TinyVAE(random_taesd_state_dict(0), 64, 64, device=dev)
z = vae.encode(torch.rand(1, 3, 64, 64, device=dev).half() * 2 - 1).latents
vae.decode(z)
# tiny streaming UNet, device-resident stream, graph + PDL
d = UNetDims(block_out_channels=(64, 128, 128, 128), cross_attention_dim=96)
unet = B200UNetStep(random_state_dict(d, seed=1), d, 2, 16, 16, use_cuda_graph=True, device=dev)
ds = B200DeviceStream(unet, [30, 40])
ds.prepare(torch.randn(1, 77, 96), unet.prepare_cache(2))
for _ in range(3):
    out = ds(r(1, 4, 1, 16, 16), r(1, 4, 1, 16, 16))
torch.cuda.synchronize()
assert torch.isfinite(out.float()).all()
print("sanitize_ops: all kernels ran")
