"""Isolated timing of the tcgen05 GEMM / implicit conv at the UNet's shapes (CUDA events, 30 launches each,
operands rotated through 4 buffers so successive launches do not hit L2-resident outputs)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200 import ops  # noqa: E402

dev = "cuda:0"


def bench(fn, iters=30):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


print(f"{'shape':34s} {'tile':>5s} {'us':>8s} {'TFLOP/s':>8s}")
for (m, n, k, tag) in [(8192, 320, 320, "l0 linear"), (8192, 960, 320, "l0 qkv"), (8192, 2560, 320, "l0 geglu"),
                       (8192, 320, 1280, "l0 ff2"), (2048, 640, 640, "l1 linear"), (2048, 1920, 640, "l1 qkv"),
                       (2048, 5120, 640, "l1 geglu"), (2048, 640, 2560, "l1 ff2"), (512, 1280, 1280, "l2 linear"),
                       (512, 3840, 1280, "l2 qkv"), (512, 10240, 1280, "l2 geglu"), (512, 1280, 5120, "l2 ff2"),
                       (128, 1280, 1280, "l3 linear"), (128, 10240, 1280, "l3 geglu"), (154, 24960, 768, "ctx kv")]:
    a = [torch.randn(m, k, device=dev).half() for _ in range(4)]
    w = [(torch.randn(n, k, device=dev) / math.sqrt(k)).half() for _ in range(4)]
    res = torch.randn(m, n, device=dev).half()
    out = torch.empty(m, n, device=dev, dtype=torch.float16)
    bias = torch.randn(n, device=dev).half()
    us = bench(lambda i: ops.gemm(a[i % 4], w[i % 4], bias=bias, residual=res, out=out))
    print(f"{tag + f' {m}x{n}x{k}':34s} {ops.gemm_tile_n(m, n, k):5d} {us:8.1f} {2 * m * n * k / us / 1e6:8.1f}")

for (nimg, h, w_, cin, cout, tag) in [(2, 64, 64, 320, 320, "l0 conv"), (2, 64, 64, 960, 320, "l0 up conv"),
                                      (2, 32, 32, 640, 640, "l1 conv"), (2, 32, 32, 1920, 640, "l1 up conv"),
                                      (2, 16, 16, 1280, 1280, "l2 conv"), (2, 16, 16, 2560, 1280, "l2 up conv"),
                                      (2, 8, 8, 1280, 1280, "l3 conv"), (2, 8, 8, 2560, 1280, "l3 up conv")]:
    x = [torch.randn(nimg * h * w_, cin, device=dev).half() for _ in range(4)]
    wt = [(torch.randn(cout, 9 * cin, device=dev) / math.sqrt(9 * cin)).half() for _ in range(2)]
    res = torch.randn(nimg * h * w_, cout, device=dev).half()
    bias = torch.randn(cout, device=dev).half()
    us = bench(lambda i: ops.conv3x3(x[i % 4], nimg, h, w_, wt[i % 2], bias=bias, residual=res))
    fl = 2 * nimg * h * w_ * 9 * cin * cout
    print(f"{tag + f' {nimg}x{h}x{w_} {cin}->{cout}':34s} {'':5s} {us:8.1f} {fl / us / 1e6:8.1f}")
