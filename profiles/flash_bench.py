"""Spatial attention kernel timing at the UNet's shapes (CUDA events, 20 launches each after warm-up).
L2D_FLASH_LEGACY=1 selects the mma.sync kernel (flash_attn.cu); default = tcgen05 kernel (flash_tcgen05.cu).

    python profiles/flash_bench.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200 import ops  # noqa: E402

dev = "cuda:0"
from live2diff_b200 import _lib  # noqa: E402

out = {"legacy": os.environ.get("L2D_FLASH_LEGACY", "0"),
       "ctas_per_sm": {hd: _lib.lib().l2d_debug_flash_ctas_per_sm(hd) for hd in (40, 80)}}
for name, b, heads, sq, skv, hd in [("l0_self", 2, 8, 4096, 4096, 40), ("l0_cross", 2, 8, 4096, 77, 40),
                                    ("l1_self", 2, 8, 1024, 1024, 80), ("l1_cross", 2, 8, 1024, 77, 80),
                                    ("cfg3_l0_self", 2, 8, 6144, 6144, 40), ("cfg4_l0_self", 4, 8, 4096, 4096, 40)]:
    c = heads * hd
    g = torch.Generator(device=dev).manual_seed(1)
    qkv = torch.randn(b * sq, 3 * c, device=dev, generator=g).half()
    kv = torch.randn(b * skv, 2 * c, device=dev, generator=g).half()
    q = qkv[:, :c]
    k, v = (qkv[:, c:2 * c], qkv[:, 2 * c:]) if skv == sq else (kv[:, :c], kv[:, c:])
    for _ in range(3):
        ops.attention(q, k, v, b, heads, sq, skv, hd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.attention(q, k, v, b, heads, sq, skv, hd)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    flops = 4.0 * b * heads * sq * skv * hd
    out[name] = {"us": round(us, 2), "tflops": round(flops / us / 1e6, 1)}
print(json.dumps(out))
