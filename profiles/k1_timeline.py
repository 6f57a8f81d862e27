"""Per-CTA phase breakdown of the K1 (L == 16) kernel via l2d_kv_attn_set_debug, plus isolated timing per level."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
lib = _lib.lib()
for (n, hw, c) in [(2, 4096, 320), (2, 1024, 640), (2, 256, 1280), (2, 64, 1280)]:
    L, heads = 16, 8
    q, k, v = [torch.randn(n, hw, c, device=dev).half() for _ in range(3)]
    caches = [torch.randn(n, 2, hw, L, c, device=dev).half() for _ in range(3)]     # rotate: no L2 reuse at level 0
    pe = [torch.randn(L, c, device=dev).half() for _ in range(3)]
    mask = torch.zeros(n, L, device=dev).half()
    pi = torch.arange(L, device=dev).repeat(n, 1)
    up = torch.tensor([9, 12][:n], device=dev)
    for i in range(3):
        ops.kv_attn(q, k, v, caches[i % 3], pe[0], pe[1], pe[2], mask, pi, up, heads)
    torch.cuda.synchronize()
    # 12 launches replayed from a CUDA graph: no Python / launch-queue time inside the measurement
    gr = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr, stream=side):
            for i in range(12):
                ops.kv_attn(q, k, v, caches[i % 3], pe[0], pe[1], pe[2], mask, pi, up, heads)
    torch.cuda.current_stream().wait_stream(side)
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 12 * 1e3
    gb = n * hw * c * (2 * L + 4) * 2 / 1e9
    tl = torch.zeros(296, 16, dtype=torch.int64, device=dev)
    lib.l2d_kv_attn_set_debug(tl.data_ptr())
    ops.kv_attn(q, k, v, caches[0], pe[0], pe[1], pe[2], mask, pi, up, heads)
    torch.cuda.synchronize()
    lib.l2d_kv_attn_set_debug(0)
    t = tl.cpu().double()
    t = t[t[:, 4] > 0]
    per = t.sum(0) / t[:, 4].sum()
    names = {0: "waitK", 1: "patchK+bar", 2: "qk", 3: "waitV+patchV", 5: "bar2", 6: "softmax", 7: "pv+bar", 8: "gather", 9: "release"}
    body = " ".join(f"{v} {per[k]:.0f}" for k, v in names.items())
    tot = sum(float(per[k]) for k in names)
    span = (t[:, 13].max() - t[:, 12].min()) / 1e3
    cta = t[:, 11].mean()
    print(f"   kernel span (globaltimer) {span:.1f} us; CTA lifetime {cta:.0f} cycles, prologue {t[:, 10].mean():.0f}, "
          f"loop {float((t[:, [0,1,2,3,5,6,7,8,9]].sum(1)).mean()):.0f}")
    print(f"N{n} hw{hw} C{c}: {us:7.1f} us  {gb / us * 1e6:7.0f} GB/s | cycles per tile (thread 0): {body} | total {tot:.0f}  "
          f"(tiles/CTA {float(t[:, 4].mean()):.1f})")
