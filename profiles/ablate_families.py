"""True critical-path cost of each kernel family inside the CUDA-graph replay of one UNet step (config 2): replay the
step with that family's launches removed (l2d_unet_set_ablation) and report the drop in ms/step.  ncu's per-launch
times are cold-cache and serialised; CUDA events around eager launches include launch gaps; this is the number that
says what a faster kernel family would actually buy.  Outputs are garbage while a family is ablated (timing only).

    python profiles/ablate_families.py [steps]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200 import _lib  # noqa: E402
from live2diff_b200.stream_pipeline import B200StreamPipeline  # noqa: E402
from live2diff_b200.unet_step import B200UNetStep  # noqa: E402
from live2diff_b200.weights import UNetDims, random_state_dict  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda:0")
d = UNetDims()
unet = B200UNetStep(random_state_dict(d, seed=0), d, 2, 64, 64, use_cuda_graph=True, device=dev)
pipe = B200StreamPipeline(unet, [30, 40])
kv = unet.prepare_cache(2)
for c in kv:
    c.normal_()
pipe.prepare(torch.randn(1, 77, 768), kv)
for _ in range(48):
    pipe.schedule.advance()
x = torch.randn(1, 4, 1, 64, 64, device=dev).half()
dep = torch.randn(1, 4, 1, 64, 64, device=dev).half()


def ms_per_step(mask):
    _lib.lib().l2d_unet_set_ablation(unet._handle, mask)
    for _ in range(4):
        pipe(x, dep)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        pipe(x, dep)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, unet.launches_per_step


full, n_full = ms_per_step(0)
out = {"full_ms": round(full, 4), "launches": int(n_full), "families": {}}
for name, mask in (("kv_attn", 1), ("gemm", 2), ("spatial_attn", 4), ("norm_all", 8), ("layernorm", 64), ("groupnorm", 128),
                   ("im2col", 16)):
    ms, n = ms_per_step(mask)
    out["families"][name] = {"ms_without": round(ms, 4), "cost_ms": round(full - ms, 4), "share": round((full - ms) / full, 4),
                             "launches_removed": int(n_full - n)}
_lib.lib().l2d_unet_set_ablation(unet._handle, 0)
print(json.dumps(out))
