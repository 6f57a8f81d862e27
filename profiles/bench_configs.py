"""Throughput of the other BASELINE.json single-GPU configurations (parity-tested in tests/test_modules_gpu.py; the
bench line itself is config 2): config 3 = 768x512 (96x64 latent), N=2, L=16; config 4 = 512x512, N=4 denoise rows,
KV window 32 (long-cache stress: 11.3 GiB of cache, K1 runs its general-window kernel).  Same method as bench.py:
B200DeviceStream, whole-frame CUDA graph, CUDA events, steady state; K1 time from the event-bracketed family profile.

    python profiles/bench_configs.py [frames]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from live2diff_b200.device_stream import B200DeviceStream  # noqa: E402
from live2diff_b200.stream_pipeline import B200StreamPipeline  # noqa: E402
from live2diff_b200.unet_step import B200UNetStep  # noqa: E402
from live2diff_b200.weights import UNetDims, random_state_dict  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
CONFIGS = [("config3_768x512_N2_L16", UNetDims(), [30, 40], 64, 96),
           ("config4_512x512_N4_L32", UNetDims(window_size=32, sink_size=8, pe_max_len=32), [25, 31, 37, 43], 64, 64)]
out = {}
for name, d, t_index, h, w in CONFIGS:
    n = len(t_index)
    unet = B200UNetStep(random_state_dict(d, seed=0), d, n, h, w, use_cuda_graph=True, device=dev)
    ds = B200DeviceStream(unet, t_index)
    kv = unet.prepare_cache(n)
    for c in kv:
        c.normal_()
    prompt = torch.randn(1, 77, 768)
    ds.prepare(prompt, kv)
    x = torch.randn(1, 4, 1, h, w, device=dev).half()
    dep = torch.randn(1, 4, 1, h, w, device=dev).half()
    for _ in range(2 * d.window_size + 4):                 # real frames until every ring slot is valid
        ds(x, dep)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(frames):
        ds(x, dep)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / frames
    # K1 by the event-bracketed family profile of an eager step on the same state
    pipe = B200StreamPipeline(unet, t_index)
    pipe.prepare(prompt, kv)
    for _ in range(3 * d.window_size):
        pipe.schedule.advance()
    pipe._x_cat[0:1].copy_(x)
    pipe._d_cat[0:1].copy_(dep)
    pipe._upload_schedule()
    res = None
    for _ in range(2):
        res = unet.profile_step(pipe._x_cat, pipe.sub_timesteps_tensor, encoder_hidden_states=pipe.prompt_embeds,
                                temporal_attention_mask=pipe.attn_bias, depth_sample=pipe._d_cat,
                                kv_cache=pipe.kv_cache_list, pe_idx=pipe.pe_idx, update_idx=pipe.update_idx)
    k1_bytes = sum(s[0] * s[2] * s[4] * (2 * s[3] + 4) * 2 for s in d.kv_cache_shapes(n, h, w))
    k1_ms = res["kv_attn"][0]
    out[name] = {"ms_per_frame": round(ms, 3), "frames_per_s": round(1e3 / ms, 2), "launches_per_frame": int(ds.launches_per_frame),
                 "kv_cache_gib": round(sum(c.numel() for c in kv) * 2 / 2 ** 30, 2),
                 "k1": {"algorithmic_gb_per_step": round(k1_bytes / 1e9, 3), "ms_per_step": round(k1_ms, 3),
                        "gbs": round(k1_bytes / k1_ms / 1e6, 1), "frac_of_hbm_peak": round(k1_bytes / k1_ms / 1e6 / HBM, 3)},
                 "families_ms": {f: round(v[0], 3) for f, v in res.items()}}
    del ds, pipe, unet, kv
    torch.cuda.empty_cache()
print(json.dumps(out))
