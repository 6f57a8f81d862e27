"""Benchmark of the hot path: streaming frames/s (BASELINE.json configs; default = configs[1], the headline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4] [--impl native|reference|torch-fp16]

A "step" = one frame of the stream = one `predict_x0_batch`: stream-batch assembly, the whole UNet step (SD1.5 widths,
fp16), the LCM x0 prediction with re-noise, the buffer shift and the ring-schedule advance -- by default through
`B200DeviceStream` (state resident in HBM, one CUDA graph per frame; `--pipeline host` selects the host-scheduled
`B200StreamPipeline`).  Synthetic latents, seeded random weights with the real shapes (no checkpoints can be downloaded).
  --config 2   512x512 (64x64 latent), 2 denoise rows, KV window 16      (BASELINE configs[1], the metric's config)
  --config 3   768x512 (64x96 latent), 2 denoise rows, KV window 16      (configs[2]; the depth latent is synthetic)
  --config 4   512x512, 4 denoise rows, KV window 32 (long-cache stress) (configs[3])

Printed JSON (one line, rank 0):
  value        frames/s with inputs resident in HBM, CUDA-event timed, max over ranks, all ranks' frames
  e2e          frames/s through the public API from pinned HOST buffers (H2D of x_t + depth latent, frame graph,
               D2H of x0 into pinned memory, stream sync per frame) -- the number to compare with the reference arm
  roofline     the temporal KV-cache attention kernel (K1): algorithmic bytes / CUDA-event time of its 40 launches
               inside a real (eager, event-bracketed) step, against the measured HBM peak (MEASURED_PEAKS.json);
               `frac_in_graph` = the same bytes / the kernel's marginal cost inside the whole-frame graph replay (frame
               time with and without its launches, `l2d_unet_set_ablation`), where PDL overlaps its prologue with the
               QKV GEMM's tail; `traffic` = ncu DRAM bytes per launch (profiles/k1_traffic.json, regenerated per round)
  cpu_baseline the oracle (CPU restatement of the reference UNet step, torch fp32) on the host cores
  torch_fp16_eager / vs_torch_fp16_eager (N=1): the reference's evaluation order restated with torch ops, run EAGERLY
               in fp16 on the same GPU (cuDNN / cuBLAS / SDPA kernels, host-synchronising ring schedule like the
               reference's update_attn_bias) = the stand-in for "the reference, TensorRT off, fp16, on this box"
               (the reference package itself cannot travel to the GPU box); BASELINE's target is >= 2x this.
With --impl reference the same workload runs on the CPU oracle port only (rank 0: ONE CPU stream; the line says so in
`streams`, so a multi-GPU ratio against it is N GPU streams vs 1 CPU stream); --impl torch-fp16 prints the eager-torch
GPU arm alone, same JSON shape.
Multi-GPU: one independent stream per rank (weak scaling); the only collective is the NCCL broadcast of the
prompt embedding before the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

UNIT = "frames/s"
CONFIGS = {
    2: dict(metric="stream_fps_512x512_2step", h=64, w=64, t_index=[30, 40], window=16, sink=8, pe_max=24,
            name="BASELINE configs[1]: 512x512 stream (64x64 latent), 2 denoise steps (t_index [30,40] of 50), KV window 16 "
                 "(8 sink + 8 rolling), SD1.5 UNet widths (320,640,1280,1280), fp16, steady state"),
    3: dict(metric="stream_fps_768x512_2step", h=64, w=96, t_index=[30, 40], window=16, sink=8, pe_max=24,
            name="BASELINE configs[2]: 768x512 stream (64x96 latent), 2 denoise steps, KV window 16, fp16, steady state; "
                 "the depth latent is synthetic (the MiDaS prior is outside the UNet step)"),
    4: dict(metric="stream_fps_512x512_4step_L32", h=64, w=64, t_index=[25, 31, 37, 43], window=32, sink=8, pe_max=32,
            name="BASELINE configs[3]: 512x512 stream, 4 denoise steps (t_index [25,31,37,43]), KV window 32 "
                 "(8 sink + 24 rolling; long-cache stress), fp16, steady state"),
}
CFG = CONFIGS[2]


def unet_dims():
    from live2diff_b200.weights import UNetDims

    return UNetDims(window_size=CFG["window"], sink_size=CFG["sink"], pe_max_len=CFG["pe_max"])


def workload_config(n_gpus, cpu_streams=None):
    from live2diff_b200.weights import UNetDims  # noqa: F401

    d = unet_dims()
    n = len(CFG["t_index"])
    kv_gib = sum(s[0] * s[1] * s[2] * s[3] * s[4] for s in d.kv_cache_shapes(n, CFG["h"], CFG["w"])) * 2 / 2 ** 30
    return {"workload": CFG["name"],
            "streams": n_gpus if cpu_streams is None else cpu_streams, "stream_batch_rows": n,
            "latent": [CFG["h"], CFG["w"]], "window": CFG["window"],
            "l2": f"per-step working set ({kv_gib:.2f} GiB KV cache + 2.56 GB weights) exceeds the 126 MB L2; no flush needed",
            "parallelism": f"{n_gpus} independent stream replicas, 1 per GPU; prompt embedding broadcast once (NCCL)"
                           if cpu_streams is None else f"{cpu_streams} CPU stream on rank 0 (N GPUs requested: {n_gpus})"}


def unet_gemm_flops(d, n, h, w):
    """Dense-contraction flops of one UNet step that run on the tcgen05 GEMM kernel (every Linear, 1x1 and 3x3 conv;
    attention matmuls and the once-per-prompt projections excluded), counted from the layer shapes."""
    c = d.block_out_channels
    nlev = len(c)

    def lin(m, k, nn):
        return 2.0 * m * k * nn

    def resnet(m, cin, cout):
        return lin(m, 9 * cin, cout) + lin(m, 9 * cout, cout) + (lin(m, cin, cout) if cin != cout else 0.0)

    m0 = n * h * w
    mc = d.mapping_channels
    fl = lin(m0, 64, c[0]) + lin(m0, 64, mc[0]) + lin(m0, 9 * mc[-1], c[0]) + lin(m0, 9 * c[0], 8)
    for i in range(len(mc) - 1):
        fl += lin(m0, 9 * mc[i], mc[i]) + lin(m0, 9 * mc[i], mc[i + 1])
    out_ch = c[0]
    for bi in range(nlev):
        in_ch, out_ch = out_ch, c[bi]
        m = n * (h >> bi) * (w >> bi)
        for li in range(d.layers_per_block):
            fl += resnet(m, in_ch if li == 0 else out_ch, out_ch) + 22 * lin(m, out_ch, out_ch)   # + temporal transformer
            if d.down_has_attn[bi]:
                fl += 20 * lin(m, out_ch, out_ch)                                                # spatial transformer
        if bi != nlev - 1:
            fl += lin(m // 4, 9 * out_ch, out_ch)
    m = n * (h >> (nlev - 1)) * (w >> (nlev - 1))
    fl += 2 * resnet(m, c[-1], c[-1]) + 20 * lin(m, c[-1], c[-1])
    out_ch = c[-1]
    for bi in range(nlev):
        prev_out, out_ch = out_ch, c[nlev - 1 - bi]
        in_ch = c[max(nlev - 2 - bi, 0)]
        lvl = nlev - 1 - bi
        m = n * (h >> lvl) * (w >> lvl)
        for li in range(d.layers_per_block + 1):
            fl += resnet(m, (prev_out if li == 0 else out_ch) + (in_ch if li == d.layers_per_block else out_ch), out_ch)
            fl += 22 * lin(m, out_ch, out_ch)
            if d.up_has_attn[bi]:
                fl += 20 * lin(m, out_ch, out_ch)
        if bi != nlev - 1:
            fl += lin(m * 4, 9 * out_ch, out_ch)
    return fl


# --------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 0))), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def k1_algorithmic_bytes(dims, n_rows, h, w):
    """per launch N*hw*C*(2L+4)*2 B (read K and V windows, write new k,v, read q, write out)."""
    return sum(s[0] * s[2] * s[4] * (2 * s[3] + 4) * 2 for s in dims.kv_cache_shapes(n_rows, h, w))


# --------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline and --impl reference)
# --------------------------------------------------------------------------------------------------
def pick_cpu_threads():
    """Threads the CPU leg actually uses: the host may expose far more logical CPUs than the container may
    run on (cgroup quota / affinity), and oversubscribing torch's pool is catastrophically slow -- so the
    count is chosen by a 2-second fp32 GEMM probe over candidate thread counts."""
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            avail = max(1, min(avail, int(float(q) / float(per) + 0.5)))
    except (OSError, ValueError):
        pass
    cands = sorted({c for c in (1, 2, 4, 8, 16, 32, 64, 128, avail) if c <= avail})
    a = torch.randn(1024, 1024)
    best, best_t = 1, float("inf")
    for c in cands:
        torch.set_num_threads(c)
        a @ a
        t0 = time.perf_counter()
        for _ in range(3):
            a @ a
        dt = time.perf_counter() - t0
        if dt < best_t * 0.9:
            best, best_t = c, dt
    return best


def _oracle_inputs(n, gen, device="cpu", dtype=torch.float32):
    d = unet_dims()
    x = torch.randn(n, 4, 1, CFG["h"], CFG["w"], generator=gen)
    dep = torch.randn(n, 4, 1, CFG["h"], CFG["w"], generator=gen)
    ctx = torch.randn(n, 77, d.cross_attention_dim, generator=gen)
    return x.to(device=device, dtype=dtype), dep.to(device=device, dtype=dtype), ctx.to(device=device, dtype=dtype)


def cpu_oracle_fps(steps, warmup, budget_s):
    from live2diff_b200.weights import random_state_dict
    from oracle import schedule_oracle as S
    from oracle import unet_oracle as O

    cores = pick_cpu_threads()
    torch.set_num_threads(cores)
    d = unet_dims()
    od = O.UNetDims(**d.__dict__)
    sd = random_state_dict(d, seed=0)
    n = len(CFG["t_index"])
    L, W0 = CFG["window"], CFG["sink"]
    gen = torch.Generator().manual_seed(1)
    kv = O.alloc_kv_cache(od, n, CFG["h"], CFG["w"])
    for c in kv:
        c.normal_(generator=gen)
    ab, pe, up = S.init_schedule(n, L, W0)
    for _ in range(3 * L):
        S.update_schedule(ab, pe, up, L, W0)
    x, dep, ctx = _oracle_inputs(n, gen)
    sub, c_skip, c_out, a, b = S.stream_constants(CFG["t_index"])

    def one():
        with torch.no_grad():
            eps = O.unet_forward(sd, od, x, sub, ctx, ab, dep, kv, pe, up)
            S.scheduler_step_batch(eps, x, c_skip, c_out, a, b)
        S.update_schedule(ab, pe, up, L, W0)

    what = f"full UNet steps (N={n}, {CFG['h']}x{CFG['w']} latent, L={L}) of the CPU oracle, torch fp32, {cores} threads"
    t0 = time.perf_counter()
    one()                                           # first call (allocator / thread-pool warm-up), normally discarded
    first = time.perf_counter() - t0
    if first > budget_s:                            # bounded sample: do not spend minutes of box time on the CPU leg
        sample = (f"1 of the {what}, first call (no warm-up: it alone exceeded the {budget_s:.0f} s budget); "
                  f"{first:.2f} s/step")
        return 1.0 / first, first, cores, 1, sample
    w_eff = max(0, min(warmup - 1, int(budget_s * 0.2 / max(first, 1e-3))))
    for _ in range(w_eff):
        one()
    k_eff = max(1, min(steps, int(budget_s / max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(k_eff):
        one()
    dt = (time.perf_counter() - t0) / k_eff
    sample = f"{k_eff} timed + {w_eff + 1} discarded {what}; {dt:.2f} s/step"
    return 1.0 / dt, dt, cores, k_eff, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fps, dt, cores, k_eff, sample = cpu_oracle_fps(args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": CFG["metric"], "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": k_eff,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus, cpu_streams=1),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "streams_measured": 1, "value_per_stream": fps,
            "note": "reference arm = the reference's UNet step restated for CPU (oracle/, torch fp32): the reference "
                    "package itself hard-codes CUDA in its pipeline and is absent from the GPU box.  ONE CPU stream on "
                    "all usable host cores whatever --gpus says (the cores are shared, N CPU streams would each run 1/N "
                    "as fast): at N > 1 compare this value with the native arm's value / n_gpus (per-stream)"}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# torch fp16 eager arm on the GPU (the measured stand-in for "reference, TensorRT off, fp16, same box")
# --------------------------------------------------------------------------------------------------
def torch_fp16_eager_fps(dev, steps, warmup):
    """The reference's own evaluation order (oracle/unet_oracle.py restates live2diff's modules op by op with torch
    calls) run eagerly in fp16 on `dev`: conv2d -> cuDNN, Linear -> cuBLAS, attention -> F.scaled_dot_product_attention,
    the ring schedule advanced on the host with `.any()` syncs exactly like update_attn_bias (pipeline:416-438)."""
    from live2diff_b200.weights import random_state_dict
    from oracle import schedule_oracle as S
    from oracle import unet_oracle as O

    d = unet_dims()
    od = O.UNetDims(**d.__dict__)
    n = len(CFG["t_index"])
    L, W0 = CFG["window"], CFG["sink"]
    sd16 = {k: v.to(device=dev, dtype=torch.float16) for k, v in random_state_dict(d, seed=0).items()}
    gen = torch.Generator().manual_seed(1)
    kv = [torch.randn(s, device=dev, dtype=torch.float16) for s in d.kv_cache_shapes(n, CFG["h"], CFG["w"])]
    x1, d1, ctx = _oracle_inputs(1, gen, dev, torch.float16)
    noise = torch.randn(max(n - 1, 1), 4, 1, CFG["h"], CFG["w"], generator=gen).to(device=dev, dtype=torch.float16)
    orc = S.StreamOracle(lambda sm, t, **kw: O.unet_forward(sd16, od, sm, t, kw["encoder_hidden_states"],
                                                            kw["temporal_attention_mask"], kw["depth_sample"],
                                                            kw["kv_cache"], kw["pe_idx"], kw["update_idx"]),
                         kv, ctx.repeat(n, 1, 1), CFG["t_index"], (CFG["h"], CFG["w"]), window=L, warmup=W0,
                         dtype=torch.float16, device=dev)
    with torch.no_grad():
        for _ in range(max(warmup, 3) + 2 * L):      # fill the ring: steady state like the native arm
            orc.step(x1, d1, noise[: n - 1] if n > 1 else None)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = orc.step(x1, d1, noise[: n - 1] if n > 1 else None)
        e1.record()
        torch.cuda.synchronize(dev)
    assert torch.isfinite(out).all()
    ms = e0.elapsed_time(e1) / steps
    del sd16, kv, orc
    torch.cuda.empty_cache()
    return 1e3 / ms, ms


def run_torch_fp16(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    with ClockSampler(dev.index) as clk:
        fps, ms = torch_fp16_eager_fps(dev, args.steps, args.warmup)
    line = {"impl": "torch-fp16", "metric": CFG["metric"], "value": fps, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": workload_config(1), "clocks": clk.summary(),
            "gpu_launches": 0,
            "note": "eager torch fp16 restatement of the reference's UNet step + LCM step + host ring schedule on one GPU "
                    "(library kernels only: cuDNN / cuBLAS / SDPA); none of this repo's kernels run in this arm"}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def run_native(args):
    import torch.distributed as dist

    from live2diff_b200 import _lib
    from live2diff_b200.device_stream import B200DeviceStream
    from live2diff_b200.stream_pipeline import B200StreamPipeline, broadcast_prompt
    from live2diff_b200.unet_step import B200UNetStep
    from live2diff_b200.weights import random_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    d = unet_dims()
    T_INDEX, LAT_H, LAT_W, WINDOW = CFG["t_index"], CFG["h"], CFG["w"], CFG["window"]
    n = len(T_INDEX)
    sd = random_state_dict(d, seed=0)
    unet = B200UNetStep(sd, d, n, LAT_H, LAT_W, use_cuda_graph=True, device=dev)
    del sd
    pipe = B200StreamPipeline(unet, T_INDEX, seed=2 + rank)
    gen = torch.Generator().manual_seed(100 + rank)
    kv = unet.prepare_cache(n)
    for c in kv:                                   # steady state: every slot holds plausible K/V (warm-up + history)
        c.normal_(generator=torch.Generator(device=dev).manual_seed(7 + rank))
    prompt = torch.randn(1, 77, d.cross_attention_dim, generator=torch.Generator().manual_seed(42)) if rank == 0 else None
    prompt = broadcast_prompt(prompt, (1, 77, d.cross_attention_dim), dev, src=0)
    pipe.prepare(prompt, kv)                       # host-scheduled pipeline: only used for the per-family profile below
    for _ in range(3 * WINDOW):                    # run the ring schedule into steady state (all L slots valid)
        pipe.schedule.advance()
    # the measured path: device-resident stream state, one CUDA graph per frame
    stream = B200DeviceStream(unet, T_INDEX, seed=2 + rank) if args.pipeline == "device" else pipe
    if stream is not pipe:
        stream.prepare(prompt, kv)

    n_frames_pool = 16
    host_x = [torch.randn(1, 4, 1, LAT_H, LAT_W, generator=gen).half().pin_memory() for _ in range(n_frames_pool)]
    host_d = [torch.randn(1, 4, 1, LAT_H, LAT_W, generator=gen).half().pin_memory() for _ in range(n_frames_pool)]
    dev_x = [t.to(dev) for t in host_x]
    dev_d = [t.to(dev) for t in host_d]
    dev_out = torch.empty(1, 4, 1, LAT_H, LAT_W, dtype=torch.float16, device=dev)
    host_out = torch.empty(1, 4, 1, LAT_H, LAT_W, dtype=torch.float16).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def frame_dev(i):
        if stream is not pipe:
            stream(dev_x[i % n_frames_pool], dev_d[i % n_frames_pool], out=dev_out)
        else:
            stream(dev_x[i % n_frames_pool], dev_d[i % n_frames_pool])

    def timed_frames(k):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(k):
            frame_dev(i)
        ev1.record()
        torch.cuda.synchronize(dev)
        return ev0.elapsed_time(ev1)

    if stream is not pipe:                         # real frames until every ring slot is valid (steady state)
        for i in range(3 * WINDOW):
            frame_dev(i)
    # ---------------- device-resident throughput ----------------
    for i in range(max(args.warmup, 3)):
        frame_dev(i)
    barrier()
    l0 = lib.l2d_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        for i in range(args.steps):
            frame_dev(i)
        ev1.record()
        barrier()
    launches = lib.l2d_launch_count() - l0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    value = world * args.steps / (ms_total / 1e3)

    # ---------------- end to end from host buffers ----------------
    barrier()
    ev0.record()
    for i in range(args.steps):
        if stream is not pipe:                               # pinned host in -> H2D, frame graph, D2H -> pinned host out
            stream(host_x[i % n_frames_pool], host_d[i % n_frames_pool], out=host_out)
        else:
            x = host_x[i % n_frames_pool].to(dev, non_blocking=True)
            dd = host_d[i % n_frames_pool].to(dev, non_blocking=True)
            out = pipe(x, dd)
            host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()            # the caller consumes frame i before sending frame i+1
    ev1.record()
    barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1))
    e2e = world * args.steps / (e2e_ms / 1e3)
    bytes_in = 2 * host_x[0].numel() * 2
    bytes_out = host_out.numel() * 2

    # ---------------- image in -> image out (a1: TAESD encode x2 + frame + TAESD decode + uint8 pre/post), N = 1 ----------------
    e2e_image = None
    if world == 1 and stream is not pipe and not args.no_image:
        from live2diff_b200.image_pipeline import B200ImageStream
        from live2diff_b200.taesd import B200TinyVAE, random_taesd_state_dict

        vae = B200TinyVAE(random_taesd_state_dict(0), LAT_H * 8, LAT_W * 8, max_batch=1, device=dev)
        img_stream = B200ImageStream(unet, vae, T_INDEX, seed=2)
        img_stream.prepare(prompt, kv)
        frames_u8 = [torch.randint(0, 256, (LAT_H * 8, LAT_W * 8, 3), generator=gen, dtype=torch.uint8).pin_memory()
                     for _ in range(4)]
        depth_maps = [torch.rand(1, LAT_H * 8, LAT_W * 8, generator=gen).half().to(dev) for _ in range(4)]
        out_u8 = torch.empty(LAT_H * 8, LAT_W * 8, 3, dtype=torch.uint8).pin_memory()
        for i in range(3 * WINDOW + 3):                       # fill this stream's ring, warm the VAE kernels
            img_stream.frame_u8(frames_u8[i % 4], depth_maps[i % 4], out=out_u8)
        torch.cuda.synchronize(dev)
        k_img = max(10, min(args.steps, 40))
        l_img0 = lib.l2d_launch_count()
        ev0.record()
        for i in range(k_img):
            img_stream.frame_u8(frames_u8[i % 4], depth_maps[i % 4], out=out_u8)
            torch.cuda.current_stream().synchronize()
        ev1.record()
        torch.cuda.synchronize(dev)
        img_ms = ev0.elapsed_time(ev1) / k_img
        e2e_image = {"value": 1e3 / img_ms, "unit": UNIT, "ms_per_step": img_ms, "h2d_bytes_per_step": frames_u8[0].numel(),
                     "d2h_bytes_per_step": out_u8.numel(), "launches_per_step": int((lib.l2d_launch_count() - l_img0) / k_img),
                     "what": "uint8 frame (pinned host) -> preprocess -> TAESD encode (frame) + TAESD encode (depth map, supplied: "
                             "MiDaS is not built) -> add_noise -> stream frame -> TAESD decode -> uint8 frame (pinned host), "
                             "sync per frame; seeded random TAESD weights with the real shapes"}
        del img_stream, vae

    line = None
    if rank == 0:
        # ---------------- per-family CUDA-event profile of real steps (eager), K1 roofline ----------------
        fam_ms, fam_cnt = {}, {}
        prof_steps = 3
        for i in range(prof_steps + 1):
            pipe._x_cat[0:1].copy_(dev_x[i % n_frames_pool])
            pipe._d_cat[0:1].copy_(dev_d[i % n_frames_pool])
            pipe._upload_schedule()
            res = unet.profile_step(pipe._x_cat, pipe.sub_timesteps_tensor, encoder_hidden_states=pipe.prompt_embeds,
                                    temporal_attention_mask=pipe.attn_bias, depth_sample=pipe._d_cat,
                                    kv_cache=pipe.kv_cache_list, pe_idx=pipe.pe_idx, update_idx=pipe.update_idx)
            pipe.schedule.advance()
            if i == 0:
                continue                                   # first profiled step creates the events
            for f, (ms, cnt) in res.items():
                fam_ms[f] = fam_ms.get(f, 0.0) + ms / prof_steps
                fam_cnt[f] = cnt
        hbm_peak, tf_peak, which = peaks()
        k1_bytes = k1_algorithmic_bytes(d, n, LAT_H, LAT_W)
        k1_ms = fam_ms["kv_attn"]
        achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
        # marginal cost of K1 / the GEMM family inside the whole-frame graph replay (what they cost in the measured frame)
        in_graph = {}
        if stream is not pipe and not args.no_ablation:
            k_abl = max(10, min(args.steps, 40))
            timed_frames(3)
            base_ms = timed_frames(k_abl) / k_abl
            for fam, bit in (("kv_attn", 0), ("gemm", 1), ("spatial_attn", 2), ("norm", 3)):
                lib.l2d_unet_set_ablation(unet._handle, 1 << bit)
                stream.invalidate_graph()
                timed_frames(3)
                in_graph[fam] = base_ms - timed_frames(k_abl) / k_abl
            lib.l2d_unet_set_ablation(unet._handle, 0)
            stream.invalidate_graph()
            timed_frames(3)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tp) and args.config == 2:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch_avg")     # ncu dram read+write, per launch
        roofline = {"kernel": "kv_attn_mma_kernel (K1, %d launches/step)" % fam_cnt["kv_attn"], "bound": "hbm",
                    "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": which,
                    "algorithmic_bytes_per_step": k1_bytes, "algorithmic_bytes_per_launch": k1_bytes / fam_cnt["kv_attn"],
                    "ms_per_step_in_kernel": k1_ms, "us_per_launch": k1_ms * 1e3 / fam_cnt["kv_attn"], "traffic": traffic}
        if in_graph.get("kv_attn", 0) > 0:
            roofline["ms_per_step_in_graph"] = in_graph["kv_attn"]
            roofline["frac_in_graph"] = k1_bytes / (in_graph["kv_attn"] * 1e-3) / 1e9 / hbm_peak
        gemm_flops = unet_gemm_flops(d, n, LAT_H, LAT_W)
        breakdown = {f: {"ms": round(fam_ms[f], 4), "launches": fam_cnt[f]} for f in fam_ms}
        for f, v in in_graph.items():
            breakdown[f]["ms_in_graph"] = round(v, 4)
        tensor = {"kernel": "gemm_f16_tcgen05_kernel (all linears + convs)", "bound": "tensor",
                  "achieved": gemm_flops / (fam_ms["gemm"] * 1e-3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
                  "frac": gemm_flops / (fam_ms["gemm"] * 1e-3) / 1e12 / tf_peak if tf_peak else None,
                  "flops_per_step": gemm_flops,
                  "note": "flops counted from the layer shapes (bench.unet_gemm_flops), time = CUDA events around the GEMM "
                          "launches of an eager step; peak = measured cuBLAS bf16 sustained"}
        if in_graph.get("gemm", 0) > 0 and tf_peak:
            tensor["frac_in_graph"] = gemm_flops / (in_graph["gemm"] * 1e-3) / 1e12 / tf_peak
        line = {"metric": CFG["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": workload_config(world),
                "clocks": clk.summary(),
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_out,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(launches),
                "launches_per_step": int(stream.launches_per_frame if stream is not pipe else unet.launches_per_step),
                "pipeline": "device-resident stream state, whole frame = 1 CUDA graph (l2d_stream_frame)"
                            if stream is not pipe else "host-scheduled B200StreamPipeline, UNet step = 1 CUDA graph",
                "roofline": roofline, "roofline_tensor": tensor, "kernel_time_breakdown_ms": breakdown,
                "engine_device_gib": round(unet.device_bytes / 2 ** 30, 2), "value_per_stream": value / world}
        if e2e_image is not None:
            line["e2e_image"] = e2e_image
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_torch_baseline:
            del stream, pipe, unet, kv
            torch.cuda.empty_cache()
            t_fps, t_ms = torch_fp16_eager_fps(dev, max(5, min(args.steps, 30)), 3)
            line["torch_fp16_eager"] = {"value": t_fps, "unit": UNIT, "ms_per_step": t_ms,
                                        "what": "eager torch fp16 restatement of the reference step on the same GPU "
                                                "(cuDNN/cuBLAS/SDPA), same config; see --impl torch-fp16"}
            line["vs_torch_fp16_eager"] = e2e / t_fps
        if world == 1 and not args.no_cpu_baseline:
            fps, dt, cores, k_eff, sample = cpu_oracle_fps(2, 1, budget_s=25.0)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    global CFG
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "torch-fp16"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json configuration (2 = the metric's own, default)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the (~1 min) CPU oracle leg")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the eager torch fp16 GPU leg")
    ap.add_argument("--no-ablation", action="store_true", help="skip the in-graph marginal-cost passes")
    ap.add_argument("--no-image", action="store_true", help="skip the image-in/image-out (TAESD) leg")
    ap.add_argument("--pipeline", default="device", choices=["device", "host"],
                    help="device: B200DeviceStream (state in HBM, whole-frame graph); host: B200StreamPipeline")
    args = ap.parse_args()
    CFG = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-fp16":
        run_torch_fp16(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
