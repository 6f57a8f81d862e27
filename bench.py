"""Benchmark of the hot path: streaming frames/s at 512x512, 2 denoise steps (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" = one frame of the stream = one `predict_x0_batch`: stream-batch assembly, the whole UNet
step (N=2 rows, 64x64 latent, L=16 KV window, SD1.5 widths, fp16), the LCM x0 prediction with re-noise, the
buffer shift and the ring-schedule advance -- by default through `B200DeviceStream` (state resident in HBM,
one CUDA graph per frame; `--pipeline host` selects the host-scheduled `B200StreamPipeline`).  VAE/MiDaS are outside the hot path.  Synthetic latents, seeded random
weights with the real shapes (no checkpoints can be downloaded here).

Printed JSON (one line, rank 0):
  value        frames/s with inputs resident in HBM, CUDA-event timed, max over ranks, all ranks' frames
  e2e          frames/s through the public API from pinned HOST buffers (H2D of x_t + depth latent, frame graph,
               D2H of x0 into pinned memory, stream sync per frame) -- the number to compare with the reference arm
  roofline     the temporal KV-cache attention kernel (K1): algorithmic bytes / CUDA-event time of its 40
               launches inside a real (eager, event-bracketed) step, against the measured HBM peak
               (MEASURED_PEAKS.json); `traffic` = ncu DRAM bytes per launch (profiles/k1_traffic.json)
  cpu_baseline the oracle (CPU restatement of the reference UNet step, torch fp32) on the host cores
With --impl reference the same workload runs on the CPU oracle port only (rank 0), same JSON shape.
Multi-GPU: one independent stream per rank (weak scaling); the only collective is the
NCCL broadcast of the prompt embedding before the timed region.
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

METRIC = "stream_fps_512x512_2step"
UNIT = "frames/s"
T_INDEX = [30, 40]
LAT_H = LAT_W = 64
WINDOW, WARMUP_SLOTS = 16, 8


def workload_config(n_gpus):
    return {"workload": "BASELINE configs[1]: 512x512 stream (64x64 latent), 2 denoise steps (t_index [30,40] of 50), "
                        "KV window 16 (8 sink + 8 rolling), SD1.5 UNet widths (320,640,1280,1280), fp16, steady state",
            "streams": n_gpus, "stream_batch_rows": 2, "latent": [LAT_H, LAT_W], "window": WINDOW,
            "l2": "per-step working set (2.83 GiB KV cache + 2.56 GB weights) exceeds the 126 MB L2; no flush needed",
            "parallelism": f"{n_gpus} independent stream replicas, 1 per GPU; prompt embedding broadcast once (NCCL)"}


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


def cpu_oracle_fps(steps, warmup, budget_s):
    from live2diff_b200.weights import UNetDims, random_state_dict
    from oracle import schedule_oracle as S
    from oracle import unet_oracle as O

    cores = pick_cpu_threads()
    torch.set_num_threads(cores)
    d = UNetDims()
    od = O.UNetDims(**d.__dict__)
    sd = random_state_dict(d, seed=0)
    n = len(T_INDEX)
    gen = torch.Generator().manual_seed(1)
    kv = O.alloc_kv_cache(od, n, LAT_H, LAT_W)
    for c in kv:
        c.normal_(generator=gen)
    ab, pe, up = S.init_schedule(n, WINDOW, WARMUP_SLOTS)
    for _ in range(3 * WINDOW):
        S.update_schedule(ab, pe, up, WINDOW, WARMUP_SLOTS)
    ctx = torch.randn(n, 77, d.cross_attention_dim, generator=gen)
    sub, c_skip, c_out, a, b = S.stream_constants(T_INDEX)
    x = torch.randn(n, 4, 1, LAT_H, LAT_W, generator=gen)
    dep = torch.randn(n, 4, 1, LAT_H, LAT_W, generator=gen)

    def one():
        with torch.no_grad():
            eps = O.unet_forward(sd, od, x, sub, ctx, ab, dep, kv, pe, up)
            S.scheduler_step_batch(eps, x, c_skip, c_out, a, b)
        S.update_schedule(ab, pe, up, WINDOW, WARMUP_SLOTS)

    t0 = time.perf_counter()
    one()                                           # first call (allocator / thread-pool warm-up), normally discarded
    first = time.perf_counter() - t0
    if first > budget_s:                            # bounded sample: do not spend minutes of box time on the CPU leg
        sample = (f"1 full UNet step (N=2, 64x64 latent, L=16) of the CPU oracle, torch fp32, {cores} threads, first call "
                  f"(no warm-up: it alone exceeded the {budget_s:.0f} s budget); {first:.2f} s/step")
        return 1.0 / first, first, cores, 1, sample
    w_eff = max(0, min(warmup - 1, int(budget_s * 0.2 / max(first, 1e-3))))
    for _ in range(w_eff):
        one()
    k_eff = max(1, min(steps, int(budget_s / max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(k_eff):
        one()
    dt = (time.perf_counter() - t0) / k_eff
    sample = (f"{k_eff} timed + {w_eff + 1} discarded full UNet steps (N=2, 64x64 latent, L=16) of the CPU oracle, "
              f"torch fp32, {cores} threads; {dt:.2f} s/step")
    return 1.0 / dt, dt, cores, k_eff, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fps, dt, cores, k_eff, sample = cpu_oracle_fps(args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": k_eff,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference arm = the reference's UNet step restated for CPU (oracle/, torch fp32): the reference "
                    "package itself hard-codes CUDA in its pipeline and is absent from the GPU box"}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def run_native(args):
    import torch.distributed as dist

    from live2diff_b200 import _lib
    from live2diff_b200.device_stream import B200DeviceStream
    from live2diff_b200.stream_pipeline import B200StreamPipeline, broadcast_prompt
    from live2diff_b200.unet_step import B200UNetStep
    from live2diff_b200.weights import UNetDims, random_state_dict

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

    d = UNetDims()
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

    if stream is not pipe:                         # real frames until every ring slot is valid (steady state)
        for i in range(3 * WINDOW):
            stream(dev_x[i % n_frames_pool], dev_d[i % n_frames_pool])
    # ---------------- device-resident throughput ----------------
    for i in range(max(args.warmup, 3)):
        stream(dev_x[i % n_frames_pool], dev_d[i % n_frames_pool])
    barrier()
    l0 = lib.l2d_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        for i in range(args.steps):
            stream(dev_x[i % n_frames_pool], dev_d[i % n_frames_pool])
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
        traffic = None
        tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch_avg")     # ncu dram read+write, per launch
        roofline = {"kernel": "kv_attn_mma_kernel (K1, 40 launches/step)", "bound": "hbm", "achieved": achieved,
                    "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": which,
                    "algorithmic_bytes_per_step": k1_bytes, "algorithmic_bytes_per_launch": k1_bytes / fam_cnt["kv_attn"],
                    "ms_per_step_in_kernel": k1_ms, "us_per_launch": k1_ms * 1e3 / fam_cnt["kv_attn"], "traffic": traffic}
        gemm_flops = 2.227e12 * 0.93                        # all dense contractions except the 3 attention families
        breakdown = {f: {"ms": round(fam_ms[f], 4), "launches": fam_cnt[f]} for f in fam_ms}
        tensor = {"kernel": "gemm_f16_tcgen05_kernel (all linears + convs)", "bound": "tensor",
                  "achieved": gemm_flops / (fam_ms["gemm"] * 1e-3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
                  "frac": gemm_flops / (fam_ms["gemm"] * 1e-3) / 1e12 / tf_peak if tf_peak else None,
                  "note": "flops = 0.93 x 2.227 TFLOP/step (SURVEY §6 split), time = CUDA events around the GEMM launches "
                          "of an eager step; peak = measured cuBLAS bf16 sustained"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
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
                "engine_device_gib": round(unet.device_bytes / 2 ** 30, 2)}
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            fps, dt, cores, k_eff, sample = cpu_oracle_fps(2, 1, budget_s=25.0)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the (~1 min) CPU oracle leg")
    ap.add_argument("--pipeline", default="device", choices=["device", "host"],
                    help="device: B200DeviceStream (state in HBM, whole-frame graph); host: B200StreamPipeline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
