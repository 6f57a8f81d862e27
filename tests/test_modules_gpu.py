"""B2 (TemporalTransformer3DModel drop-in) and B1 (whole UNet step) parity on the GPU."""
import pytest
import torch

from helpers import dims_from, load_golden, prefixed, regen_weights, sub_spec
from live2diff_b200.weights import UNetDims, random_state_dict, unet_param_spec
from oracle import schedule_oracle as S
from oracle import unet_oracle as O
from parity import referee

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def frames(n_rows, window, warmup, count):
    ab, pe, up = S.init_schedule(n_rows, window, warmup)
    for _ in range(count):
        yield ab.clone(), pe.clone(), up.clone()
        S.update_schedule(ab, pe, up, window, warmup)


def odims(d):
    return O.UNetDims(**d.__dict__)


@pytest.mark.parametrize("tag", ["c64", "c320"])
def test_temporal_transformer_dropin_vs_reference_golden(tag):
    from live2diff_b200.modules import B200TemporalTransformer3DModel

    g = load_golden(f"temporal_transformer_{tag}.pt")
    ch, heads, L, W0 = g["ch"], g["heads"], g["window"], g["warmup"]
    d = UNetDims(block_out_channels=(ch,), heads=heads, window_size=L, sink_size=W0, pe_max_len=g["pe_max"],
                 down_has_attn=(False,), up_has_attn=(False,))
    pre = "down_blocks.0.motion_modules.0.temporal_transformer"
    w = regen_weights(sub_spec(d, pre), g["seed"], g["fingerprint"])
    tt = B200TemporalTransformer3DModel(in_channels=ch, num_attention_heads=heads, attention_head_dim=ch // heads,
                                        num_layers=1, temporal_position_encoding=True,
                                        temporal_position_encoding_max_len=g["pe_max"], attention_class_name="stream",
                                        attention_kwargs=dict(window_size=L, sink_size=W0), enable_streaming=True)
    missing, unexpected = tt.load_state_dict(w, strict=True)
    tt = tt.to(DEV).half()
    caches = []
    for i, a in enumerate(tt.transformer_blocks[0].attention_blocks):
        a.set_info(g["h"], g["w"])
        a.set_index(i)
        c = a.set_cache(g["n_rows"])
        c.copy_(g["cache0"][i].half())
        caches.append(c)
    sd16 = {k: v.to(DEV).half() for k, v in prefixed(w, "t").items()}
    caches16 = [c.half().to(DEV) for c in g["cache0"]]
    for f, (mask, pe_idx, update_idx) in enumerate(frames(g["n_rows"], L, W0, g["x"].shape[0])):
        x = g["x"][f].half().to(DEV)
        y = tt(x, temporal_attention_mask=mask.half().to(DEV), kv_cache=caches, pe_idx=pe_idx.to(DEV),
               update_idx=update_idx.to(DEV))
        y16 = O.temporal_transformer(sd16, "t", x[:, :, 0], caches16, mask.half().to(DEV), pe_idx.to(DEV),
                                     update_idx.to(DEV), odims(d))
        referee(y[:, :, 0], g["y"][f][:, :, 0], y16, f"temporal_transformer[{tag}] frame {f}")
    for i in range(2):
        referee(caches[i], g["cache_final"][i], caches16[i], f"temporal_transformer[{tag}] cache {i}")


def test_unet_tiny_stream_vs_reference_golden():
    """Whole streaming UNet (tiny channels, real topology) through fill phase and first wrap, vs the
    fixture produced by the reference's UNet3DConditionStreamingModel."""
    from live2diff_b200.unet_step import B200UNetStep

    g = load_golden("unet_tiny_stream.pt")
    d = dims_from(g["dims"])
    sd = regen_weights(unet_param_spec(d), g["seed"], g["fingerprint"])
    n, h, w = g["n_rows"], g["h"], g["w"]
    unet = B200UNetStep(sd, d, n, h, w, use_cuda_graph=False)
    kv = unet.prepare_cache(n)
    sd16 = {k: v.to(DEV).half() for k, v in sd.items()}
    kv16 = [torch.zeros_like(c) for c in kv]
    gen = torch.Generator().manual_seed(g["seed"] + 1)
    for c, c16 in zip(kv, kv16):
        r = torch.randn(c[:, :, :, : d.sink_size].shape, generator=gen).half().to(DEV)
        c[:, :, :, : d.sink_size] = r
        c16[:, :, :, : d.sink_size] = r
    ctx = g["ctx"].half().to(DEV)
    t = g["timesteps"].to(DEV)
    for f, (mask, pe_idx, update_idx) in enumerate(frames(n, d.window_size, d.sink_size, g["x"].shape[0])):
        x, dep = g["x"][f].half().to(DEV), g["depth"][f].half().to(DEV)
        m16, pi, ui = mask.half().to(DEV), pe_idx.to(DEV), update_idx.to(DEV)
        out = unet(x, t, depth_sample=dep, encoder_hidden_states=ctx, temporal_attention_mask=m16, kv_cache=kv,
                   pe_idx=pi, update_idx=ui)
        assert out["kv_cache"] is kv
        y16 = O.unet_forward(sd16, odims(d), x, t, ctx, m16, dep, kv16, pi, ui)
        referee(out["sample"], g["y"][f], y16, f"unet_tiny frame {f}", slack=2.5)
    referee(kv[12], g["kv_final_12"], kv16[12], "unet_tiny kv[12]", slack=2.5)
    sums = torch.tensor([float(c.double().sum()) for c in kv])
    sums16 = torch.tensor([float(c.double().sum()) for c in kv16])
    referee(sums, g["kv_sums"], sums16, "unet_tiny kv sums", slack=3.0, floor=2e-3)


def test_unet_sd15_widths_vs_reference_golden():
    """The engine at the real SD1.5 widths against OUTPUTS OF THE REFERENCE ITSELF (fixture unet_sd15_widths.pt: the
    reference's UNet3DConditionStreamingModel with `random_state_dict(UNetDims(), seed=0)` on a 16x16 latent, 3 frames of
    the fill phase).  The small latent also exercises the degenerate geometries: 2x2 pixels at the last level."""
    from live2diff_b200.unet_step import B200UNetStep

    g = load_golden("unet_sd15_widths.pt")
    d = UNetDims()
    n, h, w = g["n_rows"], g["h"], g["w"]
    sd = random_state_dict(d, seed=g["seed"])
    unet = B200UNetStep(sd, d, n, h, w, use_cuda_graph=False)
    sd16 = {k: v.to(DEV).half() for k, v in sd.items()}
    del sd
    kv = unet.prepare_cache(n)
    kv16 = [torch.zeros_like(c) for c in kv]
    gen = torch.Generator().manual_seed(g["seed"] + 101)
    for c, c16 in zip(kv, kv16):
        r = torch.randn(c[:, :, :, : d.sink_size].shape, generator=gen).half().to(DEV)
        c[:, :, :, : d.sink_size] = r
        c16[:, :, :, : d.sink_size] = r
    ctx = g["ctx"].half().to(DEV)
    t = g["timesteps"].to(DEV)
    for f, (mask, pe_idx, update_idx) in enumerate(frames(n, d.window_size, d.sink_size, g["x"].shape[0])):
        x, dep = g["x"][f].half().to(DEV), g["depth"][f].half().to(DEV)
        m16, pi, ui = mask.half().to(DEV), pe_idx.to(DEV), update_idx.to(DEV)
        out = unet(x, t, depth_sample=dep, encoder_hidden_states=ctx, temporal_attention_mask=m16, kv_cache=kv,
                   pe_idx=pi, update_idx=ui)["sample"]
        y16 = O.unet_forward(sd16, odims(d), x, t, ctx, m16, dep, kv16, pi, ui)
        referee(out, g["y"][f], y16, f"unet_sd15_widths frame {f}", slack=2.5)
    fr = g["x"].shape[0]
    for i, ref in g["kv_probe"].items():
        referee(kv[i][:, :, :4, d.sink_size:d.sink_size + fr + 1], ref, kv16[i][:, :, :4, d.sink_size:d.sink_size + fr + 1],
                f"unet_sd15_widths kv[{i}]", slack=2.5)


def _unet_case(d, n, h, w, graph, steps, tag, t_list, kv_probe=(0, 13, 39)):
    from live2diff_b200.unet_step import B200UNetStep

    L, sink = d.window_size, d.sink_size
    sd = random_state_dict(d, seed=0)
    unet = B200UNetStep(sd, d, n, h, w, use_cuda_graph=graph)
    sd32 = {k: v.to(DEV) for k, v in sd.items()}
    sd16 = {k: v.half() for k, v in sd32.items()}
    del sd
    gen = torch.Generator().manual_seed(1)
    kv = unet.prepare_cache(n)
    for c in kv:
        c.copy_(torch.randn(c.shape, generator=gen).half())
    kv32 = [c.float() for c in kv]
    kv16 = [c.clone() for c in kv]
    ab, pe, up = S.init_schedule(n, L, sink)
    for _ in range(L + 4):
        S.update_schedule(ab, pe, up, L, sink)
    ctx = torch.randn(n, 77, 768, generator=gen).half().to(DEV)
    t = torch.tensor(t_list, device=DEV)
    for step in range(steps):
        x = torch.randn(n, 4, 1, h, w, generator=gen).half().to(DEV)
        dep = torch.randn(n, 4, 1, h, w, generator=gen).half().to(DEV)
        m16, pi, ui = ab.half().to(DEV), pe.to(DEV), up.to(DEV)
        out = unet(x, t, depth_sample=dep, encoder_hidden_states=ctx, temporal_attention_mask=m16, kv_cache=kv,
                   pe_idx=pi, update_idx=ui)["sample"]
        y32 = O.unet_forward(sd32, odims(d), x.float(), t, ctx.float(), m16.float(), dep.float(), kv32, pi, ui)
        y16 = O.unet_forward(sd16, odims(d), x, t, ctx, m16, dep, kv16, pi, ui)
        referee(out, y32, y16, f"{tag} graph={graph} step {step}", slack=2.5)
        S.update_schedule(ab, pe, up, L, sink)
    # the caches evolved identically (same slots written, values within fp16 chain error)
    for i in kv_probe:
        referee(kv[i], kv32[i], kv16[i], f"{tag} kv[{i}]", slack=2.5)
    assert unet.launches_per_step > 0
    print(f"[info] {tag}: launches/step={unet.launches_per_step} engine bytes={unet.device_bytes / 2**30:.2f} GiB")


@pytest.mark.parametrize("graph", [False, True])
def test_unet_sd15_size_vs_oracle(graph):
    """BASELINE config 2 (512x512 -> 64x64 latent, N=2, L=16, SD1.5 widths, random weights): two steady-state
    steps against the fp32 oracle (evaluated on the GPU for speed), torch-fp16 restatement as referee."""
    _unet_case(UNetDims(), 2, 64, 64, graph, 3 if graph else 2, "unet_sd15", [399, 199])


def test_unet_config3_768x512_vs_oracle():
    """BASELINE config 3 geometry (768x512 -> 96x64 latent, non-square: h != w through conv halos, GroupNorm, the
    2x down/up samplers and the KV-cache pixel order), depth latent on, N=2, L=16."""
    _unet_case(UNetDims(), 2, 64, 96, True, 2, "unet_cfg3_96x64", [399, 199])


def test_unet_config4_long_cache_vs_oracle():
    """BASELINE config 4 (4 denoise rows, KV window 32 = 8 sink + 24 rolling; 11.3 GiB of cache): K1 runs its
    general-window kernel, every per-row schedule differs."""
    d = UNetDims(window_size=32, sink_size=8, pe_max_len=32)   # base_config.yaml has max_len 24: config 4 needs >= 32
    _unet_case(d, 4, 64, 64, True, 2, "unet_cfg4_N4_L32", [699, 499, 299, 99], kv_probe=(0, 39))


def test_stream_pipeline_vs_stream_oracle():
    """predict_x0_batch loop (stream batch + LCM step + ring schedule) against StreamOracle on the tiny UNet."""
    from live2diff_b200.stream_pipeline import B200StreamPipeline
    from live2diff_b200.unet_step import B200UNetStep

    d = UNetDims(block_out_channels=(64, 128, 128, 128), cross_attention_dim=96)
    n, h, w = 2, 16, 16
    sd = random_state_dict(d, seed=7)
    unet = B200UNetStep(sd, d, n, h, w, use_cuda_graph=True)
    pipe = B200StreamPipeline(unet, [30, 40])
    gen = torch.Generator().manual_seed(3)
    prompt = torch.randn(1, 77, 96, generator=gen)
    kv = unet.prepare_cache(n)
    for c in kv:
        c[:, :, :, :8] = torch.randn(c[:, :, :, :8].shape, generator=gen).half().to(DEV)
    kv32 = [c.float().cpu() for c in kv]
    pipe.prepare(prompt, kv)
    od = odims(d)

    def unet_fn(sample, timestep, encoder_hidden_states, temporal_attention_mask, depth_sample, kv_cache, pe_idx, update_idx):
        return O.unet_forward(sd, od, sample, timestep, encoder_hidden_states, temporal_attention_mask, depth_sample,
                              kv_cache, pe_idx, update_idx)

    orc = S.StreamOracle(unet_fn, kv32, prompt.half().float().repeat(n, 1, 1), [30, 40], (h, w))
    # the oracle runs with the fp16-rounded constants the reference would hold (prepare() casts them to fp16)
    orc.c_skip, orc.c_out, orc.a, orc.b = [v.half().float() for v in (orc.c_skip, orc.c_out, orc.a, orc.b)]
    for f in range(12):
        x = torch.randn(1, 4, 1, h, w, generator=gen).half()
        dep = torch.randn(1, 4, 1, h, w, generator=gen).half()
        noise = torch.randn(n - 1, 4, 1, h, w, generator=gen).half()
        out = pipe(x.to(DEV), dep.to(DEV), noise=noise.to(DEV))
        ref = orc.step(x.float(), dep.float(), noise.float())
        err = float((out.float().cpu() - ref).abs().max())
        scale = float(ref.abs().max())
        print(f"[parity] stream frame {f}: max-abs err {err:.3e} (scale {scale:.3e})")
        assert err <= 2e-2 * max(scale, 1.0), f"frame {f}"
        assert pipe.schedule.update_idx == orc.update_idx.tolist() and pipe.schedule.pe_idx == orc.pe_idx.tolist()
