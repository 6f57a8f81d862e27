"""the warm-up pass on the device -- bidirectional temporal attention + sink-slot fill (kernel),
the warm-up UNet engine (`B200UNetWarmup` at `stream.unet_warmup`), and the hand-off to the streaming step."""
import pytest
import torch
import torch.nn.functional as F

from helpers import dims_from, load_golden, prefixed, regen_weights, sub_spec
from live2diff_b200.weights import UNetDims, random_state_dict, unet_param_spec
from oracle import schedule_oracle as S
from oracle import unet_oracle as O
from parity import referee

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def odims(d):
    return O.UNetDims(**d.__dict__)


@pytest.mark.parametrize("tag", ["c64_f8", "c320_f8", "c128_f2_L4"])
def test_warmup_attention_kernel_vs_reference_golden(tag):
    """l2d_warmup_attn fed the bias-free projections, against the fixture of the reference's VersatileAttention."""
    from live2diff_b200 import ops

    g = load_golden(f"warmup_attention_{tag}.pt")
    ch, heads, hw, f, L = g["ch"], g["heads"], g["hw"], g["frames"], g["window"]
    d = UNetDims(block_out_channels=(ch,), heads=heads, window_size=L, pe_max_len=g["pe_max"], down_has_attn=(False,),
                 up_has_attn=(False,))
    pre = "down_blocks.0.motion_modules.0.temporal_transformer.transformer_blocks.0.attention_blocks.0"
    w32 = prefixed(regen_weights(sub_spec(d, pre), g["seed"], g["fingerprint"]), "a")
    w16 = {k: v.half().to(DEV) for k, v in w32.items()}
    x16 = g["x"].half().to(DEV)
    # projections + PE tables as fp16 Linears (what the engine's fused QKV GEMM / prepare_pe_buffer produce)
    wqkv = torch.cat([w16["a.to_q.weight"], w16["a.to_k.weight"], w16["a.to_v.weight"]], 0)
    qkv = F.linear(x16, wqkv).contiguous()                                   # [F,hw,3C]
    pe_tab = F.linear(w16["a.pos_encoder.pe"][0, :f], wqkv).contiguous()     # [F,3C]
    row = torch.zeros(2, hw, L, ch, dtype=torch.float16, device=DEV)
    o = ops.warmup_attn(qkv[..., :ch], qkv[..., ch:2 * ch], qkv[..., 2 * ch:], row, pe_tab[:, :ch], pe_tab[:, ch:2 * ch],
                        pe_tab[:, 2 * ch:], heads, qkv_ld=3 * ch, pe_ld=3 * ch)
    y = F.linear(o, w16["a.to_out.0.weight"], w16["a.to_out.0.bias"])
    # referee: the oracle in fp16 on the GPU (the reference's own fp16 evaluation order)
    row16 = torch.zeros_like(row)
    y16 = O.warmup_temporal_attention(w16, "a", x16, row16, odims(d))
    referee(y, g["y"], y16, f"warmup_attn {tag}")
    # the cache fill is a copy of the projected k / v: bit-exact, and slots >= F stay untouched
    assert torch.equal(row[0, :, :f], qkv[..., ch:2 * ch].transpose(0, 1))
    assert torch.equal(row[1, :, :f], qkv[..., 2 * ch:].transpose(0, 1))
    assert float(row[:, :, f:].abs().max()) == 0.0 if f < L else True
    referee(row, g["kv_row"], row16, f"warmup_attn {tag} cache row")


def test_unet_tiny_warmup_then_stream_vs_reference_golden():
    """Warm-up UNet (F = 8 frames on the batch axis) once per denoise row on `cache[idx]`, then streaming steps on
    the caches it filled -- against the fixture from the reference's UNet3DConditionWarmupModel + StreamingModel."""
    from live2diff_b200.unet_step import B200UNetStep
    from live2diff_b200.unet_warmup import B200UNetWarmup

    g = load_golden("unet_tiny_warmup.pt")
    d = dims_from(g["dims"])
    sd = regen_weights(unet_param_spec(d), g["seed"], g["fingerprint"])
    n, h, w, f = g["n_rows"], g["h"], g["w"], g["frames"]
    warm = B200UNetWarmup(sd, d, f, h, w)
    unet = B200UNetStep(sd, d, n, h, w, use_cuda_graph=False)
    kv = unet.prepare_cache(n)
    sd16 = {k: v.to(DEV).half() for k, v in sd.items()}
    kv16 = [torch.zeros_like(c) for c in kv]
    ctx = g["ctx"].half().to(DEV)
    for idx in range(n):
        x, dep = g["x"][idx].half().to(DEV), g["depth"][idx].half().to(DEV)
        t = g["timesteps"][idx].view(1).to(DEV)
        out = warm(x, t, temporal_attention_mask=None, depth_sample=dep, encoder_hidden_states=ctx,
                   kv_cache=[c[idx] for c in kv], return_dict=True)["sample"]
        y16 = O.unet_forward_warmup(sd16, odims(d), x, t, ctx, dep, [c[idx] for c in kv16])
        referee(out, g["y"][idx], y16, f"unet_tiny warmup row {idx}", slack=2.5)
    for i, ref in g["kv_after_warmup"].items():
        referee(kv[i][:, :, :16], ref, kv16[i][:, :, :16], f"unet_tiny warmup kv[{i}]", slack=2.5)
        assert float(kv[i][:, :, :, f:].abs().max()) == 0.0                      # only the sink slots were written
    sums = torch.tensor([float(c.double().abs().sum()) for c in kv])
    sums16 = torch.tensor([float(c.double().abs().sum()) for c in kv16])
    referee(sums, g["kv_abs_sums"], sums16, "unet_tiny warmup kv |sums|", slack=3.0, floor=2e-3)
    # hand-off: the streaming engine continues on the caches the warm-up engine filled
    ctx_s = ctx.repeat(n, 1, 1)
    ts = g["timesteps"].to(DEV)
    ab, pe, up = S.init_schedule(n, d.window_size, d.sink_size)
    for fr in range(g["stream_x"].shape[0]):
        x, dep = g["stream_x"][fr].half().to(DEV), g["stream_depth"][fr].half().to(DEV)
        m16, pi, ui = ab.half().to(DEV), pe.to(DEV), up.to(DEV)
        out = unet(x, ts, depth_sample=dep, encoder_hidden_states=ctx_s, temporal_attention_mask=m16, kv_cache=kv,
                   pe_idx=pi, update_idx=ui)["sample"]
        y16 = O.unet_forward(sd16, odims(d), x, ts, ctx_s, m16, dep, kv16, pi, ui)
        referee(out, g["stream_y"][fr], y16, f"unet_tiny stream-after-warmup frame {fr}", slack=2.5)
        S.update_schedule(ab, pe, up, d.window_size, d.sink_size)


def test_unet_sd15_warmup_vs_oracle():
    """SD1.5 widths, 8-frame clip at 32x32 latent: one warm-up pass against the fp32 oracle evaluated on the GPU."""
    from live2diff_b200.unet_warmup import B200UNetWarmup

    d = UNetDims()
    f, h, w = d.sink_size, 32, 32
    sd = random_state_dict(d, seed=0)
    warm = B200UNetWarmup(sd, d, f, h, w)
    sd32 = {k: v.to(DEV) for k, v in sd.items()}
    sd16 = {k: v.half() for k, v in sd32.items()}
    del sd
    gen = torch.Generator().manual_seed(5)
    rows = [torch.zeros(s[1:], dtype=torch.float16, device=DEV) for s in d.kv_cache_shapes(1, h, w)]
    rows32 = [r.float() for r in rows]
    rows16 = [r.clone() for r in rows]
    x = torch.randn(1, 4, f, h, w, generator=gen).half().to(DEV)
    dep = torch.randn(1, 4, f, h, w, generator=gen).half().to(DEV)
    ctx = torch.randn(1, 77, 768, generator=gen).half().to(DEV)
    t = torch.tensor([399], device=DEV)
    out = warm(x, t, depth_sample=dep, encoder_hidden_states=ctx, kv_cache=rows)["sample"]
    y32 = O.unet_forward_warmup(sd32, odims(d), x.float(), t, ctx.float(), dep.float(), rows32)
    y16 = O.unet_forward_warmup(sd16, odims(d), x, t, ctx, dep, rows16)
    referee(out, y32, y16, "unet_sd15 warmup", slack=2.5)
    for i in (0, 13, 39):
        referee(rows[i], rows32[i], rows16[i], f"unet_sd15 warmup kv row[{i}]", slack=2.5)
    print(f"[info] warm-up engine bytes={warm.device_bytes / 2**30:.2f} GiB")


def test_unet_sd15_widths_warmup_vs_reference_golden():
    """The warm-up engine at the real SD1.5 widths against OUTPUTS OF THE REFERENCE's UNet3DConditionWarmupModel
    (fixture unet_sd15_widths_warmup.pt: 8-frame clip, 16x16 latent, row 0 of zero caches)."""
    from helpers import load_golden
    from live2diff_b200.unet_warmup import B200UNetWarmup

    g = load_golden("unet_sd15_widths_warmup.pt")
    d = UNetDims()
    f, h, w = g["frames"], g["h"], g["w"]
    sd = random_state_dict(d, seed=g["seed"])
    warm = B200UNetWarmup(sd, d, f, h, w)
    sd16 = {k: v.to(DEV).half() for k, v in sd.items()}
    del sd
    rows = [torch.zeros(s_[1:], dtype=torch.float16, device=DEV) for s_ in d.kv_cache_shapes(1, h, w)]
    rows16 = [r.clone() for r in rows]
    x, dep, ctx = g["x"].half().to(DEV), g["depth"].half().to(DEV), g["ctx"].half().to(DEV)
    t = g["timestep"].to(DEV)
    out = warm(x, t, depth_sample=dep, encoder_hidden_states=ctx, kv_cache=rows)["sample"]
    y16 = O.unet_forward_warmup(sd16, odims(d), x, t, ctx, dep, rows16)
    referee(out, g["y"], y16, "unet_sd15_widths warmup", slack=2.5)
    for i, ref in g["kv_probe"].items():
        referee(rows[i][:, :4, :f], ref, rows16[i][:, :4, :f], f"unet_sd15_widths warmup kv row[{i}]", slack=2.5)


def test_pipeline_warmup_loop_vs_oracle():
    """B200StreamPipeline.warmup_denoise (pipeline:315-338): N passes with LCM x0 prediction and injected re-noise, against the
    same loop over the CPU oracle; then one streaming frame on the warmed caches."""
    from live2diff_b200.stream_pipeline import B200StreamPipeline
    from live2diff_b200.unet_step import B200UNetStep
    from live2diff_b200.unet_warmup import B200UNetWarmup

    d = UNetDims(block_out_channels=(64, 128, 128, 128), cross_attention_dim=96)
    n, h, w, f = 2, 16, 16, d.sink_size
    sd = random_state_dict(d, seed=9)
    unet = B200UNetStep(sd, d, n, h, w, use_cuda_graph=False)
    # the warm-up engine is a view over the streaming engine's repacked weights (l2d_unet_create_shared): workspace only
    warm = B200UNetWarmup(None, d, f, h, w, share_weights_with=unet)
    own = B200UNetWarmup(sd, d, f, h, w)
    assert warm.device_bytes < 0.6 * own.device_bytes, (warm.device_bytes, own.device_bytes)
    pipe = B200StreamPipeline(unet, [30, 40])
    gen = torch.Generator().manual_seed(4)
    prompt = torch.randn(1, 77, 96, generator=gen).half()
    pipe.prepare(prompt)
    x = torch.randn(1, 4, f, h, w, generator=gen).half()
    dep = torch.randn(1, 4, f, h, w, generator=gen).half()
    noise = [torch.randn(1, 4, f, h, w, generator=gen).half()]
    x0 = pipe.warmup_denoise(warm, x.to(DEV), dep.to(DEV), noise=[z.to(DEV) for z in noise])
    # the engine with its own weight copy gives the same bits as the shared-weight view
    rows_a = [torch.zeros(s_[1:], dtype=torch.float16, device=DEV) for s_ in d.kv_cache_shapes(1, h, w)]
    rows_b = [r.clone() for r in rows_a]
    t1 = torch.tensor([399], device=DEV)
    ya = warm(x.to(DEV), t1, depth_sample=dep.to(DEV), encoder_hidden_states=prompt.to(DEV), kv_cache=rows_a)["sample"]
    yb = own(x.to(DEV), t1, depth_sample=dep.to(DEV), encoder_hidden_states=prompt.to(DEV), kv_cache=rows_b)["sample"]
    assert torch.equal(ya, yb) and all(torch.equal(p_, q_) for p_, q_ in zip(rows_a, rows_b))
    del own
    # oracle loop (fp32, fp16-rounded constants like the reference's prepare())
    od = odims(d)
    sub, c_skip, c_out, a, b = [v.half().float() if v.is_floating_point() else v for v in S.stream_constants([30, 40])]
    kv32 = O.alloc_kv_cache(od, n, h, w)
    xt = x.float()
    for idx in range(n):
        eps = O.unet_forward_warmup(sd, od, xt, sub[idx].view(1), prompt.float(), dep.float(), [c[idx] for c in kv32])
        f_theta = (xt - b[idx] * eps) / a[idx]
        x0_ref = c_out[idx] * f_theta + c_skip[idx] * xt
        if idx < n - 1:
            xt = a[idx + 1] * x0_ref + b[idx + 1] * noise[idx].float()
    ref = x0_ref[0].transpose(0, 1)
    err = float((x0.float().cpu() - ref).abs().max())
    scale = float(ref.abs().max())
    print(f"[parity] pipeline warm-up x0: max-abs err {err:.3e} (scale {scale:.3e})")
    assert tuple(x0.shape) == (f, 4, h, w) and err <= 2e-2 * max(scale, 1.0)
    for i in (0, 21, 39):
        e = float((pipe.kv_cache_list[i].float().cpu() - kv32[i]).abs().max())
        assert e <= 2e-2 * max(float(kv32[i].abs().max()), 1.0), f"kv[{i}] after warm-up: {e}"
    out = pipe(torch.randn(1, 4, 1, h, w, generator=gen).half().to(DEV), torch.randn(1, 4, 1, h, w, generator=gen).half().to(DEV))
    assert torch.isfinite(out).all()
