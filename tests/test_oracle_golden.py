"""Pin the CPU oracle against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  fp32 vs fp32: tolerance covers only summation-order noise."""
import json
import os

import pytest
import torch

from helpers import GOLDEN, dims_from, load_golden, mask_from_valid, prefixed, regen_weights, sub_spec
from live2diff_b200.weights import UNetDims, unet_param_spec
from oracle import schedule_oracle as S
from oracle import unet_oracle as O

TOL = dict(rtol=1e-4, atol=2e-5)


def odims(d: UNetDims) -> O.UNetDims:
    return O.UNetDims(**d.__dict__)


def schedule_frames(n_rows, window, warmup, frames):
    ab, pe, up = S.init_schedule(n_rows, window, warmup)
    for _ in range(frames):
        yield ab.clone(), pe.clone(), up.clone()
        S.update_schedule(ab, pe, up, window, warmup)


@pytest.mark.parametrize("tag", ["c64_fill_wrap", "c320_hd40", "c128_L32_N4", "c64_N1_L4"])
def test_stream_attention_matches_reference(tag):
    g = load_golden(f"stream_attention_{tag}.pt")
    d = UNetDims(block_out_channels=(g["ch"],), heads=g["heads"], window_size=g["window"], sink_size=g["warmup"],
                 pe_max_len=g["pe_max"], down_has_attn=(False,), up_has_attn=(False,))
    pre = "down_blocks.0.motion_modules.0.temporal_transformer.transformer_blocks.0.attention_blocks.0"
    sd = prefixed(regen_weights(sub_spec(d, pre), g["seed"], g["fingerprint"]), "a")
    cache = g["cache0"].clone()
    frames = g["x"].shape[0]
    for f, (mask, pe_idx, update_idx) in enumerate(schedule_frames(g["n_rows"], g["window"], g["warmup"], frames)):
        y = O.stream_temporal_attention(sd, "a", g["x"][f], cache, mask, pe_idx, update_idx, odims(d))
        torch.testing.assert_close(y, g["y"][f], **TOL)
    torch.testing.assert_close(cache, g["cache_final"], **TOL)


@pytest.mark.parametrize("tag", ["c64", "c320"])
def test_temporal_transformer_matches_reference(tag):
    g = load_golden(f"temporal_transformer_{tag}.pt")
    d = UNetDims(block_out_channels=(g["ch"],), heads=g["heads"], window_size=g["window"], sink_size=g["warmup"],
                 pe_max_len=g["pe_max"], down_has_attn=(False,), up_has_attn=(False,))
    pre = "down_blocks.0.motion_modules.0.temporal_transformer"
    sd = prefixed(regen_weights(sub_spec(d, pre), g["seed"], g["fingerprint"]), "t")
    caches = [c.clone() for c in g["cache0"]]
    for f, (mask, pe_idx, update_idx) in enumerate(schedule_frames(g["n_rows"], g["window"], g["warmup"], g["x"].shape[0])):
        y = O.temporal_transformer(sd, "t", g["x"][f][:, :, 0], caches, mask, pe_idx, update_idx, odims(d))
        torch.testing.assert_close(y, g["y"][f][:, :, 0], **TOL)
    for c, cf in zip(caches, g["cache_final"]):
        torch.testing.assert_close(c, cf, **TOL)


def test_blocks_match_reference():
    g = load_golden("blocks_small.pt")
    d = dims_from(g["dims"])
    od = odims(d)
    for tag in ("res_same", "res_short"):
        c = g[tag]
        sd = prefixed(regen_weights(sub_spec(d, c["prefix"]), g["seed"], c["fingerprint"]), "r")
        y = O.resnet_block(sd, "r", c["x"][:, :, 0], c["temb"], od)
        torch.testing.assert_close(y, c["y"][:, :, 0], **TOL)
    c = g["down"]
    sd = prefixed(regen_weights(sub_spec(d, c["prefix"]), g["seed"], c["fingerprint"]), "s")
    torch.testing.assert_close(O.downsample(sd, "s", c["x"][:, :, 0]), c["y"][:, :, 0], **TOL)
    c = g["up"]
    sd = prefixed(regen_weights(sub_spec(d, c["prefix"]), g["seed"], c["fingerprint"]), "s")
    torch.testing.assert_close(O.upsample(sd, "s", c["x"][:, :, 0]), c["y"][:, :, 0], **TOL)
    c = g["mapping"]
    sd = prefixed(regen_weights(sub_spec(d, c["prefix"]), g["seed"], c["fingerprint"]), "m")
    torch.testing.assert_close(O.mapping_network(sd, "m", c["x"][:, :, 0], 6), c["y"][:, :, 0], **TOL)
    c = g["spatial"]
    sd = prefixed(regen_weights(sub_spec(d, c["prefix"]), g["seed"], c["fingerprint"]), "t")
    torch.testing.assert_close(O.spatial_transformer(sd, "t", c["x"][:, :, 0], c["ctx"], od), c["y"][:, :, 0], **TOL)


def test_unet_tiny_stream_matches_reference():
    g = load_golden("unet_tiny_stream.pt")
    d = dims_from(g["dims"])
    od = odims(d)
    sd = regen_weights(unet_param_spec(d), g["seed"], g["fingerprint"])
    kv = O.alloc_kv_cache(od, g["n_rows"], g["h"], g["w"])
    gen = torch.Generator().manual_seed(g["seed"] + 1)
    for c in kv:
        c[:, :, :, : d.sink_size] = torch.randn(c[:, :, :, : d.sink_size].shape, generator=gen)
    ctx = torch.randn(g["n_rows"], 77, d.cross_attention_dim, generator=gen)
    torch.testing.assert_close(ctx, g["ctx"])
    frames = g["x"].shape[0]
    for f, (mask, pe_idx, update_idx) in enumerate(schedule_frames(g["n_rows"], d.window_size, d.sink_size, frames)):
        y = O.unet_forward(sd, od, g["x"][f], g["timesteps"], ctx, mask, g["depth"][f], kv, pe_idx, update_idx)
        torch.testing.assert_close(y, g["y"][f], rtol=2e-4, atol=5e-5)
    sums = torch.tensor([float(c.double().sum()) for c in kv])
    torch.testing.assert_close(sums, g["kv_sums"], rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(kv[12], g["kv_final_12"], **TOL)
    torch.testing.assert_close(kv[39][0, :, :64], g["kv_final_39_row0"], **TOL)


_SD15 = {}


def sd15_weights(seed, fingerprint):
    """`random_state_dict(UNetDims(), seed)` (1.28 G parameters, ~15 s to generate) shared by the SD1.5-width tests."""
    from live2diff_b200.weights import random_state_dict, spec_fingerprint

    if seed not in _SD15:
        _SD15.clear()
        _SD15[seed] = random_state_dict(UNetDims(), seed=seed)
    sd = _SD15[seed]
    fp = spec_fingerprint({k: sd[k] for k in list(sd)[:40]})
    assert abs(fp - fingerprint) <= 1e-9 * abs(fingerprint), "seeded weights differ from the fixture's"
    return sd


def test_unet_sd15_widths_warmup_matches_reference():
    """The warm-up pass of the oracle at the real SD1.5 widths against the reference's UNet3DConditionWarmupModel."""
    g = load_golden("unet_sd15_widths_warmup.pt")
    d = UNetDims()
    od = odims(d)
    sd = sd15_weights(g["seed"], g["fingerprint"])
    rows = [torch.zeros(s_[1:]) for s_ in d.kv_cache_shapes(1, g["h"], g["w"])]
    y = O.unet_forward_warmup(sd, od, g["x"], g["timestep"], g["ctx"], g["depth"], rows)
    torch.testing.assert_close(y, g["y"], rtol=5e-4, atol=1e-4)
    torch.testing.assert_close(torch.tensor([float(c.double().abs().sum()) for c in rows]), g["kv_abs_sums"], rtol=1e-4,
                               atol=1e-2)
    for i, ref in g["kv_probe"].items():
        torch.testing.assert_close(rows[i][:, :4, : g["frames"]], ref, rtol=5e-4, atol=1e-4)


def test_unet_sd15_widths_matches_reference():
    """The oracle at the real SD1.5 widths (1.28 G parameters, head dims 40/80/160) on a 16x16 latent, against the
    fixture from the reference's UNet3DConditionStreamingModel loaded with the very weights the GPU parity tests and
    bench.py use (`random_state_dict(UNetDims(), seed=0)`): reference <-> oracle here, oracle <-> CUDA engine at the same
    widths in tests/test_modules_gpu.py."""
    g = load_golden("unet_sd15_widths.pt")
    d = UNetDims()
    od = odims(d)
    sd = sd15_weights(g["seed"], g["fingerprint"])
    n, h, w = g["n_rows"], g["h"], g["w"]
    kv = O.alloc_kv_cache(od, n, h, w)
    gen = torch.Generator().manual_seed(g["seed"] + 101)
    for c in kv:
        c[:, :, :, : d.sink_size] = torch.randn(c[:, :, :, : d.sink_size].shape, generator=gen)
    ctx = torch.randn(n, 77, d.cross_attention_dim, generator=gen)
    torch.testing.assert_close(ctx, g["ctx"])
    frames = g["x"].shape[0]
    for f, (mask, pe_idx, update_idx) in enumerate(schedule_frames(n, d.window_size, d.sink_size, frames)):
        y = O.unet_forward(sd, od, g["x"][f], g["timesteps"], ctx, mask, g["depth"][f], kv, pe_idx, update_idx)
        torch.testing.assert_close(y, g["y"][f], rtol=5e-4, atol=1e-4)
    torch.testing.assert_close(torch.tensor([float(c.double().abs().sum()) for c in kv]), g["kv_abs_sums"], rtol=1e-4,
                               atol=1e-2)
    for i, ref in g["kv_probe"].items():
        torch.testing.assert_close(kv[i][:, :, :4, d.sink_size:d.sink_size + frames + 1], ref, rtol=5e-4, atol=1e-4)


@pytest.mark.parametrize("tag", ["c64_f8", "c320_f8", "c128_f2_L4"])
def test_warmup_attention_matches_reference(tag):
    """VersatileAttention (bidirectional, writes cache slots 0..F-1) -- SURVEY.md §8f-1."""
    g = load_golden(f"warmup_attention_{tag}.pt")
    d = UNetDims(block_out_channels=(g["ch"],), heads=g["heads"], window_size=g["window"], pe_max_len=g["pe_max"],
                 down_has_attn=(False,), up_has_attn=(False,))
    pre = "down_blocks.0.motion_modules.0.temporal_transformer.transformer_blocks.0.attention_blocks.0"
    sd = prefixed(regen_weights(sub_spec(d, pre), g["seed"], g["fingerprint"]), "a")
    kv_row = torch.zeros(2, g["hw"], g["window"], g["ch"])
    y = O.warmup_temporal_attention(sd, "a", g["x"], kv_row, odims(d))
    torch.testing.assert_close(y, g["y"], **TOL)
    torch.testing.assert_close(kv_row, g["kv_row"], **TOL)
    assert float(kv_row[:, :, g["frames"]:].abs().max()) == 0.0          # slots >= F untouched


def test_unet_tiny_warmup_then_stream_matches_reference():
    """Warm-up UNet once per denoise row on `cache[idx]`, then streaming steps on the caches it filled."""
    g = load_golden("unet_tiny_warmup.pt")
    d = dims_from(g["dims"])
    od = odims(d)
    sd = regen_weights(unet_param_spec(d), g["seed"], g["fingerprint"])
    n = g["n_rows"]
    kv = O.alloc_kv_cache(od, n, g["h"], g["w"])
    for idx in range(n):
        y = O.unet_forward_warmup(sd, od, g["x"][idx], g["timesteps"][idx].view(1), g["ctx"], g["depth"][idx],
                                  [c[idx] for c in kv])
        torch.testing.assert_close(y, g["y"][idx], rtol=2e-4, atol=5e-5)
    torch.testing.assert_close(torch.tensor([float(c.double().sum()) for c in kv]), g["kv_sums"], rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(torch.tensor([float(c.double().abs().sum()) for c in kv]), g["kv_abs_sums"], rtol=1e-4,
                               atol=1e-2)
    for i, ref in g["kv_after_warmup"].items():
        torch.testing.assert_close(kv[i][:, :, :16], ref, **TOL)
    ctx = g["ctx"].repeat(n, 1, 1)
    frames = g["stream_x"].shape[0]
    for f, (mask, pe_idx, update_idx) in enumerate(schedule_frames(n, d.window_size, d.sink_size, frames)):
        y = O.unet_forward(sd, od, g["stream_x"][f], g["timesteps"], ctx, mask, g["stream_depth"][f], kv, pe_idx,
                           update_idx)
        torch.testing.assert_close(y, g["stream_y"][f], rtol=2e-4, atol=5e-5)


@pytest.mark.parametrize("tag", ["tiny", "sd15"])
def test_param_spec_matches_reference_state_dict(tag):
    ref = json.load(open(os.path.join(GOLDEN, f"state_dict_spec_{tag}.json")))
    d = UNetDims() if tag == "sd15" else UNetDims(block_out_channels=(64, 128, 128, 128), cross_attention_dim=96)
    mine = {k: list(v) for k, v in unet_param_spec(d).items()}
    assert set(mine) == set(ref)
    for k in ref:
        assert mine[k] == ref[k], k
    assert len(d.kv_cache_shapes(2, 64, 64)) == 40
