"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

* Model-level fixtures: the reference's own classes (imported unmodified through
  tests/shim/diffusers, see _ref_import.py) are instantiated, loaded with the seeded weights of
  `live2diff_b200.weights.random_tensors`, and run in fp32 on CPU on seeded inputs.  Fixtures
  store {seed, dims, inputs, outputs}; weights are regenerated from the seed by the tests (a
  fingerprint is stored to detect RNG drift).
* Pipeline-level fixtures (ring schedule, scheduler pointwise): the pipeline module cannot be
  imported (it needs the real diffusers LCMScheduler and hard-codes CUDA), so the *source text* of
  the individual methods is extracted from live2diff/pipeline_stream_animation_depth.py with `ast`
  and executed as-is against a stub `self` -- the reference's code is run, not copied.
"""
import ast
import json
import os
import sys
import textwrap
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from _ref_import import REF_ROOT, import_reference_models  # noqa: E402

from live2diff_b200.weights import UNetDims, random_tensors, spec_fingerprint, unet_param_spec  # noqa: E402

torch.set_grad_enabled(False)
torch.manual_seed(0)
M = import_reference_models()


def mm_kwargs(d: UNetDims):
    return dict(num_attention_heads=d.heads, num_transformer_block=1,
                attention_block_types=["Temporal_Self", "Temporal_Self"], temporal_position_encoding=True,
                temporal_position_encoding_max_len=d.pe_max_len, temporal_attention_dim_div=1,
                zero_initialize=False, attention_class_name="stream",
                attention_kwargs=dict(window_size=d.window_size, sink_size=d.sink_size))


def build_ref_unet(d: UNetDims, device="cpu"):
    U = M["unet_depth_streaming"].UNet3DConditionStreamingModel
    with torch.device(device):
        return U(block_out_channels=d.block_out_channels, cross_attention_dim=d.cross_attention_dim,
                 attention_head_dim=d.heads, cond_mapping=True, use_inflated_groupnorm=True, use_motion_module=True,
                 motion_module_resolutions=(1, 2, 4, 8), motion_module_type="Streaming",
                 motion_module_kwargs=mm_kwargs(d), norm_eps=d.norm_eps, layers_per_block=d.layers_per_block)


# ---------------------------------------------------------------------------------------------
# pipeline methods, executed from their source text
# ---------------------------------------------------------------------------------------------

def load_pipeline_methods(window: int, warmup: int):
    src = open(os.path.join(REF_ROOT, "live2diff", "pipeline_stream_animation_depth.py")).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "StreamAnimateDiffusionDepth")
    want = {"initialize_attn_bias_pe_and_update_idx", "update_attn_bias", "scheduler_step_batch", "add_noise"}
    ns = {"torch": torch, "WARMUP_FRAMES": warmup, "WINDOW_SIZE": window, "Optional": __import__("typing").Optional}
    for node in cls.body:
        if isinstance(node, ast.FunctionDef) and node.name in want:
            code = textwrap.dedent(ast.get_source_segment(src, node))
            exec(compile(code, f"<reference:{node.name}>", "exec"), ns)
    return {k: ns[k] for k in want}


def schedule_trace(n_rows, window, warmup, frames):
    fns = load_pipeline_methods(window, warmup)
    stub = types.SimpleNamespace(denoising_steps_num=n_rows, device="cpu", dtype=torch.float32)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self          # :410 calls .cuda(); no GPU here
    try:
        if n_rows == 1:
            # the reference raises IndexError for N == 1 (SURVEY A-1); record that fact
            try:
                fns["initialize_attn_bias_pe_and_update_idx"](stub)
                raised = False
            except IndexError:
                raised = True
            return {"n_rows": 1, "window": window, "warmup": warmup, "reference_raises_index_error": raised}
        ab, pe, up = fns["initialize_attn_bias_pe_and_update_idx"](stub)
    finally:
        torch.Tensor.cuda = orig_cuda
    tr = []
    for _ in range(frames):
        tr.append({"valid": (ab == 0).int().tolist(), "pe_idx": pe.tolist(), "update_idx": up.tolist()})
        ab, pe, up = fns["update_attn_bias"](stub, ab, pe, up)
    return {"n_rows": n_rows, "window": window, "warmup": warmup, "frames": tr}


def gen_schedule():
    out = [schedule_trace(2, 16, 8, 22), schedule_trace(4, 32, 8, 60), schedule_trace(3, 16, 8, 22),
           schedule_trace(2, 4, 2, 8), schedule_trace(1, 4, 2, 1)]
    json.dump(out, open(os.path.join(HERE, "schedule_trace.json"), "w"))
    print("schedule_trace.json", len(out))


def gen_scheduler_pointwise():
    fns = load_pipeline_methods(16, 8)
    g = torch.Generator().manual_seed(11)
    n = 3
    a = torch.rand(n, 1, 1, 1, 1, generator=g) * 0.5 + 0.4
    stub = types.SimpleNamespace(alpha_prod_t_sqrt=a, beta_prod_t_sqrt=(1 - a * a).sqrt(),
                                 c_skip=torch.rand(n, 1, 1, 1, 1, generator=g), c_out=torch.rand(n, 1, 1, 1, 1, generator=g))
    x = torch.randn(n, 4, 1, 8, 8, generator=g)
    eps = torch.randn(n, 4, 1, 8, 8, generator=g)
    out = fns["scheduler_step_batch"](stub, eps, x, None)
    noisy = fns["add_noise"](stub, x[:1], eps[:1], 1)
    torch.save({"a": a.flatten(), "b": stub.beta_prod_t_sqrt.flatten(), "c_skip": stub.c_skip.flatten(),
                "c_out": stub.c_out.flatten(), "x": x, "eps": eps, "x0": out, "add_noise_t1": noisy},
               os.path.join(HERE, "scheduler_pointwise.pt"))
    print("scheduler_pointwise.pt")


# ---------------------------------------------------------------------------------------------
# module-level fixtures
# ---------------------------------------------------------------------------------------------

def sub_spec(d: UNetDims, prefix: str):
    return {k[len(prefix) + 1:]: v for k, v in unet_param_spec(d).items() if k.startswith(prefix + ".")}


def run_schedule(n_rows, window, warmup, frames):
    tr = schedule_trace(n_rows, window, warmup, frames)["frames"]
    for f in tr:
        valid = torch.tensor(f["valid"], dtype=torch.bool)
        yield (torch.zeros(valid.shape).masked_fill_(~valid, float("-inf")), torch.tensor(f["pe_idx"]),
               torch.tensor(f["update_idx"]))


def gen_stream_attention(tag, ch, heads, h, w, n_rows, window, warmup, pe_max, frames, seed):
    d = UNetDims(block_out_channels=(ch,), heads=heads, window_size=window, sink_size=warmup, pe_max_len=pe_max,
                 down_has_attn=(False,), up_has_attn=(False,))
    pre = "down_blocks.0.motion_modules.0.temporal_transformer.transformer_blocks.0.attention_blocks.0"
    spec = sub_spec(d, pre)
    w_ = random_tensors(spec, seed=seed)
    A = M["stream_motion_module"].StreamTemporalAttention
    attn = A(attention_mode="Temporal", cross_attention_dim=None, query_dim=ch, heads=heads, dim_head=ch // heads,
             dropout=0.0, bias=False, upcast_attention=False, temporal_position_encoding=True,
             temporal_position_encoding_max_len=pe_max, window_size=window, sink_size=warmup)
    attn.load_state_dict(w_, strict=True)
    attn.set_info(h, w)
    attn.set_index(0)
    cache = attn.set_cache(n_rows)
    attn.prepare_pe_buffer()
    g = torch.Generator().manual_seed(seed + 1)
    cache[:, :, :, :warmup] = torch.randn(n_rows, 2, h * w, warmup, ch, generator=g)      # "warm-up" K/V
    cache0 = cache.clone()
    xs, ys = [], []
    for mask, pe_idx, update_idx in run_schedule(n_rows, window, warmup, frames):
        x = torch.randn(n_rows, h * w, ch, generator=g)
        y = attn(x, video_length=1, temporal_attention_mask=mask, kv_cache=cache, pe_idx=pe_idx, update_idx=update_idx)
        xs.append(x)
        ys.append(y.clone())
    torch.save({"ch": ch, "heads": heads, "h": h, "w": w, "n_rows": n_rows, "window": window, "warmup": warmup,
                "pe_max": pe_max, "seed": seed, "fingerprint": spec_fingerprint(w_), "cache0": cache0,
                "x": torch.stack(xs), "y": torch.stack(ys), "cache_final": cache.clone()},
               os.path.join(HERE, f"stream_attention_{tag}.pt"))
    print("stream_attention", tag, float(torch.stack(ys).abs().mean()))


def gen_temporal_transformer(tag, ch, heads, h, w, n_rows, window, warmup, pe_max, frames, seed):
    d = UNetDims(block_out_channels=(ch,), heads=heads, window_size=window, sink_size=warmup, pe_max_len=pe_max,
                 down_has_attn=(False,), up_has_attn=(False,))
    pre = "down_blocks.0.motion_modules.0.temporal_transformer"
    w_ = random_tensors(sub_spec(d, pre), seed=seed)
    T = M["motion_module"].TemporalTransformer3DModel
    kw = mm_kwargs(d)
    tt = T(in_channels=ch, num_attention_heads=heads, attention_head_dim=ch // heads, num_layers=1,
           attention_block_types=kw["attention_block_types"], temporal_position_encoding=True,
           temporal_position_encoding_max_len=pe_max, attention_class_name="stream",
           attention_kwargs=kw["attention_kwargs"], enable_streaming=True)
    tt.load_state_dict(w_, strict=True)
    caches = []
    for i, a in enumerate(tt.transformer_blocks[0].attention_blocks):
        a.set_info(h, w)
        a.set_index(i)
        caches.append(a.set_cache(n_rows))
        a.prepare_pe_buffer()
    g = torch.Generator().manual_seed(seed + 1)
    for c in caches:
        c[:, :, :, :warmup] = torch.randn(n_rows, 2, h * w, warmup, ch, generator=g)
    cache0 = [c.clone() for c in caches]
    xs, ys = [], []
    for mask, pe_idx, update_idx in run_schedule(n_rows, window, warmup, frames):
        x = torch.randn(n_rows, ch, 1, h, w, generator=g)
        y = tt(x, temporal_attention_mask=mask, kv_cache=caches, pe_idx=pe_idx, update_idx=update_idx)
        xs.append(x)
        ys.append(y.clone())
    torch.save({"ch": ch, "heads": heads, "h": h, "w": w, "n_rows": n_rows, "window": window, "warmup": warmup,
                "pe_max": pe_max, "seed": seed, "fingerprint": spec_fingerprint(w_), "cache0": cache0,
                "x": torch.stack(xs), "y": torch.stack(ys), "cache_final": [c.clone() for c in caches]},
               os.path.join(HERE, f"temporal_transformer_{tag}.pt"))
    print("temporal_transformer", tag, float(torch.stack(ys).abs().mean()))


def gen_resnet_and_friends(seed=5):
    d = UNetDims(block_out_channels=(64, 128), cross_attention_dim=96, down_has_attn=(True, False),
                 up_has_attn=(False, True))
    g = torch.Generator().manual_seed(seed)
    out = {"dims": d.__dict__, "seed": seed}
    R = M["resnet"]
    # ResnetBlock3D with shortcut (64 -> 128) and without (64 -> 64)
    for tag, pre, cin, cout in (("res_same", "down_blocks.0.resnets.0", 64, 64),
                                ("res_short", "down_blocks.1.resnets.0", 64, 128)):
        w_ = random_tensors(sub_spec(d, pre), seed=seed)
        blk = R.ResnetBlock3D(in_channels=cin, out_channels=cout, temb_channels=d.temb_dim, eps=d.norm_eps,
                              groups=32, use_inflated_groupnorm=True)
        blk.load_state_dict(w_, strict=True)
        x = torch.randn(2, cin, 1, 8, 8, generator=g)
        temb = torch.randn(2, d.temb_dim, generator=g)
        out[tag] = {"prefix": pre, "x": x, "temb": temb, "y": blk(x, temb), "fingerprint": spec_fingerprint(w_)}
    # samplers
    for tag, pre, cls in (("down", "down_blocks.0.downsamplers.0", R.Downsample3D),
                          ("up", "up_blocks.0.upsamplers.0", R.Upsample3D)):
        ch = 64 if tag == "down" else 128
        w_ = random_tensors(sub_spec(d, pre), seed=seed)
        m = cls(ch, use_conv=True, out_channels=ch) if tag == "up" else cls(ch, use_conv=True, out_channels=ch, padding=1)
        m.load_state_dict(w_, strict=True)
        x = torch.randn(2, ch, 1, 8, 8, generator=g)
        out[tag] = {"prefix": pre, "x": x, "y": m(x), "fingerprint": spec_fingerprint(w_)}
    # mapping network
    w_ = random_tensors(sub_spec(d, "flow_conv_in"), seed=seed)
    mp = R.MappingNetwork(conditioning_embedding_channels=64, conditioning_channels=4)
    mp.load_state_dict(w_, strict=True)
    x = torch.randn(2, 4, 1, 8, 8, generator=g)
    out["mapping"] = {"prefix": "flow_conv_in", "x": x, "y": mp(x), "fingerprint": spec_fingerprint(w_)}
    # spatial transformer
    pre = "down_blocks.0.attentions.0"
    w_ = random_tensors(sub_spec(d, pre), seed=seed)
    T = M["attention"].Transformer3DModel
    st = T(d.heads, 64 // d.heads, in_channels=64, num_layers=1, cross_attention_dim=96, norm_num_groups=32,
           unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
    st.load_state_dict(w_, strict=True)
    x = torch.randn(2, 64, 1, 6, 5, generator=g)
    ctx = torch.randn(2, 77, 96, generator=g)
    out["spatial"] = {"prefix": pre, "x": x, "ctx": ctx, "y": st(x, encoder_hidden_states=ctx).sample,
                      "fingerprint": spec_fingerprint(w_)}
    torch.save(out, os.path.join(HERE, "blocks_small.pt"))
    print("blocks_small.pt")


TINY = UNetDims(block_out_channels=(64, 128, 128, 128), cross_attention_dim=96)


def gen_unet_tiny(frames=11, seed=3, h=16, w=16, n_rows=2):
    d = TINY
    sd = random_tensors(unet_param_spec(d), seed=seed)
    u = build_ref_unet(d)
    missing, unexpected = u.load_state_dict(sd, strict=True)
    u.eval()
    u.set_info_for_attn(h, w)
    kv = u.prepare_cache(n_rows)
    g = torch.Generator().manual_seed(seed + 1)
    for c in kv:
        c[:, :, :, : d.sink_size] = torch.randn(c[:, :, :, : d.sink_size].shape, generator=g)
    ctx = torch.randn(n_rows, 77, d.cross_attention_dim, generator=g)
    t = torch.tensor([399, 199])
    xs, ds, ys = [], [], []
    for mask, pe_idx, update_idx in run_schedule(n_rows, d.window_size, d.sink_size, frames):
        x = torch.randn(n_rows, 4, 1, h, w, generator=g)
        dep = torch.randn(n_rows, 4, 1, h, w, generator=g)
        o = u(x, t, encoder_hidden_states=ctx, temporal_attention_mask=mask, depth_sample=dep, kv_cache=kv,
              pe_idx=pe_idx, update_idx=update_idx)
        xs.append(x)
        ds.append(dep)
        ys.append(o["sample"].clone())
    torch.save({"dims": d.__dict__, "seed": seed, "h": h, "w": w, "n_rows": n_rows, "timesteps": t, "ctx": ctx,
                "fingerprint": spec_fingerprint(sd), "x": torch.stack(xs), "depth": torch.stack(ds),
                "y": torch.stack(ys), "kv_sums": torch.tensor([float(c.double().sum()) for c in kv]),
                "kv_abs_sums": torch.tensor([float(c.double().abs().sum()) for c in kv]),
                "kv_final_12": kv[12].clone(), "kv_final_39_row0": kv[39][0, :, :64].clone()},
               os.path.join(HERE, "unet_tiny_stream.pt"))
    print("unet_tiny_stream.pt", float(torch.stack(ys).abs().mean()), float(torch.stack(ys).abs().max()))


def build_ref_warmup_unet(d: UNetDims):
    """UNet3DConditionWarmupModel configured like its from_pretrained_2d (unet_depth_warmup.py:611-614): motion
    module type "Vanilla", attention class "versatile", no attention kwargs."""
    U = M["unet_depth_warmup"].UNet3DConditionWarmupModel
    kw = mm_kwargs(d)
    kw["attention_class_name"] = "versatile"
    kw["attention_kwargs"] = {}
    return U(block_out_channels=d.block_out_channels, cross_attention_dim=d.cross_attention_dim,
             attention_head_dim=d.heads, cond_mapping=True, use_inflated_groupnorm=True, use_motion_module=True,
             motion_module_resolutions=(1, 2, 4, 8), motion_module_type="Vanilla",
             motion_module_kwargs=kw, norm_eps=d.norm_eps, layers_per_block=d.layers_per_block)


def gen_warmup_attention(tag, ch, heads, hw, frames, window, pe_max, seed):
    """VersatileAttention alone (motion_module.py:438-530): tokens [(b f), d, c] of one clip, cache row [2,hw,L,C]."""
    d = UNetDims(block_out_channels=(ch,), heads=heads, window_size=window, pe_max_len=pe_max,
                 down_has_attn=(False,), up_has_attn=(False,))
    pre = "down_blocks.0.motion_modules.0.temporal_transformer.transformer_blocks.0.attention_blocks.0"
    w_ = random_tensors(sub_spec(d, pre), seed=seed)
    A = M["motion_module"].VersatileAttention
    attn = A(attention_mode="Temporal", cross_attention_dim=None, query_dim=ch, heads=heads, dim_head=ch // heads,
             dropout=0.0, bias=False, upcast_attention=False, temporal_position_encoding=True,
             temporal_position_encoding_max_len=pe_max)
    attn.load_state_dict(w_, strict=True)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(frames, hw, ch, generator=g)
    kv_row = torch.zeros(2, hw, window, ch)
    y = attn(x, video_length=frames, kv_cache=kv_row)
    torch.save({"ch": ch, "heads": heads, "hw": hw, "frames": frames, "window": window, "pe_max": pe_max, "seed": seed,
                "fingerprint": spec_fingerprint(w_), "x": x, "y": y.clone(), "kv_row": kv_row.clone()},
               os.path.join(HERE, f"warmup_attention_{tag}.pt"))
    print("warmup_attention", tag, float(y.abs().mean()))


def gen_unet_tiny_warmup(seed=3, h=16, w=16, n_rows=2, stream_frames=3):
    """Warm-up hand-off (pipeline:315-338 around unet_warmup, then the streaming UNet on the caches it filled):
    the warm-up model runs once per denoise row idx on `cache[idx]` views, then `stream_frames` streaming steps
    continue on the same caches.  Same seeded weights for both models (identical state_dict key set)."""
    d = TINY
    sd = random_tensors(unet_param_spec(d), seed=seed)
    uw = build_ref_warmup_unet(d)
    uw.load_state_dict(sd, strict=True)
    uw.eval()
    us = build_ref_unet(d)
    us.load_state_dict(sd, strict=True)
    us.eval()
    us.set_info_for_attn(h, w)
    uw.set_info_for_attn(h, w)
    kv = us.prepare_cache(n_rows)
    f = d.sink_size
    g = torch.Generator().manual_seed(seed + 7)
    ctx = torch.randn(1, 77, d.cross_attention_dim, generator=g)
    ts = torch.tensor([399, 199])
    xs, ds, ys = [], [], []
    for idx in range(n_rows):
        x = torch.randn(1, 4, f, h, w, generator=g)
        dep = torch.randn(1, 4, f, h, w, generator=g)
        o = uw(x, ts[idx].view(1), temporal_attention_mask=None, depth_sample=dep, encoder_hidden_states=ctx,
               kv_cache=[c[idx] for c in kv], return_dict=True)["sample"]
        xs.append(x)
        ds.append(dep)
        ys.append(o.clone())
    kv_after_warmup = {i: kv[i][:, :, :16].clone() for i in (0, 1, 12, 27, 39)}     # first 16 pixels; sums cover the rest
    kv_sums = torch.tensor([float(c.double().sum()) for c in kv])
    kv_abs = torch.tensor([float(c.double().abs().sum()) for c in kv])
    sx, sdp, sy = [], [], []
    ctx_s = ctx.repeat(n_rows, 1, 1)
    for mask, pe_idx, update_idx in run_schedule(n_rows, d.window_size, d.sink_size, stream_frames):
        x = torch.randn(n_rows, 4, 1, h, w, generator=g)
        dep = torch.randn(n_rows, 4, 1, h, w, generator=g)
        o = us(x, ts, encoder_hidden_states=ctx_s, temporal_attention_mask=mask, depth_sample=dep, kv_cache=kv,
               pe_idx=pe_idx, update_idx=update_idx)
        sx.append(x)
        sdp.append(dep)
        sy.append(o["sample"].clone())
    torch.save({"dims": d.__dict__, "seed": seed, "h": h, "w": w, "n_rows": n_rows, "frames": f, "timesteps": ts,
                "ctx": ctx, "fingerprint": spec_fingerprint(sd), "x": torch.stack(xs), "depth": torch.stack(ds),
                "y": torch.stack(ys), "kv_sums": kv_sums, "kv_abs_sums": kv_abs, "kv_after_warmup": kv_after_warmup,
                "stream_x": torch.stack(sx), "stream_depth": torch.stack(sdp), "stream_y": torch.stack(sy)},
               os.path.join(HERE, "unet_tiny_warmup.pt"))
    print("unet_tiny_warmup.pt", float(torch.stack(ys).abs().mean()), float(torch.stack(sy).abs().mean()))


def gen_unet_sd15_widths(frames=3, seed=0, h=16, w=16, n_rows=2):
    """The reference's streaming UNet at the REAL SD1.5 widths (320/640/1280/1280, head dims 40/80/160, 1.28 G
    parameters) on a small 16x16 latent: pins the oracle at the channel geometry the GPU engine runs (the tiny fixture
    covers the topology, this one the widths).  Weights = live2diff_b200.weights.random_state_dict(seed) -- the same
    ones the GPU parity tests and bench.py use; only inputs, outputs and cache checksums are stored."""
    from live2diff_b200.weights import random_state_dict

    d = UNetDims()
    sd = random_state_dict(d, seed=seed)
    u = build_ref_unet(d)
    u.load_state_dict(sd, strict=True)
    u.eval()
    u.set_info_for_attn(h, w)
    kv = u.prepare_cache(n_rows)
    g = torch.Generator().manual_seed(seed + 101)
    for c in kv:
        c[:, :, :, : d.sink_size] = torch.randn(c[:, :, :, : d.sink_size].shape, generator=g)
    ctx = torch.randn(n_rows, 77, d.cross_attention_dim, generator=g)
    t = torch.tensor([399, 199])
    xs, ds, ys = [], [], []
    for mask, pe_idx, update_idx in run_schedule(n_rows, d.window_size, d.sink_size, frames):
        x = torch.randn(n_rows, 4, 1, h, w, generator=g)
        dep = torch.randn(n_rows, 4, 1, h, w, generator=g)
        o = u(x, t, encoder_hidden_states=ctx, temporal_attention_mask=mask, depth_sample=dep, kv_cache=kv,
              pe_idx=pe_idx, update_idx=update_idx)
        xs.append(x)
        ds.append(dep)
        ys.append(o["sample"].clone())
    probe = {i: kv[i][:, :, :4, d.sink_size:d.sink_size + frames + 1].clone() for i in (0, 13, 26, 39)}
    torch.save({"seed": seed, "h": h, "w": w, "n_rows": n_rows, "timesteps": t, "ctx": ctx,
                "fingerprint": spec_fingerprint({k: sd[k] for k in list(sd)[:40]}),
                "x": torch.stack(xs), "depth": torch.stack(ds), "y": torch.stack(ys),
                "kv_abs_sums": torch.tensor([float(c.double().abs().sum()) for c in kv]), "kv_probe": probe},
               os.path.join(HERE, "unet_sd15_widths.pt"))
    print("unet_sd15_widths.pt", float(torch.stack(ys).abs().mean()), float(torch.stack(ys).abs().max()))


def gen_unet_sd15_widths_warmup(seed=0, h=16, w=16):
    """The reference's WARM-UP UNet at the real SD1.5 widths: one pass over an 8-frame clip on row 0 of zero caches."""
    from live2diff_b200.weights import random_state_dict

    d = UNetDims()
    sd = random_state_dict(d, seed=seed)
    uw = build_ref_warmup_unet(d)
    uw.load_state_dict(sd, strict=True)
    uw.eval()
    uw.set_info_for_attn(h, w)
    f = d.sink_size
    rows = [torch.zeros(s_[1:]) for s_ in d.kv_cache_shapes(1, h, w)]
    g = torch.Generator().manual_seed(seed + 202)
    x = torch.randn(1, 4, f, h, w, generator=g)
    dep = torch.randn(1, 4, f, h, w, generator=g)
    ctx = torch.randn(1, 77, d.cross_attention_dim, generator=g)
    t = torch.tensor([399])
    y = uw(x, t, temporal_attention_mask=None, depth_sample=dep, encoder_hidden_states=ctx, kv_cache=rows,
           return_dict=True)["sample"]
    probe = {i: rows[i][:, :4, :f].clone() for i in (0, 13, 26, 39)}
    torch.save({"seed": seed, "h": h, "w": w, "frames": f, "timestep": t, "ctx": ctx, "x": x, "depth": dep, "y": y.clone(),
                "fingerprint": spec_fingerprint({k: sd[k] for k in list(sd)[:40]}),
                "kv_abs_sums": torch.tensor([float(c.double().abs().sum()) for c in rows]), "kv_probe": probe},
               os.path.join(HERE, "unet_sd15_widths_warmup.pt"))
    print("unet_sd15_widths_warmup.pt", float(y.abs().mean()), float(y.abs().max()))


def gen_specs():
    for tag, d in (("tiny", TINY), ("sd15", UNetDims())):
        u = build_ref_unet(d, device="meta")
        spec = {k: list(v.shape) for k, v in u.state_dict().items()}
        json.dump(spec, open(os.path.join(HERE, f"state_dict_spec_{tag}.json"), "w"))
        print("spec", tag, len(spec), sum(torch.Size(v).numel() for v in spec.values()) / 1e6, "M")


def gen_warmup():
    gen_warmup_attention("c64_f8", 64, 8, 12, 8, 16, 24, seed=41)
    gen_warmup_attention("c320_f8", 320, 8, 6, 8, 16, 24, seed=42)
    gen_warmup_attention("c128_f2_L4", 128, 8, 4, 2, 4, 24, seed=43)
    gen_unet_tiny_warmup()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sd15":        # only the SD1.5-width fixture (needs ~15 GB of host memory)
        gen_unet_sd15_widths()
        gen_unet_sd15_widths_warmup()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "warmup":      # only the warm-up fixtures (the others are unchanged)
        gen_warmup()
        sys.exit(0)
    gen_specs()
    gen_schedule()
    gen_scheduler_pointwise()
    gen_stream_attention("c64_fill_wrap", 64, 8, 3, 4, 2, 16, 8, 24, 12, seed=21)
    gen_stream_attention("c320_hd40", 320, 8, 2, 3, 2, 16, 8, 24, 10, seed=22)
    gen_stream_attention("c128_L32_N4", 128, 8, 2, 2, 4, 32, 8, 32, 30, seed=23)
    gen_stream_attention("c64_N1_L4", 64, 8, 2, 2, 2, 4, 2, 24, 6, seed=24)
    gen_temporal_transformer("c64", 64, 8, 4, 4, 2, 16, 8, 24, 10, seed=31)
    gen_temporal_transformer("c320", 320, 8, 2, 2, 2, 16, 8, 24, 3, seed=32)
    gen_resnet_and_friends()
    gen_unet_tiny()
    gen_warmup()

