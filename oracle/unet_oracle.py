"""CPU oracle for the Live2Diff per-frame streaming UNet step.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, not the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  The product package
(`live2diff_b200`) never does, and fails loudly when its CUDA library is missing.

It restates, as plain functional PyTorch (fp32 on CPU by default; dtype/device generic so the
GPU tests can also run it in fp16 as the "reference eager path" comparand), the algorithm of

  UNet3DConditionStreamingModel.forward        live2diff/animatediff/models/unet_depth_streaming.py:429-627
  blocks                                       .../unet_blocks_streaming.py:157-850
  ResnetBlock3D / MappingNetwork / samplers    .../resnet.py:17-259
  Transformer3DModel / BasicTransformerBlock   .../attention.py:28-270
  TemporalTransformer3DModel / ...Block        .../motion_module.py:153-435
  StreamTemporalAttention                      .../stream_motion_module.py:57-213
  PositionalEncoding                           .../positional_encoding.py:8-18

over a `state_dict` with the reference's own key names (so reference checkpoints drive it
unchanged).  Arithmetic that lives in the un-vendored dependency diffusers==0.25.0 (setup.py:5)
-- `Attention`/`AttnProcessor2_0`, `FeedForward`/`GEGLU`, `Timesteps`, `TimestepEmbedding` --
is restated from its published algorithm (see each function).

Parity pin: the reference ships no tests / golden vectors for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: `tests/golden/make_golden.py` imports the
unmodified reference model files in the build container and commits seeded input/output
fixtures under tests/golden/; `tests/test_oracle_golden.py` checks this file against them.
Because the diffusers pieces of those fixtures come from a restatement (diffusers is absent from
the image), the diffusers-owned arithmetic is "parity unpinned" in the strict sense; everything
under /root/reference is pinned.

Streaming always runs with one frame per call (frame_buffer_size must be 1), so
5-D `b c f h w` tensors with f == 1 are handled as 4-D NCHW here; every `(b f) d c <-> (b d) f c`
rearrange of the reference is then a pure view.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass
class UNetDims:
    """Static geometry of the streaming UNet (defaults = SD1.5 + configs/base_config.yaml:6-28)."""

    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    heads: int = 8                      # SD1.5 "attention_head_dim: 8" is the head COUNT (SURVEY A-10)
    cross_attention_dim: int = 768
    in_channels: int = 4
    out_channels: int = 4
    norm_groups: int = 32
    norm_eps: float = 1e-5              # resnets / conv_norm_out (unet_depth_streaming.py:67)
    window_size: int = 16               # L   (base_config.yaml:27)
    sink_size: int = 8                  # W0  (base_config.yaml:28)
    pe_max_len: int = 24                # base_config.yaml:20
    mapping_channels: Tuple[int, ...] = (16, 32, 96, 256)   # resnet.py:26
    # which down blocks carry spatial transformers (unet_depth_streaming.py:47-59)
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    up_has_attn: Tuple[bool, ...] = (False, True, True, True)

    @property
    def temb_dim(self) -> int:
        return self.block_out_channels[0] * 4

    def level_hw(self, h: int, w: int) -> List[Tuple[int, int]]:
        return [(h >> i, w >> i) for i in range(len(self.block_out_channels))]


# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------

def positional_encoding(max_len: int, d_model: int) -> Tensor:
    """Sinusoidal table [max_len, d_model] (positional_encoding.py:12-17)."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def timestep_sinusoid(timesteps: Tensor, dim: int) -> Tensor:
    """diffusers 0.25.0 `get_timestep_embedding` with flip_sin_to_cos=True, freq_shift=0
    (unet_depth_streaming.py:45-46,102): fp32 [N, dim] = [cos | sin] of t * exp(-ln(1e4) i/half)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    ang = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


def conv2d(sd: SD, p: str, x: Tensor, stride: int = 1, padding: int = 1) -> Tensor:
    """InflatedConv3d with f == 1 (resnet.py:57-65)."""
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def group_norm(sd: SD, p: str, x: Tensor, groups: int, eps: float) -> Tensor:
    """nn.GroupNorm / InflatedGroupNorm with f == 1 (resnet.py:68-76)."""
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def layer_norm(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def linear(sd: SD, p: str, x: Tensor, bias: bool = True) -> Tensor:
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"] if bias else None)


def split_heads(t: Tensor, heads: int) -> Tensor:
    """[B, S, C] -> [B, heads, S, C/heads]  (attention.py:346-351 keeps B*heads flat; same math)."""
    b, s, c = t.shape
    return t.view(b, s, heads, c // heads).transpose(1, 2)


def merge_heads(t: Tensor) -> Tensor:
    b, h, s, d = t.shape
    return t.transpose(1, 2).reshape(b, s, h * d)


def sdpa(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    """softmax(q k^T / sqrt(hd) + mask) v -- what F.scaled_dot_product_attention computes
    (attention.py:552-560).  Written out so the fp32 oracle does not depend on SDPA back-ends."""
    if q.dtype == torch.float32:
        s = torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
        if mask is not None:
            s = s + mask
        return torch.matmul(torch.softmax(s, dim=-1), v)
    return F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0, is_causal=False)


def geglu_ff(sd: SD, p: str, x: Tensor) -> Tensor:
    """diffusers 0.25.0 FeedForward(activation_fn="geglu"): net.0 = GEGLU (proj: C -> 8C,
    h * gelu(g), exact-erf gelu), net.2 = Linear(4C -> C)  (attention.py:204, motion_module.py:360)."""
    h, g = linear(sd, p + ".net.0.proj", x).chunk(2, dim=-1)
    return linear(sd, p + ".net.2", h * F.gelu(g))


# --------------------------------------------------------------------------------------
# ResNet / samplers / mapping network
# --------------------------------------------------------------------------------------

def resnet_block(sd: SD, p: str, x: Tensor, temb: Tensor, d: UNetDims) -> Tensor:
    """ResnetBlock3D.forward (resnet.py:229-259), time_embedding_norm="default",
    output_scale_factor=1."""
    h = F.silu(group_norm(sd, p + ".norm1", x, d.norm_groups, d.norm_eps))
    h = conv2d(sd, p + ".conv1", h)
    h = h + linear(sd, p + ".time_emb_proj", F.silu(temb))[:, :, None, None]
    h = F.silu(group_norm(sd, p + ".norm2", h, d.norm_groups, d.norm_eps))
    h = conv2d(sd, p + ".conv2", h)
    if (p + ".conv_shortcut.weight") in sd:
        x = conv2d(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def downsample(sd: SD, p: str, x: Tensor) -> Tensor:
    """Downsample3D: conv3x3 stride 2 pad 1 (resnet.py:141,145-153)."""
    return conv2d(sd, p + ".conv", x, stride=2, padding=1)


def upsample(sd: SD, p: str, x: Tensor) -> Tensor:
    """Upsample3D: nearest x2 then conv3x3 (resnet.py:112,125)."""
    x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    return conv2d(sd, p + ".conv", x)


def mapping_network(sd: SD, p: str, x: Tensor, n_blocks: int) -> Tensor:
    """MappingNetwork.forward (resnet.py:44-54): conv_in, SiLU, (conv, SiLU) x n, conv_out."""
    e = F.silu(conv2d(sd, p + ".conv_in", x))
    for i in range(n_blocks):
        e = F.silu(conv2d(sd, f"{p}.blocks.{i}", e))
    return conv2d(sd, p + ".conv_out", e)


# --------------------------------------------------------------------------------------
# spatial transformer
# --------------------------------------------------------------------------------------

def dense_attention(sd: SD, p: str, x: Tensor, ctx: Optional[Tensor], heads: int) -> Tensor:
    """diffusers 0.25.0 `Attention` + AttnProcessor2_0 as used for attn1/attn2
    (attention.py:173-194, 243, 251-253): bias-free q/k/v, biased to_out.0, SDPA scale hd^-1/2."""
    src = x if ctx is None else ctx
    q = split_heads(linear(sd, p + ".to_q", x, bias=False), heads)
    k = split_heads(linear(sd, p + ".to_k", src, bias=False), heads)
    v = split_heads(linear(sd, p + ".to_v", src, bias=False), heads)
    return linear(sd, p + ".to_out.0", merge_heads(sdpa(q, k, v)))


def spatial_transformer(sd: SD, p: str, x: Tensor, ctx: Tensor, d: UNetDims) -> Tensor:
    """Transformer3DModel.forward, conv-projection variant (attention.py:91-135), one
    BasicTransformerBlock (attention.py:221-270): GN(eps 1e-6) -> 1x1 conv -> tokens ->
    [LN, self-attn, +res] [LN, cross-attn, +res] [LN, GEGLU-FF, +res] -> 1x1 conv -> +input."""
    n, c, hh, ww = x.shape
    h = group_norm(sd, p + ".norm", x, d.norm_groups, 1e-6)
    h = conv2d(sd, p + ".proj_in", h, padding=0)
    t = h.permute(0, 2, 3, 1).reshape(n, hh * ww, c)
    b = p + ".transformer_blocks.0"
    t = dense_attention(sd, b + ".attn1", layer_norm(sd, b + ".norm1", t), None, d.heads) + t
    t = dense_attention(sd, b + ".attn2", layer_norm(sd, b + ".norm2", t), ctx, d.heads) + t
    t = geglu_ff(sd, b + ".ff", layer_norm(sd, b + ".norm3", t)) + t
    h = t.reshape(n, hh, ww, c).permute(0, 3, 1, 2)
    return conv2d(sd, p + ".proj_out", h, padding=0) + x


# --------------------------------------------------------------------------------------
# streaming temporal attention (the north-star op) and the motion module around it
# --------------------------------------------------------------------------------------

def pe_tables(sd: SD, p: str, L: int) -> Tuple[Tensor, Tensor, Tensor]:
    """prepare_pe_buffer (stream_motion_module.py:79-97): q_pe,k_pe,v_pe = pe[:L] @ W_{q,k,v}^T."""
    pe = sd[p + ".pos_encoder.pe"][0, :L]
    return (F.linear(pe, sd[p + ".to_q.weight"]), F.linear(pe, sd[p + ".to_k.weight"]),
            F.linear(pe, sd[p + ".to_v.weight"]))


def kv_cache_attention(q: Tensor, k_new: Tensor, v_new: Tensor, kv_cache: Tensor, q_pe: Tensor, k_pe: Tensor,
                       v_pe: Tensor, mask: Tensor, pe_idx: Tensor, update_idx: Tensor, heads: int) -> Tensor:
    """The K1 contract (stream_motion_module.py:117-147,172-194).

    q,k_new,v_new [N,hw,C]; kv_cache [N,2,hw,L,C] (mutated in place, PE-free); *_pe [L,C];
    mask [N,L] additive {0,-inf}; pe_idx [N,L]; update_idx [N].  Returns merged heads [N,hw,C].
    """
    n_rows, hw, c = q.shape
    for n in range(n_rows):                                            # :117-119
        u = int(update_idx[n])
        kv_cache[n, 0, :, u] = k_new[n]
        kv_cache[n, 1, :, u] = v_new[n]
    q_sel = torch.stack([pe_idx[n, int(update_idx[n])] for n in range(n_rows)])          # :124-127
    k_full = kv_cache[:, 0] + k_pe[pe_idx].unsqueeze(1)               # [N,hw,L,C]   :129-131,140
    v_full = kv_cache[:, 1] + v_pe[pe_idx].unsqueeze(1)               #              :132-134,141
    q_full = q + q_pe[q_sel].unsqueeze(1)                             # [N,hw,C]     :135-139
    L = k_full.shape[2]
    hd = c // heads
    qh = q_full.reshape(n_rows * hw, 1, heads, hd).transpose(1, 2)    # [(N hw), heads, 1, hd]
    kh = k_full.reshape(n_rows * hw, L, heads, hd).transpose(1, 2)
    vh = v_full.reshape(n_rows * hw, L, heads, hd).transpose(1, 2)
    m = mask.to(q.dtype)[:, None, None, None, :].expand(n_rows, hw, 1, 1, L).reshape(n_rows * hw, 1, 1, L)  # :181-186
    o = sdpa(qh, kh, vh, m)                                           # :191-194
    return o.transpose(1, 2).reshape(n_rows, hw, c)


def stream_temporal_attention(sd: SD, p: str, x: Tensor, kv_cache: Tensor, mask: Tensor, pe_idx: Tensor,
                              update_idx: Tensor, d: UNetDims) -> Tensor:
    """StreamTemporalAttention.forward (stream_motion_module.py:149-213) for tokens x [N,hw,C]."""
    q = linear(sd, p + ".to_q", x, bias=False)
    k = linear(sd, p + ".to_k", x, bias=False)
    v = linear(sd, p + ".to_v", x, bias=False)
    q_pe, k_pe, v_pe = pe_tables(sd, p, d.window_size)
    o = kv_cache_attention(q, k, v, kv_cache, q_pe.to(x.dtype), k_pe.to(x.dtype), v_pe.to(x.dtype), mask,
                           pe_idx, update_idx, d.heads)
    return linear(sd, p + ".to_out.0", o)


def temporal_transformer(sd: SD, p: str, x: Tensor, kv_caches: Sequence[Tensor], mask: Tensor, pe_idx: Tensor,
                         update_idx: Tensor, d: UNetDims) -> Tensor:
    """TemporalTransformer3DModel.forward_streaming (motion_module.py:256-299) with one
    TemporalTransformerBlock.forward_streaming (:401-435).  `p` is the temporal_transformer
    prefix; `kv_caches` = the two cache tensors of this module's two attention blocks."""
    n, c, hh, ww = x.shape
    h = group_norm(sd, p + ".norm", x, d.norm_groups, 1e-6)
    t = h.permute(0, 2, 3, 1).reshape(n, hh * ww, c)
    t = linear(sd, p + ".proj_in", t)
    b = p + ".transformer_blocks.0"
    for i in range(2):
        a = stream_temporal_attention(sd, f"{b}.attention_blocks.{i}", layer_norm(sd, f"{b}.norms.{i}", t),
                                      kv_caches[i], mask, pe_idx, update_idx, d)
        t = a + t
    t = geglu_ff(sd, b + ".ff", layer_norm(sd, b + ".ff_norm", t)) + t
    t = linear(sd, p + ".proj_out", t)
    return t.reshape(n, hh, ww, c).permute(0, 3, 1, 2) + x


# --------------------------------------------------------------------------------------
# warm-up pass: the same UNet over F frames with bidirectional temporal attention
# --------------------------------------------------------------------------------------

def warmup_temporal_attention(sd: SD, p: str, x: Tensor, kv_row: Tensor, d: UNetDims) -> Tensor:
    """VersatileAttention.forward (motion_module.py:469-530) for tokens x [F,hw,C] of ONE clip (b = 1).

    kv_row [2,hw,L,C] is one denoise row of the module's cache (`cache[idx]`, pipeline:323); slots 0..F-1
    receive the PE-free k / v of the F warm-up frames (:488-489).  q/k/v get the PE of positions 0..F-1 through
    the same bias-free projections (:491-499); attention is full (no mask) over the F frames of a pixel."""
    f, hw, c = x.shape
    xt = x.transpose(0, 1)                                            # "(b f) d c -> (b d) f c"   :481
    q = linear(sd, p + ".to_q", xt, bias=False)
    k = linear(sd, p + ".to_k", xt, bias=False)
    v = linear(sd, p + ".to_v", xt, bias=False)
    kv_row[0, :, :f] = k                                              # :488-489
    kv_row[1, :, :f] = v
    q_pe, k_pe, v_pe = pe_tables(sd, p, f)                            # :491-495
    q = q + q_pe.to(x.dtype)
    k = k + k_pe.to(x.dtype)
    v = v + v_pe.to(x.dtype)
    o = merge_heads(sdpa(split_heads(q, d.heads), split_heads(k, d.heads), split_heads(v, d.heads)))   # :501-516
    return linear(sd, p + ".to_out.0", o).transpose(0, 1)             # :519-525


def temporal_transformer_warmup(sd: SD, p: str, x: Tensor, kv_rows: Sequence[Tensor], d: UNetDims) -> Tensor:
    """TemporalTransformer3DModel.forward_orig (motion_module.py:215-254) + TemporalTransformerBlock.forward_orig
    (:369-399) for x [F,C,h,w] (GroupNorm per frame: the reference normalises the "(b f) c h w" view)."""
    n, c, hh, ww = x.shape
    h = group_norm(sd, p + ".norm", x, d.norm_groups, 1e-6)
    t = h.permute(0, 2, 3, 1).reshape(n, hh * ww, c)
    t = linear(sd, p + ".proj_in", t)
    b = p + ".transformer_blocks.0"
    for i in range(2):
        a = warmup_temporal_attention(sd, f"{b}.attention_blocks.{i}", layer_norm(sd, f"{b}.norms.{i}", t), kv_rows[i], d)
        t = a + t
    t = geglu_ff(sd, b + ".ff", layer_norm(sd, b + ".ff_norm", t)) + t
    t = linear(sd, p + ".proj_out", t)
    return t.reshape(n, hh, ww, c).permute(0, 3, 1, 2) + x


def unet_forward_warmup(sd: SD, d: UNetDims, sample: Tensor, timestep: Tensor, encoder_hidden_states: Tensor,
                        depth_sample: Optional[Tensor], kv_rows: List[Tensor]) -> Tensor:
    """UNet3DConditionWarmupModel.forward (unet_depth_warmup.py:405-590) for one clip: sample/depth_sample
    [1,4,F,h,w], timestep [1], encoder_hidden_states [1,77,D], kv_rows = 40 views [2,hw,L,C] (`cache[idx]`).
    Every non-temporal layer treats the F frames as batch rows (all `(b f)` rearranges), so the streaming walk
    is reused with the frames on the batch axis and the bidirectional motion module swapped in."""
    assert sample.dim() == 5 and sample.shape[0] == 1, "warm-up runs one clip (pipeline:310-312)"
    f = sample.shape[2]
    x = sample[0].transpose(0, 1)
    dep = depth_sample[0].transpose(0, 1) if depth_sample is not None else None
    ctx = encoder_hidden_states.expand(f, -1, -1)

    def temporal(p, xx, caches):
        return temporal_transformer_warmup(sd, p, xx, caches, d)

    y = unet_forward(sd, d, x, timestep.reshape(-1)[:1], ctx, None, dep, kv_rows, None, None, temporal=temporal)
    return y.transpose(0, 1)[None]


# --------------------------------------------------------------------------------------
# whole UNet step
# --------------------------------------------------------------------------------------

@dataclass
class _Cursor:
    """motion_module_idx bookkeeping: traversal order of set_info_for_attn
    (unet_depth_streaming.py:252-281): down0..3 (2 modules each), then up0..3 (3 each)."""
    i: int = 0

    def take2(self, kv: Sequence[Tensor]) -> Tuple[Tensor, Tensor]:
        a, b = kv[self.i], kv[self.i + 1]
        self.i += 2
        return a, b


def unet_forward(sd: SD, d: UNetDims, sample: Tensor, timestep: Tensor, encoder_hidden_states: Tensor,
                 temporal_attention_mask: Optional[Tensor], depth_sample: Optional[Tensor], kv_cache: List[Tensor],
                 pe_idx: Optional[Tensor], update_idx: Optional[Tensor], temporal=None) -> Tensor:
    """UNet3DConditionStreamingModel.forward (unet_depth_streaming.py:429-627).

    sample/depth_sample [N,4,1,h,w] or [N,4,h,w]; returns the same rank.  `kv_cache` (40 tensors
    [N,2,hw,L,C]) is mutated in place exactly like the reference (stream_motion_module.py:117-119).
    """
    five_d = sample.dim() == 5
    if five_d:
        assert sample.shape[2] == 1, "streaming UNet runs one frame per call (SURVEY.md A-7)"
        sample = sample[:, :, 0]
        depth_sample = depth_sample[:, :, 0] if depth_sample is not None else None
    dt = sample.dtype
    ctx = encoder_hidden_states
    mask = temporal_attention_mask
    nlev = len(d.block_out_channels)
    if temporal is None:                       # streaming motion module; the warm-up pass swaps in the bidirectional one
        def temporal(p, x, caches):
            return temporal_transformer(sd, p, x, caches, mask, pe_idx, update_idx, d)

    # time (unet_depth_streaming.py:497-505): sinusoid in fp32 -> model dtype -> MLP
    t_emb = timestep_sinusoid(timestep.expand(sample.shape[0]), d.block_out_channels[0]).to(dt)
    emb = linear(sd, "time_embedding.linear_2", F.silu(linear(sd, "time_embedding.linear_1", t_emb)))

    # pre-process (:523-526)
    x = conv2d(sd, "conv_in", sample)
    if depth_sample is not None:
        x = mapping_network(sd, "flow_conv_in", depth_sample, 2 * (len(d.mapping_channels) - 1)) + x

    cur = _Cursor()
    skips = [x]
    # down (:529-553; unet_blocks_streaming.py:381-445, 516-569)
    for bi in range(nlev):
        bp = f"down_blocks.{bi}"
        for li in range(d.layers_per_block):
            x = resnet_block(sd, f"{bp}.resnets.{li}", x, emb, d)
            if d.down_has_attn[bi]:
                x = spatial_transformer(sd, f"{bp}.attentions.{li}", x, ctx, d)
            x = temporal(f"{bp}.motion_modules.{li}.temporal_transformer", x, cur.take2(kv_cache))
            skips.append(x)
        if bi != nlev - 1:
            x = downsample(sd, f"{bp}.downsamplers.0", x)
            skips.append(x)

    # mid (:564-573; unet_blocks_streaming.py:253-280) -- no motion module (motion_module_mid_block=False)
    x = resnet_block(sd, "mid_block.resnets.0", x, emb, d)
    x = spatial_transformer(sd, "mid_block.attentions.0", x, ctx, d)
    x = resnet_block(sd, "mid_block.resnets.1", x, emb, d)

    # up (:582-617; unet_blocks_streaming.py:666-731, 798-850)
    for bi in range(nlev):
        bp = f"up_blocks.{bi}"
        for li in range(d.layers_per_block + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(sd, f"{bp}.resnets.{li}", x, emb, d)
            if d.up_has_attn[bi]:
                x = spatial_transformer(sd, f"{bp}.attentions.{li}", x, ctx, d)
            x = temporal(f"{bp}.motion_modules.{li}.temporal_transformer", x, cur.take2(kv_cache))
        if bi != nlev - 1:
            x = upsample(sd, f"{bp}.upsamplers.0", x)
    assert not skips and cur.i == len(kv_cache)

    # post-process (:620-622)
    x = F.silu(group_norm(sd, "conv_norm_out", x, d.norm_groups, d.norm_eps))
    x = conv2d(sd, "conv_out", x)
    return x[:, :, None] if five_d else x


def alloc_kv_cache(d: UNetDims, n_rows: int, h: int, w: int, dtype=torch.float32, device="cpu") -> List[Tensor]:
    """prepare_cache / set_cache (unet_depth_streaming.py:283-302; stream_motion_module.py:57-77):
    40 zero tensors [N,2,hw,L,C] in motion_module_idx order."""
    lv = d.level_hw(h, w)
    order = [(bi, d.layers_per_block) for bi in range(len(lv))] + \
            [(len(lv) - 1 - bi, d.layers_per_block + 1) for bi in range(len(lv))]
    out = []
    for lvl, nmod in order:
        hh, ww = lv[lvl]
        for _ in range(2 * nmod):
            out.append(torch.zeros(n_rows, 2, hh * ww, d.window_size, d.block_out_channels[lvl], dtype=dtype,
                                   device=device))
    return out
