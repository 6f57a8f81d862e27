"""Parameter inventory of the streaming UNet (names + shapes) and seeded random initialisation.

The names are the reference's `state_dict` keys (UNet3DConditionStreamingModel,
live2diff/animatediff/models/unet_depth_streaming.py:88-250 and the block/module constructors it
calls), so a reference checkpoint can be handed to `B200UNetStep` unchanged; the native engine
looks tensors up by these names (csrc/engine.cu).  `tests/test_oracle_golden.py` checks
the inventory against the key/shape list dumped from the reference model itself
(tests/golden/state_dict_spec_*.json).

No checkpoint can be downloaded in this environment, so benchmarks and parity tests use
`random_state_dict` -- seeded, real SD1.5 shapes, torch-default-like fan-in scaling.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch


@dataclass(frozen=True)
class UNetDims:
    """Static geometry (defaults: SD1.5 + Live2Diff configs/base_config.yaml:6-28)."""

    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    heads: int = 8
    cross_attention_dim: int = 768
    in_channels: int = 4
    out_channels: int = 4
    norm_groups: int = 32
    norm_eps: float = 1e-5
    window_size: int = 16
    sink_size: int = 8
    pe_max_len: int = 24
    mapping_channels: Tuple[int, ...] = (16, 32, 96, 256)
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    up_has_attn: Tuple[bool, ...] = (False, True, True, True)

    @property
    def temb_dim(self) -> int:
        return self.block_out_channels[0] * 4

    def level_hw(self, h: int, w: int) -> List[Tuple[int, int]]:
        return [(h >> i, w >> i) for i in range(len(self.block_out_channels))]

    def kv_cache_shapes(self, n_rows: int, h: int, w: int) -> List[Tuple[int, ...]]:
        """40 cache shapes [N,2,hw,L,C] in motion_module_idx order (set_info_for_attn traversal,
        unet_depth_streaming.py:252-302): down0..3 (2 modules x 2 attns), then up0..3 (3 x 2)."""
        lv = self.level_hw(h, w)
        nlev = len(lv)
        order = [(i, self.layers_per_block) for i in range(nlev)] + \
                [(nlev - 1 - i, self.layers_per_block + 1) for i in range(nlev)]
        out = []
        for lvl, nmod in order:
            hh, ww = lv[lvl]
            out += [(n_rows, 2, hh * ww, self.window_size, self.block_out_channels[lvl])] * (2 * nmod)
        return out


def sinusoid_table(max_len: int, d_model: int) -> torch.Tensor:
    """`pos_encoder.pe` buffer [1,max_len,C] (positional_encoding.py:12-18)."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def unet_param_spec(d: UNetDims) -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape for every parameter/buffer of the streaming UNet, in module order."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    c = d.block_out_channels
    temb = d.temb_dim

    def conv(p, cin, cout, k=3):
        s[p + ".weight"] = (cout, cin, k, k)
        s[p + ".bias"] = (cout,)

    def lin(p, cin, cout, bias=True):
        s[p + ".weight"] = (cout, cin)
        if bias:
            s[p + ".bias"] = (cout,)

    def norm(p, ch):
        s[p + ".weight"] = (ch,)
        s[p + ".bias"] = (ch,)

    def dense_attn(p, ch, kv):
        lin(p + ".to_q", ch, ch, False)
        lin(p + ".to_k", kv, ch, False)
        lin(p + ".to_v", kv, ch, False)
        lin(p + ".to_out.0", ch, ch)

    def ff(p, ch):
        lin(p + ".net.0.proj", ch, 8 * ch)
        lin(p + ".net.2", 4 * ch, ch)

    def spatial(p, ch):
        norm(p + ".norm", ch)
        conv(p + ".proj_in", ch, ch, 1)
        b = p + ".transformer_blocks.0"
        dense_attn(b + ".attn1", ch, ch)
        norm(b + ".norm1", ch)
        dense_attn(b + ".attn2", ch, d.cross_attention_dim)
        norm(b + ".norm2", ch)
        ff(b + ".ff", ch)
        norm(b + ".norm3", ch)
        conv(p + ".proj_out", ch, ch, 1)

    def resnet(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cin, cout)
        lin(p + ".time_emb_proj", temb, cout)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout)
        if cin != cout:
            conv(p + ".conv_shortcut", cin, cout, 1)

    def motion(p, ch):
        t = p + ".temporal_transformer"
        norm(t + ".norm", ch)
        lin(t + ".proj_in", ch, ch)
        b = t + ".transformer_blocks.0"
        for i in range(2):
            a = f"{b}.attention_blocks.{i}"
            lin(a + ".to_q", ch, ch, False)
            lin(a + ".to_k", ch, ch, False)
            lin(a + ".to_v", ch, ch, False)
            lin(a + ".to_out.0", ch, ch)
            s[a + ".pos_encoder.pe"] = (1, d.pe_max_len, ch)
        for i in range(2):
            norm(f"{b}.norms.{i}", ch)
        ff(b + ".ff", ch)
        norm(b + ".ff_norm", ch)
        lin(t + ".proj_out", ch, ch)

    conv("conv_in", d.in_channels, c[0])
    m = d.mapping_channels
    conv("flow_conv_in.conv_in", d.in_channels, m[0])
    for i in range(len(m) - 1):
        conv(f"flow_conv_in.blocks.{2 * i}", m[i], m[i])
        conv(f"flow_conv_in.blocks.{2 * i + 1}", m[i], m[i + 1])
    conv("flow_conv_in.conv_out", m[-1], c[0])
    lin("time_embedding.linear_1", c[0], temb)
    lin("time_embedding.linear_2", temb, temb)

    nlev = len(c)
    out_ch = c[0]
    for bi in range(nlev):
        in_ch, out_ch = out_ch, c[bi]
        bp = f"down_blocks.{bi}"
        # nn.Module registration order inside the block: attentions, resnets, motion_modules, downsamplers
        if d.down_has_attn[bi]:
            for li in range(d.layers_per_block):
                spatial(f"{bp}.attentions.{li}", out_ch)
        for li in range(d.layers_per_block):
            resnet(f"{bp}.resnets.{li}", in_ch if li == 0 else out_ch, out_ch)
        for li in range(d.layers_per_block):
            motion(f"{bp}.motion_modules.{li}", out_ch)
        if bi != nlev - 1:
            conv(f"{bp}.downsamplers.0.conv", out_ch, out_ch)

    rev = list(reversed(c))
    out_ch = rev[0]
    for bi in range(nlev):
        prev_out, out_ch = out_ch, rev[bi]
        in_ch = rev[min(bi + 1, nlev - 1)]
        bp = f"up_blocks.{bi}"
        nl = d.layers_per_block + 1
        if d.up_has_attn[bi]:
            for li in range(nl):
                spatial(f"{bp}.attentions.{li}", out_ch)
        for li in range(nl):
            skip = in_ch if li == nl - 1 else out_ch
            rin = prev_out if li == 0 else out_ch
            resnet(f"{bp}.resnets.{li}", rin + skip, out_ch)
        for li in range(nl):
            motion(f"{bp}.motion_modules.{li}", out_ch)
        if bi != nlev - 1:
            conv(f"{bp}.upsamplers.0.conv", out_ch, out_ch)

    spatial("mid_block.attentions.0", c[-1])
    resnet("mid_block.resnets.0", c[-1], c[-1])
    resnet("mid_block.resnets.1", c[-1], c[-1])
    norm("conv_norm_out", c[0])
    conv("conv_out", c[0], d.out_channels)
    return s


def random_state_dict(d: UNetDims, seed: int = 0, dtype=torch.float32, device="cpu",
                      gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded random weights with the reference's names/shapes.

    Matrices/filters ~ U(-b, b), b = gain / sqrt(fan_in) (the bound torch's default Linear/Conv
    init uses); biases ~ U(-b, b); norm scales 1 + 0.1 U(-1,1), norm shifts 0.1 U(-1,1);
    `pos_encoder.pe` is the deterministic sinusoid.  Generated on CPU in fp32 in spec order
    from one torch.Generator, then cast/moved, so the same seed gives the same weights anywhere.
    """
    return random_tensors(unet_param_spec(d), seed=seed, dtype=dtype, device=device, gain=gain)


def random_tensors(spec: "OrderedDict[str, Tuple[int, ...]]", seed: int = 0, dtype=torch.float32, device="cpu",
                   gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded init of an arbitrary name->shape spec (same rules as `random_state_dict`)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    fan_in_of: Dict[str, int] = {}
    for name, shape in spec.items():
        if name.endswith(".weight") and len(shape) >= 2:
            fan_in_of[name[: -len(".weight")]] = int(math.prod(shape[1:]))
    for name, shape in spec.items():
        if name.endswith("pos_encoder.pe"):
            t = sinusoid_table(shape[1], shape[2])
        elif len(shape) >= 2:
            b = gain / math.sqrt(math.prod(shape[1:]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        else:
            base = name.rsplit(".", 1)[0]
            u = torch.rand(shape, generator=g) * 2 - 1
            if base in fan_in_of:                       # bias of a linear/conv
                t = u * (gain / math.sqrt(fan_in_of[base]))
            elif name.endswith(".weight"):              # norm scale
                t = 1.0 + 0.1 * u
            else:                                       # norm shift
                t = 0.1 * u
        out[name] = t.to(dtype=dtype, device=device)
    return out


def spec_fingerprint(sd: Dict[str, torch.Tensor]) -> float:
    """Cheap order-independent checksum used by golden fixtures to prove the weights were
    regenerated identically (guards against a silent RNG-stream change, not a parity check)."""
    tot = 0.0
    for k in sorted(sd):
        tot += float(sd[k].double().abs().sum())
    return tot
