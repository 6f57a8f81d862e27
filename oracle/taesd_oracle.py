"""CPU oracle for the tiny VAE legs of the frame.  TEST INFRASTRUCTURE ONLY: imported by tests/,
smoke() and bench.py's baseline legs, never by the product (live2diff_b200/).

The reference uses `diffusers.AutoencoderTiny` ("madebyollin/taesd") as `stream.vae`
(live2diff/utils/wrapper.py:468-470) and calls it in
  encode_image   live2diff/pipeline_stream_animation_depth.py:517-535   vae.encode(x) -> latents * scaling_factor, add_noise(.., 0)
  encode_depth   :544-571 (last three lines)                            vae.encode(depth_map_norm) -> latents * scaling_factor
  decode_image   :537-542                                               vae.decode(x0 / scaling_factor)[0].clip(-1, 1)
  __call__       :625-660                                               preprocess -> encode x2 -> predict_x0_batch -> decode

PARITY UNPINNED: the arithmetic lives in the un-vendored dependency diffusers==0.25.0 (models/autoencoder_tiny.py,
models/vae.py: EncoderTiny / DecoderTiny / AutoencoderTinyBlock), which is neither in /root/reference nor installed
here, and the reference holds no golden vectors for it.  Restated from the published architecture:
  AutoencoderTinyBlock(c): fuse = ReLU( conv3x3 -> ReLU -> conv3x3 -> ReLU -> conv3x3  +  skip ),  skip = identity (c -> c)
  EncoderTiny:  x <- (x + 1) / 2;  conv3x3(3,64); Block; [conv3x3 stride 2, no bias; Block x3] x3; conv3x3(64,4)
  DecoderTiny:  x <- tanh(x / 3) * 3;  conv3x3(4,64); ReLU; [Block x3; Upsample(2, nearest); conv3x3(64,64, no bias)] x3;
                Block; conv3x3(64,3);  out <- 2 * out - 1
  config: scaling_factor = 1.0, latent_channels = 4, block_out_channels (64,64,64,64), num_encoder_blocks (1,3,3,3),
          num_decoder_blocks (3,3,3,1), act_fn "relu"
State-dict keys are diffusers' (`encoder.layers.N...`, `decoder.layers.N...`), so a real TAESD checkpoint loads unchanged.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
CH = 64
ENC_BLOCKS = (1, 3, 3, 3)
DEC_BLOCKS = (3, 3, 3, 1)
SCALING_FACTOR = 1.0


def taesd_param_spec() -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape, in diffusers AutoencoderTiny.state_dict() order (nn.Sequential indices incl. ReLU / Upsample slots)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def conv(p, cin, cout, bias=True):
        s[p + ".weight"] = (cout, cin, 3, 3)
        if bias:
            s[p + ".bias"] = (cout,)

    def block(p):
        for i in (0, 2, 4):
            conv(f"{p}.conv.{i}", CH, CH)

    i = 0
    for stage, nb in enumerate(ENC_BLOCKS):
        conv(f"encoder.layers.{i}", 3 if stage == 0 else CH, CH, bias=stage == 0)
        i += 1
        for _ in range(nb):
            block(f"encoder.layers.{i}")
            i += 1
    conv(f"encoder.layers.{i}", CH, 4)
    conv("decoder.layers.0", 4, CH)
    i = 2                                   # layers.1 is the ReLU
    for stage, nb in enumerate(DEC_BLOCKS):
        for _ in range(nb):
            block(f"decoder.layers.{i}")
            i += 1
        last = stage == len(DEC_BLOCKS) - 1
        if not last:
            i += 1                          # nn.Upsample slot
        conv(f"decoder.layers.{i}", CH, 3 if last else CH, bias=last)
        i += 1
    return s


def _block(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    h = F.relu(F.conv2d(x, sd[f"{p}.conv.0.weight"], sd[f"{p}.conv.0.bias"], padding=1))
    h = F.relu(F.conv2d(h, sd[f"{p}.conv.2.weight"], sd[f"{p}.conv.2.bias"], padding=1))
    h = F.conv2d(h, sd[f"{p}.conv.4.weight"], sd[f"{p}.conv.4.bias"], padding=1)
    return F.relu(h + x)


def encode(sd: Dict[str, Tensor], x: Tensor) -> Tensor:
    """EncoderTiny.forward on images in [-1, 1], [N,3,H,W] -> latents [N,4,H/8,W/8] (AutoencoderTiny.encode(...).latents)."""
    x = x.add(1).div(2)
    i = 0
    for stage, nb in enumerate(ENC_BLOCKS):
        p = f"encoder.layers.{i}"
        x = F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), padding=1, stride=1 if stage == 0 else 2)
        i += 1
        for _ in range(nb):
            x = _block(sd, f"encoder.layers.{i}", x)
            i += 1
    p = f"encoder.layers.{i}"
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)


def decode(sd: Dict[str, Tensor], z: Tensor) -> Tensor:
    """DecoderTiny.forward: latents [N,4,h,w] -> images [N,3,8h,8w] in (about) [-1, 1] (AutoencoderTiny.decode(...)[0])."""
    x = torch.tanh(z / 3) * 3
    x = F.relu(F.conv2d(x, sd["decoder.layers.0.weight"], sd["decoder.layers.0.bias"], padding=1))
    i = 2
    for stage, nb in enumerate(DEC_BLOCKS):
        for _ in range(nb):
            x = _block(sd, f"decoder.layers.{i}", x)
            i += 1
        last = stage == len(DEC_BLOCKS) - 1
        if not last:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            i += 1
        p = f"decoder.layers.{i}"
        x = F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), padding=1)
        i += 1
    return x.mul(2).sub(1)


# ---- the pipeline legs around the VAE (pipeline_stream_animation_depth.py) ---------------------------------------------

def preprocess_u8(img_u8_hwc: Tensor) -> Tensor:
    """VaeImageProcessor.preprocess of a uint8 [N,H,W,3] frame already at the stream's size: /255 -> [0,1] -> 2x-1, NCHW."""
    return img_u8_hwc.permute(0, 3, 1, 2).to(torch.float32).div(255.0).mul(2).sub(1)


def postprocess_u8(img: Tensor) -> Tensor:
    """image_utils.postprocess_image (denormalize :9-13, numpy_to_pil :24-30): (x/2+0.5).clamp(0,1) * 255, round, uint8 NHWC."""
    return (img.float() / 2 + 0.5).clamp(0, 1).mul(255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def encode_image(sd, x: Tensor, noise: Tensor, sqrt_abar0: float, sqrt_1m_abar0: float) -> Tensor:
    """encode_image (:517-535): latents * scaling_factor, then add_noise(.., t_index 0) with an injected `noise`."""
    lat = encode(sd, x) * SCALING_FACTOR
    return sqrt_abar0 * lat + sqrt_1m_abar0 * noise


def encode_depth_map(sd, depth_map_norm: Tensor) -> Tensor:
    """Tail of encode_depth (:565-571) from the normalised depth map [N,H,W] in [0,1] (the MiDaS part is §8 f2):
    repeat to 3 channels, * 2 - 1, vae.encode, * scaling_factor."""
    x = depth_map_norm[:, None].repeat(1, 3, 1, 1) * 2 - 1
    return encode(sd, x) * SCALING_FACTOR


def decode_image(sd, x0: Tensor) -> Tensor:
    """decode_image (:537-542)."""
    return decode(sd, x0 / SCALING_FACTOR).clip(-1, 1)
