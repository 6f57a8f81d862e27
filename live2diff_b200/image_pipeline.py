"""`B200ImageStream`: a1 of SURVEY.md §8 -- `StreamAnimateDiffusionDepth.__call__`
(live2diff/pipeline_stream_animation_depth.py:625-660) image in -> image out, on the B200 kernels:

    x            = preprocess(frame)                               :630   uint8 [H,W,3] -> fp16 [1,3,H,W] in [-1,1]   (1 kernel)
    x_t_latent   = encode_image(x)                                 :517-535  TAESD encode, * scaling_factor, add_noise(., randn, t_index 0)
    depth_latent = encode_depth(x)                                 :544-571  depth prior -> 3-channel [-1,1] map -> TAESD encode
    x_0_pred     = predict_x0_batch(x_t_latent, depth_latent)      :573-601  `B200DeviceStream` (whole frame = one CUDA graph)
    x_output     = decode_image(x_0_pred)                          :537-542  TAESD decode, .clip(-1, 1)
    (wrapper.postprocess_image -> uint8)                           utils/wrapper.py:273-292, image_utils.py:9-30

The MiDaS depth network itself (§8 f2) is not part of this package: `depth_detector` is any callable mapping the
[1,3,384,384] fp16 image in [-1,1] to a depth map [1,384,384] (the reference's `self.depth_detector`), or the caller passes
a ready `depth_map` in [0,1]; everything around it (bilinear resizes, min/max normalisation, 3-channel repeat, VAE encode)
follows encode_depth line by line.  With neither, the depth latent is the encoding of a constant mid-grey map.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
import torch.nn.functional as F

from .device_stream import B200DeviceStream
from .taesd import B200TinyVAE
from .unet_step import B200UNetStep


class B200ImageStream:
    def __init__(self, unet: B200UNetStep, vae: B200TinyVAE, t_index_list: Sequence[int], num_inference_steps: int = 50,
                 depth_detector: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, seed: int = 2,
                 do_add_noise: bool = True, use_cuda_graph: bool = True):
        if (vae.height // 8, vae.width // 8) != (unet.h, unet.w):
            raise ValueError(f"VAE image size {vae.height}x{vae.width} does not match the UNet latent {unet.h}x{unet.w}")
        self.unet, self.vae, self.depth_detector = unet, vae, depth_detector
        self.device = unet.device
        self.height, self.width = vae.height, vae.width
        self.stream = B200DeviceStream(unet, t_index_list, num_inference_steps, do_add_noise=do_add_noise, seed=seed,
                                       use_cuda_graph=use_cuda_graph)
        c = self.stream.consts_host
        self.sqrt_abar0, self.sqrt_1m_abar0 = float(c.sqrt_abar[0]), float(c.sqrt_1m_abar[0])
        self.generator = torch.Generator(device=self.device)
        self.generator.manual_seed(seed)
        self.prev_image_result = None

    # ---- reference-named pieces ----------------------------------------------------------------------------------
    def prepare(self, prompt_embeds: torch.Tensor, kv_cache_list=None) -> None:
        self.stream.prepare(prompt_embeds, kv_cache_list)

    def update_prompt(self, prompt_embeds: torch.Tensor) -> None:
        self.stream.update_prompt(prompt_embeds)

    @torch.no_grad()
    def encode_image(self, image_tensors: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """:517-535.  image_tensors [f,3,H,W] in [-1,1] -> noisy latent at t_index 0; `noise` overrides the generator."""
        lat = self.vae.encode(image_tensors).latents * self.vae.config.scaling_factor
        if noise is None:
            noise = torch.randn(lat.shape, device=lat.device, dtype=lat.dtype, generator=self.generator)
        return self.sqrt_abar0 * lat + self.sqrt_1m_abar0 * noise.to(lat)

    @torch.no_grad()
    def encode_depth(self, image_tensors: torch.Tensor, depth_map: Optional[torch.Tensor] = None) -> torch.Tensor:
        """:544-571.  depth_map (optional) [f,H,W] in [0,1] replaces detector + min/max normalisation."""
        h, w = image_tensors.shape[2], image_tensors.shape[3]
        if depth_map is None and self.depth_detector is not None:
            images_input = F.interpolate(image_tensors, (384, 384), mode="bilinear", align_corners=False)
            dm = self.depth_detector(images_input)
            dm = (dm - dm.min()) / (dm.max() - dm.min())
            dm3 = dm[:, None].repeat(1, 3, 1, 1) * 2 - 1
            dm3 = F.interpolate(dm3, (h, w), mode="bilinear", align_corners=False)
        else:
            if depth_map is None:
                depth_map = torch.full((image_tensors.shape[0], h, w), 0.5, device=self.device, dtype=torch.float16)
            dm3 = depth_map.to(device=self.device, dtype=torch.float16)[:, None].repeat(1, 3, 1, 1) * 2 - 1
        return self.vae.encode(dm3.to(torch.float16)).latents * self.vae.config.scaling_factor

    @torch.no_grad()
    def decode_image(self, x_0_pred_out: torch.Tensor) -> torch.Tensor:
        """:537-542 (the clip runs inside the decoder's last kernel)."""
        return self.vae.decode(x_0_pred_out / self.vae.config.scaling_factor, return_dict=False, clip=True)[0]

    # ---- the frame -------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, x: torch.Tensor, depth_map: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                 renoise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: uint8 frame [H,W,3] / [1,H,W,3] (CUDA) or an fp16/fp32 tensor [1,3,H,W] in [-1,1].  Returns x_output
        [1,3,H,W] fp16 in [-1,1] like the reference; `frame_u8` wraps it for uint8 in / uint8 out."""
        if x.dtype == torch.uint8:
            x = self.vae.preprocess_u8(x.to(self.device, non_blocking=True).reshape(1, self.height, self.width, 3))
        else:
            x = x.to(device=self.device, dtype=torch.float16)
        x_t_latent = self.encode_image(x, noise)
        depth_latent = self.encode_depth(x, depth_map)
        x_0_pred_out = self.stream(x_t_latent.unsqueeze(2).contiguous(), depth_latent.unsqueeze(2).contiguous(), noise=renoise)
        x_output = self.decode_image(x_0_pred_out[:, :, 0].contiguous())
        self.prev_image_result = x_output
        return x_output

    @torch.no_grad()
    def frame_u8(self, frame_u8: torch.Tensor, depth_map: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 [H,W,3] in (CUDA or pinned host) -> uint8 [H,W,3] out (`out`: CUDA or pinned host tensor, else a new
        CUDA tensor).  Host tensors are copied asynchronously on the current stream."""
        res = self.vae.postprocess_u8(self(frame_u8, depth_map))[0]
        if out is not None:
            out.copy_(res, non_blocking=True)
            return out
        return res
