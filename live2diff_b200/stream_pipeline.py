"""`B200StreamPipeline`: the per-frame denoising part of StreamAnimateDiffusionDepth, B200-native.

Mirrors the reference pipeline's stream-batch logic (live2diff/pipeline_stream_animation_depth.py):
  prepare()            :171-301   zero latent/depth buffers, ring-schedule init, prompt embeds, LCM constants
  update_prompt()      :368-376
  predict_x0_batch()   :573-601   cat new frame with the (N-1) buffered rows, UNet step, LCM x0 prediction,
                                  schedule advance, output = x0[-1], buffer = sqrt(abar)[1:] x0[:-1] + sqrt(1-abar)[1:] noise
VAE encode/decode, MiDaS and the one-shot bidirectional warm-up UNet are outside this path:
the caller supplies the noisy latent x_t and the depth latent, and (optionally) KV caches whose sink
slots 0..W0-1 were filled by the reference's warm-up.

Differences from the reference that do not change results: the ring schedule lives on the host as
integers (no `.any()/.argmax()` device syncs), and the scheduler pointwise + buffer shift is one kernel
(l2d_lcm_step) instead of ~10 elementwise launches.

Multi-GPU: streams are independent, one stream per rank/GPU; the only collective is the
broadcast of the shared prompt embedding (`broadcast_prompt`, NCCL).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops
from .schedule import RingSchedule, stream_constants
from .unet_step import B200UNetStep


def broadcast_prompt(prompt_embeds: Optional[torch.Tensor], shape, device, src: int = 0) -> torch.Tensor:
    """Rank `src` supplies the text embedding [1 or N, 77, 768]; every rank returns a copy (NCCL broadcast
    over NVLink; 118-236 KB, latency-bound).  With no process group this is the identity."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert prompt_embeds is not None
        return prompt_embeds.to(device=device, dtype=torch.float16)
    buf = torch.empty(shape, dtype=torch.float16, device=device)
    if dist.get_rank() == src:
        buf.copy_(prompt_embeds.to(device=device, dtype=torch.float16))
    dist.broadcast(buf, src=src)
    return buf


class B200StreamPipeline:
    def __init__(self, unet: B200UNetStep, t_index_list: Sequence[int], num_inference_steps: int = 50,
                 window: Optional[int] = None, warmup: Optional[int] = None, do_add_noise: bool = True,
                 seed: int = 2):
        self.unet = unet
        self.device = unet.device
        self.n = len(t_index_list)
        if self.n != unet.n_rows:
            raise ValueError(f"t_index_list has {self.n} steps but the UNet engine was built for {unet.n_rows} rows")
        self.window = unet.dims.window_size if window is None else window
        self.warmup = unet.dims.sink_size if warmup is None else warmup
        self.do_add_noise = do_add_noise
        self.consts_host = stream_constants(t_index_list, num_inference_steps)
        self.generator = torch.Generator(device=self.device)
        self.generator.manual_seed(seed)
        self.latent_shape = (4, 1, unet.h, unet.w)
        self.prompt_embeds = None
        self.kv_cache_list: List[torch.Tensor] = []
        self.frames_done = 0

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def prepare(self, prompt_embeds: torch.Tensor, kv_cache_list: Optional[List[torch.Tensor]] = None) -> None:
        dev = self.device
        n = self.n
        self.x_t_latent_buffer = torch.zeros((n - 1,) + self.latent_shape, dtype=torch.float16, device=dev) if n > 1 else None
        self.depth_latent_buffer = torch.zeros_like(self.x_t_latent_buffer) if n > 1 else None
        self.schedule = RingSchedule(n, self.window, self.warmup)
        self.update_prompt(prompt_embeds)
        c = self.consts_host
        self.sub_timesteps_tensor = torch.tensor(c.timesteps, dtype=torch.int64, device=dev)
        self.consts = torch.tensor(c.table(), dtype=torch.float32, device=dev).contiguous()
        self.kv_cache_list = kv_cache_list if kv_cache_list is not None else self.unet.prepare_cache(n)
        # pinned staging for the three tiny schedule tensors: a ring of slots, each guarded by a CUDA event, because the
        # host runs ahead of the GPU (graph replay, no sync per frame) and must not rewrite a slot whose H2D copy is queued
        self._stage_slots = []
        for _ in range(4):
            self._stage_slots.append((torch.empty(n, self.window, dtype=torch.float16).pin_memory(),
                                      torch.empty(n, self.window, dtype=torch.int64).pin_memory(),
                                      torch.empty(n, dtype=torch.int64).pin_memory(), torch.cuda.Event()))
        self._stage_i = 0
        self.attn_bias = torch.empty(n, self.window, dtype=torch.float16, device=dev)
        self.pe_idx = torch.empty(n, self.window, dtype=torch.int64, device=dev)
        self.update_idx = torch.empty(n, dtype=torch.int64, device=dev)
        self._x_cat = torch.empty((n,) + self.latent_shape, dtype=torch.float16, device=dev)
        self._d_cat = torch.empty_like(self._x_cat)
        self.frames_done = 0

    @torch.no_grad()
    def update_prompt(self, prompt_embeds: torch.Tensor) -> None:
        pe = prompt_embeds.to(device=self.device, dtype=torch.float16)
        if pe.dim() == 2:
            pe = pe[None]
        self.prompt_embeds = (pe if pe.shape[0] == self.n else pe[:1].repeat(self.n, 1, 1)).contiguous()   # :231

    @torch.no_grad()
    def warmup_denoise(self, unet_warmup, x_t_latent: torch.Tensor, depth_latent: torch.Tensor,
               noise: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
        """The denoising loop of the reference's warm-up (pipeline:315-338) after `prepare()`: one pass of the
        warm-up UNet per denoise row idx over the clip x_t_latent / depth_latent [1,4,F,h,w], each filling sink
        slots 0..F-1 of row idx of every cache; between passes x_t = sqrt(abar)[idx+1] x0 + sqrt(1-abar)[idx+1] randn.
        Returns x_0_pred as [F,4,h,w] (what the reference hands to the VAE decoder, :341-342).  `noise[idx]`
        overrides the generator for the re-noise after pass idx (parity tests)."""
        c = self.consts_host
        x_t = x_t_latent.to(device=self.device, dtype=torch.float16).contiguous()
        dep = depth_latent.to(device=self.device, dtype=torch.float16).contiguous()
        if x_t.shape[2] != self.warmup:
            raise ValueError(f"warm-up clip has {x_t.shape[2]} frames, the schedule expects {self.warmup} sink slots")
        x0 = None
        for idx, t in enumerate(c.timesteps):
            out = unet_warmup(x_t, torch.tensor([t], dtype=torch.int64, device=self.device), temporal_attention_mask=None,
                              depth_sample=dep, encoder_hidden_states=self.prompt_embeds[0:1],
                              kv_cache=[cache[idx] for cache in self.kv_cache_list], return_dict=True)
            consts = torch.tensor([[c.sqrt_abar[idx]], [c.sqrt_1m_abar[idx]], [c.c_skip[idx]], [c.c_out[idx]]],
                                  dtype=torch.float32, device=self.device)
            x0, _, _ = ops.lcm_step(x_t, out["sample"], consts)          # scheduler_step_batch(..., idx)  :330
            if idx < self.n - 1:                                          # :331-337
                eps = noise[idx].to(x0) if noise is not None else torch.randn(
                    x0.shape, dtype=torch.float16, device=self.device, generator=self.generator)
                a = torch.tensor(c.sqrt_abar[idx + 1], dtype=torch.float16, device=self.device)
                b = torch.tensor(c.sqrt_1m_abar[idx + 1], dtype=torch.float16, device=self.device)
                x_t = a * x0 + b * eps
        return x0[0].transpose(0, 1).contiguous()                         # "b c f h w -> b f c h w"[0]  :341

    def _upload_schedule(self):
        s = self.schedule
        h_mask, h_pe, h_up, ev = self._stage_slots[self._stage_i % len(self._stage_slots)]
        if self._stage_i >= len(self._stage_slots):
            ev.synchronize()                       # the copies issued from this slot 4 frames ago have run
        self._stage_i += 1
        h_mask.copy_(torch.tensor(s.mask_rows(), dtype=torch.float16))
        h_pe.copy_(torch.tensor(s.pe_idx, dtype=torch.int64))
        h_up.copy_(torch.tensor(s.update_idx, dtype=torch.int64))
        self.attn_bias.copy_(h_mask, non_blocking=True)
        self.pe_idx.copy_(h_pe, non_blocking=True)
        self.update_idx.copy_(h_up, non_blocking=True)
        ev.record()

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def predict_x0_batch(self, x_t_latent: torch.Tensor, depth_latent: torch.Tensor,
                         noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x_t_latent, depth_latent: [1,4,1,h,w] fp16 CUDA.  `noise` [(N-1),4,1,h,w] overrides the internal
        generator for the re-noise term (parity tests inject it, SURVEY A-6)."""
        n = self.n
        self._x_cat[0:1].copy_(x_t_latent.reshape((1,) + self.latent_shape))
        self._d_cat[0:1].copy_(depth_latent.reshape((1,) + self.latent_shape))
        if n > 1:                                                    # :579-581
            self._x_cat[1:].copy_(self.x_t_latent_buffer)
            self._d_cat[1:].copy_(self.depth_latent_buffer)
        self._upload_schedule()
        out = self.unet(self._x_cat, self.sub_timesteps_tensor, depth_sample=self._d_cat,
                        encoder_hidden_states=self.prompt_embeds, temporal_attention_mask=self.attn_bias,
                        kv_cache=self.kv_cache_list, pe_idx=self.pe_idx, update_idx=self.update_idx)
        self.kv_cache_list = out["kv_cache"]                          # :468-469
        self.schedule.advance()                                       # :585-587
        if n > 1 and self.do_add_noise and noise is None:
            noise = torch.randn((n - 1,) + self.latent_shape, dtype=torch.float16, device=self.device,
                                generator=self.generator)             # :596-598
        x0_last, next_buf, _ = ops.lcm_step(self._x_cat, out["sample"], self.consts, noise if self.do_add_noise else None)
        if n > 1:
            self.x_t_latent_buffer = next_buf
            self.depth_latent_buffer = self._d_cat[:-1].clone()       # :601
        self.frames_done += 1
        return x0_last

    __call__ = predict_x0_batch
