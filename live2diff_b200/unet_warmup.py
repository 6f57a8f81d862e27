"""`B200UNetWarmup`: the object installed at `stream.unet_warmup`.

Call contract = the one `StreamAnimateDiffusionDepth.prepare` uses once per denoise row `idx`
(live2diff/pipeline_stream_animation_depth.py:315-329) on `UNet3DConditionWarmupModel.forward`
(live2diff/animatediff/models/unet_depth_warmup.py:405-590):

    out = unet_warmup(x_t_latent[1,4,F,h,w], t[1], temporal_attention_mask=None, depth_sample=[1,4,F,h,w],
                      encoder_hidden_states=[1,77,768], kv_cache=[cache[idx] for cache in kv_cache_list],
                      return_dict=True)
    out["sample"] -> [1,4,F,h,w];  every `cache[idx]` ([2,hw,L,C] view) has slots 0..F-1 filled in place with the
    PE-free k / v of the F warm-up frames (motion_module.py:488-489) -- the sink slots the streaming step reads.

The warm-up UNet shares the streaming UNet's state_dict (identical key set); every non-temporal layer treats the
F frames as batch rows, so the same native engine (csrc/engine.cu) runs it with `warmup_frames = F`: the frames ride
on the batch axis and the motion modules call the bidirectional kernel (csrc/kv_warmup.cu) instead of K1.
The reference keeps this model on the CPU and moves 2.6 GB to the GPU and back around every warm-up
(pipeline:316, 340); here it is a second engine object that the caller may keep or drop.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from ._lib import L2DUnetStepArgs, check, current_stream, lib
from .unet_step import UNetStepOutput, create_engine, create_shared_engine
from .weights import UNetDims


class B200UNetWarmup:
    def __init__(self, state_dict: Optional[Dict[str, torch.Tensor]], dims: UNetDims, frames: int, latent_h: int, latent_w: int,
                 ctx_len: int = 77, device: Optional[torch.device] = None, share_weights_with=None):
        """`share_weights_with`: a `B200UNetStep` built from the same state_dict -- the warm-up engine then reads that
        engine's repacked weights (no second 2.6 GiB copy; `state_dict` may be None) and keeps it alive."""
        if not torch.cuda.is_available():
            raise RuntimeError("B200UNetWarmup needs a CUDA device; live2diff_b200 has no CPU fallback")
        if not 1 <= frames <= min(8, dims.window_size):
            raise ValueError("warm-up clip length must be in 1..min(8, window)")
        self.dims, self.frames, self.h, self.w, self.ctx_len = dims, frames, latent_h, latent_w, ctx_len
        self.device = torch.device(device or "cuda")
        self._base = share_weights_with
        if share_weights_with is not None:
            if share_weights_with.dims != dims:
                raise ValueError("share_weights_with: the streaming engine was built for different UNetDims")
            self._handle = create_shared_engine(share_weights_with._handle, dims, frames, latent_h, latent_w, ctx_len, False,
                                                frames, self.device)
        else:
            self._handle = create_engine(state_dict, dims, frames, latent_h, latent_w, ctx_len, False, frames, self.device)
        self.dtype = torch.float16
        dev = self.device
        self._sample = torch.empty(frames, 4, 1, latent_h, latent_w, dtype=torch.float16, device=dev)
        self._depth = torch.empty_like(self._sample)
        self._out = torch.empty_like(self._sample)
        self._t = torch.empty(frames, dtype=torch.int64, device=dev)
        self._ctx = torch.empty(frames, ctx_len, dims.cross_attention_dim, dtype=torch.float16, device=dev)
        # one denoise row of every cache: [2, hw, L, C]
        self._row_shapes = [tuple(s[1:]) for s in dims.kv_cache_shapes(1, latent_h, latent_w)]

    def to(self, *args, **kwargs):            # the pipeline shuttles the reference model CPU<->GPU (:316, :340)
        return self

    def __del__(self):
        if getattr(self, "_handle", None):
            lib().l2d_unet_destroy(self._handle)
            self._handle = None

    @property
    def device_bytes(self) -> int:
        return lib().l2d_unet_device_bytes(self._handle)

    @torch.no_grad()
    def __call__(self, sample, timestep, temporal_attention_mask=None, depth_sample=None, encoder_hidden_states=None,
                 kv_cache: Optional[List[torch.Tensor]] = None, return_dict: bool = True, **kwargs):
        if depth_sample is None or encoder_hidden_states is None or kv_cache is None:
            raise ValueError("unet_warmup(...) needs depth_sample, encoder_hidden_states and kv_cache")
        f = self.frames
        if sample.dim() != 5 or sample.shape[0] != 1 or sample.shape[2] != f:
            raise ValueError(f"expected sample [1,4,{f},h,w], got {tuple(sample.shape)}")
        if len(kv_cache) != len(self._row_shapes):
            raise ValueError(f"expected {len(self._row_shapes)} kv-cache rows, got {len(kv_cache)}")
        for t, s in zip(kv_cache, self._row_shapes):
            if tuple(t.shape) != s or t.dtype != torch.float16 or not t.is_cuda or not t.is_contiguous():
                raise ValueError(f"kv-cache row {tuple(t.shape)} {t.dtype}: expected contiguous CUDA fp16 {s} (cache[idx])")
        # "b c f h w -> (b f) c h w": the frames become the engine's batch rows
        self._sample[:, :, 0].copy_(sample[0].transpose(0, 1))
        self._depth[:, :, 0].copy_(depth_sample[0].transpose(0, 1))
        self._t.copy_(timestep.reshape(-1)[:1].to(torch.int64).expand(f))
        ctx = encoder_hidden_states.to(device=self.device, dtype=torch.float16)
        self._ctx.copy_(ctx[:1].expand(f, -1, -1))                  # repeat_interleave over frames (attention.py:112)
        ptrs = (C.c_void_p * len(kv_cache))(*[t.data_ptr() for t in kv_cache])
        args = L2DUnetStepArgs()
        args.sample, args.timestep = self._sample.data_ptr(), self._t.data_ptr()
        args.encoder_hidden_states, args.depth_sample = self._ctx.data_ptr(), self._depth.data_ptr()
        args.temporal_attention_mask = args.pe_idx = args.update_idx = None
        args.kv_cache = C.cast(ptrs, C.POINTER(C.c_void_p))
        args.n_kv = len(kv_cache)
        args.out_sample = self._out.data_ptr()
        check(lib().l2d_unet_step(self._handle, C.byref(args), current_stream()))
        out = self._out[:, :, 0].transpose(0, 1)[None].contiguous()   # "(b f) c h w -> b c f h w"
        if not return_dict:
            return (out,)
        return UNetStepOutput(sample=out)
