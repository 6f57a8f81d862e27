"""`B200DeviceStream`: predict_x0_batch with the whole stream state resident in HBM.

Same contract as `B200StreamPipeline` (stream_pipeline.py; reference: live2diff/pipeline_stream_animation_depth.py
:171-301 prepare, :368-376 update_prompt, :573-601 predict_x0_batch), but nothing of the per-frame state machine runs
in Python or through torch ops: the latent/depth buffers, the KV ring schedule (:403-438), the LCM constants and the
re-noise generator live inside `l2d_stream` (csrc/stream_state.cu) and one frame is one CUDA-graph launch between two
small copies.  `__call__` accepts CUDA tensors or pinned host tensors for the new latent pair and writes x0 into a
caller-supplied (device or pinned host) tensor, so an end-to-end frame is: H2D 2 x 32 KiB, graph, D2H 32 KiB.

Differences from the reference that do not change results: the schedule is advanced by a device kernel (the reference
does it with `.any()/.argmax()` host syncs); the re-noise `randn` comes from a counter-based Philox4x32-10 keyed by
(seed, frame, row, element) instead of torch's global generator (same distribution; parity tests inject the noise).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from ._lib import check, current_stream, lib
from .schedule import stream_constants
from .unet_step import B200UNetStep


class B200DeviceStream:
    def __init__(self, unet: B200UNetStep, t_index_list: Sequence[int], num_inference_steps: int = 50,
                 warmup: Optional[int] = None, do_add_noise: bool = True, seed: int = 2, use_cuda_graph: bool = True):
        self.unet = unet
        self.device = unet.device
        self.n = len(t_index_list)
        if self.n != unet.n_rows:
            raise ValueError(f"t_index_list has {self.n} steps but the UNet engine was built for {unet.n_rows} rows")
        self.window = unet.dims.window_size
        self.warmup = unet.dims.sink_size if warmup is None else warmup
        self.consts_host = stream_constants(t_index_list, num_inference_steps)
        self.latent_shape = (4, 1, unet.h, unet.w)
        ts = (C.c_int64 * self.n)(*self.consts_host.timesteps)
        flat = [v for row in self.consts_host.table() for v in row]
        consts = (C.c_float * (4 * self.n))(*flat)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().l2d_stream_create(C.byref(handle), unet._handle, ts, consts, self.warmup, seed, int(do_add_noise),
                                          int(use_cuda_graph)))
        self._handle = handle
        self.kv_cache_list: List[torch.Tensor] = []
        self._out = torch.empty((1,) + self.latent_shape, dtype=torch.float16, device=self.device)

    def __del__(self):
        if getattr(self, "_handle", None):
            lib().l2d_stream_destroy(self._handle)
            self._handle = None

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def prepare(self, prompt_embeds: torch.Tensor, kv_cache_list: Optional[List[torch.Tensor]] = None) -> None:
        check(lib().l2d_stream_reset(self._handle, current_stream()))
        self.update_prompt(prompt_embeds)
        self.kv_cache_list = kv_cache_list if kv_cache_list is not None else self.unet.prepare_cache(self.n)
        arr = self.unet._kv_table(self.kv_cache_list)                   # validates shapes / dtype / contiguity
        check(lib().l2d_stream_set_cache(self._handle, C.cast(arr, C.POINTER(C.c_void_p)), len(self.kv_cache_list)))

    @torch.no_grad()
    def update_prompt(self, prompt_embeds: torch.Tensor) -> None:
        pe = prompt_embeds.to(device=self.device, dtype=torch.float16)
        if pe.dim() == 2:
            pe = pe[None]
        pe = pe.contiguous()
        rows = self.n if pe.shape[0] == self.n else 1
        if tuple(pe.shape[1:]) != (self.unet.ctx_len, self.unet.dims.cross_attention_dim):
            raise ValueError(f"prompt_embeds {tuple(pe.shape)}: expected [1 or N, {self.unet.ctx_len}, "
                             f"{self.unet.dims.cross_attention_dim}]")
        check(lib().l2d_stream_set_prompt(self._handle, pe.data_ptr(), rows, current_stream()))
        self.prompt_embeds = pe                                          # keep alive until the copy has run

    @staticmethod
    def _addr(t: torch.Tensor, name: str, numel: int) -> int:
        if t.dtype != torch.float16 or not t.is_contiguous() or t.numel() != numel:
            raise ValueError(f"{name}: expected a contiguous fp16 tensor of {numel} elements")
        if not (t.is_cuda or t.is_pinned()):
            raise ValueError(f"{name}: must be a CUDA tensor or a pinned host tensor")
        return t.data_ptr()

    @torch.no_grad()
    def predict_x0_batch(self, x_t_latent: torch.Tensor, depth_latent: torch.Tensor,
                         noise: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x_t_latent, depth_latent: [1,4,1,h,w] fp16, CUDA or pinned host.  Returns x0 [1,4,1,h,w]: a fresh CUDA tensor
        per frame like the reference (safe to queue), or `out` if given (CUDA or pinned host; for a host `out` the caller
        synchronises the stream before reading it).  Pinned-host inputs are read asynchronously by the frame's H2D copies:
        do not rewrite them before the stream has been synchronised (or an event recorded after this call has fired)."""
        per = self._out.numel()
        if out is None:
            out = torch.empty_like(self._out)
        n_ptr = None if noise is None else self._addr(noise, "noise", (self.n - 1) * per)
        check(lib().l2d_stream_frame(self._handle, self._addr(x_t_latent, "x_t_latent", per),
                                     self._addr(depth_latent, "depth_latent", per), n_ptr, self._addr(out, "out", per),
                                     current_stream()))
        return out

    __call__ = predict_x0_batch

    # ------------------------------------------------------------------------------------------
    @property
    def launches_per_frame(self) -> int:
        return lib().l2d_stream_launches_per_frame(self._handle)

    def invalidate_graph(self) -> None:
        """Developer hook (profiles/, bench.py's ablation passes): re-capture the frame graph on the next frame."""
        lib().l2d_stream_invalidate_graph(self._handle)

    def schedule(self) -> Dict[str, list]:
        """Read the ring schedule back from the device (synchronises): valid [N], pe_idx [N][L], update_idx [N], frame."""
        n, L = self.n, self.window
        valid, pe, up, fr = (C.c_int32 * n)(), (C.c_int64 * (n * L))(), (C.c_int64 * n)(), C.c_uint64()
        check(lib().l2d_stream_get_schedule(self._handle, valid, pe, up, C.byref(fr)))
        return {"valid": list(valid), "pe_idx": [list(pe[r * L:(r + 1) * L]) for r in range(n)], "update_idx": list(up),
                "frame": int(fr.value)}

    def save_state(self) -> bytes:
        """Latent/depth buffers + schedule + frame counter + seed as one blob (stream migration; the KV caches are
        ordinary tensors owned by the caller and travel separately)."""
        nbytes = lib().l2d_stream_state_bytes(self._handle)
        buf = C.create_string_buffer(nbytes)
        check(lib().l2d_stream_save_state(self._handle, buf, nbytes))
        return buf.raw

    def load_state(self, blob: bytes) -> None:
        buf = C.create_string_buffer(blob, len(blob))
        check(lib().l2d_stream_load_state(self._handle, buf, len(blob)))
