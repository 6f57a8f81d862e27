"""CPU oracle for the pipeline-level stream logic.  TEST INFRASTRUCTURE ONLY (see unet_oracle.py).

Restates the parts of StreamAnimateDiffusionDepth that the reference hard-codes to CUDA
(`.cuda()` pipeline_stream_animation_depth.py:410, `.to("cuda")` :315, torch.cuda.Event :627) and
therefore cannot be imported and run in the CPU build container:

  initialize_attn_bias_pe_and_update_idx   live2diff/pipeline_stream_animation_depth.py:403-414
  update_attn_bias                         :416-438
  prepare() constants                      :242-301
  add_noise / scheduler_step_batch         :378-401
  predict_x0_batch (use_denoising_batch)   :573-601

LCM constants live in the un-vendored dependency diffusers==0.25.0 (`LCMScheduler`,
schedulers/scheduling_lcm.py): restated from its published algorithm -- "linear" betas
(configs/base_config.yaml:30-36), `set_timesteps` with original_inference_steps=50,
`get_scalings_for_boundary_condition_discrete` with sigma_data=0.5, timestep_scaling=10.

The only reference-provided expectation for the ring schedule is the docstring sketch at
pipeline_stream_animation_depth.py:417-421 (attn_bias [[0,0,0,inf],[0,0,inf,inf]], pe_idx
[[0,1,2,3]]*2, update_idx [2,1] for W0=2, L=4) -- tests/test_schedule.py checks it, plus the
trace in SURVEY.md Appendix A-2.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor

WARMUP_FRAMES = 8   # pipeline_stream_animation_depth.py:20
WINDOW_SIZE = 16    # :21


# --------------------------------------------------------------------------------------
# LCM scheduler constants (diffusers 0.25.0 LCMScheduler)
# --------------------------------------------------------------------------------------

def alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> Tensor:
    """beta_schedule == "linear" (base_config.yaml:30-36; SURVEY A-11)."""
    betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    return torch.cumprod(1.0 - betas, dim=0)


def lcm_timesteps(num_inference_steps: int = 50, num_train_timesteps: int = 1000,
                  original_inference_steps: int = 50) -> np.ndarray:
    """LCMScheduler.set_timesteps (strength=1): evenly skip through the 50 'origin' timesteps."""
    c = num_train_timesteps // original_inference_steps
    origin = np.asarray(list(range(1, original_inference_steps + 1))) * c - 1
    origin = origin[::-1].copy()
    idx = np.floor(np.linspace(0, len(origin), num=num_inference_steps, endpoint=False)).astype(np.int64)
    return origin[idx]


def boundary_scalings(timestep: int, sigma_data: float = 0.5, timestep_scaling: float = 10.0) -> Tuple[float, float]:
    """LCMScheduler.get_scalings_for_boundary_condition_discrete."""
    s = timestep * timestep_scaling
    c_skip = sigma_data ** 2 / (s ** 2 + sigma_data ** 2)
    c_out = s / (s ** 2 + sigma_data ** 2) ** 0.5
    return c_skip, c_out


def stream_constants(t_index_list: Sequence[int], num_inference_steps: int = 50, dtype=torch.float32):
    """prepare() :242-301 -> (sub_timesteps [N] int64, c_skip, c_out, sqrt(abar), sqrt(1-abar)), each [N]."""
    ts = lcm_timesteps(num_inference_steps)
    sub = [int(ts[t]) for t in t_index_list]
    ac = alphas_cumprod()
    c_skip = torch.tensor([boundary_scalings(t)[0] for t in sub], dtype=torch.float32).to(dtype)
    c_out = torch.tensor([boundary_scalings(t)[1] for t in sub], dtype=torch.float32).to(dtype)
    a = torch.stack([ac[t].sqrt() for t in sub]).to(dtype)
    b = torch.stack([(1 - ac[t]).sqrt() for t in sub]).to(dtype)
    return torch.tensor(sub, dtype=torch.int64), c_skip, c_out, a, b


# --------------------------------------------------------------------------------------
# KV ring schedule
# --------------------------------------------------------------------------------------

def init_schedule(n_rows: int, window: int = WINDOW_SIZE, warmup: int = WARMUP_FRAMES, dtype=torch.float32):
    """initialize_attn_bias_pe_and_update_idx (:403-414).  The reference's unguarded
    `update_idx[1] = WARMUP_FRAMES + 1` raises IndexError for N == 1 (SURVEY A-1); guarded here."""
    valid = torch.zeros(n_rows, window, dtype=torch.bool)
    valid[:, :warmup] = True
    valid[0, warmup] = True
    attn_bias = torch.zeros(n_rows, window, dtype=dtype).masked_fill_(~valid, float("-inf"))
    pe_idx = torch.arange(window).unsqueeze(0).repeat(n_rows, 1)
    update_idx = torch.full((n_rows,), warmup, dtype=torch.int64)
    if n_rows > 1:
        update_idx[1] = warmup + 1
    return attn_bias, pe_idx, update_idx


def update_schedule(attn_bias: Tensor, pe_idx: Tensor, update_idx: Tensor, window: int = WINDOW_SIZE,
                    warmup: int = WARMUP_FRAMES):
    """update_attn_bias (:416-438), in place like the reference."""
    for n in range(attn_bias.shape[0]):
        if torch.isinf(attn_bias[n]).any():
            update_idx[n] = (attn_bias[n] == 0).sum()
        else:
            pe_idx[n, warmup:] = pe_idx[n, warmup:].roll(shifts=1, dims=0)
            update_idx[n] = pe_idx[n].argmax()
        num_unmask = int((attn_bias[n] == 0).sum())
        attn_bias[n, : min(num_unmask + 1, window)] = 0
    return attn_bias, pe_idx, update_idx


# --------------------------------------------------------------------------------------
# scheduler pointwise + stream batch
# --------------------------------------------------------------------------------------

def _col(v: Tensor, like: Tensor) -> Tensor:
    return v.view(-1, *([1] * (like.dim() - 1))).to(like.dtype)


def add_noise(x0: Tensor, noise: Tensor, a: Tensor, b: Tensor, t_index: int) -> Tensor:
    """:378-385"""
    return a[t_index] * x0 + b[t_index] * noise


def scheduler_step_batch(model_pred: Tensor, x_t: Tensor, c_skip: Tensor, c_out: Tensor, a: Tensor, b: Tensor) -> Tensor:
    """:387-401, idx=None branch: F = (x - sqrt(1-abar) eps) / sqrt(abar);  x0 = c_out F + c_skip x."""
    f_theta = (x_t - _col(b, x_t) * model_pred) / _col(a, x_t)
    return _col(c_out, x_t) * f_theta + _col(c_skip, x_t) * x_t


class StreamOracle:
    """Per-stream state machine = predict_x0_batch (:573-601) around an injected UNet callable.

    `unet(sample, timestep, encoder_hidden_states=, temporal_attention_mask=, depth_sample=,
    kv_cache=, pe_idx=, update_idx=)` must return the noise prediction tensor and mutate
    `kv_cache` in place.  Noise is injected (SURVEY A-6) so two implementations can be compared.
    """

    def __init__(self, unet: Callable, kv_cache: List[Tensor], prompt_embeds: Tensor, t_index_list: Sequence[int],
                 latent_hw: Tuple[int, int], window: int = WINDOW_SIZE, warmup: int = WARMUP_FRAMES,
                 dtype=torch.float32, device="cpu"):
        self.unet = unet
        self.kv_cache = kv_cache
        self.n = len(t_index_list)
        self.window, self.warmup = window, warmup
        self.dtype, self.device = dtype, device
        (self.timesteps, self.c_skip, self.c_out, self.a, self.b) = [
            t.to(device) for t in stream_constants(t_index_list, dtype=dtype)]
        self.prompt_embeds = prompt_embeds.to(device=device, dtype=dtype)
        h, w = latent_hw
        if self.n > 1:                                                        # :193-206
            self.x_buf = torch.zeros(self.n - 1, 4, 1, h, w, dtype=dtype, device=device)
            self.d_buf = torch.zeros_like(self.x_buf)
        else:
            self.x_buf = self.d_buf = None
        ab, pe, up = init_schedule(self.n, window, warmup, dtype)             # :211
        self.attn_bias, self.pe_idx, self.update_idx = ab.to(device), pe.to(device), up.to(device)

    def step(self, x_t: Tensor, depth: Tensor, noise: Optional[Tensor]) -> Tensor:
        """x_t, depth [1,4,1,h,w]; noise [(N-1),4,1,h,w] for the re-noise of :596-598."""
        if self.n > 1:                                                        # :579-581
            x_t = torch.cat((x_t, self.x_buf), dim=0)
            depth = torch.cat((depth, self.d_buf), dim=0)
        eps = self.unet(x_t, self.timesteps, encoder_hidden_states=self.prompt_embeds,
                        temporal_attention_mask=self.attn_bias, depth_sample=depth, kv_cache=self.kv_cache,
                        pe_idx=self.pe_idx, update_idx=self.update_idx)
        x0 = scheduler_step_batch(eps, x_t, self.c_skip, self.c_out, self.a, self.b)          # :489
        update_schedule(self.attn_bias, self.pe_idx, self.update_idx, self.window, self.warmup)  # :585-587
        if self.n > 1:                                                        # :589-601
            out = x0[-1:].clone()
            self.x_buf = _col(self.a[1:], x0) * x0[:-1] + _col(self.b[1:], x0) * noise
            self.d_buf = depth[:-1]
        else:
            out = x0
        return out
