"""Host-side stream schedule: LCM constants and the KV ring schedule, kept as plain Python state.

Mirrors (reference: live2diff/pipeline_stream_animation_depth.py)
  * prepare() constants :242-301  -- sub-timesteps, c_skip/c_out, sqrt(abar), sqrt(1-abar) of the
    diffusers==0.25.0 LCMScheduler (beta_schedule "linear", configs/base_config.yaml:30-36)
  * initialize_attn_bias_pe_and_update_idx :403-414 and update_attn_bias :416-438

The reference keeps attn_bias/pe_idx/update_idx as CUDA tensors and advances them with
`.any()/.sum()/.argmax()` in Python control flow -- >= 3 device->host syncs per row per frame
(SURVEY.md §3.3).  The schedule is a pure function of the frame counter, so here it lives on the host as
integers and is uploaded as three tiny tensors per frame (no sync); the kernels consume the indices on
the device.  `tests/test_schedule.py` checks this state machine against the trace produced by the
reference's own methods.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple


def lcm_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012) -> List[float]:
    """Linear betas in fp32 like torch.linspace + cumprod (diffusers LCMScheduler.__init__)."""
    import torch

    betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    return torch.cumprod(1.0 - betas, dim=0).tolist()


def lcm_timesteps(num_inference_steps=50, num_train_timesteps=1000, original_inference_steps=50) -> List[int]:
    c = num_train_timesteps // original_inference_steps
    origin = [i * c - 1 for i in range(1, original_inference_steps + 1)][::-1]
    return [origin[int(math.floor(i * len(origin) / num_inference_steps))] for i in range(num_inference_steps)]


@dataclass
class StreamConstants:
    timesteps: List[int]
    sqrt_abar: List[float]
    sqrt_1m_abar: List[float]
    c_skip: List[float]
    c_out: List[float]

    def table(self) -> List[List[float]]:
        """[4][N] in the order l2d_lcm_step expects."""
        return [self.sqrt_abar, self.sqrt_1m_abar, self.c_skip, self.c_out]


def stream_constants(t_index_list: Sequence[int], num_inference_steps=50, sigma_data=0.5,
                     timestep_scaling=10.0) -> StreamConstants:
    ts = lcm_timesteps(num_inference_steps)
    ac = lcm_alphas_cumprod()
    sub = [ts[t] for t in t_index_list]
    c_skip = [sigma_data ** 2 / ((t * timestep_scaling) ** 2 + sigma_data ** 2) for t in sub]
    c_out = [(t * timestep_scaling) / math.sqrt((t * timestep_scaling) ** 2 + sigma_data ** 2) for t in sub]
    return StreamConstants(sub, [math.sqrt(ac[t]) for t in sub], [math.sqrt(1 - ac[t]) for t in sub], c_skip, c_out)


@dataclass
class RingSchedule:
    """Per denoise row: `valid` slots (a prefix), PE index of every slot, and the slot written next."""

    n_rows: int
    window: int = 16
    warmup: int = 8
    valid: List[int] = field(default_factory=list)          # number of unmasked (leading) slots per row
    pe_idx: List[List[int]] = field(default_factory=list)
    update_idx: List[int] = field(default_factory=list)

    def __post_init__(self):
        if not 0 < self.warmup < self.window:
            raise ValueError("need 0 < warmup < window")
        # row 0 already sees the slot it is about to fill; rows >= 1 see only the sink slots (:404-408)
        self.valid = [self.warmup + (1 if r == 0 else 0) for r in range(self.n_rows)]
        self.pe_idx = [list(range(self.window)) for _ in range(self.n_rows)]
        # :411-412 (row 1 starts one slot further; the unguarded reference line crashes for N == 1, SURVEY A-1)
        self.update_idx = [self.warmup + (1 if r == 1 else 0) for r in range(self.n_rows)]

    def mask_rows(self) -> List[List[float]]:
        return [[0.0] * v + [float("-inf")] * (self.window - v) for v in self.valid]

    def advance(self) -> None:
        """One frame: the state transition of update_attn_bias (:423-436)."""
        for r in range(self.n_rows):
            if self.valid[r] < self.window:            # still filling: write the first masked slot, PE unchanged
                self.update_idx[r] = self.valid[r]
                self.valid[r] += 1
            else:                                      # full: rotate the rolling PEs, overwrite the oldest slot
                tail = self.pe_idx[r][self.warmup:]
                self.pe_idx[r][self.warmup:] = tail[-1:] + tail[:-1]
                row = self.pe_idx[r]
                self.update_idx[r] = row.index(max(row))
