"""Drop-in replacements for the reference's temporal modules, backed by libl2d_b200.so.

  B200StreamTemporalAttention    <- StreamTemporalAttention      live2diff/animatediff/models/stream_motion_module.py:9-213
  B200TemporalTransformer3DModel <- TemporalTransformer3DModel   live2diff/animatediff/models/motion_module.py:153-299

Same constructor keywords, same method names (`set_info`, `set_index`, `set_cache`,
`prepare_pe_buffer`, `forward`), same `state_dict` keys, same in-place KV-cache mutation.  Parameters
stay `nn.Parameter`s so checkpoint loading / LoRA fusing keep working; the native object re-reads
them whenever their version counters change.  CUDA + fp16 only -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
from torch import nn

from . import ops
from ._lib import L2DTensor, check, current_stream, lib, make_tensor_table, ptr
from .weights import sinusoid_table


class _PositionalEncoding(nn.Module):
    """Holder of the `pos_encoder.pe` buffer [1,max_len,C] (positional_encoding.py:8-18)."""

    def __init__(self, d_model: int, max_len: int):
        super().__init__()
        self.register_buffer("pe", sinusoid_table(max_len, d_model))


def _require_cuda_half(t: torch.Tensor, what: str):
    if not t.is_cuda or t.dtype != torch.float16:
        raise RuntimeError(f"{what}: live2diff_b200 runs on CUDA fp16 tensors only (got {t.dtype} on {t.device}); "
                           "there is no CPU fallback")


class B200StreamTemporalAttention(nn.Module):
    def __init__(self, attention_mode=None, cross_frame_attention_mode=None, temporal_position_encoding=False,
                 temporal_position_encoding_max_len=32, window_size=8, sink_size=0, *, query_dim: int,
                 cross_attention_dim=None, heads: int = 8, dim_head: int = 64, dropout: float = 0.0, bias=False,
                 upcast_attention=False, **kwargs):
        super().__init__()
        if cross_attention_dim is not None:
            raise NotImplementedError("temporal cross-attention is not used by Live2Diff (Temporal_Self only)")
        if bias:
            raise NotImplementedError("attention_bias=True is not used by Live2Diff")
        inner = heads * dim_head
        assert inner == query_dim, "Live2Diff uses dim_head = C / heads"
        self.attention_mode = self._orig_attention_mode = attention_mode
        self.is_cross_attention = False
        self.heads, self.scale = heads, dim_head ** -0.5
        self.window_size, self.sink_size = window_size, sink_size
        self.cache_size = window_size - sink_size
        assert self.cache_size >= 0
        self.kv_channels = query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(query_dim, inner, bias=False)
        self.to_v = nn.Linear(query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(dropout)])
        self.pos_encoder = _PositionalEncoding(query_dim, temporal_position_encoding_max_len)
        self.motion_module_idx = None
        self._fused = None
        self._fused_key = None

    # -- reference API ---------------------------------------------------------------------
    def set_info(self, h: int, w: int, *args, **kwargs):
        self.h, self.w = h, w

    def set_index(self, idx):
        self.motion_module_idx = idx

    @torch.no_grad()
    def set_cache(self, denoising_steps_num: int):
        p = next(self.parameters())
        self.denoising_steps_num = denoising_steps_num
        return torch.zeros(denoising_steps_num, 2, self.h * self.w, self.window_size, self.kv_channels,
                           device=p.device, dtype=p.dtype)

    def _wqkv(self):
        key = (self.to_q.weight._version, self.to_k.weight._version, self.to_v.weight._version,
               self.to_q.weight.data_ptr())
        if self._fused is None or self._fused_key != key:
            self._fused = torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], dim=0).contiguous()
            self._fused_key = key
        return self._fused

    @torch.no_grad()
    def prepare_pe_buffer(self):
        _require_cuda_half(self.to_q.weight, "prepare_pe_buffer")
        c = self.kv_channels
        pe = self.pos_encoder.pe[0, : self.window_size].contiguous()
        tab = ops.gemm(pe, self._wqkv())                       # [L, 3C] = (q_pe | k_pe | v_pe), an fp16 Linear
        self.register_buffer("q_pe", tab[:, :c].contiguous()[None])
        self.register_buffer("k_pe", tab[:, c:2 * c].contiguous()[None])
        self.register_buffer("v_pe", tab[:, 2 * c:].contiguous()[None])

    @torch.no_grad()
    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, video_length=None,
                temporal_attention_mask=None, kv_cache=None, pe_idx=None, update_idx=None, *args, **kwargs):
        _require_cuda_half(hidden_states, "StreamTemporalAttention.forward")
        if video_length not in (None, 1):
            raise RuntimeError("streaming attention runs one frame per call (frame_buffer_size == 1)")
        n, hw, c = hidden_states.shape
        x = hidden_states.reshape(n * hw, c).contiguous()
        qkv = ops.gemm(x, self._wqkv())
        o = ops.kv_attn(qkv, qkv[:, c:], qkv[:, 2 * c:], kv_cache, self.q_pe[0], self.k_pe[0], self.v_pe[0],
                        temporal_attention_mask.to(torch.float16).contiguous(), pe_idx.contiguous(),
                        update_idx.contiguous(), self.heads, qkv_ld=3 * c)
        y = ops.gemm(o.reshape(n * hw, c), self.to_out[0].weight, bias=self.to_out[0].bias)
        return y.reshape(n, hw, c)


class _FF(nn.Module):
    """diffusers FeedForward(geglu) parameter layout: net.0.proj, net.2."""

    def __init__(self, dim):
        super().__init__()
        proj = nn.Module()
        proj.proj = nn.Linear(dim, 8 * dim)
        self.net = nn.ModuleList([proj, nn.Dropout(0.0), nn.Linear(4 * dim, dim)])


class _TemporalBlock(nn.Module):
    def __init__(self, dim, heads, head_dim, block_types, max_len, attn_kwargs):
        super().__init__()
        self.attention_blocks = nn.ModuleList([
            B200StreamTemporalAttention(attention_mode=bt.split("_")[0], query_dim=dim, heads=heads, dim_head=head_dim,
                                        temporal_position_encoding=True, temporal_position_encoding_max_len=max_len,
                                        **attn_kwargs) for bt in block_types])
        self.norms = nn.ModuleList([nn.LayerNorm(dim) for _ in block_types])
        self.ff = _FF(dim)
        self.ff_norm = nn.LayerNorm(dim)


class B200TemporalTransformer3DModel(nn.Module):
    def __init__(self, in_channels, num_attention_heads, attention_head_dim, num_layers=1,
                 attention_block_types=("Temporal_Self", "Temporal_Self"), dropout=0.0, norm_num_groups=32,
                 cross_attention_dim=1280, activation_fn="geglu", attention_bias=False, upcast_attention=False,
                 cross_frame_attention_mode=None, temporal_position_encoding=False,
                 temporal_position_encoding_max_len=32, attention_class_name="stream", attention_kwargs=None,
                 enable_streaming=True):
        super().__init__()
        if not (num_layers == 1 and len(attention_block_types) == 2 and attention_class_name == "stream"
                and enable_streaming and activation_fn == "geglu"):
            raise NotImplementedError("only the Live2Diff streaming configuration is implemented "
                                      "(1 block, 2 Temporal_Self stream attentions, GEGLU)")
        inner = num_attention_heads * attention_head_dim
        assert inner == in_channels
        self.in_channels, self.heads, self.groups = in_channels, num_attention_heads, norm_num_groups
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([
            _TemporalBlock(inner, num_attention_heads, attention_head_dim, attention_block_types,
                           temporal_position_encoding_max_len, dict(attention_kwargs or {}))])
        self.proj_out = nn.Linear(inner, in_channels)
        self.enable_streaming = True
        self._handle = None
        self._handle_key = None
        self._param_list = None

    def _apply(self, fn, *args, **kwargs):       # .to() / .half() / .cuda() replace the parameter tensors
        self._param_list = None
        return super()._apply(fn, *args, **kwargs)

    def __del__(self):
        self._free()

    def _free(self):
        if getattr(self, "_handle", None):
            lib().l2d_tt_destroy(self._handle)
            self._handle = None

    def _native(self, n_rows, h, w):
        # the key walks the cached parameter list only (no state_dict() per forward): the native object is rebuilt when the
        # geometry changes or a parameter was replaced (.to()/.half()/load_state_dict change data_ptr or bump _version)
        if self._param_list is None:
            self._param_list = list(self.parameters())
        key = (n_rows, h, w) + tuple((v._version, v.data_ptr()) for v in self._param_list)
        if self._handle is None or key != self._handle_key:
            self._free()
            self._param_list = list(self.parameters())
            key = (n_rows, h, w) + tuple((v._version, v.data_ptr()) for v in self._param_list)
            sd = {k: v for k, v in self.state_dict().items() if not k.endswith(("q_pe", "k_pe", "v_pe"))}
            window = self.transformer_blocks[0].attention_blocks[0].window_size
            named = {k: v.detach().contiguous() for k, v in sd.items()}
            for k, v in named.items():
                _require_cuda_half(v, k)
            arr, keep = make_tensor_table(named)
            h_out = C.c_void_p()
            check(lib().l2d_tt_create(C.byref(h_out), arr, len(named), self.in_channels, self.heads, self.groups, window,
                                      n_rows, h, w))
            self._handle, self._handle_key, self._keep = h_out, key, (named, keep)
        return self._handle

    @torch.no_grad()
    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, temporal_attention_mask=None,
                kv_cache: Optional[List[torch.Tensor]] = None, pe_idx=None, update_idx=None):
        assert hidden_states.dim() == 5, f"Expected hidden_states to have ndim=5, but got ndim={hidden_states.dim()}."
        _require_cuda_half(hidden_states, "TemporalTransformer3DModel.forward")
        n, c, f, h, w = hidden_states.shape
        if f != 1:
            raise RuntimeError("streaming mode processes one frame per call")
        blocks = self.transformer_blocks[0].attention_blocks
        caches = [kv_cache[b.motion_module_idx] for b in blocks]       # motion_module.py:416
        x = hidden_states.contiguous()
        y = torch.empty_like(x)
        check(lib().l2d_tt_forward(self._native(n, h, w), ptr(x), ptr(y), ptr(caches[0]), ptr(caches[1]),
                                   ptr(temporal_attention_mask.to(torch.float16).contiguous()),
                                   ptr(pe_idx.contiguous()), ptr(update_idx.contiguous()), current_stream()))
        return y
