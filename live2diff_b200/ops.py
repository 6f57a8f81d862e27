"""Tensor-level Python entry points of the op kernels (thin: argument checks + the C-ABI call).

Used by the module drop-ins and by the parity tests, which therefore exercise the same ABI
(include/l2d_b200.h) a non-Python host would bind.  All tensors: CUDA, fp16, contiguous unless a
stride argument says otherwise.  Nothing here computes on the CPU or through torch ops.
"""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import check, current_stream, lib, ptr

ACT_NONE, ACT_SILU, ACT_GEGLU = 0, 1, 2


def _chk(t: torch.Tensor, name: str, dtype=torch.float16):
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError(f"{name}: expected a contiguous CUDA {dtype} tensor, got {t.dtype} {t.device} "
                         f"contiguous={t.is_contiguous()}")


def kv_attn(q, k_new, v_new, kv_cache, q_pe, k_pe, v_pe, mask, pe_idx, update_idx, heads, qkv_ld=None, out=None):
    """K1.  q/k_new/v_new [N,hw,C] (or views into a fused buffer with row pitch qkv_ld); kv_cache [N,2,hw,L,C]."""
    n, _, hw, L, c = kv_cache.shape
    _chk(kv_cache, "kv_cache")
    for t, nm in ((q_pe, "q_pe"), (k_pe, "k_pe"), (v_pe, "v_pe"), (mask, "mask")):
        _chk(t, nm)
    _chk(pe_idx, "pe_idx", torch.int64)
    _chk(update_idx, "update_idx", torch.int64)
    if out is None:
        out = torch.empty(n, hw, c, dtype=torch.float16, device=kv_cache.device)
    ld = qkv_ld if qkv_ld is not None else c
    check(lib().l2d_kv_attn(ptr(q), ptr(k_new), ptr(v_new), ld, ptr(kv_cache), ptr(q_pe), ptr(k_pe), ptr(v_pe),
                            ptr(mask), ptr(pe_idx), ptr(update_idx), ptr(out), n, hw, L, c, heads, current_stream()))
    return out


def warmup_attn(q, k, v, kv_cache_row, q_pe, k_pe, v_pe, heads, qkv_ld=None, pe_ld=None, out=None):
    """Warm-up (bidirectional) temporal attention of one clip + fill of cache slots 0..F-1 (csrc/kv_warmup.cu).
    q/k/v [F,hw,C] (or views into a fused buffer with row pitch qkv_ld); kv_cache_row [2,hw,L,C] = cache[idx];
    q_pe/k_pe/v_pe rows 0..F-1 (pitch pe_ld).  Returns out [F,hw,C]."""
    f, hw, c = q.shape
    _chk(kv_cache_row, "kv_cache_row")
    if tuple(kv_cache_row.shape[:2]) != (2, hw) or kv_cache_row.shape[3] != c:
        raise ValueError(f"kv_cache_row {tuple(kv_cache_row.shape)}: expected [2,{hw},L,{c}]")
    if out is None:
        out = torch.empty(f, hw, c, dtype=torch.float16, device=q.device)
    check(lib().l2d_warmup_attn(ptr(q), ptr(k), ptr(v), qkv_ld if qkv_ld is not None else c, ptr(kv_cache_row), ptr(q_pe),
                                ptr(k_pe), ptr(v_pe), pe_ld if pe_ld is not None else c, ptr(out), c, f, hw,
                                kv_cache_row.shape[2], c, heads, current_stream()))
    return out


def layernorm(x, gamma, beta, eps=1e-5):
    _chk(x, "x")
    y = torch.empty_like(x)
    c = x.shape[-1]
    check(lib().l2d_layernorm(ptr(x), ptr(gamma), ptr(beta), ptr(y), x.numel() // c, c, eps, current_stream()))
    return y


def groupnorm(x1, gamma, beta, n_img, h, w, groups, eps, silu=False, x2=None, im2col=False, stride=1):
    """x1 [N*h*w, C1] (+ optional x2 [N*h*w, C2], channel concat).  Returns NHWC rows or the 3x3 im2col matrix."""
    _chk(x1, "x1")
    c1 = x1.shape[-1]
    c2 = 0 if x2 is None else x2.shape[-1]
    c = c1 + c2
    ws = torch.zeros(lib().l2d_groupnorm_workspace_bytes(n_img, groups), dtype=torch.uint8, device=x1.device)
    if im2col:
        y = torch.empty(n_img * (h // stride) * (w // stride), 9 * c, dtype=torch.float16, device=x1.device)
    else:
        y = torch.empty(n_img * h * w, c, dtype=torch.float16, device=x1.device)
    check(lib().l2d_groupnorm(ptr(x1), c1, ptr(x2), c2, ptr(gamma), ptr(beta), ptr(y), ptr(ws), n_img, h, w, groups,
                              eps, int(silu), int(im2col), stride, current_stream()))
    return y


def im2col3x3(x, n_img, h, w, stride=1, upsample2x=False, silu=False):
    _chk(x, "x")
    c = x.shape[-1]
    hs, ws_ = (2 * h, 2 * w) if upsample2x else (h, w)
    y = torch.empty(n_img * (hs // stride) * (ws_ // stride), 9 * c, dtype=torch.float16, device=x.device)
    check(lib().l2d_im2col3x3(ptr(x), ptr(y), n_img, h, w, c, stride, int(upsample2x), int(silu), current_stream()))
    return y


def im2col3x3_nchw4(x):
    _chk(x, "x")
    n, c, h, w = x.shape
    assert c == 4
    y = torch.empty(n * h * w, 64, dtype=torch.float16, device=x.device)
    check(lib().l2d_im2col3x3_nchw4(ptr(x), ptr(y), n, h, w, current_stream()))
    return y


def nchw_to_nhwc(x):
    _chk(x, "x")
    n, c = x.shape[0], x.shape[1]
    hw = x.numel() // (n * c)
    y = torch.empty(n * hw, c, dtype=torch.float16, device=x.device)
    check(lib().l2d_nchw_to_nhwc(ptr(x), ptr(y), n, c, hw, current_stream()))
    return y


def nhwc_to_nchw(x, n_img, shape, residual=None):
    _chk(x, "x")
    c = x.shape[-1]
    y = torch.empty(shape, dtype=torch.float16, device=x.device)
    check(lib().l2d_nhwc_to_nchw(ptr(x), ptr(residual), ptr(y), n_img, c, x.shape[0] // n_img, current_stream()))
    return y


def gemm_tile_n(m, n, k):
    return lib().l2d_gemm_tile_n(m, n, k)


def geglu_interleave(w, b, tile_n):
    _chk(w, "w")
    two_f, k = w.shape
    wo, bo = torch.empty_like(w), (torch.empty_like(b) if b is not None else None)
    check(lib().l2d_geglu_interleave(ptr(w), ptr(b), ptr(wo), ptr(bo), two_f, k, tile_n, current_stream()))
    return wo, bo


def gemm(a, w, bias=None, rowgroup_bias=None, rows_per_group=0, residual=None, act=ACT_NONE, out=None, lda=None,
         m=None, k=None):
    """out = epilogue(a @ w.T).  a [M,K] (row pitch lda), w [N,K]."""
    _chk(w, "w")
    n, kk = w.shape
    k = kk if k is None else k
    m = a.shape[0] if m is None else m
    lda = (a.stride(0) if a.dim() == 2 else k) if lda is None else lda
    n_out = n // 2 if act == ACT_GEGLU else n
    if out is None:
        out = torch.empty(m, n_out, dtype=torch.float16, device=w.device)
    ldr = residual.stride(0) if residual is not None else 0
    check(lib().l2d_gemm(ptr(a), lda, ptr(w), ptr(out), out.stride(0), m, n, k, ptr(bias), ptr(rowgroup_bias),
                         rows_per_group, ptr(residual), ldr, act, current_stream()))
    return out


def conv3x3(x, n_img, h, w, weight, bias=None, rowgroup_bias=None, residual=None, act=ACT_NONE):
    """Implicit-GEMM conv3x3 (pad 1, stride 1).  x [N*h*w, Cin] channels-last, weight [Cout, 9*Cin] (tap, cin)."""
    _chk(x, "x")
    _chk(weight, "weight")
    cin = x.shape[-1]
    cout = weight.shape[0]
    out = torch.empty(n_img * h * w, cout, dtype=torch.float16, device=x.device)
    ldr = residual.stride(0) if residual is not None else 0
    check(lib().l2d_conv3x3(ptr(x), n_img, h, w, cin, ptr(weight), ptr(out), out.stride(0), cout, ptr(bias),
                            ptr(rowgroup_bias), ptr(residual), ldr, act, current_stream()))
    return out


def small_linear(x, w, b=None, silu_in=False, silu_out=False):
    _chk(x, "x")
    _chk(w, "w")
    m, k = x.shape
    n = w.shape[0]
    out = torch.empty(m, n, dtype=torch.float16, device=x.device)
    check(lib().l2d_small_linear(ptr(x), ptr(w), ptr(b), ptr(out), m, n, k, int(silu_in), int(silu_out),
                                 current_stream()))
    return out


def timestep_embedding(t, dim):
    _chk(t, "t", torch.int64)
    out = torch.empty(t.shape[0], dim, dtype=torch.float16, device=t.device)
    check(lib().l2d_timestep_embedding(ptr(t), ptr(out), t.shape[0], dim, current_stream()))
    return out


def attention(q, k, v, batch, heads, sq, skv, hd, ldq=None, ldk=None, ldv=None):
    """q rows (b*sq+i) with pitch ldq; returns [batch*sq, heads*hd]."""
    out = torch.empty(batch * sq, heads * hd, dtype=torch.float16, device=q.device)
    ldq = q.stride(0) if ldq is None else ldq
    ldk = k.stride(0) if ldk is None else ldk
    ldv = v.stride(0) if ldv is None else ldv
    check(lib().l2d_attention(ptr(q), ldq, ptr(k), ldk, ptr(v), ldv, ptr(out), out.stride(0), batch, heads, sq, skv,
                              hd, current_stream()))
    return out


def lcm_step(x_t, eps, consts, noise=None, want_x0=False):
    """x_t, eps [N,4,1,h,w] fp16; consts fp32 [4,N] = (sqrt(abar), sqrt(1-abar), c_skip, c_out).
    Returns (out_last [1,...], next_buf [(N-1),...] or None, x0_all or None)."""
    _chk(x_t, "x_t")
    _chk(eps, "eps")
    _chk(consts, "consts", torch.float32)
    n = x_t.shape[0]
    per = x_t.numel() // n
    out_last = torch.empty((1,) + tuple(x_t.shape[1:]), dtype=torch.float16, device=x_t.device)
    nxt = torch.empty((n - 1,) + tuple(x_t.shape[1:]), dtype=torch.float16, device=x_t.device) if n > 1 else None
    x0 = torch.empty_like(x_t) if want_x0 else None
    check(lib().l2d_lcm_step(ptr(x_t), ptr(eps), ptr(consts), ptr(noise), ptr(x0), ptr(out_last), ptr(nxt), n, per,
                             current_stream()))
    return out_last, nxt, x0
