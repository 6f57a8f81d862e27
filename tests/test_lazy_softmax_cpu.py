"""CPU model of the attention kernel's online softmax (flash_tcgen05.cu): key tiles of 128, the exp of tile j taken against the
row max of the EARLIER tiles (lazy reference), O and l rescaled only when a row outgrows its reference by more than 2^8, P
rounded to fp16 before P.V, fp32 accumulation.  The GPU tests check the kernel against SDPA; this file checks the ALGORITHM
against exact softmax attention on the host, including the inputs that drive every branch (reference never moves, moves a
little every tile, jumps past the threshold, ragged last tile), so a wrong rescale rule cannot hide behind a tolerance."""
import math

import pytest
import torch

LAZY = 8.0


def lazy_attention(q, k, v, tile=128):
    """q [sq, hd], k / v [skv, hd] (fp16 values held in fp32); returns (out fp32 [sq, hd], rescale events, max P seen)."""
    sq, hd = q.shape
    skv = k.shape[0]
    c = math.log2(math.e) / math.sqrt(hd)
    m_ref = torch.full((sq,), -math.inf)
    l_run = torch.zeros(sq)
    o = torch.zeros(sq, hd)
    rescales, p_max = 0, 0.0
    for j0 in range(0, skv, tile):
        s = q @ k[j0:j0 + tile].T                         # fp32 scores of the tile (the tensor core accumulates in fp32)
        mx = s.max(dim=1).values
        if j0 == 0:
            m_ref = mx.clone()
        else:
            # the kernel decides per WARP (32 rows): any row over the threshold moves every row of its warp to its new max
            over = (mx * c - m_ref * c) > LAZY
            for w0 in range(0, sq, 32):
                if bool(over[w0:w0 + 32].any()):
                    rows = slice(w0, w0 + 32)
                    m_new = torch.maximum(m_ref[rows], mx[rows])
                    corr = torch.exp2((m_ref[rows] - m_new) * c)
                    o[rows] *= corr[:, None]
                    l_run[rows] *= corr
                    m_ref[rows] = m_new
                    rescales += 1
        p = torch.exp2(s * c - (m_ref * c)[:, None])
        p_max = max(p_max, float(p.max()))
        l_run += p.sum(dim=1)                             # the row sum uses the unrounded exponentials, like the kernel
        o += p.half().float() @ v[j0:j0 + tile]           # P is rounded to fp16 for the tensor core
    return o / l_run[:, None], rescales, p_max


def exact_attention(q, k, v):
    s = (q.double() @ k.double().T) / math.sqrt(q.shape[1])
    return (torch.softmax(s, dim=1) @ v.double()).float()


def _inputs(sq, skv, hd, growth, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.sign(torch.randn(sq, hd, generator=g)).half().float()
    k = torch.randn(skv, hd, generator=g)
    tile = torch.arange(skv) // 128
    k = (k * (growth * (tile.float() + 1.0) * 0.7)[:, None]).half().float()
    v = torch.randn(skv, hd, generator=g).half().float()
    return q, k, v


@pytest.mark.parametrize("growth,expect_rescale", [(0.0, False), (0.25, False), (12.0, True)])
@pytest.mark.parametrize("sq,skv,hd", [(64, 1024, 40), (32, 333, 80)])
def test_lazy_reference_matches_exact_softmax(sq, skv, hd, growth, expect_rescale):
    if growth == 0.0:
        g = torch.Generator().manual_seed(5)
        q, k, v = (torch.randn(n, hd, generator=g).half().float() for n in (sq, skv, skv))
    else:
        q, k, v = _inputs(sq, skv, hd, growth, seed=9)
    out, rescales, p_max = lazy_attention(q, k, v)
    ref = exact_attention(q, k, v)
    err = float((out - ref).abs().max())
    scale = float(ref.abs().max())
    print(f"[lazy softmax] sq{sq} skv{skv} hd{hd} growth {growth}: max err {err:.2e} (scale {scale:.2f}), {rescales} warp rescales, "
          f"largest P {p_max:.1f}")
    assert torch.isfinite(out).all()
    assert p_max <= 2.0 ** LAZY * 1.0001, "P must stay below 2^LAZY: that is what keeps it inside fp16"
    assert err <= 2e-3 * max(scale, 1.0)                  # fp16 rounding of P: 2^-11 relative per term
    assert (rescales > 0) == expect_rescale or growth == 0.25   # small growth may or may not cross the threshold


def test_stale_reference_equals_exact_reference_up_to_rounding():
    """With no fp16 rounding of P the lazy reference is algebraically exact: any reference cancels in O / l."""
    q, k, v = _inputs(32, 640, 40, 2.0, seed=3)
    sq, hd = q.shape
    c = math.log2(math.e) / math.sqrt(hd)
    m_ref = (q @ k[:128].T).max(dim=1).values            # the first tile's max, never updated: reference up to 2^? stale
    s = q @ k.T
    p = torch.exp2((s.double() * c) - (m_ref.double() * c)[:, None])
    out = (p @ v.double()) / p.sum(dim=1, keepdim=True)
    torch.testing.assert_close(out.float(), exact_attention(q, k, v), rtol=1e-5, atol=1e-6)
