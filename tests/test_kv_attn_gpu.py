"""K1 (temporal KV-cache attention) parity: golden fixtures from the reference, the oracle at full
BASELINE sizes, and exactness properties of the cache append."""
import pytest
import torch

from helpers import load_golden, prefixed, regen_weights, sub_spec
from live2diff_b200.weights import UNetDims
from oracle import schedule_oracle as S
from oracle import unet_oracle as O
from parity import referee

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def frames(n_rows, window, warmup, count):
    ab, pe, up = S.init_schedule(n_rows, window, warmup)
    for _ in range(count):
        yield ab.clone(), pe.clone(), up.clone()
        S.update_schedule(ab, pe, up, window, warmup)


@pytest.mark.parametrize("tag", ["c64_fill_wrap", "c320_hd40", "c128_L32_N4", "c64_N1_L4"])
def test_stream_attention_module_vs_reference_golden(tag):
    """B200StreamTemporalAttention (drop-in B3) through fill phase, first wrap and steady state."""
    from live2diff_b200.modules import B200StreamTemporalAttention

    g = load_golden(f"stream_attention_{tag}.pt")
    ch, heads, L, W0 = g["ch"], g["heads"], g["window"], g["warmup"]
    d = UNetDims(block_out_channels=(ch,), heads=heads, window_size=L, sink_size=W0, pe_max_len=g["pe_max"],
                 down_has_attn=(False,), up_has_attn=(False,))
    pre = "down_blocks.0.motion_modules.0.temporal_transformer.transformer_blocks.0.attention_blocks.0"
    w = regen_weights(sub_spec(d, pre), g["seed"], g["fingerprint"])
    attn = B200StreamTemporalAttention(attention_mode="Temporal", query_dim=ch, heads=heads, dim_head=ch // heads,
                                       temporal_position_encoding=True, temporal_position_encoding_max_len=g["pe_max"],
                                       window_size=L, sink_size=W0)
    attn.load_state_dict(w, strict=True)
    attn = attn.to(DEV).half()
    attn.set_info(g["h"], g["w"])
    attn.set_index(0)
    cache = attn.set_cache(g["n_rows"])
    attn.prepare_pe_buffer()
    cache.copy_(g["cache0"].half())
    # torch-fp16 evaluation of the oracle restatement = the reference's own fp16 rounding order (referee)
    sd16 = {k: v.to(DEV).half() for k, v in prefixed(w, "a").items()}
    cache16 = g["cache0"].half().to(DEV)
    od = O.UNetDims(**d.__dict__)
    for f, (mask, pe_idx, update_idx) in enumerate(frames(g["n_rows"], L, W0, g["x"].shape[0])):
        x = g["x"][f].half().to(DEV)
        y = attn(x, video_length=1, temporal_attention_mask=mask.half().to(DEV), kv_cache=cache,
                 pe_idx=pe_idx.to(DEV), update_idx=update_idx.to(DEV))
        y16 = O.stream_temporal_attention(sd16, "a", x, cache16, mask.half().to(DEV), pe_idx.to(DEV), update_idx.to(DEV), od)
        referee(y, g["y"][f], y16, f"stream_attention[{tag}] frame {f}")
    referee(cache, g["cache_final"], cache16, f"stream_attention[{tag}] final cache")


def _random_case(n, hw, L, c, seed, steady=True):
    gen = torch.Generator().manual_seed(seed)
    mk = lambda *s: torch.randn(*s, generator=gen).half().to(DEV)
    q, k, v = mk(n, hw, c), mk(n, hw, c), mk(n, hw, c)
    cache = mk(n, 2, hw, L, c)
    pe = [mk(L, c) * 0.5 for _ in range(3)]
    ab, pi, up = S.init_schedule(n, L, 8 if L >= 16 else 2)
    for _ in range(3 * L if steady else 2):
        S.update_schedule(ab, pi, up, L, 8 if L >= 16 else 2)
    return q, k, v, cache, pe, ab.half().to(DEV), pi.to(DEV), up.to(DEV)


@pytest.mark.parametrize("n,hw,L,c,heads", [(2, 4096, 16, 320, 8), (2, 1024, 16, 640, 8), (2, 256, 16, 1280, 8),
                                            (2, 64, 16, 1280, 8), (4, 1024, 32, 640, 8), (1, 4096, 4, 320, 8),
                                            (4, 4096, 32, 320, 8), (4, 256, 32, 1280, 8), (4, 64, 32, 1280, 8)])
def test_kv_attn_full_size_vs_oracle(n, hw, L, c, heads):
    """BASELINE.json sizes (configs 1, 2, 4): steady state, all slots valid."""
    from live2diff_b200 import ops

    q, k, v, cache, pe, mask, pi, up = _random_case(n, hw, L, c, seed=hw + L)
    cache_ref32, cache_ref16 = cache.float(), cache.clone()
    out = ops.kv_attn(q, k, v, cache, pe[0], pe[1], pe[2], mask, pi, up, heads)
    ref32 = O.kv_cache_attention(q.float(), k.float(), v.float(), cache_ref32, pe[0].float(), pe[1].float(), pe[2].float(),
                                 mask.float(), pi, up, heads)
    ref16 = O.kv_cache_attention(q, k, v, cache_ref16, pe[0], pe[1], pe[2], mask, pi, up, heads)
    referee(out, ref32, ref16, f"kv_attn N{n} hw{hw} L{L} C{c}")
    assert torch.equal(cache, cache_ref16), "cache append must be bit-exact"


@pytest.mark.parametrize("c,hw", [(320, 67), (640, 33), (320, 1)])
def test_kv_attn_window32_through_fill_and_wrap(c, hw):
    """Window 32 on the tensor-core path (two 16-slot blocks per pixel, softmax exchanged between the two pixel groups):
    every frame from the initial schedule (9 valid slots: the second block fully masked) through the fill phase, the
    block boundary (write slot 15 -> 16) and the first wrap of the ring, incl. ragged tiles (odd hw) -- against the oracle
    on a cache both sides advance in lock-step; the append must stay bit-exact."""
    from live2diff_b200 import ops

    n, L, heads, W0 = 4, 32, 8, 8
    gen = torch.Generator().manual_seed(c + hw)
    mk = lambda *s: torch.randn(*s, generator=gen).half().to(DEV)
    cache = torch.zeros(n, 2, hw, L, c, dtype=torch.float16, device=DEV)
    cache[:, :, :, :W0] = mk(n, 2, hw, W0, c)
    cache32, cache16 = cache.float(), cache.clone()
    pe = [mk(L, c) * 0.5 for _ in range(3)]
    for f, (ab, pi, up) in enumerate(frames(n, L, W0, 30)):
        q, k, v = mk(n, hw, c), mk(n, hw, c), mk(n, hw, c)
        mask, pi, up = ab.half().to(DEV), pi.to(DEV), up.to(DEV)
        out = ops.kv_attn(q, k, v, cache, pe[0], pe[1], pe[2], mask, pi, up, heads)
        ref32 = O.kv_cache_attention(q.float(), k.float(), v.float(), cache32, pe[0].float(), pe[1].float(), pe[2].float(),
                                     mask.float(), pi, up, heads)
        ref16 = O.kv_cache_attention(q, k, v, cache16, pe[0], pe[1], pe[2], mask, pi, up, heads)
        referee(out, ref32, ref16, f"kv_attn L32 C{c} hw{hw} frame {f}")
        assert torch.equal(cache, cache16), f"frame {f}: cache append must be bit-exact"


def test_kv_attn_append_properties():
    """Size-independent invariants: only slot update_idx[n] of row n changes, it holds k/v bit-exactly,
    masked slots never influence the output, fused-QKV strided input gives identical bits."""
    from live2diff_b200 import ops

    n, hw, L, c, heads = 2, 512, 16, 320, 8
    q, k, v, cache, pe, mask, pi, up = _random_case(n, hw, L, c, seed=5, steady=False)   # fill phase: some slots masked
    before = cache.clone()
    out = ops.kv_attn(q, k, v, cache, pe[0], pe[1], pe[2], mask, pi, up, heads)
    for r in range(n):
        u = int(up[r])
        assert torch.equal(cache[r, 0, :, u], k[r]) and torch.equal(cache[r, 1, :, u], v[r])
        keep = [j for j in range(L) if j != u]
        assert torch.equal(cache[r][:, :, keep], before[r][:, :, keep])
    # poison masked slots (that are not the write slot) with huge values: output must not change
    cache2 = before.clone()
    for r in range(n):
        for j in range(L):
            if torch.isinf(mask[r, j]) and j != int(up[r]):
                cache2[r, :, :, j] = 6e4
    out2 = ops.kv_attn(q, k, v, cache2, pe[0], pe[1], pe[2], mask, pi, up, heads)
    assert torch.equal(out, out2)
    # fused [M,3C] buffer, strided views
    qkv = torch.cat([q, k, v], dim=-1).reshape(n * hw, 3 * c).contiguous()
    cache3 = before.clone()
    out3 = ops.kv_attn(qkv, qkv[:, c:], qkv[:, 2 * c:], cache3, pe[0], pe[1], pe[2], mask, pi, up, heads, qkv_ld=3 * c)
    assert torch.equal(out, out3) and torch.equal(cache, cache3)
