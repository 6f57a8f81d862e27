"""Op-level parity of the CUDA kernels, called through the C ABI (live2diff_b200.ops -> libl2d_b200.so),
against fp32 torch evaluations of the same op on the same fp16 inputs (STRICT criterion, tests/parity.py)."""
import math

import pytest
import torch
import torch.nn.functional as F

from parity import referee, strict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def rnd(*shape, seed=0, scale=1.0, dev="cuda:0"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(device=dev, dtype=torch.float16)


# ------------------------------------------------------------------------------------------------
# tcgen05 GEMM
# ------------------------------------------------------------------------------------------------
GEMM_SHAPES = [
    # (M, N, K)  -- UNet shapes at 512^2 and edge cases
    (8192, 320, 320), (8192, 960, 320), (2048, 640, 640), (512, 1280, 1280), (128, 1280, 1280),
    (154, 640, 768),              # cross-attention K/V projection of the 77-token context
    (8192, 320, 2880),            # conv3x3 320->320 as GEMM over im2col
    (128, 1280, 11520),           # level-3 conv, long K
    (300, 96, 144),               # partial M tile, K not a multiple of 64, small N
    (16, 192, 64), (8, 64, 64), (4, 8, 64),   # tiny-config shapes: box larger than the tensor
    (2048, 16, 64),               # mapping network conv_in
]


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
def test_gemm_plain(m, n, k):
    from live2diff_b200 import ops

    a, w = rnd(m, k, seed=1), rnd(n, k, seed=2, scale=1 / math.sqrt(k))
    out = ops.gemm(a, w)
    strict(out, a.float() @ w.float().t(), f"gemm {m}x{n}x{k}")


@pytest.mark.parametrize("m,n,k", [(8192, 320, 320), (300, 96, 144), (512, 1280, 2560), (16, 192, 64)])
def test_gemm_epilogues(m, n, k):
    from live2diff_b200 import ops

    a, w = rnd(m, k, seed=3), rnd(n, k, seed=4, scale=1 / math.sqrt(k))
    bias, res = rnd(n, seed=5), rnd(m, n, seed=6)
    groups = 2 if m % 2 == 0 else 1
    rg = rnd(groups, n, seed=7)
    ref = a.float() @ w.float().t() + bias.float()
    strict(ops.gemm(a, w, bias=bias), ref, "bias")
    strict(ops.gemm(a, w, bias=bias, residual=res), ref + res.float(), "bias+residual")
    strict(ops.gemm(a, w, bias=bias, act=ops.ACT_SILU), F.silu(ref), "bias+silu")
    rgr = rg.float().repeat_interleave(m // groups, dim=0)
    strict(ops.gemm(a, w, bias=bias, rowgroup_bias=rg, rows_per_group=m // groups), ref + rgr, "bias+rowgroup")
    # in-place residual (out aliases residual), as the engine uses for  h += f(h)
    buf = res.clone()
    ops.gemm(a, w, bias=bias, residual=buf, out=buf)
    strict(buf, ref + res.float(), "in-place residual")


@pytest.mark.parametrize("m,c", [(8192, 320), (2048, 640), (512, 1280), (100, 64)])
def test_gemm_geglu(m, c):
    """FeedForward GEGLU (diffusers 0.25.0): proj C->8C, h * gelu(g); weight rows tile-interleaved."""
    from live2diff_b200 import ops

    x, w, b = rnd(m, c, seed=8), rnd(8 * c, c, seed=9, scale=1 / math.sqrt(c)), rnd(8 * c, seed=10, scale=0.1)
    tile = ops.gemm_tile_n(m, 8 * c, c)
    wi, bi = ops.geglu_interleave(w, b, tile)
    out = ops.gemm(x, wi, bias=bi, act=ops.ACT_GEGLU)
    proj = (x.float() @ w.float().t() + b.float())
    proj16 = proj.half().float()                      # the reference's Linear output is an fp16 tensor
    h, g = proj16.chunk(2, dim=-1)
    # two roundings (projection -> fp16, product -> fp16): an fp32 sum-order difference can flip the first
    # one by an ulp, so the bound is 2 fp16 ulps (~2e-3) instead of 1
    strict(out, h * F.gelu(g), f"geglu {m}x{c}", rtol=2.5e-3, atol=2e-4)


def test_gemm_strided_views():
    """A and W given as column slices of wider buffers (fused QKV, concat-shortcut K segments)."""
    from live2diff_b200 import ops

    m, c = 512, 320
    buf = rnd(m, 3 * c, seed=11)
    w = rnd(c, c, seed=12, scale=1 / math.sqrt(c))
    a = buf[:, c:2 * c]
    out = ops.gemm(a, w, lda=3 * c, m=m, k=c)
    strict(out, a.float() @ w.float().t(), "strided A")


# ------------------------------------------------------------------------------------------------
# norms / im2col / layout
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,c", [(8192, 320), (2048, 640), (512, 1280), (37, 64)])
def test_layernorm(rows, c):
    from live2diff_b200 import ops

    x, g, b = rnd(rows, c, seed=1, scale=2.0) + 0.5, rnd(c, seed=2) * 0.1 + 1, rnd(c, seed=3) * 0.1
    strict(ops.layernorm(x, g, b), F.layer_norm(x.float(), (c,), g.float(), b.float(), 1e-5), "layernorm")


def _nhwc(x_nchw):
    n, c, h, w = x_nchw.shape
    return x_nchw.permute(0, 2, 3, 1).reshape(n * h * w, c).contiguous()


@pytest.mark.parametrize("n,c,h,w,silu,eps", [(2, 320, 64, 64, True, 1e-5), (2, 1280, 8, 8, True, 1e-5),
                                              (2, 640, 32, 32, False, 1e-6), (1, 64, 4, 4, True, 1e-5)])
def test_groupnorm_nhwc(n, c, h, w, silu, eps):
    from live2diff_b200 import ops

    x = rnd(n, c, h, w, seed=4, scale=1.5) + 0.3
    g, b = rnd(c, seed=5) * 0.1 + 1, rnd(c, seed=6) * 0.1
    ref = F.group_norm(x.float(), 32, g.float(), b.float(), eps)
    ref = F.silu(ref) if silu else ref
    strict(ops.groupnorm(_nhwc(x), g, b, n, h, w, 32, eps, silu=silu), _nhwc(ref), "groupnorm")


@pytest.mark.parametrize("n,c1,c2,h,w", [(2, 640, 320, 32, 32), (2, 640, 320, 64, 64), (4, 320, 0, 64, 64), (2, 1280, 1280, 16, 16),
                                         (1, 320, 0, 64, 64), (3, 64, 0, 5, 7), (2, 1280, 640, 32, 32), (2, 320, 0, 96, 64)])
def test_groupnorm_one_kernel_path(n, c1, c2, h, w):
    """NHWC output (no im2col) takes the one-kernel path when every slab of every image can be resident at once: slab in
    shared memory via bulk copies, partial statistics, a grid barrier per image, merge + normalise from shared memory.
    Shapes: concat sources, the largest level-0 slab (960 channels), four images, odd pixel counts, config 3's plane."""
    from live2diff_b200 import ops

    x1, x2 = rnd(n, c1, h, w, seed=41, scale=1.3) + 0.2, (rnd(n, c2, h, w, seed=42) - 0.4 if c2 else None)
    x = x1 if x2 is None else torch.cat([x1, x2], dim=1)
    c = c1 + c2
    g, b = rnd(c, seed=43) * 0.1 + 1, rnd(c, seed=44) * 0.1
    ref = F.silu(F.group_norm(x.float(), 32, g.float(), b.float(), 1e-5))
    for _ in range(2):
        out = ops.groupnorm(_nhwc(x1), g, b, n, h, w, 32, 1e-5, silu=True, x2=None if x2 is None else _nhwc(x2))
    strict(out, _nhwc(ref), f"groupnorm one-kernel n{n} c{c1}+{c2} {h}x{w}")


def _im2col_ref(x_nchw, stride=1):
    """[N,C,H,W] -> [N*Ho*Wo, 9*C] with columns ordered (tap, channel) -- the engine's weight repack order."""
    n, c, h, w = x_nchw.shape
    cols = F.unfold(x_nchw, 3, padding=1, stride=stride)            # [N, C*9, L] ordered (channel, tap)
    L = cols.shape[-1]
    return cols.view(n, c, 9, L).permute(0, 3, 2, 1).reshape(n * L, 9 * c)


@pytest.mark.parametrize("c1,c2,h,w,stride", [(320, 0, 16, 16, 1), (640, 320, 16, 16, 1), (64, 64, 8, 8, 1),
                                              (320, 0, 16, 16, 2)])
def test_groupnorm_concat_im2col(c1, c2, h, w, stride):
    from live2diff_b200 import ops

    n = 2
    x1, x2 = rnd(n, c1, h, w, seed=7), (rnd(n, c2, h, w, seed=8) if c2 else None)
    x = x1 if x2 is None else torch.cat([x1, x2], dim=1)
    c = c1 + c2
    g, b = rnd(c, seed=9) * 0.1 + 1, rnd(c, seed=10) * 0.1
    ref = F.silu(F.group_norm(x.float(), 32, g.float(), b.float(), 1e-5)).half().float()
    out = ops.groupnorm(_nhwc(x1), g, b, n, h, w, 32, 1e-5, silu=True, x2=None if x2 is None else _nhwc(x2),
                        im2col=True, stride=stride)
    strict(out, _im2col_ref(ref, stride), "gn+silu+im2col")


@pytest.mark.parametrize("stride,up,silu", [(1, False, False), (2, False, False), (1, True, False), (1, False, True)])
def test_im2col(stride, up, silu):
    from live2diff_b200 import ops

    n, c, h, w = 2, 64, 8, 12
    x = rnd(n, c, h, w, seed=11)
    src = x.float()
    if silu:
        src = F.silu(src).half().float()
    if up:
        src = F.interpolate(src, scale_factor=2.0, mode="nearest")
    out = ops.im2col3x3(_nhwc(x), n, h, w, stride=stride, upsample2x=up, silu=silu)
    strict(out, _im2col_ref(src, stride), "im2col", rtol=0, atol=0 if not silu else 1e-3)


def test_im2col_nchw4_and_conv_equivalence():
    """conv3x3 == im2col x repacked weight (the path every 3x3 conv of the UNet takes)."""
    from live2diff_b200 import ops

    n, h, w, cout = 2, 16, 16, 64
    x = rnd(n, 4, h, w, seed=12)
    wt = rnd(cout, 4, 3, 3, seed=13, scale=0.2)
    cols = ops.im2col3x3_nchw4(x)
    wr = torch.zeros(cout, 64, dtype=torch.float16, device=x.device)
    wr[:, :36] = wt.permute(0, 2, 3, 1).reshape(cout, 36)           # (tap, cin)
    out = ops.gemm(cols, wr)
    strict(out, _nhwc(F.conv2d(x.float(), wt.float(), padding=1)), "conv_in as gemm")


def test_layout_roundtrip():
    from live2diff_b200 import ops

    x = rnd(2, 320, 1, 8, 8, seed=14)
    y = ops.nchw_to_nhwc(x)
    assert torch.equal(y, x[:, :, 0].permute(0, 2, 3, 1).reshape(128, 320))
    z = ops.nhwc_to_nchw(y, 2, x.shape)
    assert torch.equal(z, x)
    z2 = ops.nhwc_to_nchw(y, 2, x.shape, residual=x)
    assert torch.equal(z2, x + x)


# ------------------------------------------------------------------------------------------------
# small linears / timestep embedding / scheduler pointwise
# ------------------------------------------------------------------------------------------------
def test_time_embedding_path():
    from live2diff_b200 import ops
    from oracle import unet_oracle as O

    t = torch.tensor([399, 199], dtype=torch.int64, device="cuda:0")
    sin = ops.timestep_embedding(t, 320)
    strict(sin, O.timestep_sinusoid(t.cpu(), 320), "timestep sinusoid", rtol=1e-3, atol=1e-3)
    w1, b1 = rnd(1280, 320, seed=15, scale=1 / math.sqrt(320)), rnd(1280, seed=16, scale=0.1)
    out = ops.small_linear(sin, w1, b1, silu_out=True)
    strict(out, F.silu(sin.float() @ w1.float().t() + b1.float()), "time mlp 1")
    x = rnd(2, 1280, seed=17)
    w2, b2 = rnd(640, 1280, seed=18, scale=1 / math.sqrt(1280)), rnd(640, seed=19, scale=0.1)
    out = ops.small_linear(x, w2, b2, silu_in=True)
    strict(out, F.silu(x.float()).half().float() @ w2.float().t() + b2.float(), "time_emb_proj")


def test_lcm_step_matches_reference_golden():
    """scheduler_step_batch + stream-batch shift vs. the fixture produced by the reference's own method."""
    from helpers import load_golden
    from live2diff_b200 import ops
    from oracle import schedule_oracle as S

    g = load_golden("scheduler_pointwise.pt")
    dev = "cuda:0"
    x, eps = g["x"].half().to(dev), g["eps"].half().to(dev)
    consts = torch.stack([g["a"], g["b"], g["c_skip"], g["c_out"]]).float().to(dev).contiguous()
    noise = rnd(2, 4, 1, 8, 8, seed=20)
    out_last, nxt, x0 = ops.lcm_step(x, eps, consts, noise, want_x0=True)
    a16, b16, cs16, co16 = [v.half().float() for v in (g["a"], g["b"], g["c_skip"], g["c_out"])]
    ref = S.scheduler_step_batch(eps.float().cpu(), x.float().cpu(), cs16, co16, a16, b16)
    strict(x0, ref, "lcm x0 (fp16 path vs fp32)", rtol=4e-3, atol=4e-3)     # six fp16 roundings in the reference order
    assert torch.equal(out_last[0], x0[-1])
    ref_next = a16[1:].view(-1, 1, 1, 1, 1) * x0[:-1].float().cpu() + b16[1:].view(-1, 1, 1, 1, 1) * noise.float().cpu()
    strict(nxt, ref_next, "stream-batch shift", rtol=2e-3, atol=2e-3)
    # and against the reference's own fp32 output (golden)
    strict(x0, g["x0"], "lcm x0 vs reference golden", rtol=6e-3, atol=6e-3)


# ------------------------------------------------------------------------------------------------
# flash attention (spatial self / cross)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("b,heads,sq,skv,hd", [(2, 8, 1024, 1024, 40), (2, 8, 256, 256, 80), (2, 8, 64, 64, 160),
                                               (2, 8, 1024, 77, 40), (2, 8, 256, 77, 160), (2, 8, 16, 16, 8),
                                               (1, 8, 4, 77, 16), (2, 8, 4096, 4096, 40), (2, 8, 1024, 1024, 80),
                                               (2, 8, 1024, 77, 80), (2, 8, 256, 300, 40), (1, 8, 128, 128, 40),
                                               (3, 8, 384, 129, 80)])
def test_attention(b, heads, sq, skv, hd):
    from live2diff_b200 import ops

    c = heads * hd
    q, k, v = rnd(b * sq, c, seed=21), rnd(b * skv, c, seed=22), rnd(b * skv, c, seed=23)
    out = ops.attention(q, k, v, b, heads, sq, skv, hd)

    def split(t, s):
        return t.view(b, s, heads, hd).transpose(1, 2)

    ref32 = F.scaled_dot_product_attention(split(q.float(), sq), split(k.float(), skv), split(v.float(), skv))
    ref16 = F.scaled_dot_product_attention(split(q, sq), split(k, skv), split(v, skv))
    merge = lambda t: t.transpose(1, 2).reshape(b * sq, c)
    referee(out, merge(ref32), merge(ref16), f"attention b{b} s{sq}x{skv} hd{hd}")


@pytest.mark.parametrize("hd,sq", [(40, 1024), (80, 256)])
def test_attention_fused_qkv_views(hd, sq):
    """The engine's call pattern: q / k / v are column slices of one fused [M, 3C] projection buffer (row pitch 3C,
    16-byte aligned column offsets) -- the layout the tcgen05 kernel's TMA tensor maps are built over."""
    from live2diff_b200 import ops

    b, heads = 2, 8
    c = heads * hd
    qkv = rnd(b * sq, 3 * c, seed=27)
    q, k, v = qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:]
    out = ops.attention(q, k, v, b, heads, sq, sq, hd)

    def split(t):
        return t.reshape(b, sq, heads, hd).transpose(1, 2)

    ref32 = F.scaled_dot_product_attention(split(q.float()), split(k.float()), split(v.float()))
    ref16 = F.scaled_dot_product_attention(split(q), split(k), split(v))
    merge = lambda t: t.transpose(1, 2).reshape(b * sq, c)
    referee(out, merge(ref32), merge(ref16), f"attention fused-qkv views hd{hd} s{sq}")


@pytest.mark.parametrize("hd,sq,skv", [(40, 512, 1024), (80, 256, 640), (40, 256, 333)])
@pytest.mark.parametrize("growth", [0.25, 2.0, 12.0])
def test_attention_row_max_grows_across_key_tiles(hd, sq, skv, growth):
    """The tcgen05 kernel exponentiates key tile j against the row max of the EARLIER tiles and only redoes a tile when a
    row outgrows that reference by more than 2^8.  Keys whose scale rises tile after tile drive the row max up by `growth`
    (log2 units of the scaled score, roughly) per 128-key tile: 0.25 and 2 stay on the stale-reference path, 12 forces
    the redo (O and l rescaled in TMEM) on most tiles; a few rows also get one dominant late key."""
    from live2diff_b200 import ops

    b, heads = 2, 8
    c = heads * hd
    q, v = rnd(b * sq, c, seed=31), rnd(b * skv, c, seed=33)
    k = rnd(b * skv, c, seed=32).float()
    q = (q.float() * 0 + torch.sign(q.float())).half()          # +-1 queries: score = sum of +-k, spread ~ sqrt(hd) * |k|
    tile = (torch.arange(b * skv, device=k.device) % skv) // 128
    k = k * (growth * (tile.float() + 1.0) * 0.7)[:, None]       # score scale grows linearly with the key tile index
    k[skv - 3] *= 3.0                                             # one dominant key in the last tile of image 0
    k = k.half()
    out = ops.attention(q, k, v, b, heads, sq, skv, hd)

    def split(t, s):
        return t.view(b, s, heads, hd).transpose(1, 2)

    ref32 = F.scaled_dot_product_attention(split(q.float(), sq), split(k.float(), skv), split(v.float(), skv))
    ref16 = F.scaled_dot_product_attention(split(q, sq), split(k, skv), split(v, skv))
    merge = lambda t: t.transpose(1, 2).reshape(b * sq, c)
    assert torch.isfinite(out.float()).all()
    referee(out, merge(ref32), merge(ref16), f"attention growing max x{growth} hd{hd} s{sq}x{skv}")


def test_gemm_splitk_workspace_is_restored():
    """Small-M / long-K problems run split-K (fp32 atomics into a workspace that the last CTA of each tile
    zeroes again): back-to-back launches of different shapes must not see each other's partial sums."""
    from live2diff_b200 import ops

    for rep in range(3):
        for (m, n, k) in [(128, 1280, 11520), (512, 1280, 5760), (300, 96, 2304), (128, 1280, 11520)]:
            a, w = rnd(m, k, seed=30 + rep), rnd(n, k, seed=31, scale=1 / math.sqrt(k))
            bias, res = rnd(n, seed=32), rnd(m, n, seed=33)
            out = ops.gemm(a, w, bias=bias, residual=res)
            strict(out, a.float() @ w.float().t() + bias.float() + res.float(), f"split-K {m}x{n}x{k} rep {rep}")


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 64, 64, 320, 320), (2, 32, 32, 640, 640), (2, 8, 8, 1280, 1280),
                                            (2, 16, 16, 2560, 1280), (1, 8, 8, 64, 64), (2, 64, 96, 64, 64),
                                            (2, 64, 64, 320, 8), (2, 4, 4, 128, 128)])
def test_conv3x3_implicit_gemm(n, h, w, cin, cout):
    """conv3x3 pad 1 as an implicit GEMM (4-D TMA halo boxes, zero fill = padding), incl. temb bias + residual."""
    from live2diff_b200 import ops

    x = rnd(n, cin, h, w, seed=40)
    wt = rnd(cout, cin, 3, 3, seed=41, scale=1 / math.sqrt(9 * cin))
    bias, temb, res = rnd(cout, seed=42, scale=0.1), rnd(n, cout, seed=43), rnd(n, cout, h, w, seed=44)
    wr = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()          # (tap, cin) columns
    out = ops.conv3x3(_nhwc(x), n, h, w, wr, bias=bias, rowgroup_bias=temb, residual=_nhwc(res))
    ref = F.conv2d(x.float(), wt.float(), bias.float(), padding=1) + temb.float()[:, :, None, None] + res.float()
    strict(out, _nhwc(ref), f"conv3x3 {n}x{h}x{w} {cin}->{cout}")
