"""Device-resident stream state, the parts that can be pinned without a GPU: the ring-schedule
transition and the counter-based generator are `__host__ __device__` code (csrc/stream_state.cuh); the library exports
host evaluations of exactly that code, checked here against the trace of the reference's own `update_attn_bias`
(tests/golden/schedule_trace.json) and against the published Philox4x32-10 known-answer vectors."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN
from live2diff_b200 import _lib


def run_schedule(n_rows, window, warmup, frames):
    lib = _lib.lib()
    valid = (C.c_int32 * n_rows)()
    pe = (C.c_int64 * (n_rows * window))()
    up = (C.c_int64 * n_rows)()
    _lib.check(lib.l2d_ring_schedule_host(valid, pe, up, n_rows, window, warmup, 1, 0))
    for _ in range(frames):
        yield list(valid), [list(pe[r * window:(r + 1) * window]) for r in range(n_rows)], list(up)
        _lib.check(lib.l2d_ring_schedule_host(valid, pe, up, n_rows, window, warmup, 0, 1))


def test_device_schedule_code_matches_reference_trace():
    traces = json.load(open(os.path.join(GOLDEN, "schedule_trace.json")))
    checked = 0
    for tr in traces:
        if "frames" not in tr:
            continue
        n, L, w0 = tr["n_rows"], tr["window"], tr["warmup"]
        for f, (valid, pe, up) in enumerate(run_schedule(n, L, w0, len(tr["frames"]))):
            ref = tr["frames"][f]
            assert [sum(row) for row in ref["valid"]] == valid, (n, L, f)
            assert all(row[:v] == [1] * v for row, v in zip(ref["valid"], valid))          # unmasked slots are a prefix
            assert ref["pe_idx"] == pe, (n, L, f)
            assert ref["update_idx"] == up, (n, L, f)
            checked += 1
    assert checked > 100


def test_host_schedule_matches_python_ring_schedule():
    from live2diff_b200.schedule import RingSchedule

    for n, L, w0 in ((2, 16, 8), (4, 32, 8), (1, 4, 2), (3, 8, 7)):
        rs = RingSchedule(n, L, w0)
        for valid, pe, up in run_schedule(n, L, w0, 3 * L):
            # (N == 1: the reference's initialiser raises, SURVEY A-1; both implementations guard it the same way)
            assert (valid, pe, up) == (rs.valid, rs.pe_idx, rs.update_idx)
            rs.advance()


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors, philox4x32 10 rounds."""
    lib = _lib.lib()
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        out = (C.c_uint32 * 4)()
        lib.l2d_philox4x32_10_host((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert tuple(out) == want


def test_stream_randn_is_standard_normal_and_keyed():
    lib = _lib.lib()

    def draw(seed, frame, row, n=1 << 16):
        buf = (C.c_float * n)()
        _lib.check(lib.l2d_stream_randn_host(seed, frame, row, buf, n))
        return np.frombuffer(buf, dtype=np.float32).copy()

    a = draw(2, 0, 0)
    assert np.isfinite(a).all()
    assert abs(a.mean()) < 0.02 and abs(a.std() - 1.0) < 0.02
    assert abs(np.mean(a ** 3)) < 0.05 and abs(np.mean(a ** 4) - 3.0) < 0.15            # skew, kurtosis
    assert abs(np.corrcoef(a[:-1], a[1:])[0, 1]) < 0.02                                  # the cos/sin pair is uncorrelated
    assert np.array_equal(a, draw(2, 0, 0))                                              # deterministic
    for other in (draw(3, 0, 0), draw(2, 1, 0), draw(2, 0, 1)):                          # seed / frame / row all key it
        assert abs(np.corrcoef(a, other)[0, 1]) < 0.02


def test_config1_plumbing_stream_on_the_oracle():
    """BASELINE configs[0] geometry -- 1 denoise step, KV window 4 with 2 sink slots, CPU fp32 -- through the stream
    oracle (tiny channel widths keep it to seconds): fill phase, first wrap and steady state of a single-row stream;
    the oracle's tensor schedule, the Python RingSchedule and the device schedule code agree frame by frame."""
    import torch

    from live2diff_b200.schedule import RingSchedule
    from live2diff_b200.weights import UNetDims, random_state_dict
    from oracle import schedule_oracle as S
    from oracle import unet_oracle as O

    d = UNetDims(block_out_channels=(32, 64, 64, 64), cross_attention_dim=64, window_size=4, sink_size=2, pe_max_len=24)
    od = O.UNetDims(**d.__dict__)
    sd = random_state_dict(d, seed=3)
    n, h, w = 1, 8, 8
    gen = torch.Generator().manual_seed(0)
    kv = O.alloc_kv_cache(od, n, h, w)
    for c in kv:
        c[:, :, :, :2] = torch.randn(c[:, :, :, :2].shape, generator=gen)
    prompt = torch.randn(n, 77, 64, generator=gen)
    orc = S.StreamOracle(lambda s_, t, **kw: O.unet_forward(sd, od, s_, t, kw["encoder_hidden_states"],
                                                            kw["temporal_attention_mask"], kw["depth_sample"],
                                                            kw["kv_cache"], kw["pe_idx"], kw["update_idx"]),
                         kv, prompt, [40], (h, w), window=4, warmup=2)
    rs = RingSchedule(n, 4, 2)
    dev_sched = run_schedule(n, 4, 2, 8)
    outs = []
    for f in range(8):
        valid, pe, up = next(dev_sched)
        assert (valid, pe, up) == (rs.valid, rs.pe_idx, rs.update_idx) == (
            (orc.attn_bias == 0).sum(1).tolist(), orc.pe_idx.tolist(), orc.update_idx.tolist()), f
        x0 = orc.step(torch.randn(1, 4, 1, h, w, generator=gen), torch.randn(1, 4, 1, h, w, generator=gen), None)
        assert x0.shape == (1, 4, 1, h, w) and torch.isfinite(x0).all()
        outs.append(x0)
        rs.advance()
    assert rs.valid == [4] and sorted(rs.pe_idx[0]) == [0, 1, 2, 3]
    assert float((outs[-1] - outs[-2]).abs().max()) > 0                      # a live stream, not a constant
