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
            if n > 1:                                   # N == 1: the reference has no defined initial state (SURVEY A-1)
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
