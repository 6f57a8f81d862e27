"""Ring-schedule / scheduler-pointwise oracle vs. the reference's own methods (executed from their
source text by make_golden.py) and vs. the docstring sketch at pipeline_stream_animation_depth.py:417-421."""
import json
import os

import pytest
import torch

from helpers import GOLDEN, load_golden
from oracle import schedule_oracle as S


def traces():
    return json.load(open(os.path.join(GOLDEN, "schedule_trace.json")))


@pytest.mark.parametrize("i", range(4))
def test_schedule_trace_matches_reference(i):
    tr = traces()[i]
    ab, pe, up = S.init_schedule(tr["n_rows"], tr["window"], tr["warmup"])
    for f in tr["frames"]:
        assert (ab == 0).int().tolist() == f["valid"]
        assert pe.tolist() == f["pe_idx"]
        assert up.tolist() == f["update_idx"]
        S.update_schedule(ab, pe, up, tr["window"], tr["warmup"])


def test_reference_n1_quirk_and_guard():
    tr = traces()[4]
    assert tr["n_rows"] == 1 and tr["reference_raises_index_error"] is True      # SURVEY A-1
    ab, pe, up = S.init_schedule(1, 4, 2)
    assert up.tolist() == [2] and (ab == 0).int().tolist() == [[1, 1, 1, 0]]


def test_docstring_sketch():
    """update_attn_bias docstring: init attn_bias [[0,0,0,inf],[0,0,inf,inf]], pe_idx [[0,1,2,3]]*2,
    update_idx [2,1] -- i.e. the state for W0=2, L=4 (the sketch lists update_idx per row as
    'slots already valid', the code initialises row 1 to W0+1; we follow the code)."""
    ab, pe, up = S.init_schedule(2, 4, 2)
    assert torch.isinf(ab).int().tolist() == [[0, 0, 0, 1], [0, 0, 1, 1]]
    assert pe.tolist() == [[0, 1, 2, 3], [0, 1, 2, 3]]
    assert up.tolist() == [2, 3]


def test_steady_state_invariants():
    ab, pe, up = S.init_schedule(2, 16, 8)
    for f in range(40):
        S.update_schedule(ab, pe, up, 16, 8)
    for f in range(16):
        assert not torch.isinf(ab).any()
        for n in range(2):
            assert sorted(pe[n].tolist()) == list(range(16))          # a permutation
            assert pe[n, :8].tolist() == list(range(8))                # sinks keep PE 0..7
            assert pe[n, up[n]] == 15                                  # write slot carries the largest PE
        S.update_schedule(ab, pe, up, 16, 8)


def test_scheduler_pointwise_matches_reference():
    g = load_golden("scheduler_pointwise.pt")
    x0 = S.scheduler_step_batch(g["eps"], g["x"], g["c_skip"], g["c_out"], g["a"], g["b"])
    torch.testing.assert_close(x0, g["x0"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(S.add_noise(g["x"][:1], g["eps"][:1], g["a"], g["b"], 1), g["add_noise_t1"])


def test_lcm_constants():
    ts = S.lcm_timesteps(50)
    assert ts[0] == 999 and ts[30] == 399 and ts[40] == 199 and ts[-1] == 19
    sub, c_skip, c_out, a, b = S.stream_constants([30, 40])
    assert sub.tolist() == [399, 199]
    torch.testing.assert_close(a * a + b * b, torch.ones(2))
    assert 0 < float(c_skip[0]) < 1e-6 and 0.999 < float(c_out[0]) <= 1.0


# ---- the product's host-side state machine (live2diff_b200.schedule) against the same reference traces ----
@pytest.mark.parametrize("i", range(4))
def test_product_ring_schedule_matches_reference_trace(i):
    from live2diff_b200.schedule import RingSchedule

    tr = traces()[i]
    rs = RingSchedule(tr["n_rows"], tr["window"], tr["warmup"])
    for f in tr["frames"]:
        assert [[1] * v + [0] * (tr["window"] - v) for v in rs.valid] == f["valid"]
        assert rs.pe_idx == f["pe_idx"]
        assert rs.update_idx == f["update_idx"]
        rs.advance()


def test_product_ring_schedule_n1():
    from live2diff_b200.schedule import RingSchedule

    rs = RingSchedule(1, 4, 2)
    ab, pe, up = S.init_schedule(1, 4, 2)
    for _ in range(12):
        assert rs.mask_rows() == ab.tolist() and rs.pe_idx == pe.tolist() and rs.update_idx == up.tolist()
        rs.advance()
        S.update_schedule(ab, pe, up, 4, 2)


def test_product_lcm_constants_match_oracle():
    from live2diff_b200.schedule import stream_constants

    for tl in ([30, 40], [25, 31, 37, 43], [0]):
        c = stream_constants(tl)
        sub, c_skip, c_out, a, b = S.stream_constants(tl)
        assert c.timesteps == sub.tolist()
        torch.testing.assert_close(torch.tensor(c.table(), dtype=torch.float32), torch.stack([a, b, c_skip, c_out]),
                                   rtol=1e-6, atol=1e-7)
