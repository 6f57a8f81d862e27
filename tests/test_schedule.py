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


# Known-answer values of the published DDPM "linear" beta schedule the reference configures (configs/base_config.yaml:30-36:
# beta_start 0.00085, beta_end 0.012, 1000 steps) and of LCMScheduler.get_scalings_for_boundary_condition_discrete
# (sigma_data 0.5, timestep_scaling 10), computed independently in float64 (numpy cumprod / closed form), so that neither
# the oracle nor the product is checked against the other only:  t -> (abar, sqrt(abar), sqrt(1-abar), c_skip, c_out)
LCM_KAT = {
    999: (0.0015789629305514416, 0.039736166530648646, 0.9992102066479548, 2.5050075037374525e-09, 0.9999999987474963),
    699: (0.03560389684468509, 0.18868994897631694, 0.9820367117146461, 5.116649346211688e-09, 0.9999999974416753),
    499: (0.1618121459134018, 0.4022588046437291, 0.9155259985858393, 1.004012021999791e-08, 0.99999999497994),
    399: (0.2914485025459839, 0.5398597063552566, 0.8417550103527843, 1.5703418701776077e-08, 0.9999999921482906),
    379: (0.3234614268200419, 0.5687366937520754, 0.8225196491148148, 1.7404501197351364e-08, 0.9999999912977493),
    259: (0.5501447094244379, 0.7417174053670562, 0.6707125245405531, 3.726837564778493e-08, 0.999999981365812),
    199: (0.6753436096512047, 0.8217929238264373, 0.5697862672518489, 6.312971496113009e-08, 0.9999999684351419),
    139: (0.796285419325596, 0.8923482612330211, 0.4513475165262394, 1.2939287182432691e-07, 0.999999935303562),
    19: (0.9810520056735957, 0.9904806942457767, 0.1376517138520413, 6.92515979806234e-06, 0.9999965374141062),
}
# t_index -> timestep of LCMScheduler.set_timesteps(50) with original_inference_steps=50: 999 - 20 * t_index
LCM_T_INDEX = {0: 999, 25: 499, 30: 399, 31: 379, 37: 259, 40: 199, 43: 139, 49: 19}


def test_lcm_constants_known_answers_oracle_and_product():
    from live2diff_b200.schedule import stream_constants

    ts = S.lcm_timesteps(50)
    for ti, t in LCM_T_INDEX.items():
        assert int(ts[ti]) == t == 999 - 20 * ti
    ac = S.alphas_cumprod()
    for t, (abar, sa, sb, cs, co) in LCM_KAT.items():
        assert abs(float(ac[t]) - abar) <= 2e-5 * abar                       # float32 cumprod of 1000 factors
        got_cs, got_co = S.boundary_scalings(t)
        assert abs(got_cs - cs) <= 1e-12 * max(cs, 1e-9) + 1e-18 and abs(got_co - co) <= 1e-12
    for tl in ([30, 40], [25, 31, 37, 43], [0, 49]):
        want = [LCM_KAT[LCM_T_INDEX[i]] for i in tl]
        sub, c_skip, c_out, a, b = S.stream_constants(tl)
        c = stream_constants(tl)
        assert sub.tolist() == c.timesteps == [LCM_T_INDEX[i] for i in tl]
        for r, (abar, sa, sb, cs, co) in enumerate(want):
            for name, got_o, got_p, ref in (("sqrt_abar", a[r], c.sqrt_abar[r], sa), ("sqrt_1m_abar", b[r], c.sqrt_1m_abar[r], sb),
                                            ("c_skip", c_skip[r], c.c_skip[r], cs), ("c_out", c_out[r], c.c_out[r], co)):
                assert abs(float(got_o) - ref) <= 2e-5 * abs(ref) + 1e-12, (name, tl, r, float(got_o), ref)
                assert abs(float(got_p) - ref) <= 2e-5 * abs(ref) + 1e-12, (name, tl, r, float(got_p), ref)


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
