"""The bench contract, checked on the host: (1) the lines kept under profiles/r2/ (what the docs quote) carry every key the
contract names and are internally consistent (value = steps / time, roofline.frac = achieved / peak, achieved = algorithmic
bytes / kernel time, ...); (2) bench.py's own helpers that need no GPU (flop count of the UNet's GEMMs, workload
description) agree with those lines.  A number in README.md that no longer matches its JSON line fails here."""
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R2 = os.path.join(ROOT, "profiles", "r2")

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"]


def _load(name):
    path = os.path.join(R2, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not present")
    txt = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return json.loads(txt[-1])


@pytest.mark.parametrize("name,world", [("bench_final_cfg2.json", 1), ("bench_final_cfg3.json", 1), ("bench_final_cfg4.json", 1),
                                        ("bench_final_2gpu.json", 2)])
def test_kept_bench_lines_follow_the_contract(name, world):
    d = _load(name)
    for k in REQUIRED:
        assert k in d, f"{name}: missing {k}"
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f16"
    assert d["n_gpus"] == world and "workload" in d["config"] and d["vs_baseline"] is None
    # value = all ranks' frames / max-over-ranks time
    assert d["value"] == pytest.approx(world * 1e3 / d["ms_per_step"], rel=1e-6)
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0 < e["d2h_bytes_per_step"]
    assert e["value"] <= d["value"] * 1.02, "end to end cannot beat the device-resident number"
    assert d["gpu_launches"] >= d["steps"] * 500, "the line must claim its own kernels"
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-6)
    assert r["achieved"] == pytest.approx(r["algorithmic_bytes_per_step"] / (r["ms_per_step_in_kernel"] * 1e-3) / 1e9, rel=1e-3)
    if r.get("traffic"):
        assert 0.9 <= r["traffic"] / r["algorithmic_bytes_per_launch"] <= 1.05, "DRAM traffic must stay at the algorithmic bytes"
    if world == 1 and name.endswith("cfg2.json"):
        cb = d["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
        assert d["vs_torch_fp16_eager"] == pytest.approx(e["value"] / d["torch_fp16_eager"]["value"], rel=1e-6)
        assert d["vs_torch_fp16_eager"] >= 2.0, "BASELINE.md's target for this path"


def test_readme_headline_matches_the_kept_line():
    d = _load("bench_final_cfg2.json")
    readme = open(os.path.join(ROOT, "README.md")).read()
    m = re.search(r"\*\*([0-9.]+) / ([0-9.]+)\*\* \(round 1", readme)
    assert m, "headline row not found"
    assert float(m.group(1)) == pytest.approx(d["value"], abs=0.06) and float(m.group(2)) == pytest.approx(d["e2e"]["value"], abs=0.06)
    assert f"{d['launches_per_step']} kernels per frame" in readme


def test_gemm_flop_count_matches_the_line():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(bench)
    except Exception as exc:  # bench.py imports the product at call time only; a hard import error is a real failure
        pytest.fail(f"bench.py does not import on the host: {exc}")
    d = _load("bench_final_cfg2.json")
    from live2diff_b200.weights import UNetDims

    flops = bench.unet_gemm_flops(UNetDims(), 2, 64, 64)      # config 2: two stream-batch rows of a 64 x 64 latent, SD1.5 widths
    assert flops == pytest.approx(d["roofline_tensor"]["flops_per_step"], rel=1e-9)
