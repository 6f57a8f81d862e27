"""The golden fixtures were generated with the unmodified reference running against tests/shim/diffusers (a restatement of
the few diffusers 0.25.0 classes the hot path touches), so parity for the diffusers-owned arithmetic is pinned to the shim.
This test closes that loop wherever the real package exists: it runs the REAL diffusers classes in a clean subprocess (the
shim shadows the package name inside the test process) on seeded weights / inputs and compares with the shim and with the
schedule constants of the oracle.  Skipped when diffusers is not installed (it is not in the build image: no network)."""
import importlib.metadata
import json
import os
import subprocess
import sys
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_REAL = r"""
import json, sys, torch
from diffusers.models.attention_processor import Attention, AttnProcessor2_0
from diffusers.models.attention import FeedForward
from diffusers.models.embeddings import Timesteps, TimestepEmbedding
from diffusers import LCMScheduler
out_path = sys.argv[1]
torch.manual_seed(0)
res = {}
attn = Attention(query_dim=64, cross_attention_dim=48, heads=4, dim_head=16, bias=False, processor=AttnProcessor2_0())
ff = FeedForward(64, mult=4, activation_fn="geglu")
ts = Timesteps(320, True, 0)
te = TimestepEmbedding(320, 128)
x, ctx, t = torch.randn(2, 10, 64), torch.randn(2, 7, 48), torch.tensor([999.0, 379.0, 19.0])
res["state"] = {"attn": attn.state_dict(), "ff": ff.state_dict(), "te": te.state_dict()}
res["inputs"] = {"x": x, "ctx": ctx, "t": t}
sa = Attention(query_dim=64, heads=4, dim_head=16, processor=AttnProcessor2_0())
res["state"]["sa"] = sa.state_dict()
with torch.no_grad():
    res["sa_out"] = sa(x)
    res["attn_out"] = attn(x, encoder_hidden_states=ctx)
    res["ff_out"] = ff(x)
    res["ts_out"] = ts(t)
    res["te_out"] = te(ts(t))
sch = LCMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000)
sch.set_timesteps(50, original_inference_steps=50)
res["timesteps"] = sch.timesteps.tolist()
res["alphas_cumprod"] = sch.alphas_cumprod.double()
res["scalings"] = {int(t): [float(v) for v in sch.get_scalings_for_boundary_condition_discrete(int(t))] for t in (999, 499, 399, 199, 19)}
torch.save(res, out_path)
"""


def _real_diffusers_version():
    try:
        return importlib.metadata.version("diffusers")
    except importlib.metadata.PackageNotFoundError:
        return None


@pytest.mark.refcontainer
def test_shim_and_schedule_constants_match_real_diffusers():
    ver = _real_diffusers_version()
    if ver is None:
        pytest.skip("diffusers is not installed (the fixtures stay pinned to tests/shim/diffusers)")
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "real.pt")
        env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}      # no tests/shim on the path: the real package
        r = subprocess.run([sys.executable, "-c", _REAL, out], cwd=td, env=env, capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            pytest.skip(f"real diffusers {ver} could not run the probe: {r.stderr[-400:]}")
        real = torch.load(out, weights_only=False)
    sys.path.insert(0, os.path.join(ROOT, "tests", "shim"))
    try:
        from diffusers.models.attention import Attention, FeedForward          # the shim
        from diffusers.models.embeddings import Timesteps, TimestepEmbedding
    finally:
        sys.path.pop(0)
    x, ctx, t = real["inputs"]["x"], real["inputs"]["ctx"], real["inputs"]["t"]
    sa = Attention(query_dim=64, heads=4, dim_head=16)
    sa.load_state_dict(real["state"]["sa"], strict=True)
    ca = Attention(query_dim=64, cross_attention_dim=48, heads=4, dim_head=16)
    ca.load_state_dict(real["state"]["attn"], strict=True)
    ff = FeedForward(64, mult=4, activation_fn="geglu")
    ff.load_state_dict(real["state"]["ff"], strict=True)
    te = TimestepEmbedding(320, 128)
    te.load_state_dict(real["state"]["te"], strict=True)
    ts = Timesteps(320, True, 0)
    with torch.no_grad():
        torch.testing.assert_close(sa(x), real["sa_out"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ca(x, encoder_hidden_states=ctx), real["attn_out"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ff(x), real["ff_out"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ts(t), real["ts_out"], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(te(ts(t)), real["te_out"], rtol=1e-5, atol=1e-6)
    from oracle import schedule_oracle as S

    assert [int(v) for v in real["timesteps"]] == [int(v) for v in S.lcm_timesteps(50)]
    torch.testing.assert_close(S.alphas_cumprod().double(), real["alphas_cumprod"], rtol=1e-6, atol=1e-9)
    print(json.dumps({"diffusers": ver, "checked": ["Attention(self, cross)", "FeedForward/GEGLU", "Timesteps", "TimestepEmbedding",
                                                   "LCM timesteps", "alphas_cumprod"]}))
