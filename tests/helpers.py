"""Shared helpers for the parity tests (oracle side + fixture loading)."""
import os

import torch

from live2diff_b200.weights import UNetDims, random_tensors, spec_fingerprint, unet_param_spec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def sub_spec(d: UNetDims, prefix: str):
    return {k[len(prefix) + 1:]: v for k, v in unet_param_spec(d).items() if k.startswith(prefix + ".")}


def prefixed(sd, prefix):
    return {f"{prefix}.{k}": v for k, v in sd.items()}


def regen_weights(spec, seed, fingerprint):
    w = random_tensors(spec, seed=seed)
    fp = spec_fingerprint(w)
    assert abs(fp - fingerprint) <= 1e-9 * abs(fingerprint), (
        f"seeded weights differ from the ones the golden fixture was generated with ({fp} vs {fingerprint}): "
        "torch RNG stream changed -- regenerate fixtures with tests/golden/make_golden.py")
    return w


def dims_from(dct):
    return UNetDims(**{k: (tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in dct.items()})


def mask_from_valid(valid, dtype=torch.float32):
    valid = torch.as_tensor(valid, dtype=torch.bool)
    return torch.zeros(valid.shape, dtype=dtype).masked_fill_(~valid, float("-inf"))
