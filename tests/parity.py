"""Parity criteria used by the GPU tests.

STRICT  -- a single kernel fed fp16 inputs, compared with an fp32 evaluation of the same op on the same
           (fp16-rounded) inputs: every element within rtol=1e-3 / atol=1e-4 (BASELINE.json's
           tolerance).  One fp16 output rounding is 4.9e-4 relative, so this holds exactly when the
           kernel accumulates in fp32 and rounds once.

REFEREE -- chains of ops (attention with fp16-rounded P / K+pe, whole modules, the whole UNet):
           two correct fp16 implementations differ from each other by more than 1e-3 because they round
           intermediates at different points (fp16 eps = 9.8e-4; SURVEY.md §7 "hard parts" (i)).  The fp32
           oracle referees: our error against it must not exceed `slack` x the error of the reference's
           own fp16 evaluation order (the oracle restatement run in fp16 with torch ops) against it, with
           a floor of one fp16 rounding.  The strict pass-rate is still printed for information.
"""
import torch

RTOL, ATOL = 1e-3, 1e-4


def strict(ours: torch.Tensor, ref32: torch.Tensor, name="", rtol=RTOL, atol=ATOL):
    ours = ours.float().cpu()
    ref32 = ref32.float().cpu()
    assert ours.shape == ref32.shape, (name, ours.shape, ref32.shape)
    assert torch.isfinite(ours).all(), f"{name}: non-finite values in the CUDA result"
    err = (ours - ref32).abs()
    tol = atol + rtol * ref32.abs()
    bad = err > tol
    if bad.any():
        i = int(torch.argmax(err - tol))
        raise AssertionError(f"{name}: {int(bad.sum())}/{bad.numel()} elements outside rtol={rtol} atol={atol}; "
                             f"worst: ours={ours.flatten()[i]:.6f} ref={ref32.flatten()[i]:.6f}")


def rel_rms(a: torch.Tensor, ref: torch.Tensor) -> float:
    a, ref = a.double().cpu(), ref.double().cpu()
    return float(((a - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt().clamp_min(1e-30))


def referee(ours: torch.Tensor, ref32: torch.Tensor, ref16: torch.Tensor, name="", slack=2.0, floor=6e-4):
    ours, ref32, ref16 = ours.float().cpu(), ref32.float().cpu(), ref16.float().cpu()
    assert ours.shape == ref32.shape, (name, ours.shape, ref32.shape)
    assert torch.isfinite(ours).all(), f"{name}: non-finite values in the CUDA result"
    e_ours, e_ref = rel_rms(ours, ref32), rel_rms(ref16, ref32)
    m_ours = float((ours - ref32).abs().max())
    m_ref = float((ref16 - ref32).abs().max())
    scale = float(ref32.abs().max())
    within = float(((ours - ref16).abs() <= ATOL + RTOL * ref16.abs()).float().mean())
    print(f"[parity] {name}: rel-rms ours={e_ours:.2e} torch-fp16={e_ref:.2e} | max-abs ours={m_ours:.2e} "
          f"torch-fp16={m_ref:.2e} (scale {scale:.2e}) | within rtol1e-3/atol1e-4 of torch-fp16: {within:.4f}")
    assert e_ours <= max(slack * e_ref, floor), f"{name}: rel-rms error {e_ours:.3e} vs fp16 reference {e_ref:.3e}"
    assert m_ours <= max(slack * 2 * m_ref, 4 * floor * scale), f"{name}: max-abs error {m_ours:.3e} vs {m_ref:.3e}"
