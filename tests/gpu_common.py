"""Helpers for the `-m gpu` parity tests (CUDA path vs the CPU oracle / golden vectors)."""
import numpy as np
import torch

from conftest import load_golden
from idff_b200.weights import PackedWeights

_PW = {}


def packed(name, device):
    key = (name, str(device))
    if key not in _PW:
        _PW[key] = PackedWeights.from_npz(load_golden("weights_" + name), device=device)
    return _PW[key]


def drift_report(got, ref):
    """max / p99.9 / median absolute difference (the SELU kinks make isolated particles diverge: SURVEY.md section 7)."""
    d = np.abs(np.asarray(got, dtype=np.float64) - np.asarray(ref, dtype=np.float64)).reshape(-1)
    return dict(max=float(d.max()), p999=float(np.quantile(d, 0.999)), median=float(np.median(d)))


def assert_field_close(got, ref, tol=1e-5, what=""):
    """per-evaluation tolerance: |diff| <= tol * max|f| (BASELINE.json north_star: 1e-5 relative per step)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = max(1.0, float(np.abs(ref).max()))
    err = float(np.abs(got - ref).max())
    assert err <= tol * scale, f"{what}: max|diff| {err:.3e} > {tol:g} * {scale:.3g}"
    return err / scale


def assert_field_vs_refs(got, ref32, ref64, tol=1e-5, what=""):
    """Per-evaluation contract: within tol * max|f| of the reference's fp32 result and of its fp64 twin -- or, where
    the reference's own fp32 rounding exceeds that (the t == 1 branch divides by 1e-2 and amplifies the MLP's fp32
    round-off x100), about as close to the fp64 value as the reference itself (1.25 x its fp32<->fp64 gap), hence within
    gap + 1.25 gap of the fp32 reference (triangle inequality)."""
    got, ref32, ref64 = (np.asarray(a, dtype=np.float64) for a in (got, ref32, ref64))
    scale = max(1.0, float(np.abs(ref32).max()))
    gap = float(np.abs(ref32 - ref64).max())
    e32 = float(np.abs(got - ref32).max())
    e64 = float(np.abs(got - ref64).max())
    assert e32 <= max(tol * scale, 2.25 * gap), f"{what}: |got-ref32| {e32:.3e} (tol {tol * scale:.3e}, ref gap {gap:.3e})"
    assert e64 <= max(tol * scale, 1.25 * gap), f"{what}: |got-ref64| {e64:.3e} (tol {tol * scale:.3e}, ref gap {gap:.3e})"
    return e32 / scale
