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
