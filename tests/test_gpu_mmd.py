"""-m gpu: RBF-MMD kernel vs the fp64 direct evaluation (tolerance 1e-6 abs + 1e-4 rel) and vs
the reference's own fp32 result."""
import numpy as np
import pytest
import torch

import idff_oracle as O
from conftest import load_golden
from idff_b200 import mmd as M

pytestmark = pytest.mark.gpu


def _close(got, ref):
    assert abs(got - ref) <= 1e-6 + 1e-4 * abs(ref), (got, ref)


def test_mmd_2d_vs_golden(cuda_device):
    g = load_golden("mmd")
    cb, e8, gen = (torch.tensor(g[n]).to(cuda_device) for n in ("true_checkerboard", "true_8gaussians", "gen"))
    _close(M.compute_mmd_rbf(cb, gen, 1.0), float(g["mmd_cb_gen_f64"]))
    _close(M.compute_mmd_rbf(cb, e8, 1.0), float(g["mmd_cb_8g_f64"]))
    # and within the reference's own fp32 noise of its value
    assert abs(M.compute_mmd_rbf(cb, gen, 1.0) - float(g["mmd_cb_gen"])) < 5e-6
    assert M.compute_mmd_rbf(cb, cb, 1.0) == pytest.approx(0.0, abs=1e-9)


def test_mmd_polyala_shape_vs_golden(cuda_device):
    g = load_golden("mmd")
    a, b = torch.tensor(g["md_a"]).to(cuda_device), torch.tensor(g["md_b"]).to(cuda_device)
    _close(M.compute_mmd(a, b, 1.0), float(g["mmd_md_f64"]))            # ~ 1/200 + 1/1000: diagonal only
    _close(M.compute_mmd(a / 60, b / 60 + 0.1, 3.0), float(g["mmd_md_unit_f64"]))


@pytest.mark.parametrize("N,Mm,D", [(1, 1, 2), (5, 3, 1), (300, 515, 3), (2049, 700, 4), (130, 70, 5), (257, 1100, 46)])
def test_mmd_ragged_shapes_and_row_split(cuda_device, N, Mm, D):
    g = torch.Generator().manual_seed(N + Mm + D)
    x, y = torch.randn(N, D, generator=g), torch.randn(Mm, D, generator=g) * 1.3 + 0.2
    sig = 1.7
    ref = O.mmd_partial_sums_fp64(x, y, sig)
    xs, ys = x.to(cuda_device), y.to(cuda_device)
    s = M.mmd_partial_sums(xs, ys, sig).cpu().numpy()
    np.testing.assert_allclose(s, np.array(ref), rtol=3e-6)
    _close(M.compute_mmd_rbf(xs, ys, sig), O.mmd_rbf_fp64_direct(x, y, sig))
    # row-block split (what each rank of the sharded launcher evaluates) adds up to the whole
    cut_x, cut_y = N // 3, (2 * Mm) // 3
    s1 = M.mmd_partial_sums(xs, ys, sig, (0, cut_x), (0, cut_y)).cpu().numpy()
    s2 = M.mmd_partial_sums(xs, ys, sig, (cut_x, N), (cut_y, Mm)).cpu().numpy()
    # (the full-set call uses the symmetric block shortcut, the row blocks do not: fp32 run grouping differs)
    np.testing.assert_allclose(s1 + s2, s, rtol=1e-6)


def test_mmd_large_sets(cuda_device):
    """N = M = 16384 (BASELINE.md timing shape): against the chunked fp64 oracle."""
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(16384, 2, generator=g) * 2, torch.randn(16384, 2, generator=g) * 2 + 0.05
    sx, sy, sxy = O.mmd_partial_sums_fp64(x, y, 1.0)
    ref = sx / 16384 ** 2 + sy / 16384 ** 2 - 2 * sxy / 16384 ** 2
    _close(M.compute_mmd_rbf(x.to(cuda_device), y.to(cuda_device), 1.0), ref)


def test_mmd_sweep_equals_per_set_evaluation(cuda_device):
    """idff_mmd_rbf_sweep: one launch for G generated sets against one fixed set == G single evaluations."""
    from idff_b200 import _lib
    from idff_b200.mmd import mmd_from_sums, mmd_partial_sums, mmd_sweep
    g = torch.Generator(device=cuda_device).manual_seed(11)
    for (N, M, G, D) in ((2048, 2048, 5, 2), (1000, 333, 3, 2), (64, 700, 2, 4)):
        x = torch.randn(N, D, device=cuda_device, generator=g)
        ys = torch.randn(G, M, D, device=cuda_device, generator=g) * 1.3 + 0.2
        before = _lib.launch_count()
        got = mmd_sweep(x, ys, 1.0)
        assert _lib.launch_count() == before + 1 and got.shape == (G,) and got.dtype == torch.float64
        ref = torch.stack([mmd_from_sums(mmd_partial_sums(x, ys[k], 1.0), N, M) for k in range(G)])
        np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=1e-10, atol=1e-14)
    # D > 4 falls back to one launch per set
    x = torch.randn(100, 46, device=cuda_device, generator=g) * 30
    ys = torch.randn(3, 50, 46, device=cuda_device, generator=g) * 30
    got = mmd_sweep(x, ys, 1.0)
    ref = torch.stack([mmd_from_sums(mmd_partial_sums(x, ys[k], 1.0), 100, 50) for k in range(3)])
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=1e-10, atol=1e-14)
