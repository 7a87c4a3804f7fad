"""RBF-kernel MMD (biased V-statistic) on the B200 pair kernel.

    compute_mmd_rbf(x, y, sigma)   2D_examples/Autograd_example_order2.py:260-278
    compute_mmd(x, y, sigma)       timeseris_example/MD_simulation.py:50-86
"""
import torch

from . import _lib


def mmd_partial_sums(x, y, sigma=1.0, x_rows=None, y_rows=None):
    """(sum Kxx, sum Kyy, sum Kxy) over x rows [x_rows) x all x, y rows [y_rows) x all y, x rows x all y: fp64 (3,)."""
    _lib.require_cuda(x, "x")
    _lib.require_cuda(y, "y")
    x = x.detach().contiguous()
    y = y.detach().contiguous()
    if x.dim() != 2 or y.dim() != 2 or x.shape[1] != y.shape[1]:
        raise _lib.IdffError(f"x {tuple(x.shape)} and y {tuple(y.shape)} must be (N, D) and (M, D)")
    N, M, D = x.shape[0], y.shape[0], x.shape[1]
    xb, xe = (0, N) if x_rows is None else x_rows
    yb, ye = (0, M) if y_rows is None else y_rows
    sums = torch.zeros(3, device=x.device, dtype=torch.float64)
    _lib.check(_lib.lib().idff_mmd_rbf(_lib.ptr(x), N, _lib.ptr(y), M, D, float(sigma), _lib.ptr(sums), xb, xe, yb, ye,
                                       _lib.stream_of(x)), "idff_mmd_rbf")
    return sums


def mmd_from_sums(sums, N, M):
    return sums[0] / (float(N) * N) + sums[1] / (float(M) * M) - 2.0 * sums[2] / (float(N) * M)


def mmd_sweep(x, ys, sigma=1.0):
    """MMD(x, ys[g]) for every generated set of a sweep: x (N, D), ys (G, M, D) -> fp64 tensor (G,) on the device.
    One launch for all sets when D <= 4 (idff_mmd_rbf_sweep; Kxx is evaluated once), else one launch per set."""
    _lib.require_cuda(x, "x")
    _lib.require_cuda(ys, "ys")
    x = x.detach().contiguous()
    ys = ys.detach().contiguous()
    if x.dim() != 2 or ys.dim() != 3 or x.shape[1] != ys.shape[2]:
        raise _lib.IdffError(f"x {tuple(x.shape)} must be (N, D) and ys {tuple(ys.shape)} (G, M, D)")
    N, D = x.shape
    G, M = ys.shape[0], ys.shape[1]
    if D > 4:
        return torch.stack([mmd_from_sums(mmd_partial_sums(x, ys[g], sigma), N, M) for g in range(G)])
    sums = torch.zeros(1 + 2 * G, device=x.device, dtype=torch.float64)
    _lib.check(_lib.lib().idff_mmd_rbf_sweep(_lib.ptr(x), N, _lib.ptr(ys), M, G, D, float(sigma), _lib.ptr(sums),
                                             _lib.stream_of(x)), "idff_mmd_rbf_sweep")
    return sums[0] / (float(N) * N) + sums[1::2] / (float(M) * M) - 2.0 * sums[2::2] / (float(N) * M)


def compute_mmd_rbf(x, y, sigma=1.0):
    """mean(Kxx) + mean(Kyy) - 2 mean(Kxy), K = exp(-|a-b|^2/(2 sigma^2)); returns a Python float."""
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    if not torch.is_tensor(y):
        y = torch.as_tensor(y)
    x = x.to(dtype=torch.float32)
    y = y.to(dtype=torch.float32)
    if x.is_cuda and not y.is_cuda:
        y = y.to(x.device)
    if y.is_cuda and not x.is_cuda:
        x = x.to(y.device)
    s = mmd_partial_sums(x, y, sigma)
    return float(mmd_from_sums(s, x.shape[0], y.shape[0]).item())


compute_mmd = compute_mmd_rbf
