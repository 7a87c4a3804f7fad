"""Minibatch OT coupling (torchcfm/optimal_transport.py:11-142).

The reference computes the exact plan between two minibatches of equal size with uniform marginals on the host (POT's
network simplex, `pot.emd`, batch 256) and samples index pairs from it with numpy.  For those marginals the exact plan is
1/n times a permutation matrix, so the coupling is an assignment problem.  On CUDA tensors it is solved where the data lives
(`idff_ot_assign`: fp64 assignment in one launch -- augmenting paths for small batches, an epsilon-scaling auction above 64) and the index pairs are drawn on the device:
no host round trip, graph-capturable, and they feed `idff_path_sample` as fused gather indices.  `get_map` / `sample_map`
keep the reference's host-side signatures (numpy plan in, numpy indices out; POT is not installed here, so the plan comes
from scipy's assignment solver)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def ot_assign(x0: torch.Tensor, x1: torch.Tensor, return_cost: bool = False):
    """perm (n,) int64 on the device with plan[i][perm[i]] = 1/n for the exact uniform-marginal plan of (x0, x1) under the
    squared-Euclidean cost (optimal_transport.py:59-94 for equal batch sizes); optionally the plan's total cost (fp64)."""
    _lib.require_cuda(x0, "x0")
    _lib.require_cuda(x1, "x1")
    if x0.shape[0] != x1.shape[0]:
        raise _lib.IdffError("ot_assign needs equal batch sizes (the reference's uniform-marginal minibatch plan)")
    if x0.shape[0] == 0:
        return (torch.empty(0, dtype=torch.int64, device=x0.device), torch.zeros((), dtype=torch.float64, device=x0.device)) \
            if return_cost else torch.empty(0, dtype=torch.int64, device=x0.device)
    a = x0.detach().reshape(x0.shape[0], -1).contiguous()
    b = x1.detach().reshape(x1.shape[0], -1).contiguous()
    if a.shape[1] != b.shape[1]:
        raise _lib.IdffError(f"x0 and x1 differ in dimension: {a.shape[1]} vs {b.shape[1]}")
    n, D = a.shape
    perm = torch.empty(n, dtype=torch.int64, device=a.device)
    cost = torch.zeros((), dtype=torch.float64, device=a.device) if return_cost else None
    _lib.check(_lib.lib().idff_ot_assign(_lib.ptr(a), _lib.ptr(b), n, D, _lib.ptr(perm), _lib.ptr(cost), _lib.stream_of(a)),
               "idff_ot_assign")
    return (perm, cost) if return_cost else perm


def ot_assign_batch(x0: torch.Tensor, x1: torch.Tensor, return_cost: bool = False):
    """The couplings of P independent minibatch pairs in ONE launch (`idff_ot_assign_batch`, one CTA per problem): x0, x1
    (P, n, ...) -> perm (P, n) int64 [, cost (P,) fp64]; row p is exactly ot_assign(x0[p], x1[p]).  The minibatches of the next
    P training steps do not depend on the parameters, so they are coupled together: at n = 256 a launch of 1024 problems costs
    what a handful of single ones do."""
    _lib.require_cuda(x0, "x0")
    _lib.require_cuda(x1, "x1")
    if x0.dim() < 3 or x0.shape[:2] != x1.shape[:2]:
        raise _lib.IdffError(f"ot_assign_batch needs x0, x1 of shape (P, n, ...) with equal P and n, got {tuple(x0.shape)} and {tuple(x1.shape)}")
    P, n = int(x0.shape[0]), int(x0.shape[1])
    a = x0.detach().reshape(P, n, -1).contiguous()
    b = x1.detach().reshape(P, n, -1).contiguous()
    if a.shape[2] != b.shape[2]:
        raise _lib.IdffError(f"x0 and x1 differ in dimension: {a.shape[2]} vs {b.shape[2]}")
    perm = torch.empty((P, n), dtype=torch.int64, device=a.device)
    cost = torch.zeros(P, dtype=torch.float64, device=a.device) if return_cost else None
    _lib.check(_lib.lib().idff_ot_assign_batch(_lib.ptr(a), _lib.ptr(b), P, n, int(a.shape[2]), _lib.ptr(perm), _lib.ptr(cost),
                                               _lib.stream_of(a)), "idff_ot_assign_batch")
    return (perm, cost) if return_cost else perm


class OTPlanSampler:
    def __init__(self, method: str = "exact"):
        if method != "exact":
            raise ValueError(f"only the exact plan is provided (got {method!r})")
        self.method = method

    def get_map(self, x0, x1):
        """Uniform-marginal exact plan, as a (B0, B1) numpy array (reference :59-94) -- host side, for API compatibility."""
        from scipy.optimize import linear_sum_assignment
        a = x0.reshape(x0.shape[0], -1).double()
        b = x1.reshape(x1.shape[0], -1).double()
        M = (torch.cdist(a, b) ** 2).cpu().numpy()
        if M.shape[0] != M.shape[1]:
            raise ValueError("the assignment-based exact plan needs equal batch sizes")
        r, c = linear_sum_assignment(M)
        P = np.zeros_like(M)
        P[r, c] = 1.0 / M.shape[0]
        return P

    def sample_map(self, pi, batch_size, replace=True):
        """Draw index pairs from the plan with numpy's global RNG (reference :96-118)."""
        p = pi.flatten()
        p = p / p.sum()
        choices = np.random.choice(pi.shape[0] * pi.shape[1], p=p, size=batch_size, replace=replace)
        return np.divmod(choices, pi.shape[1])

    def sample_map_indices(self, x0, x1, generator=None, replace=True):
        """(i, j) int64 index tensors on x0's device, distributed as the reference's sample_map(get_map(x0, x1), B, replace):
        the plan's support is the n pairs (i, perm[i]) with equal mass, so i is uniform with replacement (replace=True) or a
        random permutation of all rows (replace=False), and j = perm[i].  CUDA inputs: assignment and draw on the device
        (torch's generator, not numpy's)."""
        if x0.is_cuda:
            perm = ot_assign(x0, x1)
            n = x0.shape[0]
            i = (torch.randint(0, n, (n,), device=x0.device, generator=generator) if replace
                 else torch.randperm(n, device=x0.device, generator=generator))
            return i, perm[i]
        i, j = self.sample_map(self.get_map(x0, x1), x0.shape[0], replace=replace)
        return torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)

    def sample_plan(self, x0, x1, replace=True):
        i, j = self.sample_map_indices(x0, x1, replace=replace)
        return x0[i], x1[j]

    def sample_plan_batch(self, x0, x1, generator=None):
        """sample_plan for P independent minibatch pairs at once (CUDA tensors (P, n, ...)): one batched assignment launch, index
        pairs drawn with replacement per problem as sample_map does -> (x0[p][i_p], x1[p][perm_p[i_p]]) stacked over p.  What a
        many-steps-per-launch trainer calls once for the draws of all its steps (train.FusedFlowTrainer(ot_pairing=True))."""
        perm = ot_assign_batch(x0, x1)
        P, n = perm.shape
        i = torch.randint(0, n, (P, n), device=x0.device, generator=generator)
        j = torch.gather(perm, 1, i)
        tail = (1,) * (x0.dim() - 2)
        gi = i.view(P, n, *tail).expand(P, n, *x0.shape[2:])
        gj = j.view(P, n, *tail).expand(P, n, *x1.shape[2:])
        return torch.gather(x0, 1, gi), torch.gather(x1, 1, gj)

    def sample_plan_with_labels(self, x0, x1, y0=None, y1=None, replace=True):
        """(x0[i], x1[j], y0[i], y1[j]) for index pairs drawn from the plan (reference :144-179)."""
        i, j = self.sample_map_indices(x0, x1, replace=replace)
        return x0[i], x1[j], (y0[i] if y0 is not None else None), (y1[j] if y1 is not None else None)


def wasserstein(x0, x1, method=None, reg: float = 0.05, power: int = 2, **kwargs):
    """Exact Wasserstein-`power` distance between two equally weighted batches of equal size (reference :214-263, pot.emd2).
    power = 2 on CUDA tensors: one launch of the assignment kernel, sqrt(cost / n).  power = 1 (plain Euclidean cost) and CPU
    tensors: scipy's exact assignment on the host.  Only the exact method is provided."""
    del reg, kwargs
    assert power == 1 or power == 2
    if method not in (None, "exact"):
        raise ValueError(f"only the exact method is provided (got {method!r})")
    if x0.shape[0] != x1.shape[0]:
        raise ValueError("equal batch sizes are required (uniform marginals reduce to an assignment)")
    n = x0.shape[0]
    if power == 2 and x0.is_cuda:
        _, cost = ot_assign(x0, x1, return_cost=True)
        return float(torch.sqrt(cost / n))
    from scipy.optimize import linear_sum_assignment
    M = torch.cdist(x0.reshape(n, -1).double(), x1.reshape(n, -1).double()).cpu().numpy()
    if power == 2:
        M = M ** 2
    r, c = linear_sum_assignment(M)
    ret = float(M[r, c].sum() / n)
    return ret ** 0.5 if power == 2 else ret
