"""Minibatch OT coupling (torchcfm/optimal_transport.py:11-142).  OUT OF the hot path: the reference solves the
plan with POT's CPU network simplex at batch size 256; only its output -- the index pairs (i, j) -- feeds the
path-sample kernel (as a fused gather).  POT is not installed here, so the exact plan for the uniform, equal-size
marginals the reference uses is obtained with scipy's assignment solver on the host."""
import numpy as np
import torch


class OTPlanSampler:
    def __init__(self, method: str = "exact"):
        if method != "exact":
            raise ValueError(f"only the exact plan is provided (got {method!r})")
        self.method = method

    def get_map(self, x0, x1):
        """Uniform-marginal exact plan, as a (B0, B1) numpy array (reference :59-94)."""
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

    def sample_map_indices(self, x0, x1):
        i, j = self.sample_map(self.get_map(x0, x1), x0.shape[0])
        dev = x0.device
        return (torch.as_tensor(i, dtype=torch.int64, device=dev), torch.as_tensor(j, dtype=torch.int64, device=dev))

    def sample_plan(self, x0, x1, replace=True):
        i, j = self.sample_map_indices(x0, x1)
        return x0[i], x1[j]
