"""Host-side toy-data generators used as fixture inputs by the 2-D IDFF scripts (torchcfm/utils.py:12-33,74-95).
They draw from the same global RNGs in the same order as the reference so seeded runs agree; plotting helpers
and the torchdyn wrapper of the reference file are presentation-only and are not provided."""
import math

import numpy as np
import torch


def eight_normal_sample(n, dim, scale=1, var=1):
    """8 modes on a circle; NB the reference passes sqrt(var)*I as the COVARIANCE, so the std is var**0.25."""
    dist = torch.distributions.multivariate_normal.MultivariateNormal(torch.zeros(dim), math.sqrt(var) * torch.eye(dim))
    r = 1.0 / np.sqrt(2)
    centers = torch.tensor([(1, 0), (-1, 0), (0, 1), (0, -1), (r, r), (r, -r), (-r, r), (-r, -r)]) * scale
    noise = dist.sample((n,))
    which = torch.multinomial(torch.ones(8), n, replacement=True)
    return centers[which] + noise


def checkerboard(n):
    u = np.random.rand(n) * 4 - 2
    v = np.random.rand(n) - np.random.randint(0, 2, n) * 2
    v = v + (np.floor(u) % 2)
    return torch.tensor(np.stack([u, v], 1) * 2)


def sample_8gaussians(n):
    return eight_normal_sample(n, 2, scale=5, var=0.1).float()


def sample_checkerboard(n):
    return checkerboard(n).float()
