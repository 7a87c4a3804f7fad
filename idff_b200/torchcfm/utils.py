"""Host-side toy-data generators used as fixture inputs by the 2-D IDFF scripts (torchcfm/utils.py:12-95) and the
`torch_wrapper` adapter (:97-105).  They draw from the same RNGs in the same order as the reference so seeded runs agree
(`pinwheel` uses the module-level `rng`, an unseeded numpy RandomState, like the reference's).  Plotting helpers and
`sample_moons` (torchdyn's generator) are not provided."""
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


rng = np.random.RandomState()


def pinwheel(n):
    """5-armed pinwheel (reference :36-53)."""
    classes, per = 5, n // 5
    feats = rng.randn(classes * per, 2) * np.array([0.3, 0.1])
    feats[:, 0] += 1.0
    labels = np.repeat(np.arange(classes), per)
    ang = np.linspace(0, 2 * np.pi, classes, endpoint=False)[labels] + 0.25 * np.exp(feats[:, 0])
    rot = np.reshape(np.stack([np.cos(ang), -np.sin(ang), np.sin(ang), np.cos(ang)]).T, (-1, 2, 2))
    return torch.tensor(2 * rng.permutation(np.einsum("ti,tij->tj", feats, rot)))


def spirals_2d(n):
    """two mirrored spirals (reference :54-72); global numpy RNG."""
    half = int(n) // 2
    theta = np.sqrt(np.random.rand(half, 1)) * 540 * (2 * np.pi) / 360
    ax = -np.cos(theta) * theta + np.random.rand(half, 1) * 0.5
    ay = np.sin(theta) * theta + np.random.rand(half, 1) * 0.5
    pts = np.vstack((np.hstack((ax, ay)), np.hstack((-ax, -ay)))) / 3
    pts += np.random.randn(*pts.shape) * 0.1
    return torch.tensor(pts, dtype=torch.float32)


def sample_pinwheel(n):
    return pinwheel(n).float()


def sample_spirals(n):
    return spirals_2d(n).float()


class torch_wrapper(torch.nn.Module):
    """model(cat[x, t]) behind a forward(t, x) signature (reference :97-105)."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, t, x, *args, **kwargs):
        return self.model(torch.cat([x, t.repeat(x.shape[0])[:, None]], 1))


def sample_8gaussians(n):
    return eight_normal_sample(n, 2, scale=5, var=0.1).float()


def sample_checkerboard(n):
    return checkerboard(n).float()
