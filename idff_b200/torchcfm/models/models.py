"""The two networks of the hot path with the reference's constructor signatures, attribute names and
state_dict keys (torchcfm/models/models.py:4-21 MLP, :23-42 MLP_Embedding), so the shipped checkpoints load.

Under autograd (training) the layers run as ordinary torch modules (or, for the 2-D MLP, inside the fused training
kernel, whose state the parameters alias).  Under torch.no_grad() on a CUDA tensor the forward is one launch of the
B200 MLP kernel (idff_mlp_forward); MLP_Embedding takes that route when every row shares one time index, which is how
the generator calls it (MD_simulation.py:501)."""
import torch
from torch import nn

from .. import _backend as B


def _selu_stack(n_in: int, width: int, n_out: int) -> nn.Sequential:
    """Linear-SELU x3 -> Linear; module indices 0, 2, 4, 6 hold the Linear layers (the checkpoints' `net.N.weight` keys)."""
    sizes = [n_in, width, width, width]
    layers = []
    for a, b in zip(sizes[:-1], sizes[1:]):
        layers += [nn.Linear(a, b), nn.SELU()]
    layers.append(nn.Linear(width, n_out))
    return nn.Sequential(*layers)


class MLP(nn.Module):
    """x = cat[state, t] (time_varying) -> net(x)."""

    def __init__(self, dim, out_dim=None, w=64, time_varying=False):
        super().__init__()
        self.time_varying = time_varying
        self.net = _selu_stack(dim + int(bool(time_varying)), w, dim if out_dim is None else out_dim)

    def forward(self, x):
        return B.mlp_forward(self, x, 0) if B.use_kernel(self, x) else self.net(x)


class MLP_Embedding(nn.Module):
    """x = cat[state, t], ti = frame index -> net(cat[x, Embedding[ti]])."""

    def __init__(self, dim, out_dim=None, w=64, time_embed=100, time_varying=False):
        super().__init__()
        self.time_varying = time_varying
        self.time_index_embedding = nn.Embedding(time_embed, w)
        self.net = _selu_stack(dim + w + int(bool(time_varying)), w, dim if out_dim is None else out_dim)

    def forward(self, x, ti):
        if B.use_kernel(self, x) and ti.numel() > 0:
            frames = ti.reshape(-1).long()
            f0 = int(frames[0])
            if bool((frames == f0).all()):      # one frame for the whole batch: folded into the layer-0 bias table
                return B.mlp_forward(self, x, f0)
        return self.net(torch.cat([x, self.time_index_embedding(ti.long()).squeeze()], dim=-1))


class GradModel(nn.Module):
    """Gradient field of a scalar potential network: x = cat[state, t] -> d/dx sum(action(x)) without its time column
    (torchcfm/models/models.py:45-53).  Plain autograd with the graph kept (the potential's parameters train through it);
    no configuration of the hot path uses it -- provided so that `from torchcfm.models.models import *` finds the same names."""

    def __init__(self, action):
        super().__init__()
        self.action = action

    def forward(self, x):
        x = x.requires_grad_(True)
        potential = self.action(x).sum()
        (g,) = torch.autograd.grad(potential, x, create_graph=True)
        return g[:, :-1]
