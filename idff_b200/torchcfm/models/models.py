"""The two networks of the hot path with the reference's constructor signatures, attribute names and
state_dict keys (torchcfm/models/models.py:4-21 MLP, :23-42 MLP_Embedding), so the shipped checkpoints load.

Under autograd (training) the layers run as ordinary torch modules -- training the MLP itself is outside the
hot path built here.  Under torch.no_grad() on a CUDA tensor the forward is one launch of the B200 MLP kernel
(idff_mlp_forward); MLP_Embedding takes that route when every row shares one time index, which is how the
generator calls it (MD_simulation.py:501)."""
import torch

from .. import _backend as B


class MLP(torch.nn.Module):
    def __init__(self, dim, out_dim=None, w=64, time_varying=False):
        super().__init__()
        self.time_varying = time_varying
        if out_dim is None:
            out_dim = dim
        self.net = torch.nn.Sequential(
            torch.nn.Linear(dim + (1 if time_varying else 0), w), torch.nn.SELU(),
            torch.nn.Linear(w, w), torch.nn.SELU(),
            torch.nn.Linear(w, w), torch.nn.SELU(),
            torch.nn.Linear(w, out_dim))

    def forward(self, x):
        if B.use_kernel(self, x):
            return B.mlp_forward(self, x, 0)
        return self.net(x)


class MLP_Embedding(torch.nn.Module):
    def __init__(self, dim, out_dim=None, w=64, time_embed=100, time_varying=False):
        super().__init__()
        self.time_varying = time_varying
        if out_dim is None:
            out_dim = dim
        self.time_index_embedding = torch.nn.Embedding(time_embed, w)
        self.net = torch.nn.Sequential(
            torch.nn.Linear(dim + w + (1 if time_varying else 0), w), torch.nn.SELU(),
            torch.nn.Linear(w, w), torch.nn.SELU(),
            torch.nn.Linear(w, w), torch.nn.SELU(),
            torch.nn.Linear(w, out_dim))

    def forward(self, x, ti):
        if B.use_kernel(self, x) and ti.numel() > 0:
            idx = ti.reshape(-1).long()
            first = int(idx[0])
            if bool((idx == first).all()):
                return B.mlp_forward(self, x, first)
        emb = self.time_index_embedding(ti.long()).squeeze()
        return self.net(torch.cat([x, emb], dim=-1))
