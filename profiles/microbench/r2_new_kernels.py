"""One launch each of the round-2 kernels that are not on the bench's headline path, for an `ncu --set full` capture:
k_tc16<score head> (2^21 particles, nfe 2), k_train_md (256 steps, B = 10), k_train_flow<score head> (64 steps, B = 256)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import torch
from idff_b200 import fields, sampler, train
from idff_b200.torchcfm.models.models import MLP, MLP_Embedding
from idff_b200.weights import PackedWeights

dev = torch.device("cuda:0")
pw = PackedWeights.from_npz(os.path.join(ROOT, "tests", "golden", "weights_scorehead256.npz"), device=dev)
m = fields.CustomNetModelScoreHead(pw, 0.2, 0.2, 0.2)
x = torch.randn(1 << 21, 2, device=dev) / 1.5
for _ in range(2):
    sampler.sample_rk4(m, x, torch.linspace(0, 1, 3))
torch.manual_seed(0)
ds = (180 * torch.sin(torch.cumsum(torch.randn(200, 46) * 0.05, 0))).to(dev)
ft = train.FusedMDTrainer(MLP_Embedding(dim=46, w=128, time_embed=200, time_varying=True).to(dev), ds, batch_size=10)
idx, t, eps = ft.draw(256)
for _ in range(2):
    ft.steps(idx, t, eps)
fs = train.FusedFlowTrainer(MLP(dim=4, w=256, time_varying=True).to(dev), batch_size=256, lr=1e-3, score_sigma=0.2,
                            data=train.eight_gaussians_device)
x0d, x1d, td, _ = fs.draw(64)
for _ in range(2):
    fs.steps(x0d, x1d, td)
torch.cuda.synchronize()
print("done")
