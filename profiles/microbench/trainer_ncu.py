"""One launch of the fused training kernel (16 steps, B = 256) for an ncu capture."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from idff_b200 import train
from idff_b200.torchcfm.models.models import MLP
dev = torch.device("cuda:0")
torch.manual_seed(0)
tr = train.FusedFlowTrainer(MLP(dim=2, w=256, time_varying=True).to(dev), batch_size=256)
x0, x1, t, eps = tr.draw(16)
print(tr.steps(x0, x1, t).cpu())
