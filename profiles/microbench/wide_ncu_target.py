"""One order-3 solve on the layered tcgen05 engine (w = 1024, 2^17 particles) for an ncu capture of k_wide_layer<3>."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import torch
from idff_b200 import fields, sampler
from idff_b200.torchcfm.models.models import MLP
from idff_b200.weights import PackedWeights
dev = torch.device("cuda:0")
torch.manual_seed(0)
w = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pw = PackedWeights.from_module(MLP(dim=2, w=w, time_varying=True), device=dev)
m = fields.CustomNetModelOrder3(pw, 0.2, 0.2, 0.2, 0.2)
x = torch.randn(1 << 17, 2, device=dev) / 1.5
sampler.sample_rk4(m, x, torch.linspace(0, 1, 3))
torch.cuda.synchronize()
print("done")
