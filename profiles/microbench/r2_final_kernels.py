"""One launch each, for an `ncu --set full` capture at the end of round 2: k_train_flow<flow net> with the chunked weight-gradient
staging (64 steps, B = 256; one CTA per row group -- ncu cannot replay the cooperative + cluster launch of the default pair mode:
LaunchFailed) and k_md_chain under idff_md_generate_sweep (64 parameter sets x 5 trajectories, 40 frames)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch
from idff_b200 import _lib, md, train
from idff_b200.torchcfm.models.models import MLP
from idff_b200.weights import PackedWeights

dev = torch.device("cuda:0")
torch.manual_seed(0)
_lib.lib().idff_debug_train_pair_mode(0)
ft = train.FusedFlowTrainer(MLP(dim=2, w=256, time_varying=True).to(dev), batch_size=256)
x0d, x1d, td, _ = ft.draw(64)
for _ in range(2):
    ft.steps(x0d, x1d, td)
pw = PackedWeights.from_npz(np.load(os.path.join(ROOT, "tests", "golden", "weights_polyala.npz")), device=dev)
sets = [(0.2 + 0.1 * a, 0.1 * b, 1.0 + c) for a in range(4) for b in range(4) for c in range(4)]
for _ in range(2):
    md.generate_trajectories_sweep(pw, sets, 5, 46, 40, 3, dt=0.3, device=dev)
torch.cuda.synchronize()
print("done")
