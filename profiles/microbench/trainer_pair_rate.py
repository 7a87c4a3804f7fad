import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from idff_b200 import _lib, train
from idff_b200.torchcfm.models.models import MLP
dev = torch.device("cuda:0")
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
for B in (256, 128, 296):
    for pair in (0, 1):
        _lib.lib().idff_debug_train_pair_mode(pair)
        torch.manual_seed(0)
        ft = train.FusedFlowTrainer(MLP(dim=2, w=256, time_varying=True).to(dev), batch_size=B)
        x0, x1, t, _ = ft.draw(1024)
        s = timeit(lambda: ft.steps(x0, x1, t)) / 1024
        print(f"B={B} pair={pair}: {s * 1e6:.2f} us/step", flush=True)
_lib.lib().idff_debug_train_pair_mode(1)
