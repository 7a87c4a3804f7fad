"""Phase times (clock64, CTA 0) of one step of the PolyALA cluster-resident trainer + rate; run on the GPU box."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from idff_b200 import _lib, train
from idff_b200.torchcfm.models.models import MLP_Embedding

dev = torch.device("cuda:0")


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


ds = (180 * torch.sin(torch.cumsum(torch.randn(200, 46) * 0.05, 0))).to(dev)
for B in (10, 16):
    torch.manual_seed(0)
    m = MLP_Embedding(dim=46, w=128, time_embed=200, time_varying=True).to(dev)
    ft = train.FusedMDTrainer(m, ds, batch_size=B)
    idx, t, eps = ft.draw(2048)
    print(f"B={B}: fused us/step", timeit(lambda: ft.steps(idx, t, eps)) / 2048 * 1e6)
    st = (C.c_int64 * 32)()
    _lib.check(_lib.lib().idff_debug_train_md_stamps(st), "stamps")
    v = [int(x) for x in st][:18]
    names = ["phase 0 + emb push", "barrier 1 (+ deferred emb Adam)", "layer 1 + push", "barrier 2", "layer 2 + push", "barrier 3",
             "layer 3 + push", "barrier 4", "output, loss, dA3 partials", "barrier 5 (+ dW3, Adam)", "dZ3, dA2 partials",
             "barrier 6 (+ dW2, Adam)", "dZ2, dA1 partials", "barrier 7 (+ dW1, Adam)", "dZ1, dE partials", "barrier 8 (+ dW0, Adam)",
             "emb Adam (rows gathered next)"]
    print("  total cycles", v[17] - v[0])
    for n, a_, b_ in zip(names, v[:-1], v[1:]):
        print(f"  {n:34s} {b_ - a_}")
torch.manual_seed(0)
m2 = MLP_Embedding(dim=46, w=128, time_embed=200, time_varying=True).to(dev)
gt = train.GraphedMDTrainer(m2, ds, batch_size=10)
print("graph us/step", timeit(gt.step, reps=200, warm=5) * 1e6)
