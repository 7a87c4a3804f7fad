"""PolyALA SDE_cal (47-128-128-128-46, folded embedding), srk, ts = linspace(0,1,3), dt = 0.3: the tcgen05 3xFP16 kernel
(impl 5) vs the FFMA2 kernel (impl 2), parity between them and throughput over the number of trajectories."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, ROOT + "/tests"]
import numpy as np, torch
from gpu_common import packed, drift_report
from idff_b200 import fields, sampler, _lib
dev = torch.device("cuda:0")
pw = packed("polyala", dev)


def timeit(fn, reps=3, warm=2):
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


tsd = torch.linspace(0, 1, 3)
nst = sampler.sde_num_steps(tsd, 0.3)
g = torch.Generator(device=dev).manual_seed(1)
for lg in (10, 13, 17, 20):
    K = 1 << lg
    y0 = torch.randn(K, 46, device=dev, generator=g) * 60
    bm = sampler.brownian_increments(nst, tsd, 0.3, (K, 46), dev)
    outs, row = {}, {}
    for name, impl in (("tensor16", 5), ("ffma2", 2)):
        sde = fields.SDE_cal(pw, input_dim=46, sigma=0.5, init_ind=3, max_length=200, dt=0.3, gamma1=0.2, gamma2=1.25)
        sde.impl = impl
        outs[name] = sampler.sdeint(sde, y0, tsd, dt=0.3, bm=bm)
        s = timeit(lambda: sampler.sdeint(sde, y0, tsd, dt=0.3, bm=bm))
        row[name] = round(K * nst * 3 / s / 1e6, 1)
    d = (outs["tensor16"][-1] - outs["ffma2"][-1]).abs()
    scale = outs["ffma2"][-1].abs().max().item()
    fin = torch.isfinite(outs["tensor16"]).all().item()
    print(f"K=2^{lg}: M drift-evals/s {row}  tensor16 vs ffma2: max {d.max().item():.3e} median {d.median().item():.3e} (scale {scale:.3g}) finite={fin}"
          f"  tensor16 {row['tensor16'] * 1e6 * 2.0 * 44672 * (nst * 3 * 4 - 2 * 3) / (nst * 3) / 1e12:.0f} TFLOP/s algorithmic", flush=True)
print("nonfinite counters", _lib.weights_nonfinite(pw.handle))
