"""Order-3 field (three jets) at width 256: fused three-jet tcgen05 kernel (default route) vs the layered engine vs FFMA2,
and the reference's 10^3-tuple gamma sweep at 2048 particles (Autograd_example_order3.py:288-314)."""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, ROOT + "/tests"]
import numpy as np, torch
from gpu_common import packed, drift_report
from idff_b200 import fields, sampler, sweep, _lib
dev = torch.device("cuda:0")
pw = packed("checkerboard", dev)


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


F = 3 * 256 + 2 * 256 * 256 + 256 * 2
for lg in (11, 16, 21):
    x = torch.randn(1 << lg, 2, device=dev) / 1.5
    for nfe in (2, 10):
        ts = torch.linspace(0, 1, nfe + 1)
        row = {}
        for name, impl in (("fused3jet", 0), ("layered", 4), ("ffma2", 2)):
            if impl == 2 and lg > 16:
                continue
            m = fields.CustomNetModelOrder3(pw, 0.2, 0.2, 0.2, 0.2)
            m.impl = impl
            s = timeit(lambda: sampler.sample_rk4(m, x, ts))
            row[name] = round(x.shape[0] * 4 * nfe / s / 1e6, 1)
        flops = 2.0 * F * ((4 * nfe - 2) * 6 + 2)
        print(f"2^{lg} particles nfe {nfe}: M particle-steps/s {row}  (fused: {row['fused3jet'] * 1e6 / (4 * nfe) * flops / 1e12:.0f} TFLOP/s algorithmic)", flush=True)
grid10 = [0.0, 0.1, 0.2, 0.3, 0.5, 0.7, 1.0, 1.5, 2.0, 3.0]
x2k = torch.randn(2048, 2, device=dev) / 1.5
true2k = torch.randn(2048, 2, device=dev)
s = timeit(lambda: sweep.find_best_gammas_mmd(pw, x2k, true2k, gamma_range=grid10, nfe=2, sigma=0.2, sigma_kernel=1.0, order=3), reps=3, warm=1)
print(f"10x10x10 order-3 sweep, 2048 particles, nfe 2 (solves + MMD): {s * 1e3:.1f} ms", flush=True)
