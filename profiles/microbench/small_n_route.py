"""Small solves of the order-2 field: k_tc16 (64-particle tiles) vs the 32-particle-tile kernel (three-jet kernel, idle third jet).
Per-call time incl. launch, and the drift against the reference's golden end points; run on the GPU box."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch
from conftest import load_golden
from gpu_common import drift_report, packed
from idff_b200 import _lib, fields, sampler

dev = torch.device("cuda:0")
pw = packed("checkerboard", dev)
L = _lib.lib()
g = load_golden("traj2d")
x0g = torch.tensor(g["x0"]).to(dev)


def timeit(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for gam in ((0.2, 0.2), (0.2, 0.0)):
    m = fields.CustomNetModel(pw, 0.2, *gam)
    for nfe in (2, 10):
        ts = torch.linspace(0, 1, nfe + 1)
        for n in (256, 1024, 2048, 3072, 4096, 6144, 8192, 12288, 16384):
            x = torch.randn(n, 2, device=dev) / 1.5
            row = []
            for thr in (0, 1 << 30):
                L.idff_tune_small_n_route(thr)
                row.append(timeit(lambda: sampler.sample_rk4(m, x, ts)))
            print(f"gamma={gam} nfe={nfe} N={n}: k_tc16 {row[0]:.1f} us, 32-particle tiles {row[1]:.1f} us  ({row[0] / row[1]:.2f}x)", flush=True)
m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
for thr in (0, 1 << 30):
    L.idff_tune_small_n_route(thr)
    out = sampler.sample_rk4(m, x0g, torch.linspace(0, 1, 3)).cpu().numpy()
    print("route", thr, "vs golden fp32", drift_report(out, g["checkerboard_o2_nfe2_g2"][-1] if g["checkerboard_o2_nfe2_g2"].ndim == 3 else g["checkerboard_o2_nfe2_g2"]),
          "vs fp64", drift_report(out, g["checkerboard_o2_nfe2_g2_f64"]))
L.idff_tune_small_n_route(0)
