"""A/B of libidff_b200 variants on one box (clocks differ between boxes: compare only within a run).
usage: IDFF_B200_LIB=<path to .so> python profiles/microbench/ab_variants.py [log2 particles]
Prints parity of the default sampler against the fp32 FMA kernel and its throughput at nfe 2 / 10."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, ROOT + "/tests"]
import numpy as np, torch
from gpu_common import packed, drift_report
from idff_b200 import fields, sampler, _lib
dev = torch.device("cuda:0")
pw = packed("checkerboard", dev)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 23
m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
mf = fields.CustomNetModel(pw, 0.2, 0.2, 0.2); mf.impl = _lib.IDFF_IMPL_FUSED
x = torch.randn(70001, 2, device=dev, generator=torch.Generator(device=dev).manual_seed(3)) / 1.5
out = {"lib": os.path.basename(_lib.LIB_PATH)}
for nfe in (2, 5):
    ts = torch.linspace(0, 1, nfe + 1)
    r = drift_report(sampler.sample_rk4(m, x, ts).cpu().numpy(), sampler.sample_rk4(mf, x, ts).cpu().numpy())
    out[f"parity_nfe{nfe}"] = {k: float(f"{v:.3g}") for k, v in r.items()}
xx = torch.randn(1 << lg, 2, device=dev) / 1.5
for nfe in (2, 10):
    ts = torch.linspace(0, 1, nfe + 1)
    for _ in range(2):
        sampler.sample_rk4(m, xx, ts)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    a.record()
    for _ in range(reps):
        sampler.sample_rk4(m, xx, ts)
    b.record()
    torch.cuda.synchronize()
    out[f"Msteps_nfe{nfe}"] = round(xx.shape[0] * 4 * nfe * reps / (a.elapsed_time(b) * 1e-3) / 1e6, 1)
out["nonfinite"] = _lib.weights_nonfinite(pw.handle)
print(out, flush=True)
