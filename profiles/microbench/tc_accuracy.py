import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests"]
import numpy as np, torch
from conftest import load_golden
from gpu_common import packed, drift_report
from idff_b200 import fields, sampler
from make_golden_consts import FIELD_TIMES, GAMMAS2
dev = torch.device("cuda:0")
IMPLS = tuple(int(a) for a in sys.argv[1:]) or (2, 3, 5)   # 2 FFMA2, 3 tcgen05 3xTF32, 5 tcgen05 3xFP16
gf = load_golden("field2d"); x = torch.tensor(gf["x"]).to(dev)
for which in ("checkerboard", "8gaussians"):
    pw = packed(which, dev)
    for impl in IMPLS:
        worst = 0; worst64 = 0
        for gi, gam in enumerate(GAMMAS2):
            for ti, tv in enumerate(FIELD_TIMES):
                m = fields.CustomNetModel(pw, 0.2, *gam); m.impl = impl
                out = m(torch.tensor(tv, dtype=torch.float32), x).cpu().numpy()
                ref = gf[f"{which}_o2_g{gi}"][ti]; r64 = gf[f"{which}_o2_g{gi}_f64"][ti]
                sc = max(1, np.abs(ref).max()); gap = np.abs(ref - r64).max() / sc
                e32 = np.abs(out - ref).max() / sc; e64 = np.abs(out - r64).max() / sc
                if tv < 1: worst = max(worst, e32); worst64 = max(worst64, e64)
                if impl != 2 and (e32 > 8e-6) and tv < 1: print("   ", which, gam, f"t={tv:.3f} e32 {e32:.2e} e64 {e64:.2e} refgap {gap:.2e}")
        print(which, "impl", impl, f"worst interior e32 {worst:.2e} e64 {worst64:.2e}")
gt = load_golden("traj2d"); x0 = torch.tensor(gt["x0"]).to(dev)
for key, gam, nfe in (("checkerboard_o2_nfe2_g2", (0.2, 0.2), 2), ("8gaussians_o2_nfe2_g2", (0.2, 0.2), 2), ("8gaussians_o2_nfe10_g1", (0.2, 0.0), 10)):
    pw = packed(key.split("_")[0], dev)
    for impl in IMPLS:
        m = fields.CustomNetModel(pw, 0.2, *gam); m.impl = impl
        out = sampler.sample_rk4(m, x0, torch.linspace(0, 1, nfe + 1)).cpu().numpy()
        ref = gt[key]; ref = ref[-1] if ref.ndim == 3 else ref
        print(key, "impl", impl, drift_report(out, ref), "vs64", drift_report(out, gt[key + "_f64"]), "base", drift_report(ref, gt[key + "_f64"]))

# throughput at 2^21 particles, nfe = 2 (CUDA events)
xs = torch.randn(1 << 21, 2, device=dev) / 1.5
pw = packed("checkerboard", dev)
for impl in IMPLS:
    m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2); m.impl = impl
    ts = torch.linspace(0, 1, 3)
    sampler.sample_rk4(m, xs, ts); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); sampler.sample_rk4(m, xs, ts); sampler.sample_rk4(m, xs, ts); b.record(); torch.cuda.synchronize()
    print("impl", impl, f"{xs.shape[0] * 8 * 2 / (a.elapsed_time(b) * 1e-3) / 1e6:.1f} M particle-steps/s")
from idff_b200 import _lib
print("tensor16 nonfinite count:", _lib.tensor16_nonfinite())
