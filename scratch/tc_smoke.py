import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests"]
import numpy as np, torch
from conftest import load_golden
from gpu_common import packed, drift_report
from idff_b200 import fields, sampler, _lib
from make_golden_consts import FIELD_TIMES, GAMMAS2
dev = torch.device("cuda:0")
gf = load_golden("field2d"); x = torch.tensor(gf["x"]).to(dev)
pw = packed("checkerboard", dev)
print("== field eval, tensor (3) vs golden", flush=True)
for gi, gam in enumerate(GAMMAS2[:4]):
    for ti, tv in enumerate(FIELD_TIMES):
        m = fields.CustomNetModel(pw, 0.2, *gam); m.impl = 3
        out = m(torch.tensor(tv, dtype=torch.float32), x); torch.cuda.synchronize()
        out = out.cpu().numpy()
        ref = gf[f"checkerboard_o2_g{gi}"][ti]; r64 = gf[f"checkerboard_o2_g{gi}_f64"][ti]
        sc = max(1, np.abs(ref).max())
        print(gam, f"t={tv:.3f} vs32 {np.abs(out-ref).max()/sc:.2e} vs64 {np.abs(out-r64).max()/sc:.2e} ref32-64 {np.abs(ref-r64).max()/sc:.2e}", flush=True)
gt = load_golden("traj2d"); x0 = torch.tensor(gt["x0"]).to(dev)
print("== rk4", flush=True)
for key, gam, nfe in (("checkerboard_o2_nfe2_g2", (0.2, 0.2), 2), ("checkerboard_o2_nfe10_g1", (0.2, 0.0), 10), ("checkerboard_o2_nfe5_g2", (0.2, 0.2), 5)):
    for impl in (2, 3):
        m = fields.CustomNetModel(pw, 0.2, *gam); m.impl = impl
        out = sampler.sample_rk4(m, x0, torch.linspace(0, 1, nfe + 1)); torch.cuda.synchronize()
        ref = gt[key]; ref = ref[-1] if ref.ndim == 3 else ref
        print(key, "impl", impl, drift_report(out.cpu().numpy(), ref), "vs64", drift_report(out.cpu().numpy(), gt[key + "_f64"]), flush=True)
xx = torch.randn(1 << 21, 2, device=dev) / 1.5
for impl in (2, 3):
    m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2); m.impl = impl
    sampler.sample_rk4(m, xx, torch.linspace(0, 1, 3)); torch.cuda.synchronize()
    t0 = time.time(); sampler.sample_rk4(m, xx, torch.linspace(0, 1, 3)); torch.cuda.synchronize(); dt = time.time() - t0
    print("impl", impl, "2M particles nfe2:", dt, "s ->", xx.shape[0] * 8 / dt / 1e6, "M particle-steps/s", flush=True)
