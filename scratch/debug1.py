import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests"]
import numpy as np, torch
import idff_oracle as O
from conftest import load_golden
from gpu_common import packed, drift_report
from idff_b200 import fields, sampler, mmd as M, _lib
from idff_b200.torchcfm import conditional_flow_matching as cfm
from make_golden_consts import FIELD_TIMES, GAMMAS2, GAMMAS3
dev = torch.device("cuda:0")
print("== path mode 1")
g = torch.Generator().manual_seed(7)
B, D = 1000, 46
x0, x1, eps = (torch.randn(B, D, generator=g) for _ in range(3)); t = torch.rand(B, generator=g)
for sigma in (0.3, 0.5):
    xt, ut = cfm._path_sample(x0.to(dev), x1.to(dev), t.to(dev), eps.to(dev), sigma, 1)
    xr, ur = O.path_sample(x0, x1, t, eps, sigma, True)
    d = (xt.cpu() - xr).abs(); print(sigma, "xt mismatches", int((d > 0).sum()), "max", d.max().item(), "ut eq", torch.equal(ut.cpu(), ur))
    z = torch.zeros(B, D); one = torch.ones(B, D)
    st, _ = cfm._path_sample(z.to(dev), z.to(dev), t.to(dev), one.to(dev), sigma, 1)
    ref = (sigma * torch.sqrt(t * (1 - t)))[:, None] * one
    d = (st.cpu() - ref).abs(); print("   sigma_t mismatches", int((d > 0).sum()), d.max().item())
    ref2 = (torch.tensor(sigma, dtype=torch.float32) * torch.sqrt(t * (1 - t)))[:, None] * one
    print("   vs fp32-scalar product", int(((st.cpu() - ref2).abs() > 0).sum()))
    sq, _ = cfm._path_sample(z.to(dev), z.to(dev), t.to(dev), one.to(dev), 1.0, 1)
    print("   sqrt mismatches", int(((sq.cpu()[:, 0] - torch.sqrt(t * (1 - t))).abs() > 0).sum()))
print("== field 8gaussians")
gf = load_golden("field2d"); x = torch.tensor(gf["x"]).to(dev)
for which in ("checkerboard", "8gaussians"):
    pw = packed(which, dev)
    for gi, gam in enumerate(GAMMAS2):
        m = fields.CustomNetModel(pw, 0.2, *gam)
        for ti, tv in enumerate(FIELD_TIMES):
            out = m(torch.tensor(tv, dtype=torch.float32), x).cpu().numpy()
            ref = gf[f"{which}_o2_g{gi}"][ti]; r64 = gf[f"{which}_o2_g{gi}_f64"][ti]
            sc = max(1, np.abs(ref).max()); d = np.abs(out - ref) / sc; d64 = np.abs(out - r64) / sc; b = np.abs(ref - r64) / sc
            if d.max() > 5e-6 or d64.max() > 5e-6:
                print(which, "o2", gam, f"t={tv:.3f} scale={sc:.3g} max={d.max():.2e} n>1e-5={(d > 1e-5).sum()} med={np.median(d):.1e} | vs64 max={d64.max():.2e} | ref32-vs-64 max={b.max():.2e}")
    for gi, gam in enumerate(GAMMAS3):
        m = fields.CustomNetModelOrder3(pw, 0.2, *gam)
        for ti, tv in enumerate(FIELD_TIMES):
            out = m(torch.tensor(tv, dtype=torch.float32), x).cpu().numpy()
            ref = gf[f"{which}_o3_g{gi}"][ti]; r64 = gf[f"{which}_o3_g{gi}_f64"][ti]
            sc = max(1, np.abs(ref).max()); d = np.abs(out - ref) / sc; d64 = np.abs(out - r64) / sc; b = np.abs(ref - r64) / sc
            if d.max() > 5e-6:
                print(which, "o3", gam, f"t={tv:.3f} scale={sc:.3g} max={d.max():.2e} n>1e-5={(d > 1e-5).sum()} med={np.median(d):.1e} | vs64 max={d64.max():.2e} | ref32-vs-64 max={b.max():.2e}")
print("== rk4")
gt = load_golden("traj2d"); x0 = torch.tensor(gt["x0"]).to(dev)
for which in ("checkerboard", "8gaussians"):
    pw = packed(which, dev)
    for nfe in (2, 3, 5, 10):
        for gi, gam in enumerate(GAMMAS2[:4]):
            key = f"{which}_o2_nfe{nfe}_g{gi}"
            if key not in gt.files: continue
            m = fields.CustomNetModel(pw, 0.2, *gam)
            ts = torch.linspace(0, 1, nfe + 1)
            traj = sampler.odeint(m, x0, ts, method="rk4"); final = sampler.sample_rk4(m, x0, ts)
            ref = gt[key]; rf = ref[-1] if ref.ndim == 3 else ref
            print(key, "eq traj0", torch.equal(traj[0], x0), "eq final", torch.equal(final, traj[-1]), drift_report(final.cpu().numpy(), rf), "vs64", drift_report(final.cpu().numpy(), gt[key + "_f64"]), "base", drift_report(rf, gt[key + "_f64"]))
            if ref.ndim == 3: print("    traj[1]", drift_report(traj[1].cpu().numpy(), ref[1]))
print("== mmd big")
for N, Mm, D in [(130, 70, 5), (257, 1100, 46), (64, 64, 16)]:
    g = torch.Generator().manual_seed(N + Mm + D)
    x, y = torch.randn(N, D, generator=g), torch.randn(Mm, D, generator=g) * 1.3 + 0.2
    ref = np.array(O.mmd_partial_sums_fp64(x, y, 1.7))
    s = M.mmd_partial_sums(x.to(dev), y.to(dev), 1.7).cpu().numpy()
    print(N, Mm, D, "got", s, "ref", ref, "rel", (s - ref) / ref)
