import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests"]
import numpy as np, torch
from conftest import load_golden
from gpu_common import packed, drift_report
from idff_b200 import fields, sampler, _lib
dev = torch.device("cuda:0")
gt = load_golden("traj2d"); x0 = torch.tensor(gt["x0"]).to(dev)
pw = packed("checkerboard", dev)
for gam, key in (((0.2, 0.2), "checkerboard_o2_nfe2_g2"), ((0.0, 0.0), "checkerboard_o2_nfe2_g0")):
    outs = {}
    for impl in (1, 2):
        m = fields.CustomNetModel(pw, 0.2, *gam); m.impl = impl
        t0 = time.time(); out = sampler.sample_rk4(m, x0, torch.linspace(0, 1, 3)); torch.cuda.synchronize()
        outs[impl] = out.cpu().numpy()
        print(key, "impl", impl, "time", time.time() - t0, drift_report(outs[impl], gt[key][-1]), flush=True)
    print("fused vs generic", drift_report(outs[2], outs[1]), flush=True)
m = fields.CustomNetModelOrder3(pw, 0.2, 0.2, 0.2, 0.2); m.impl = 2
out = sampler.sample_rk4(m, x0, torch.linspace(0, 1, 3)); torch.cuda.synchronize()
print("o3 fused", drift_report(out.cpu().numpy(), gt["checkerboard_o3_nfe2_g1"]), flush=True)
g = load_golden("md"); y = torch.tensor(g["y"]).to(dev); pwm = packed("polyala", dev)
for impl in (1, 2):
    sde = fields.SDE_cal(pwm, input_dim=46, sigma=0.5, init_ind=3, max_length=200, dt=0.3, gamma1=0.2, gamma2=5.0 / 4); sde.impl = impl
    f = sde.f(torch.tensor(0.3), y); torch.cuda.synchronize()
    print("md f impl", impl, drift_report(f.cpu().numpy(), g["f_i3"][1]), "scale", np.abs(g["f_i3"][1]).max(), flush=True)
    out = sampler.sdeint(sde, y, torch.tensor(g["ts"]), dt=0.3, bm=(torch.tensor(g["I_k"]).to(dev), torch.tensor(g["I_k0"]).to(dev)))
    print("md srk impl", impl, drift_report(out[2].cpu().numpy(), g["srk"][2]), flush=True)
x = torch.randn(1 << 22, 2, device=dev) / 1.5
for impl in (1, 2):
    m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2); m.impl = impl
    sampler.sample_rk4(m, x, torch.linspace(0, 1, 3)); torch.cuda.synchronize()
    t0 = time.time(); sampler.sample_rk4(m, x, torch.linspace(0, 1, 3)); torch.cuda.synchronize(); dt = time.time() - t0
    print("impl", impl, "4M particles nfe2:", dt, "s ->", x.shape[0] * 8 / dt / 1e6, "M particle-steps/s", flush=True)
