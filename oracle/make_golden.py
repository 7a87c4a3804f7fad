"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz by EXECUTING the reference's own code.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
Everything written here is produced by the reference's code objects (oracle/ref_loader.py);
the only restated pieces are the two un-vendored third-party integrators (torchdiffeq rk4,
torchsde srk -- see oracle/idff_oracle.py header).  tests/test_oracle_vs_reference.py checks
that oracle/idff_oracle.py reproduces these files bit-for-bit, and the `-m gpu` tests compare
the CUDA path against them on the GPU box, where /root/reference does not exist.

Files:
  weights_{checkerboard,8gaussians}.npz   W0..W3,b0..b3 of 2D_examples/sample_*/best_IDFF_model.pt
  weights_polyala.npz                     + emb, of timeseris_example/MD/polyALA/MD_trained_polyALA.pt
  weights_scorehead.npz                   MLP(dim=4,w=64) default init, torch.manual_seed(0)
  weights_scorehead256.npz scorehead256.npz   MLP(dim=4,w=256) (the order-1 script's network), `make_golden.py scorehead256`
  field2d.npz  traj2d.npz  scorehead.npz  md.npz  path.npz  loss.npz  mmd.npz
"""
from __future__ import annotations

import itertools
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import idff_oracle as O  # noqa: E402
import ref_loader as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

FIELD_TIMES = [0.0, 1.0 / 6.0, 1.0 / 3.0, 0.37, 0.5, 5.0 / 6.0, 1.0]
GAMMAS2 = [(0.0, 0.0), (0.2, 0.0), (0.2, 0.2), (1.0, 2.0), (-0.5, 0.3)]
GAMMAS3 = [(0.0, 0.0, 0.0), (0.2, 0.2, 0.2), (1.0, 2.0, 0.5)]


def linear_layers(net):
    return [m for m in net.net if isinstance(m, torch.nn.Linear)]


def export_weights(net, path, emb=None):
    d = {}
    for i, m in enumerate(linear_layers(net)):
        d[f"W{i}"] = m.weight.detach().numpy().copy()
        d[f"b{i}"] = m.bias.detach().numpy().copy()
    if emb is not None:
        d["emb"] = emb.detach().numpy().copy()
    np.savez(path, **d)


def x0_2d(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 2, generator=g) / 1.5          # Autograd_example_order2.py:136


def run_field(model, t, x):
    return model(t, x.clone()).detach()


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    models = R.models()

    # ------------------------------------------------------------------ weights
    nets = {w: R.load_2d_checkpoint(w) for w in ("checkerboard", "8gaussians")}
    for w, net in nets.items():
        export_weights(net, os.path.join(OUT, f"weights_{w}.npz"))
    md_net = R.load_md_checkpoint()
    export_weights(md_net, os.path.join(OUT, "weights_polyala.npz"), emb=md_net.time_index_embedding.weight)
    torch.manual_seed(0)
    sh_net = models.MLP(dim=4, w=64, time_varying=True).eval()
    export_weights(sh_net, os.path.join(OUT, "weights_scorehead.npz"))

    # ------------------------------------------------------------------ 2-D field evaluations (a3, a4, a6)
    o2 = R.script_symbols(R.ORDER2, ["CustomNetModel", "CFMModel", "compute_mmd_rbf"])
    o3 = R.script_symbols(R.ORDER3, ["CustomNetModel"])
    x = x0_2d(256, seed=1) * 3.0                        # spread over the data range
    fd = {"x": x.numpy(), "times": np.array(FIELD_TIMES, dtype=np.float32),
          "gammas2": np.array(GAMMAS2, dtype=np.float64), "gammas3": np.array(GAMMAS3, dtype=np.float64)}
    for w, net in nets.items():
        net64 = R.models().MLP(dim=2, w=256, time_varying=True).double()
        net64.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
        for gi, g in enumerate(GAMMAS2):
            m32 = o2["CustomNetModel"](net, sigma=0.2, gamma1=g[0], gamma2=g[1])
            m64 = o2["CustomNetModel"](net64, sigma=0.2, gamma1=g[0], gamma2=g[1])
            fd[f"{w}_o2_g{gi}"] = np.stack([run_field(m32, torch.tensor(t, dtype=torch.float32), x).numpy()
                                           for t in FIELD_TIMES])
            fd[f"{w}_o2_g{gi}_f64"] = np.stack([
                run_field(m64, torch.tensor(np.float32(t), dtype=torch.float64), x.double()).numpy()
                for t in FIELD_TIMES])
        for gi, g in enumerate(GAMMAS3):
            m32 = o3["CustomNetModel"](net, sigma=0.2, gamma1=g[0], gamma2=g[1], gamma3=g[2])
            m64 = o3["CustomNetModel"](net64, sigma=0.2, gamma1=g[0], gamma2=g[1], gamma3=g[2])
            fd[f"{w}_o3_g{gi}"] = np.stack([run_field(m32, torch.tensor(t, dtype=torch.float32), x).numpy()
                                           for t in FIELD_TIMES])
            fd[f"{w}_o3_g{gi}_f64"] = np.stack([
                run_field(m64, torch.tensor(np.float32(t), dtype=torch.float64), x.double()).numpy()
                for t in FIELD_TIMES])
        # plain MLP forward (a1) on cat[x, t]
        fd[f"{w}_mlp"] = net(torch.cat([x, torch.full((x.shape[0], 1), 0.37)], -1)).detach().numpy()
    np.savez(os.path.join(OUT, "field2d.npz"), **fd)

    # ------------------------------------------------------------------ 2-D trajectories (a8 + a3/a4/a6)
    N = 1024
    x0 = x0_2d(N, seed=0)
    td = {"x0": x0.numpy()}
    for w, net in nets.items():
        net64 = R.models().MLP(dim=2, w=256, time_varying=True).double()
        net64.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
        for nfe, (gi, g) in itertools.product([2, 3, 5, 10], enumerate(GAMMAS2[:4])):
            if nfe in (3, 5) and gi not in (2,):
                continue
            ts = torch.linspace(0, 1, nfe + 1)
            m32 = o2["CustomNetModel"](net, sigma=0.2, gamma1=g[0], gamma2=g[1])
            tr = O.odeint_rk4(m32, x0.clone(), ts).detach()
            td[f"{w}_o2_nfe{nfe}_g{gi}"] = tr.numpy() if nfe <= 3 else tr[-1].numpy()
            m64 = o2["CustomNetModel"](net64, sigma=0.2, gamma1=g[0], gamma2=g[1])
            tr64 = O.odeint_rk4(m64, x0.double(), ts.double()).detach()
            td[f"{w}_o2_nfe{nfe}_g{gi}_f64"] = tr64[-1].numpy()
        for nfe, (gi, g) in itertools.product([2, 5], enumerate(GAMMAS3[:2])):
            ts = torch.linspace(0, 1, nfe + 1)
            m32 = o3["CustomNetModel"](net, sigma=0.2, gamma1=g[0], gamma2=g[1], gamma3=g[2])
            td[f"{w}_o3_nfe{nfe}_g{gi}"] = O.odeint_rk4(m32, x0.clone(), ts).detach()[-1].numpy()
            m64 = o3["CustomNetModel"](net64, sigma=0.2, gamma1=g[0], gamma2=g[1], gamma3=g[2])
            td[f"{w}_o3_nfe{nfe}_g{gi}_f64"] = O.odeint_rk4(m64, x0.double(), ts.double()).detach()[-1].numpy()
        cfm = o2["CFMModel"](net)
        td[f"{w}_cfm_nfe5"] = O.odeint_rk4(cfm, x0.clone(), torch.linspace(0, 1, 6)).detach()[-1].numpy()
    np.savez(os.path.join(OUT, "traj2d.npz"), **td)

    # ------------------------------------------------------------------ order-1 / score head (a5)
    o1 = R.script_symbols(R.ORDER1, ["CustomNetModel", "log_prob_xt_given_x1_dimwise"])
    sd = {"x": x.numpy(), "times": np.array(FIELD_TIMES, dtype=np.float32)}
    for gi, g in enumerate(GAMMAS2[:4]):
        m = o1["CustomNetModel"](sh_net, sigma=0.2, gamma1=g[0], gamma2=g[1])
        sd[f"field_g{gi}"] = np.stack([run_field(m, torch.tensor(t, dtype=torch.float32), x).numpy()
                                       for t in FIELD_TIMES])
        sd[f"traj_nfe5_g{gi}"] = O.odeint_rk4(m, x0[:256].clone(), torch.linspace(0, 1, 6)).detach()[-1].numpy()
    sd["gammas"] = np.array(GAMMAS2[:4])
    np.savez(os.path.join(OUT, "scorehead.npz"), **sd)

    # ------------------------------------------------------------------ PolyALA SDE_cal f/g (a7) + srk/euler solves
    mdsym = R.script_symbols(R.MD, ["SDE_cal", "compute_mmd"])
    g = torch.Generator().manual_seed(3)
    y = torch.randn(64, 46, generator=g) * 60.0
    mdd = {"y": y.numpy(), "times": np.array([0.0, 0.3, 0.6, 1.0], dtype=np.float32),
           "init_inds": np.array([0, 3, 199])}
    for ii in (0, 3, 199):
        sde = mdsym["SDE_cal"](md_net, input_dim=46, sigma=0.5, init_ind=ii, max_length=200, dt=0.3,
                               gamma1=0.2, gamma2=5.0 / (1 + ii))
        mdd[f"f_i{ii}"] = np.stack([sde.f(torch.tensor(t), y.clone()).detach().numpy()
                                    for t in (0.0, 0.3, 0.6, 1.0)])
        mdd[f"g_i{ii}"] = np.stack([sde.g(torch.tensor(t), y.clone()).detach().numpy()
                                    for t in (0.0, 0.3, 0.6, 1.0)])
    # one srk + one euler solve with injected increments (ts = linspace(0,1,NFE=3), dt = 0.3)
    sde = mdsym["SDE_cal"](md_net, input_dim=46, sigma=0.5, init_ind=3, max_length=200, dt=0.3,
                           gamma1=0.2, gamma2=5.0 / 4)
    ts = torch.linspace(0, 1, 3)
    grid = O.sde_step_grid(ts, 0.3)
    ns = len(grid) - 1
    h = torch.tensor(np.diff(np.array(grid, dtype=np.float32)))
    I_k = torch.randn(ns, 64, 46, generator=g) * h.sqrt()[:, None, None]
    I_k0 = torch.randn(ns, 64, 46, generator=g) * (h ** 1.5)[:, None, None] / np.sqrt(12.0) + 0.5 * h[:, None, None] * I_k
    mdd["I_k"], mdd["I_k0"], mdd["ts"] = I_k.numpy(), I_k0.numpy(), ts.numpy()
    mdd["srk"] = O.sdeint_srk(sde, y.clone(), ts, 0.3, I_k, I_k0).detach().numpy()
    mdd["euler"] = O.sdeint_euler(sde, y.clone(), ts, 0.3, I_k).detach().numpy()
    np.savez(os.path.join(OUT, "md.npz"), **mdd)

    # ------------------------------------------------------------------ path sample (a9/a10/a11)
    pd = {}
    for dyn in (False, True):
        cfm = R.cfm(dynamic=dyn)
        for (B, D), sigma, cls in itertools.product([(256, 2), (10, 46), (33, 5)], [0.0, 0.5],
                                                    ["ConditionalFlowMatcher",
                                                     "ExactOptimalTransportConditionalFlowMatcher"]):
            if dyn and cls.startswith("Exact"):
                continue      # the dynamic ExactOT matcher needs POT for the coupling (out of scope)
            key = f"{'dyn' if dyn else 'std'}_{cls[:5]}_B{B}_D{D}_s{sigma}"
            gg = torch.Generator().manual_seed(11)
            a0 = torch.randn(B, D, generator=gg)
            a1 = torch.randn(B, D, generator=gg) * 2 + 1
            FM = getattr(cfm, cls).__new__(getattr(cfm, cls))
            FM.sigma = sigma                         # ctor of ExactOT dereferences POT; bypass it
            torch.manual_seed(5)
            t, xt, ut, eps = FM.sample_location_and_conditional_flow(a0, a1, return_noise=True)
            pd[key + "_x0"], pd[key + "_x1"] = a0.numpy(), a1.numpy()
            pd[key + "_t"], pd[key + "_xt"], pd[key + "_ut"], pd[key + "_eps"] = (
                t.numpy(), xt.numpy(), ut.numpy(), eps.numpy())
            pd[key + "_lambda"] = FM.compute_lambda(t).numpy() if torch.is_tensor(FM.compute_lambda(t)) \
                else np.float64(FM.compute_lambda(t))
    np.savez(os.path.join(OUT, "path.npz"), **pd)

    # ------------------------------------------------------------------ losses (a12)
    gg = torch.Generator().manual_seed(21)
    ld = {}
    for B, D in [(256, 2), (10, 46), (1000, 2)]:
        vt = torch.randn(B, D, generator=gg)
        st = torch.randn(B, D, generator=gg)
        b0 = torch.randn(B, D, generator=gg)
        b1 = torch.randn(B, D, generator=gg) * 2
        t = torch.rand(B, generator=gg)
        eps = torch.randn(B, D, generator=gg)
        xt = t[:, None] * b1 + (1 - t[:, None]) * b0 + 0.1 * eps
        k = f"B{B}_D{D}"
        for n, v in dict(vt=vt, st=st, x0=b0, x1=b1, t=t, xt=xt).items():
            ld[f"{k}_{n}"] = v.numpy()
        vt_ = vt.clone().requires_grad_(True)
        loss = torch.mean(((1 / (1e-2 + 1 - t[:, None])) * (vt_ - b1)) ** 2)     # order2.py:105
        loss.backward()
        ld[f"{k}_flow"], ld[f"{k}_flow_gvt"] = loss.item(), vt_.grad.numpy()
        st_ = st.clone().requires_grad_(True)
        logp = o1["log_prob_xt_given_x1_dimwise"](xt, b1, b0, t, 0.2)                # order1.py:126
        sl = torch.mean((st_ - logp) ** 2)                                           # :131
        sl.backward()
        ld[f"{k}_score"], ld[f"{k}_score_gst"], ld[f"{k}_logp"] = sl.item(), st_.grad.numpy(), logp.numpy()
        vt_ = (vt * 100).clone().requires_grad_(True)
        x1d = b1 * 90
        ml = (torch.mean((vt_ - x1d) ** 2) + torch.mean((vt_ - (x1d + 360)) ** 2)
              + torch.mean((vt_ - (x1d - 360)) ** 2))                                 # MD_simulation.py:137-141
        ml.backward()
        ld[f"{k}_md"], ld[f"{k}_md_gvt"] = ml.item(), vt_.grad.numpy()
    np.savez(os.path.join(OUT, "loss.npz"), **ld)

    # ------------------------------------------------------------------ MMD (a13)
    np.random.seed(0)
    torch.manual_seed(0)
    true_cb = O.checkerboard(2048)
    true_8g = O.eight_gaussians(2048)
    gen = torch.tensor(td["checkerboard_o2_nfe2_g2"][-1])
    mm = {"true_checkerboard": true_cb.numpy(), "true_8gaussians": true_8g.numpy(), "gen": gen.numpy()}
    mm["mmd_cb_gen"] = o2["compute_mmd_rbf"](true_cb, gen, sigma=1.0)
    mm["mmd_cb_8g"] = o2["compute_mmd_rbf"](true_cb, true_8g, sigma=1.0)
    mm["mmd_cb_gen_f64"] = O.mmd_rbf_fp64_direct(true_cb, gen)
    mm["mmd_cb_8g_f64"] = O.mmd_rbf_fp64_direct(true_cb, true_8g)
    a = torch.randn(200, 46, generator=gg) * 60
    b = torch.randn(1000, 46, generator=gg) * 60
    mm["md_a"], mm["md_b"] = a.numpy(), b.numpy()
    mm["mmd_md"] = float(mdsym["compute_mmd"](a, b, sigma=1.0))
    mm["mmd_md_f64"] = O.mmd_rbf_fp64_direct(a, b)
    a2, b2 = a / 60, b / 60 + 0.1                        # a case where the off-diagonal terms matter
    mm["mmd_md_unit"] = float(mdsym["compute_mmd"](a2, b2, sigma=3.0))
    mm["mmd_md_unit_f64"] = O.mmd_rbf_fp64_direct(a2, b2, sigma=3.0)
    np.savez(os.path.join(OUT, "mmd.npz"), **mm)

    golden_metrics()
    golden_aux_matchers()
    golden_mlp_embedding()
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"wrote {len(os.listdir(OUT))} files, {tot / 1e6:.2f} MB -> {OUT}")


def golden_metrics():
    """TrajectoryGenerator.compute_metrics (MD_simulation.py:275-295) executed on a seeded PolyALA-shaped case."""
    import types
    fn = R.script_method(R.MD, "TrajectoryGenerator", "compute_metrics")
    g = torch.Generator().manual_seed(7)
    T, K, D = 200, 5, 46
    data = 180 * torch.sin(torch.cumsum(torch.randn(T, D, generator=g) * 0.05, 0) + torch.rand(D, generator=g) * 6.28)
    y_hat = data[:, None, :] + 10 * torch.randn(T, K, D, generator=g)
    res = fn(types.SimpleNamespace(dataset=data), y_hat.numpy())
    out = {"dataset": data.numpy(), "y_hat": y_hat.numpy()}
    out.update({k: np.float64(v) for k, v in res.items()})
    np.savez(os.path.join(OUT, "metrics.npz"), **out)


def golden_aux_matchers():
    """The matchers no IDFF configuration uses (Target :315-388, SchrodingerBridge :391-547, VariancePreserving :550-608 of
    torchcfm/conditional_flow_matching.py), executed on seeded inputs with the noise injected."""
    cf = R.cfm(False)
    g = torch.Generator().manual_seed(21)
    out = {}
    for B, D in ((64, 2), (10, 46)):
        x0, x1, eps = (torch.randn(B, D, generator=g) for _ in range(3))
        t = torch.rand(B, generator=g)
        key = f"B{B}_D{D}"
        out.update({f"{key}_x0": x0.numpy(), f"{key}_x1": x1.numpy(), f"{key}_eps": eps.numpy(), f"{key}_t": t.numpy()})
        for name, FM in (("target", cf.TargetConditionalFlowMatcher(0.1)), ("sb", cf.SchrodingerBridgeConditionalFlowMatcher(0.5)),
                         ("vp", cf.VariancePreservingConditionalFlowMatcher(0.2))):
            FM.sample_noise_like = lambda x, e=eps: e
            tt, xt, ut = FM.sample_location_and_conditional_flow(x0, x1, t=t)
            out[f"{key}_{name}_xt"], out[f"{key}_{name}_ut"] = xt.numpy(), ut.numpy()
    np.savez(os.path.join(OUT, "matchers.npz"), **out)


def golden_mlp_embedding():
    """a2: the reference's MLP_Embedding.forward (torchcfm/models/models.py:23-42) itself -- the shipped PolyALA module
    (Linear(175,128) .. Linear(128,46), Embedding(200,128)) called as the scripts call it, model(cat[y, t], ti) with one
    time index per batch (MD_simulation.py:501-504) -- and the plain MLP.forward (:4-21) of the 2-D checkpoint."""
    net = R.load_md_checkpoint()
    g = torch.Generator().manual_seed(33)
    out = {}
    y = torch.randn(33, 46, generator=g) * 60
    for ti in (0, 7, 199):
        for tv in (0.0, 0.4, 1.0):
            inp = torch.cat([y, torch.full((33, 1), tv)], -1)
            with torch.no_grad():
                out[f"emb_ti{ti}_t{tv}"] = net(inp, torch.full((33, 1), float(ti))).numpy()
    out["y"] = y.numpy()
    net2 = R.load_2d_checkpoint("checkerboard")
    x = torch.cat([torch.randn(257, 2, generator=g) * 2, torch.rand(257, 1, generator=g)], -1)
    with torch.no_grad():
        out["mlp_in"], out["mlp_out"] = x.numpy(), net2(x).numpy()
    np.savez(os.path.join(OUT, "mlp_modules.npz"), **out)


def golden_scorehead256():
    """a5 at the width the order-1 script trains (example_order1_with_prediction_head.py:101: MLP(dim=4, w=256,
    time_varying=True)): the reference's module (default init, torch.manual_seed(0)) under the script's own CustomNetModel
    (:64-96) -- field evaluations and rk4 end points, fp32 and fp64 twins.  This is the shape the tcgen05 kernel serves."""
    models = R.models()
    torch.manual_seed(0)
    net = models.MLP(dim=4, w=256, time_varying=True).eval()
    export_weights(net, os.path.join(OUT, "weights_scorehead256.npz"))
    import copy
    net64 = copy.deepcopy(net).double()
    o1 = R.script_symbols(R.ORDER1, ["CustomNetModel", "CFMModel"])
    x = x0_2d(256, seed=1) * 3.0
    x0 = x0_2d(1024, seed=0)
    sd = {"x": x.numpy(), "x0": x0.numpy(), "times": np.array(FIELD_TIMES, dtype=np.float32), "gammas": np.array(GAMMAS2[:4])}
    ts = torch.linspace(0, 1, 6)
    for gi, g in enumerate(GAMMAS2[:4]):
        m = o1["CustomNetModel"](net, sigma=0.2, gamma1=g[0], gamma2=g[1])
        m64 = o1["CustomNetModel"](net64, sigma=0.2, gamma1=g[0], gamma2=g[1])
        sd[f"field_g{gi}"] = np.stack([run_field(m, torch.tensor(t, dtype=torch.float32), x).numpy() for t in FIELD_TIMES])
        sd[f"field_g{gi}_f64"] = np.stack([run_field(m64, torch.tensor(t, dtype=torch.float32).double(), x.double()).numpy()
                                           for t in FIELD_TIMES])
        sd[f"traj_nfe5_g{gi}"] = O.odeint_rk4(m, x0.clone(), ts).detach()[-1].numpy()
        sd[f"traj_nfe5_g{gi}_f64"] = O.odeint_rk4(m64, x0.double(), ts.double()).detach()[-1].numpy()
    # the script's CFM comparator on the same network (:335-356)
    sd["cfm_nfe5"] = O.odeint_rk4(o1["CFMModel"](net, sigma=0, gamma1=0, gamma2=0), x0.clone(), ts).detach()[-1].numpy()
    sd["cfm_nfe5_f64"] = O.odeint_rk4(o1["CFMModel"](net64, sigma=0, gamma1=0, gamma2=0), x0.double(), ts.double()).detach()[-1].numpy()
    np.savez(os.path.join(OUT, "scorehead256.npz"), **sd)
    print("wrote scorehead256.npz, weights_scorehead256.npz")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "scorehead256":
        golden_scorehead256()
    elif len(sys.argv) > 1 and sys.argv[1] == "metrics":
        golden_metrics()
    elif len(sys.argv) > 1 and sys.argv[1] == "mlp":
        golden_mlp_embedding()
    elif len(sys.argv) > 1 and sys.argv[1] == "matchers":
        golden_aux_matchers()
    else:
        main()
