"""Oracle (oracle/idff_oracle.py) vs the committed golden vectors, which were produced by executing
the reference's own code objects (oracle/make_golden.py).  Bit-exact unless stated."""
import numpy as np
import pytest
import torch

import idff_oracle as O
from conftest import golden_layers, load_golden
from make_golden_consts import FIELD_TIMES, GAMMAS2, GAMMAS3

torch.set_num_threads(4)


@pytest.mark.parametrize("which", ["checkerboard", "8gaussians"])
def test_field2d_matches_golden(which):
    g = load_golden("field2d")
    layers, _ = golden_layers(which)
    x = torch.tensor(g["x"])
    for gi, gam in enumerate(GAMMAS2):
        f = O.Field2D(layers, 0.2, gam, 2)
        out = np.stack([f(torch.tensor(t, dtype=torch.float32), x).numpy() for t in FIELD_TIMES])
        np.testing.assert_array_equal(out, g[f"{which}_o2_g{gi}"])
    for gi, gam in enumerate(GAMMAS3):
        f = O.Field2D(layers, 0.2, gam, 3)
        out = np.stack([f(torch.tensor(t, dtype=torch.float32), x).numpy() for t in FIELD_TIMES])
        np.testing.assert_array_equal(out, g[f"{which}_o3_g{gi}"])
    mlp = O.mlp_forward(layers, torch.cat([x, torch.full((x.shape[0], 1), 0.37)], -1)).numpy()
    np.testing.assert_array_equal(mlp, g[f"{which}_mlp"])


def test_traj2d_matches_golden():
    g = load_golden("traj2d")
    layers, _ = golden_layers("checkerboard")
    x0 = torch.tensor(g["x0"])
    f = O.Field2D(layers, 0.2, GAMMAS2[2], 2)
    tr = O.odeint_rk4(f, x0, torch.linspace(0, 1, 3)).numpy()
    np.testing.assert_array_equal(tr, g["checkerboard_o2_nfe2_g2"])
    f3 = O.Field2D(layers, 0.2, GAMMAS3[1], 3)
    tr3 = O.odeint_rk4(f3, x0, torch.linspace(0, 1, 3)).numpy()[-1]
    np.testing.assert_array_equal(tr3, g["checkerboard_o3_nfe2_g1"])
    cfm = O.FieldCFM(layers)
    np.testing.assert_array_equal(O.odeint_rk4(cfm, x0, torch.linspace(0, 1, 6)).numpy()[-1],
                                  g["checkerboard_cfm_nfe5"])


@pytest.mark.parametrize("which", ["scorehead", "scorehead256"])
def test_scorehead_matches_golden(which):
    """w = 64 and the order-1 script's own width 256 (example_order1_with_prediction_head.py:101)."""
    g = load_golden(which)
    layers, _ = golden_layers(which)
    x = torch.tensor(g["x"])
    for gi, gam in enumerate(g["gammas"]):
        f = O.FieldScoreHead(layers, 0.2, tuple(gam))
        out = np.stack([f(torch.tensor(t, dtype=torch.float32), x).numpy() for t in FIELD_TIMES])
        np.testing.assert_array_equal(out, g[f"field_g{gi}"])
    if which == "scorehead256":
        f = O.FieldScoreHead(layers, 0.2, tuple(g["gammas"][2]))
        tr = O.odeint_rk4(f, torch.tensor(g["x0"]), torch.linspace(0, 1, 6)).numpy()[-1]
        np.testing.assert_array_equal(tr, g["traj_nfe5_g2"])
        cfm = O.FieldCFM(layers, score_head=True)          # the order-1 script's CFM comparator on the same network
        np.testing.assert_array_equal(O.odeint_rk4(cfm, torch.tensor(g["x0"]), torch.linspace(0, 1, 6)).numpy()[-1], g["cfm_nfe5"])


def test_md_field_matches_golden():
    g = load_golden("md")
    layers, emb = golden_layers("polyala")
    y = torch.tensor(g["y"])
    for ii in g["init_inds"]:
        sde = O.FieldMD(layers, emb, 46, sigma=0.5, init_ind=int(ii), dt=0.3, gamma1=0.2, gamma2=5.0 / (1 + int(ii)))
        f = np.stack([sde.f(torch.tensor(t), y).numpy() for t in (0.0, 0.3, 0.6, 1.0)])
        gg = np.stack([sde.g(torch.tensor(t), y).numpy() for t in (0.0, 0.3, 0.6, 1.0)])
        np.testing.assert_array_equal(f, g[f"f_i{ii}"])
        np.testing.assert_array_equal(gg, g[f"g_i{ii}"])
    sde = O.FieldMD(layers, emb, 46, sigma=0.5, init_ind=3, dt=0.3, gamma1=0.2, gamma2=5.0 / 4)
    srk = O.sdeint_srk(sde, y, torch.tensor(g["ts"]), 0.3, torch.tensor(g["I_k"]), torch.tensor(g["I_k0"]))
    np.testing.assert_array_equal(srk.numpy(), g["srk"])
    eul = O.sdeint_euler(sde, y, torch.tensor(g["ts"]), 0.3, torch.tensor(g["I_k"]))
    np.testing.assert_array_equal(eul.numpy(), g["euler"])


def test_path_sample_matches_golden():
    g = load_golden("path")
    keys = sorted({k.rsplit("_", 1)[0] for k in g.files if k.endswith("_xt")})
    assert len(keys) >= 12
    for k in keys:
        dyn = k.startswith("dyn")
        exact = "_Exact_" in k
        sigma = float(k.rsplit("_s", 1)[1])
        x0, x1, t, eps = (torch.tensor(g[f"{k}_{n}"]) for n in ("x0", "x1", "t", "eps"))
        xt, ut = O.path_sample(x0, x1, t, eps, sigma, exact)
        np.testing.assert_array_equal(xt.numpy(), g[k + "_xt"])
        np.testing.assert_array_equal(ut.numpy(), g[k + "_ut"])
        # draw order: t first, then eps, both from the global CPU generator (seed 5)
        torch.manual_seed(5)
        t2, xt2, ut2, eps2 = O.sample_location_and_conditional_flow(x0, x1, sigma, exact, return_noise=True)
        np.testing.assert_array_equal(t2.numpy(), g[k + "_t"])
        np.testing.assert_array_equal(eps2.numpy(), g[k + "_eps"])
        np.testing.assert_array_equal(xt2.numpy(), g[k + "_xt"])
        lam = O.compute_lambda(t, sigma, exact, dyn)
        np.testing.assert_array_equal(np.asarray(lam, dtype=np.float64), np.asarray(g[k + "_lambda"], dtype=np.float64))


def test_losses_match_golden():
    g = load_golden("loss")
    for k in ("B256_D2", "B10_D46", "B1000_D2"):
        vt, st, x0, x1, t, xt = (torch.tensor(g[f"{k}_{n}"]) for n in ("vt", "st", "x0", "x1", "t", "xt"))
        assert O.flow_loss(vt, x1, t).item() == pytest.approx(float(g[f"{k}_flow"]), rel=1e-6)
        assert O.score_loss(st, xt, x1, x0, t, 0.2).item() == pytest.approx(float(g[f"{k}_score"]), rel=1e-6)
        np.testing.assert_array_equal(O.log_prob_xt_given_x1_dimwise(xt, x1, x0, t, 0.2).numpy(), g[f"{k}_logp"])
        assert O.md_periodic_loss(vt * 100, x1 * 90).item() == pytest.approx(float(g[f"{k}_md"]), rel=1e-6)


def test_mmd_matches_golden():
    g = load_golden("mmd")
    cb, e8, gen = (torch.tensor(g[n]) for n in ("true_checkerboard", "true_8gaussians", "gen"))
    # (means are summed per thread block: equality up to the thread count used)
    assert O.mmd_rbf(cb, gen) == pytest.approx(float(g["mmd_cb_gen"]), rel=2e-6)
    assert O.mmd_rbf(cb, e8) == pytest.approx(float(g["mmd_cb_8g"]), rel=2e-6)
    assert O.mmd_rbf(torch.tensor(g["md_a"]), torch.tensor(g["md_b"])) == pytest.approx(float(g["mmd_md"]), abs=1e-7)
    # fp32 reference vs the fp64 direct yardstick (this is the slack the reference itself has)
    assert abs(g["mmd_cb_gen"] - g["mmd_cb_gen_f64"]) < 2e-6
    sx, sy, sxy = O.mmd_partial_sums_fp64(cb, gen)
    n, m = cb.shape[0], gen.shape[0]
    assert sx / n / n + sy / m / m - 2 * sxy / n / m == pytest.approx(float(g["mmd_cb_gen_f64"]), abs=1e-12)
    # toy samplers are seeded fixtures: np.random.seed(0); torch.manual_seed(0)
    np.random.seed(0)
    torch.manual_seed(0)
    np.testing.assert_array_equal(O.checkerboard(2048).numpy(), g["true_checkerboard"])
    np.testing.assert_array_equal(O.eight_gaussians(2048).numpy(), g["true_8gaussians"])


def test_jet_formulation_equals_nested_autograd():
    """The kernel's algorithm (n forward Taylor-jet passes + one reverse sweep) == the reference's nested
    autograd.grad, to fp64 round-off."""
    import jet_model as J
    layers, _ = golden_layers("checkerboard", torch.float64)
    x = torch.tensor(load_golden("field2d")["x"]).double()
    for order, gam in [(2, (0.2, 0.3)), (3, (0.2, 0.3, 0.4)), (3, (1.0, 2.0, 0.5))]:
        f = O.Field2D(layers, 0.2, gam, order)
        for tv in (0.1, 0.37, 5.0 / 6.0):
            t = torch.tensor(tv, dtype=torch.float64)
            st2 = (0.2 * torch.sqrt(t * (1 - t))) ** 2
            tt = st2 / 0.2 ** 2
            coef = [0.5 * (2 * gam[0] - 1) * tt] + [g_ * tt for g_ in gam[1:]]
            pred, gsum = J.momentum_jets(layers, x, t, coef, 2)
            out = torch.sqrt(1 - sum(gam) * st2) * (pred - x) / (1 - t) + gsum
            ref = f(t, x)
            assert (out - ref).abs().max().item() < 1e-11 * max(1.0, ref.abs().max().item())


def test_md_metrics_match_golden():
    """TrajectoryGenerator.compute_metrics (MD_simulation.py:275-295): the oracle restatement equals the values the
    reference's own method produced (tests/golden/metrics.npz, written by oracle/make_golden.py metrics)."""
    g = load_golden("metrics")
    r = O.md_compute_metrics(g["y_hat"], g["dataset"])
    for k in ("CC_mean", "CC_std", "RMSE_mean", "RMSE_std", "MAE_mean", "MAE_std"):
        assert float(r[k]) == float(g[k]), k


def test_mlp_modules_match_golden():
    """a1/a2: the oracle's MLP / MLP_Embedding restatement against outputs of the reference's own modules
    (tests/golden/mlp_modules.npz, written by oracle/make_golden.py golden_mlp_embedding)."""
    g = load_golden("mlp_modules")
    layers, emb = golden_layers("polyala")
    y = torch.tensor(g["y"])
    for ti in (0, 7, 199):
        for tv in (0.0, 0.4, 1.0):
            inp = torch.cat([y, torch.full((33, 1), tv)], -1)
            out = O.mlp_embedding_forward(layers, emb, inp, torch.full((33, 1), float(ti)))
            assert torch.equal(out, torch.tensor(g[f"emb_ti{ti}_t{tv}"]))
    l2, _ = golden_layers("checkerboard")
    assert torch.equal(O.mlp_forward(l2, torch.tensor(g["mlp_in"])), torch.tensor(g["mlp_out"]))
