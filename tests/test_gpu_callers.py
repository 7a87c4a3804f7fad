"""-m gpu: the callers either side of the sampler: gamma sweep by MMD and the PolyALA
autoregressive generator, against the oracle run the way the reference scripts run them."""
import numpy as np
import pytest
import torch

import idff_oracle as O
from conftest import golden_layers, load_golden
from gpu_common import assert_field_close, packed
from idff_b200 import _lib, md, sweep

pytestmark = pytest.mark.gpu


def test_gamma_sweep_matches_oracle_loop(cuda_device):
    """find_best_gammas_mmd (Autograd_example_order2.py:281-303): same tuple order, same MMD values, same winner."""
    layers, _ = golden_layers("checkerboard")
    g = load_golden("mmd")
    true_s = torch.tensor(g["true_checkerboard"])
    x0 = torch.tensor(load_golden("traj2d")["x0"])
    grid = [0.0, 0.2, 1.0]
    best, results = sweep.find_best_gammas_mmd(packed("checkerboard", cuda_device), x0.to(cuda_device),
                                               true_s.to(cuda_device), gamma_range=grid, nfe=2, sigma=0.2,
                                               sigma_kernel=1.0, order=2)
    assert [r[0] for r in results] == [(a, b) for a in grid for b in grid]
    ref = []
    for (g1, g2) in [r[0] for r in results]:
        fin = O.odeint_rk4(O.Field2D(layers, 0.2, (g1, g2), 2), x0, torch.linspace(0, 1, 3))[-1]
        ref.append(O.mmd_rbf_fp64_direct(true_s, fin, 1.0))
    got = np.array([r[1] for r in results])
    np.testing.assert_allclose(got, np.array(ref), rtol=2e-3, atol=2e-6)     # MMD of slightly drifted samples
    assert best == results[int(np.argmin(got))][0]
    assert abs(got[int(np.argmin(ref))] - got.min()) < 1e-5
    # order-3 tuples go through the order-3 field
    _, r3 = sweep.find_best_gammas_mmd(packed("checkerboard", cuda_device), x0.to(cuda_device), true_s.to(cuda_device),
                                       gamma_range=[0.0, 0.2], nfe=2, order=3)
    assert len(r3) == 8 and all(np.isfinite(v) for _, v in r3)
    assert r3[0][1] == pytest.approx(results[0][1], abs=1e-6)              # (0,0,0) == (0,0)


def test_compute_mmd_for_gamma_configs(cuda_device):
    """Autograd_example_order3.py:376-425: explicit (gamma1, gamma2, gamma3) list and t_span; equals per-configuration solves."""
    from idff_b200 import fields, sampler
    from idff_b200.mmd import compute_mmd_rbf
    pw = packed("checkerboard", cuda_device)
    g = torch.Generator(device=cuda_device).manual_seed(3)
    x0 = torch.randn(2048, 2, device=cuda_device, generator=g) / 1.5
    true_s = torch.tensor(load_golden("mmd")["true_checkerboard"]).to(cuda_device)
    cfgs = [(0.0, 0.0, 0.0), (0.2, 0.2, 0.1), (1.0, 0.5, 0.0), (0.3, 0.0, 0.4)]
    t_span = torch.linspace(0, 1, 4)
    res = sweep.compute_mmd_for_gamma_configs(pw, cfgs, x0, t_span, true_s, sigma_model=0.2, sigma_kernel=1.0)
    assert [c for c, _ in res] == cfgs
    for c, v in res:
        out = sampler.sample_rk4(fields.CustomNetModelOrder3(pw, 0.2, *c), x0, t_span)
        assert v == pytest.approx(compute_mmd_rbf(true_s, out, 1.0), abs=2e-6)


def test_md_autoregressive_generation_vs_oracle(cuda_device):
    """generate_trajectories (MD_simulation.py:197-213) for 6 frames, 5 trajectories, injected Brownian increments."""
    layers, emb = golden_layers("polyala")
    K, D, T, NFE, dt = 5, 46, 6, 3, 0.3
    gen = torch.Generator().manual_seed(0)
    x_init = torch.randn(K, D, generator=gen)
    ts = torch.linspace(0, 1, NFE)
    grid = O.sde_step_grid(ts, dt)
    h = torch.tensor(np.diff(np.array(grid, dtype=np.float32)))
    ns = len(h)
    noises = []
    for _ in range(T):
        ik = torch.randn(ns, K, D, generator=gen) * h.sqrt()[:, None, None]
        ik0 = 0.5 * h[:, None, None] * ik + torch.randn(ns, K, D, generator=gen) * (h ** 1.5)[:, None, None] / np.sqrt(12.0)
        noises.append((ik, ik0))
    # oracle: the reference's loop, frame by frame
    ref = []
    x1 = x_init
    for t_idx in range(T):
        sde = O.FieldMD(layers, emb, D, sigma=0.5, init_ind=t_idx, dt=dt, gamma1=0.2, gamma2=5.0 / (1 + t_idx))
        x1 = O.sdeint_srk(sde, x1, ts, dt, noises[t_idx][0], noises[t_idx][1])[-1]
        ref.append(x1)
    ref = torch.stack(ref).numpy()
    for chain in (True, False):     # the one-launch chain kernel (idff_md_generate) and the launch-per-frame path
        before = _lib.launch_count()
        got = md.generate_trajectories(packed("polyala", cuda_device), K, D, T, NFE, sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt,
                                       device=cuda_device, x_init=x_init.to(cuda_device), chain=chain,
                                       noise=lambda t_idx, n: (noises[t_idx][0].to(cuda_device), noises[t_idx][1].to(cuda_device)))
        assert _lib.launch_count() - before == (1 if chain else T)
        assert got.shape == (T, K, D)
        for t_idx in range(T):
            # each frame compounds the previous frames' round-off: 1e-4 of the frame's scale
            assert_field_close(got[t_idx].cpu().numpy(), ref[t_idx], 1e-4, f"frame {t_idx} chain={chain}")
    m = md.compute_metrics(got, torch.tensor(ref.mean(1)))          # MD_simulation.py:275-295 against a stand-in data set
    mo = O.md_compute_metrics(ref, ref.mean(1))
    for k in mo:
        assert m[k] == pytest.approx(float(mo[k]), rel=1e-3, abs=1e-3), k
    one = md.generate_onestep(packed("polyala", cuda_device), K, D, T, 3, 0.5, x_init.to(cuda_device), gamma1=0.2, gamma2=5.0)
    assert one.shape == (3, K, D) and torch.isfinite(one).all()


def test_batched_sweep_equals_per_field_solves(cuda_device):
    """idff_sample_rk4_sweep: one launch for many gamma tuples (mixed: with / without reverse sweep, order-1-like sets)
    returns bit-for-bit what one idff_sample_rk4 call per tuple returns, for ragged particle counts and odd tile counts."""
    import time

    from idff_b200 import _lib, fields, sampler
    pw = packed("checkerboard", cuda_device)
    tuples = [(0.0, 0.0), (0.5, 0.0), (0.2, 0.0), (0.2, 0.2), (1.0, 2.0), (0.5, 0.3), (-0.5, 0.3)]
    models = [fields.CustomNetModel(pw, 0.2, *g) for g in tuples]
    for n in (2048 + 17, 64, 1, 4097):
        x0 = torch.randn(n, 2, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(n)) / 1.5
        for nfe in (2, 5):
            ts = torch.linspace(0, 1, nfe + 1)
            before = _lib.launch_count()
            got = sampler.sample_rk4_sweep(models, x0, ts)
            assert _lib.launch_count() == before + 2                       # ONE sampler launch for the whole sweep + the AUTO route's range safety net
            assert got.shape == (len(tuples), n, 2)
            for k, m in enumerate(models):
                assert torch.equal(got[k], sampler.sample_rk4(m, x0, ts)), (n, nfe, tuples[k])
    # sets the tensor kernel does not serve (order 3 with gamma3 != 0) fall back to one launch per set, same results
    m3 = [fields.CustomNetModelOrder3(pw, 0.2, 0.2, 0.2, g3) for g3 in (0.0, 0.2)]
    x0 = torch.randn(512, 2, device=cuda_device) / 1.5
    got = sampler.sample_rk4_sweep(m3, x0, torch.linspace(0, 1, 3))
    for k, m in enumerate(m3):
        assert torch.equal(got[k], sampler.sample_rk4(m, x0, torch.linspace(0, 1, 3)))
    # the reference's sweep shape: 8 x 8 tuples, 2048 particles, nfe = 2 -- batched vs launch-per-tuple timing (reported)
    grid8 = [0.0, 0.1, 0.2, 0.3, 0.5, 0.7, 1.0, 2.0]
    ms = [fields.CustomNetModel(pw, 0.2, a, b) for a in grid8 for b in grid8]
    x0 = torch.randn(2048, 2, device=cuda_device) / 1.5
    ts = torch.linspace(0, 1, 3)
    sampler.sample_rk4_sweep(ms, x0, ts)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sampler.sample_rk4_sweep(ms, x0, ts)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for m in ms:
        sampler.sample_rk4(m, x0, ts)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"64-tuple sweep, 2048 particles, nfe 2: batched {1e3 * (t1 - t0):.2f} ms, per tuple {1e3 * (t2 - t1):.2f} ms")
    assert (t1 - t0) < (t2 - t1)
    # the whole caller (solves + MMD of every tuple + argmin), as the reference script runs it
    true_s = torch.tensor(load_golden("mmd")["true_checkerboard"]).to(cuda_device)
    sweep.find_best_gammas_mmd(pw, x0, true_s, gamma_range=grid8, nfe=2, sigma=0.2, sigma_kernel=1.0, order=2)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    best, res = sweep.find_best_gammas_mmd(pw, x0, true_s, gamma_range=grid8, nfe=2, sigma=0.2, sigma_kernel=1.0, order=2)
    t4 = time.perf_counter()
    print(f"find_best_gammas_mmd, 8 x 8 tuples: {1e3 * (t4 - t3):.2f} ms end to end (2 launches), best {best}")
    assert len(res) == 64 and all(np.isfinite(v) for _, v in res)


def test_md_generation_as_cuda_graph(cuda_device):
    """The T-frame autoregressive chain captured as one CUDA graph replays bit-for-bit what the eager loop computes
    (injected Brownian increments), and draws fresh noise on every replay when none is injected."""
    import time
    K, D, T, NFE, dt = 5, 46, 6, 3, 0.3
    pw = packed("polyala", cuda_device)
    ts = torch.linspace(0, 1, NFE)
    from idff_b200 import sampler
    ns = sampler.sde_num_steps(ts, dt)
    g = torch.Generator(device=cuda_device).manual_seed(3)
    ik = torch.randn(T, ns, K, D, device=cuda_device, generator=g) * 0.5
    ik0 = torch.randn(T, ns, K, D, device=cuda_device, generator=g) * 0.1
    x_init = torch.randn(K, D, device=cuda_device, generator=g)
    eager = md.generate_trajectories(pw, K, D, T, NFE, sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt, device=cuda_device,
                                     x_init=x_init, chain=False, noise=lambda t_idx, n: (ik[t_idx], ik0[t_idx]))
    gen = md.GraphedGenerator(pw, K, D, T, NFE, sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt, device=cuda_device, noise=(ik, ik0))
    assert torch.isfinite(eager).all()
    assert torch.equal(gen.generate(x_init), eager)
    assert torch.equal(gen.generate(x_init * 0.5), md.generate_trajectories(
        pw, K, D, T, NFE, sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt, device=cuda_device, x_init=x_init * 0.5, chain=False,
        noise=lambda t_idx, n: (ik[t_idx], ik0[t_idx])))
    fresh = md.GraphedGenerator(pw, K, D, T, NFE, sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt, device=cuda_device)
    a = fresh.generate(x_init).clone()
    b = fresh.generate(x_init).clone()
    assert torch.isfinite(a).all() and torch.isfinite(b).all() and not torch.equal(a, b)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fresh.generate(x_init)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(5):
        md.generate_trajectories(pw, K, D, T, NFE, sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt, device=cuda_device, x_init=x_init,
                                 chain=False)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"PolyALA chain, {T} frames x {K} trajectories: graph {1e3 * (t1 - t0) / 5:.2f} ms, eager {1e3 * (t2 - t1) / 5:.2f} ms")


def test_md_chain_kernel_full_length(cuda_device):
    """The reference's generation shape (T = 200 frames, 5 trajectories, NFE 3, dt 0.3) in ONE launch against the launch-per-frame
    path on the same injected noise: the two differ only in summation order, and the frame-to-frame map is contractive enough
    that the difference stays at round-off level over the whole chain.  Also K = 9 (three CTAs, ragged) and euler."""
    import time

    from idff_b200 import sampler as sampler_mod
    pw = packed("polyala", cuda_device)
    D, T, NFE, dt = 46, 200, 3, 0.3
    g = torch.Generator(device=cuda_device).manual_seed(11)
    n_steps = 4
    for K, method in ((5, "srk"), (9, "srk"), (5, "euler")):
        x_init = torch.randn(K, D, device=cuda_device, generator=g)        # MD_simulation.py:203
        ik = torch.randn(T, n_steps, K, D, device=cuda_device, generator=g) * 0.5
        ik0 = torch.randn(T, n_steps, K, D, device=cuda_device, generator=g) * 0.05
        kw = dict(sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt, device=cuda_device, x_init=x_init, method=method,
                  noise=lambda t_idx, n: (ik[t_idx], ik0[t_idx]))
        before = _lib.launch_count()
        one = md.generate_trajectories(pw, K, D, T, NFE, chain=True, **kw)
        assert _lib.launch_count() - before == 1
        per = md.generate_trajectories(pw, K, D, T, NFE, chain=False, **kw)
        assert torch.isfinite(one).all()
        d = (one - per).abs()
        scale = per.abs().max().item()
        print(f"K={K} {method}: chain vs per-frame over {T} frames: median|d| {d.median().item():.2e} max {d.max().item():.2e} (scale {scale:.1f})")
        assert d.median().item() <= 1e-4 * scale and d.max().item() <= 1e-2 * scale
    # many CTAs (K = 301: 76 CTAs, the last one ragged), short chain, drawn noise of the right law
    K, Ts = 301, 8
    x_init = torch.randn(K, D, device=cuda_device, generator=g)
    ik, ik0 = sampler_mod.brownian_increments(n_steps, torch.linspace(0, 1, NFE), dt, (Ts, K, D), cuda_device, g)
    ik, ik0 = ik.transpose(0, 1).contiguous(), ik0.transpose(0, 1).contiguous()
    kw = dict(sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt, device=cuda_device, x_init=x_init,
              noise=lambda t_idx, n: (ik[t_idx], ik0[t_idx]))
    one = md.generate_trajectories(pw, K, D, Ts, NFE, chain=True, **kw)
    per = md.generate_trajectories(pw, K, D, Ts, NFE, chain=False, **kw)
    assert torch.isfinite(one).all() and (one - per).abs().max().item() <= 1e-4 * per.abs().max().item()
    for _ in range(3):                     # fixed reduction orders: bit-reproducible run to run
        assert torch.equal(md.generate_trajectories(pw, K, D, Ts, NFE, chain=True, **kw), one)
    x_init = torch.randn(5, D, device=cuda_device, generator=g)
    md.generate_trajectories(pw, 5, D, T, NFE, sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt, device=cuda_device, x_init=x_init)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        md.generate_trajectories(pw, 5, D, T, NFE, sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt, device=cuda_device, x_init=x_init)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    md.generate_trajectories(pw, 5, D, T, NFE, sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt, device=cuda_device, x_init=x_init, chain=False)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"PolyALA generation, 200 frames x 5 trajectories: chain kernel {1e3 * (t1 - t0) / 3:.1f} ms (incl. noise draws), "
          f"launch per frame {1e3 * (t2 - t1):.1f} ms")
    # more trajectories than one wave of CTAs (2 x 148 x 4 + 3): the CTAs loop over groups of 4
    K2 = 2 * 592 + 3
    x2 = torch.randn(K2, D, device=cuda_device, generator=g)
    ik, ik0 = sampler_mod.brownian_increments(n_steps, torch.linspace(0, 1, NFE), dt, (3, K2, D), cuda_device, g)
    nz = (ik.transpose(0, 1).contiguous(), ik0.transpose(0, 1).contiguous())
    kw2 = dict(sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt, device=cuda_device, x_init=x2, noise=lambda t_idx, n: (nz[0][t_idx], nz[1][t_idx]))
    a = md.generate_trajectories(pw, K2, D, 3, NFE, chain=True, **kw2)
    b = md.generate_trajectories(pw, K2, D, 3, NFE, chain=False, **kw2)
    assert torch.isfinite(a).all() and (a - b).abs().max().item() <= 1e-4 * b.abs().max().item()


def test_md_compute_metrics_matches_reference_golden(cuda_device):
    """compute_metrics (MD_simulation.py:275-295) on the device against the values the reference's own method produced:
    CC in fp64 (np.corrcoef promotes), RMSE / MAE within fp32 rounding of the reference's fp32 nanmean."""
    from conftest import load_golden
    from idff_b200 import md
    g = load_golden("metrics")
    m = md.compute_metrics(torch.tensor(g["y_hat"]).to(cuda_device), torch.tensor(g["dataset"]))
    for k in ("CC_mean", "CC_std"):
        assert m[k] == pytest.approx(float(g[k]), rel=1e-6), k
    for k in ("RMSE_mean", "RMSE_std", "MAE_mean", "MAE_std"):
        assert m[k] == pytest.approx(float(g[k]), rel=1e-5), k
    y = torch.tensor(g["y_hat"]).to(cuda_device)
    y[3, 1, 2] = float("nan")
    m2 = md.compute_metrics(y, torch.tensor(g["dataset"]))
    assert np.isnan(m2["CC_mean"]) and np.isfinite(m2["RMSE_mean"]) and np.isfinite(m2["MAE_mean"])
    with pytest.raises(RuntimeError):
        md.compute_metrics(torch.tensor(g["y_hat"]), torch.tensor(g["dataset"]))


@pytest.mark.parametrize("impl", [0, 4])
def test_order3_sweep_batched(cuda_device, impl):
    """The order-3 gamma sweep (Autograd_example_order3.py:288-314).  Default route (impl 0): sets with gamma3 != 0 go out as
    ONE launch of the fused three-jet tcgen05 kernel (k_tc16o3), per-set stage tables in global memory.  impl 4 (layered
    engine, the route of wide networks): at N >= 2048 (multiple of 128) they are solved as ONE batch of n_sets * N particles,
    per-set coefficients looked up by particle row.  Either way bit-identical to one idff_sample_rk4 call per tuple."""
    import time

    from idff_b200 import _lib, fields, sampler
    pw = packed("checkerboard", cuda_device)
    tuples = [(0.2, 0.2, 0.2), (0.0, 0.0, 0.1), (1.0, 2.0, 0.5), (0.3, 0.1, -0.2), (0.2, 0.0, 0.7)]
    models = [fields.CustomNetModelOrder3(pw, 0.2, *g) for g in tuples]
    for m in models:
        m.impl = impl
    n = 2048
    x0 = torch.randn(n, 2, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(5)) / 1.5
    for nfe in (2, 3):
        ts = torch.linspace(0, 1, nfe + 1)
        per = [sampler.sample_rk4(m, x0, ts) for m in models]
        before = _lib.launch_count()
        one = sampler.sample_rk4(models[0], x0, ts)
        per_set_launches = _lib.launch_count() - before
        before = _lib.launch_count()
        got = sampler.sample_rk4_sweep(models, x0, ts)
        assert _lib.launch_count() - before <= per_set_launches + 2          # the whole sweep costs one solve's launches (+ replicate + safety net)
        assert torch.equal(one, per[0])
        for k in range(len(tuples)):
            assert torch.equal(got[k], per[k]), (nfe, tuples[k])
    # a mixed sweep (the reference's order-3 grid contains gamma3 == 0 tuples, which the fused tensor kernel serves): the sets
    # are split into classes, each class solved as one batch and scattered to its slots -- still bit-identical per tuple
    mixed = [(0.2, 0.2, 0.0), (0.2, 0.2, 0.2), (0.5, 0.0, 0.0), (0.0, 0.0, 0.1), (1.0, 2.0, 0.0), (0.3, 0.1, -0.2), (0.5, 0.0, 0.3)]
    mm = [fields.CustomNetModelOrder3(pw, 0.2, *g) for g in mixed]
    for m in mm:
        m.impl = impl
    ts = torch.linspace(0, 1, 3)
    per = [sampler.sample_rk4(m, x0, ts) for m in mm]
    before = _lib.launch_count()
    got = sampler.sample_rk4_sweep(mm, x0, ts)
    # (the layered engine takes ~8 launches per evaluation and class; one solve per tuple would be len(mm) times that)
    assert _lib.launch_count() - before < len(mm) * (20 if impl == 0 else 40)
    for k in range(len(mixed)):
        assert torch.equal(got[k], per[k]), mixed[k]
    # ragged N falls back to per-set solves, same results
    x1 = x0[:2048 - 64 + 3].contiguous()
    got = sampler.sample_rk4_sweep(models[:2], x1, torch.linspace(0, 1, 3))
    for k in range(2):
        assert torch.equal(got[k], sampler.sample_rk4(models[k], x1, torch.linspace(0, 1, 3)))
    # timing, reported: 5 x 5 x 5 tuples (the reference sweeps 10^3), 2048 particles, nfe 2
    g5 = [0.0, 0.2, 0.5, 1.0, 2.0]
    ms = [fields.CustomNetModelOrder3(pw, 0.2, a, b, c if c else 0.05) for a in g5 for b in g5 for c in g5]
    for m in ms:
        m.impl = impl
    ts = torch.linspace(0, 1, 3)
    sampler.sample_rk4_sweep(ms, x0, ts)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sampler.sample_rk4_sweep(ms, x0, ts)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for m in ms:
        sampler.sample_rk4(m, x0, ts)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"125-tuple order-3 sweep, 2048 particles, nfe 2: batched {1e3 * (t1 - t0):.2f} ms, per tuple {1e3 * (t2 - t1):.2f} ms")
    assert (t1 - t0) < (t2 - t1)


def test_new_entry_points_edge_cases(cuda_device):
    """Empty and degenerate inputs of the round's later entry points: zero steps / rows are no-ops, bad shapes raise."""
    import ctypes as C

    from idff_b200 import train
    from idff_b200.torchcfm import optimal_transport as ot
    from idff_b200.torchcfm.models.models import MLP
    dev = cuda_device
    L = _lib.lib()
    # idff_ot_assign: n = 0 is a no-op; D mismatch and unequal batches raise
    assert ot.ot_assign(torch.zeros(0, 2, device=dev), torch.zeros(0, 2, device=dev)).numel() == 0
    with pytest.raises(_lib.IdffError):
        ot.ot_assign(torch.zeros(4, 2, device=dev), torch.zeros(5, 2, device=dev))
    with pytest.raises(_lib.IdffError):
        ot.ot_assign(torch.zeros(4, 2, device=dev), torch.zeros(4, 3, device=dev))
    # idff_train_flow_steps: zero steps leave the state untouched
    tr = train.FusedFlowTrainer(MLP(dim=2, w=256, time_varying=True).to(dev), batch_size=8)
    before = tr.state.clone()
    out = tr.steps(torch.zeros(0, 8, 2, device=dev), torch.zeros(0, 8, 2, device=dev), torch.zeros(0, 8, device=dev))
    assert out.numel() == 0 and torch.equal(before, tr.state) and tr.steps_done == 0
    # idff_md_generate: K = 0 is a no-op; more frames than embedding rows is an error
    pw = packed("polyala", dev)
    y = md.generate_trajectories_chain(pw, 0, 46, 4, 3, sigma=0.5, dt=0.3, device=dev, x_init=torch.zeros(0, 46, device=dev),
                                       noise=(torch.zeros(4, 4, 0, 46, device=dev), torch.zeros(4, 4, 0, 46, device=dev)))
    assert tuple(y.shape) == (4, 0, 46)
    with pytest.raises(_lib.IdffError):
        md.generate_trajectories_chain(pw, 2, 46, 201, 3, sigma=0.5, dt=0.3, device=dev)
    # the sweep entry point with a single set and with N = 0
    from idff_b200 import fields, sampler
    m3 = fields.CustomNetModelOrder3(packed("checkerboard", dev), 0.2, 0.2, 0.2, 0.2)
    x0 = torch.randn(2048, 2, device=dev)
    ts = torch.linspace(0, 1, 3)
    assert torch.equal(sampler.sample_rk4_sweep([m3], x0, ts)[0], sampler.sample_rk4(m3, x0, ts))
    assert sampler.sample_rk4_sweep([m3, m3], x0[:0], ts).shape == (2, 0, 2)


def test_md_chain_fma_and_mma_evaluators_agree(cuda_device):
    """idff_md_generate evaluates the layers on mma.sync (3 x tf32) by default; the fp32-FMA evaluator behind
    idff_debug_md_chain_fma is the yardstick: the two stay within 1e-5 of the frame's scale over a 40-frame chain."""
    pw = packed("polyala", cuda_device)
    K, D, T, NFE, dt = 7, 46, 40, 3, 0.3
    g = torch.Generator(device=cuda_device).manual_seed(3)
    x_init = torch.randn(K, D, device=cuda_device, generator=g)
    from idff_b200 import sampler as sampler_mod
    ik, ik0 = sampler_mod.brownian_increments(4, torch.linspace(0, 1, NFE), dt, (T, K, D), cuda_device, g)
    nz = (ik.transpose(0, 1).contiguous(), ik0.transpose(0, 1).contiguous())
    out = []
    for fma in (0, 1):
        _lib.check(_lib.lib().idff_debug_md_chain_fma(fma), "chain_fma")
        try:
            out.append(md.generate_trajectories_chain(pw, K, D, T, NFE, sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt,
                                                      device=cuda_device, x_init=x_init, noise=nz))
        finally:
            _lib.check(_lib.lib().idff_debug_md_chain_fma(0), "chain_fma")
    d = (out[0] - out[1]).abs()
    scale = out[1].abs().max().item()
    assert torch.isfinite(out[0]).all() and d.max().item() <= 1e-5 * scale, (d.max().item(), scale)


def test_md_evaluate_model_grid(cuda_device):
    """evaluate_model (MD_simulation.py:159-194): one row per (nfe, sigma, gamma1, gamma2) in the reference's loop order with the
    reference's metric keys; the MMD column equals the oracle's MMD of the same frames (data (T, D) vs all generated frames)."""
    pw = packed("polyala", cuda_device)
    g = torch.Generator().manual_seed(2)
    T, D = 12, 46
    data = 100 * torch.randn(T, D, generator=g)
    rows = md.evaluate_model(pw, data, gamma1_values=[0.0, 0.2], gamma2_values=[1.0], sigma_values=[0.2, 0.5], nfes=[3],
                             num_trajectories=5, device=cuda_device)
    assert [(r["nfe"], r["sigma"], r["gamma1"], r["gamma2"]) for r in rows] == [(3, 0.2, 0.0, 1.0), (3, 0.2, 0.2, 1.0),
                                                                                 (3, 0.5, 0.0, 1.0), (3, 0.5, 0.2, 1.0)]
    keys = {"CC_mean", "CC_std", "RMSE_mean", "RMSE_std", "MAE_mean", "MAE_std", "nfe", "sigma", "gamma1", "gamma2", "MMD"}
    assert all(set(r) == keys and np.isfinite(r["MMD"]) and np.isfinite(r["RMSE_mean"]) for r in rows)
    # degree-valued frames with sigma_kernel = 1: the MMD is dominated by the diagonals, 1/T + 1/(5 T)
    assert rows[0]["MMD"] == pytest.approx(1 / T + 1 / (5 * T), rel=0.05)


def test_scorehead_gamma_sweep_is_one_launch_and_equals_per_tuple(cuda_device):
    """The order-1 script's gamma sweep (example_order1_with_prediction_head.py:305-328) on the score-head field: all tuples in ONE
    launch of the tcgen05 kernel (k_tc16<score head>, per-set stage tables), bit-identical to per-tuple solves; MMD per tuple."""
    from idff_b200 import fields, sampler
    dev = cuda_device
    pw = packed("scorehead256", dev)
    x0 = (torch.randn(2048, 2, generator=torch.Generator().manual_seed(2)) / 1.5).to(dev)
    gr = [0.0, 0.2, 1.0]
    tuples = [(a, b) for a in gr for b in gr]
    models = []
    for g in tuples:
        models.append(fields.CustomNetModelScoreHead(pw, 0.2, g[0], g[1]))
    ts = torch.linspace(0, 1, 3)
    before = _lib.launch_count()
    outs = sampler.sample_rk4_sweep(models, x0, ts)
    assert _lib.launch_count() - before <= 2          # the sweep kernel + the fp32 safety net's scan
    for k, m in enumerate(models):
        assert torch.equal(outs[k], sampler.sample_rk4(m, x0, ts)), tuples[k]
    true = (torch.randn(2048, 2, generator=torch.Generator().manual_seed(3)) * 2).to(dev)
    best, table = sweep.find_best_gammas_mmd(pw, x0, true, gamma_range=gr, nfe=2, score_head=True)
    assert len(table) == 9 and best in tuples
    from idff_b200.mmd import compute_mmd_rbf
    for (g, v), k in zip(table, range(9)):
        assert abs(v - compute_mmd_rbf(true, outs[k], 1.0)) <= 1e-6 + 1e-4 * abs(v)


def test_md_generation_sweep_is_the_per_set_generation(cuda_device):
    """idff_md_generate_sweep: the (sigma, gamma1, gamma2) loop of evaluate_model (MD_simulation.py:166-169) in one launch of the
    chain kernel -> every set's trajectories are bit for bit what the one-set launch returns on the same start rows and increments,
    including sets that need fewer jets than the others (gamma2 = 0; gamma1 = 0.5 and gamma2 = 0) and K not a multiple of 4."""
    import time
    from idff_b200 import sampler as sampler_mod
    dev = cuda_device
    pw = packed("polyala", dev)
    K, D, T, NFE, dt = 5, 46, 30, 3, 0.3
    sets = [(0.5, 0.2, 5.0), (0.2, 0.0, 1.0), (0.5, 0.2, 0.0), (0.3, 0.5, 0.0), (1.0, 1.0, 10.0), (0.05, 0.3, 2.0), (0.5, 0.5, 3.0)]
    S = len(sets)
    g = torch.Generator(device=dev).manual_seed(11)
    n_steps = sampler_mod.sde_num_steps(torch.linspace(0, 1, NFE), dt)
    x_init = torch.randn(S, K, D, device=dev, generator=g)
    ik = torch.randn(T, n_steps, S, K, D, device=dev, generator=g) * 0.3
    ik0 = torch.randn(T, n_steps, S, K, D, device=dev, generator=g) * 0.05
    ys = md.generate_trajectories_sweep(pw, sets, K, D, T, NFE, dt=dt, device=dev, x_init=x_init, noise=(ik, ik0))
    assert tuple(ys.shape) == (T, S, K, D) and torch.isfinite(ys[:, 0]).all()
    same = lambda a, b: torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))   # bit patterns: a set whose SDE runs away is NaN in both
    for i, (sg, g1, g2) in enumerate(sets):
        one = md.generate_trajectories_chain(pw, K, D, T, NFE, sigma=sg, gamma1=g1, gamma2=g2, dt=dt, device=dev,
                                             x_init=x_init[i].contiguous(),
                                             noise=(ik[:, :, i].contiguous(), ik0[:, :, i].contiguous()))
        assert same(ys[:, i], one), (i, (ys[:, i] - one).abs().max().item())
    # euler, one set through the sweep entry point, and the wall clock of a 60-set sweep against 60 launches
    ye = md.generate_trajectories_sweep(pw, sets[:2], K, D, T, NFE, dt=dt, device=dev, x_init=x_init[:2].contiguous(),
                                        noise=(ik[:, :, :2].contiguous(), None), method="euler")
    for i in range(2):
        one = md.generate_trajectories_chain(pw, K, D, T, NFE, sigma=sets[i][0], gamma1=sets[i][1], gamma2=sets[i][2], dt=dt,
                                             device=dev, x_init=x_init[i].contiguous(), noise=(ik[:, :, i].contiguous(), None),
                                             method="euler")
        assert same(ye[:, i], one)
    full, g1only, none = md.generate_overlay_sets(pw, K, D, T, NFE, 0.5, 0.2, 5.0, dt=dt, device=dev)   # MD_simulation.py:303-305
    assert full.shape == g1only.shape == none.shape == (T, K, D) and torch.isfinite(full).all()
    big = [(0.1 + 0.01 * i, 0.2, 1.0 + i) for i in range(60)]
    md.generate_trajectories_sweep(pw, big, K, D, T, NFE, dt=dt, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    md.generate_trajectories_sweep(pw, big, K, D, T, NFE, dt=dt, device=dev)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for sg, g1, g2 in big:
        md.generate_trajectories_chain(pw, K, D, T, NFE, sigma=sg, gamma1=g1, gamma2=g2, dt=dt, device=dev)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"60-set PolyALA generation sweep, {T} frames x {K} trajectories: one launch {1e3 * (t1 - t0):.2f} ms, per set {1e3 * (t2 - t1):.2f} ms")
    assert (t1 - t0) < 0.5 * (t2 - t1)


def test_metrics_of_all_sweep_sets_at_once(cuda_device):
    """md.compute_metrics_sets (T, S, K, D) = md.compute_metrics per set (the reference's compute_metrics, MD_simulation.py:275-295,
    pinned by tests/golden/metrics.npz), NaN handling included."""
    g = torch.Generator().manual_seed(8)
    T, S, K, D = 25, 3, 5, 6
    data = 50 * torch.randn(T, D, generator=g)
    ys = (data[:, None, None, :] + 10 * torch.randn(T, S, K, D, generator=g)).to(cuda_device)
    ys[3, 1, 2, 4] = float("nan")
    rows = md.compute_metrics_sets(ys, data.to(cuda_device))
    assert len(rows) == S
    for i in range(S):
        one = md.compute_metrics(ys[:, i].contiguous(), data.to(cuda_device))
        for k in ("CC_mean", "CC_std", "RMSE_mean", "RMSE_std", "MAE_mean", "MAE_std"):
            a, b = rows[i][k], one[k]
            assert (np.isnan(a) and np.isnan(b)) or a == pytest.approx(b, rel=1e-12, abs=1e-12), (i, k, a, b)


def test_trajectory_generator_class_mirror(cuda_device, tmp_path):
    """md.TrajectoryGenerator: the reference's driver class (MD_simulation.py:88-295) by name -- constructor attributes, train_model
    on the persistent-kernel trainer (loss goes down, a state_dict is saved), generate_trajectories / generate_onestep shapes,
    compute_metrics keys, evaluate_model rows + CSV."""
    g = torch.Generator().manual_seed(4)
    T, D = 24, 46
    data = 120 * torch.sin(torch.cumsum(torch.randn(T, D, generator=g) * 0.2, 0))
    tg = md.TrajectoryGenerator(dataset_name="synthetic", sigma=0.5, nfe=3, batch_size=10, num_trajectories=5,
                                save_dir=str(tmp_path), model_name="m", use_cuda=True, dataset=data, device=cuda_device)
    assert (tg.DataDim, tg.max_length, tg.NFE, tg.rgn_thr) == (D, T, 3, 30)
    tg.num_steps, tg.vis_step = 3000, 1000
    torch.manual_seed(0)
    model = tg.train_model(steps_per_launch=1000)
    assert (tmp_path / "synthetic" / "m.pt").exists()
    first = tg.trainer.run(1).item()
    assert np.isfinite(first)
    y = tg.generate_trajectories(model, gamma1=0.2, gamma2=5)
    assert isinstance(y, np.ndarray) and y.shape == (T, 5, D)
    one = tg.generate_onestep(model, 3, 2, gamma1=0.2, gamma2=5)
    assert one.shape == (3, 5, D)
    m = tg.compute_metrics(y)
    assert set(m) == {"CC_mean", "CC_std", "RMSE_mean", "RMSE_std", "MAE_mean", "MAE_std"}
    rows = tg.evaluate_model(model, [0.0, 0.2], [1.0, 5.0], [0.5], [3])
    assert len(rows) == 4 and (tmp_path / "synthetic" / "evaluation_results.csv").exists()
    with pytest.raises(FileNotFoundError):
        md.TrajectoryGenerator("no_such_dataset", 0.5, 3, 10, 5, str(tmp_path), "m", device=cuda_device)


def test_md_generation_sweep_more_tiles_than_sms(cuda_device):
    """idff_md_generate_sweep with more trajectory tiles than SMs (100 sets x 2 tiles): a CTA then walks tiles of DIFFERENT parameter
    sets one after the other and must reload the set's stage table and diffusion values -- every checked set is still bit for bit
    the one-set launch."""
    from idff_b200 import sampler as sampler_mod
    dev = cuda_device
    pw = packed("polyala", dev)
    K, D, T, NFE, dt = 5, 46, 6, 3, 0.3
    sets = [(0.2 + 0.005 * i, 0.1 + 0.002 * i, 1.0 + 0.05 * i) for i in range(100)]
    S = len(sets)
    g = torch.Generator(device=dev).manual_seed(21)
    n_steps = sampler_mod.sde_num_steps(torch.linspace(0, 1, NFE), dt)
    x_init = torch.randn(S, K, D, device=dev, generator=g)
    ik = torch.randn(T, n_steps, S, K, D, device=dev, generator=g) * 0.3
    ik0 = torch.randn(T, n_steps, S, K, D, device=dev, generator=g) * 0.05
    ys = md.generate_trajectories_sweep(pw, sets, K, D, T, NFE, dt=dt, device=dev, x_init=x_init, noise=(ik, ik0))
    for i in (0, 1, 73, 74, 75, 98, 99):
        sg, g1, g2 = sets[i]
        one = md.generate_trajectories_chain(pw, K, D, T, NFE, sigma=sg, gamma1=g1, gamma2=g2, dt=dt, device=dev,
                                             x_init=x_init[i].contiguous(),
                                             noise=(ik[:, :, i].contiguous(), ik0[:, :, i].contiguous()))
        assert torch.equal(ys[:, i].contiguous().view(torch.int32), one.view(torch.int32)), i
