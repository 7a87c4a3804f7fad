"""-m gpu: the fused samplers (whole RK4 / SRK solve in one launch) against trajectories produced by the reference's
field code under the restated torchdiffeq / torchsde schemes.

Accumulated-drift contract (SURVEY.md section 7 "kink chaos"): the fp32 reference itself differs from its fp64 twin by
max 7e-5..5e-4 / median ~1e-6 over a solve (order 2) because SELU' and SELU'' jump at 0, and isolated order-3
particles by up to 0.2.  So the bound is stated on the median and the 99.9th percentile; the max is reported."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_common import assert_field_close, drift_report, packed
from idff_b200 import _lib, fields, sampler
from make_golden_consts import GAMMAS2, GAMMAS3

pytestmark = pytest.mark.gpu

IMPLS = [_lib.IDFF_IMPL_GENERIC, _lib.IDFF_IMPL_FUSED, _lib.IDFF_IMPL_TENSOR, _lib.IDFF_IMPL_TENSOR16, _lib.IDFF_IMPL_LAYERED,
         _lib.IDFF_IMPL_AUTO]


def _check_drift(got, ref32, ref64, what, med=2e-5, p999=2e-3):
    r32, r64 = drift_report(got, ref32), drift_report(got, ref64)
    base = drift_report(ref32, ref64)
    print(f"{what}: vs fp32 ref {r32} | vs fp64 ref {r64} | fp32 ref vs fp64 ref {base}")
    # bounds are relative to what the reference's own fp32 arithmetic does on the same solve (base).  Measured on the
    # most kink-sensitive golden solve (gamma = (1, 2), nfe = 10; reference fp32 vs fp64 median 8.5e-6): fp32-FMA kernels
    # 8.5e-6, tensor-core kernels 1.9e-5 .. 2.6e-5 (their fp32 accumulators truncate instead of rounding) -- hence 4x
    assert r32["median"] <= max(med, 4 * base["median"]) and r64["median"] <= max(med, 4 * base["median"]), what
    assert r32["p999"] <= max(p999, 10 * base["p999"]), what
    assert r64["p999"] <= max(p999, 10 * base["p999"]), what


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("which", ["checkerboard", "8gaussians"])
def test_rk4_order2_vs_golden(cuda_device, which, impl):
    g = load_golden("traj2d")
    x0 = torch.tensor(g["x0"]).to(cuda_device)
    pw = packed(which, cuda_device)
    for nfe in (2, 3, 5, 10):
        for gi, gam in enumerate(GAMMAS2[:4]):
            key = f"{which}_o2_nfe{nfe}_g{gi}"
            if key not in g.files:
                continue
            m = fields.CustomNetModel(pw, 0.2, *gam)
            m.impl = impl
            ts = torch.linspace(0, 1, nfe + 1)
            traj = sampler.odeint(m, x0, ts, rtol=1e-5, atol=1e-5, method="rk4")
            assert traj.shape == (nfe + 1, x0.shape[0], 2)
            assert torch.equal(traj[0], x0)
            final = sampler.sample_rk4(m, x0, ts)
            assert torch.equal(final, traj[-1])
            ref = g[key]
            ref_final = ref[-1] if ref.ndim == 3 else ref
            _check_drift(final.cpu().numpy(), ref_final, g[key + "_f64"], f"{key} impl={impl}")
            if ref.ndim == 3:
                # first interval only: one rk4 step, no accumulated kink divergence yet
                assert drift_report(traj[1].cpu().numpy(), ref[1])["median"] <= 1e-5


@pytest.mark.parametrize("impl", IMPLS)
def test_rk4_order3_and_cfm_vs_golden(cuda_device, impl):
    g = load_golden("traj2d")
    x0 = torch.tensor(g["x0"]).to(cuda_device)
    pw = packed("checkerboard", cuda_device)
    for nfe in (2, 5):
        for gi, gam in enumerate(GAMMAS3[:2]):
            key = f"checkerboard_o3_nfe{nfe}_g{gi}"
            m = fields.CustomNetModelOrder3(pw, 0.2, *gam)
            m.impl = impl
            if impl == _lib.IDFF_IMPL_TENSOR and gam[2] != 0:      # the 3xTF32 fused kernel holds two jets; 3xFP16 has a three-jet kernel
                with pytest.raises(_lib.IdffError):
                    sampler.sample_rk4(m, x0, torch.linspace(0, 1, nfe + 1))
                continue
            final = sampler.sample_rk4(m, x0, torch.linspace(0, 1, nfe + 1))
            _check_drift(final.cpu().numpy(), g[key], g[key + "_f64"], f"{key} impl={impl}", med=5e-5, p999=0.05)
    c = fields.CFMModel(pw)
    c.impl = impl
    final = sampler.sample_rk4(c, x0, torch.linspace(0, 1, 6))
    assert drift_report(final.cpu().numpy(), g["checkerboard_cfm_nfe5"])["max"] <= 1e-4


def test_rk4_scorehead_vs_golden(cuda_device):
    g = load_golden("scorehead")
    x0 = torch.tensor(load_golden("traj2d")["x0"][:256]).to(cuda_device)
    pw = packed("scorehead", cuda_device)
    for gi, gam in enumerate(g["gammas"]):
        m = fields.CustomNetModelScoreHead(pw, 0.2, float(gam[0]), float(gam[1]))
        final = sampler.sample_rk4(m, x0, torch.linspace(0, 1, 6))
        r = drift_report(final.cpu().numpy(), g[f"traj_nfe5_g{gi}"])
        assert r["median"] <= 1e-5 and r["p999"] <= 1e-3, r
    g = load_golden("scorehead256")   # the script's width: every kernel family that serves it
    x0 = torch.tensor(g["x0"]).to(cuda_device)
    pw = packed("scorehead256", cuda_device)
    for gi, gam in enumerate(g["gammas"]):
        for impl in (_lib.IDFF_IMPL_AUTO, _lib.IDFF_IMPL_GENERIC, _lib.IDFF_IMPL_TENSOR16):
            m = fields.CustomNetModelScoreHead(pw, 0.2, float(gam[0]), float(gam[1]))
            m.impl = impl
            final = sampler.sample_rk4(m, x0, torch.linspace(0, 1, 6))
            _check_drift(final.cpu().numpy(), g[f"traj_nfe5_g{gi}"], g[f"traj_nfe5_g{gi}_f64"], f"scorehead256 g{gi} impl={impl}",
                         med=1e-5, p999=1e-3)


def test_cfm_comparator_on_the_scorehead_network(cuda_device):
    """The order-1 script's CFMModel (example_order1_with_prediction_head.py:335-356): predictions[:, :2] - x0 on the score-head
    network fed cat[xt, xt, t] -- generic kernel and k_tc16<score head>, one-launch solve and host-stepped forward(t, x) calls,
    against the end points the reference's own class produces."""
    g = load_golden("scorehead256")
    x0 = torch.tensor(g["x0"]).to(cuda_device)
    pw = packed("scorehead256", cuda_device)
    ts = torch.linspace(0, 1, 6)
    for impl in (_lib.IDFF_IMPL_AUTO, _lib.IDFF_IMPL_GENERIC, _lib.IDFF_IMPL_TENSOR16):
        c = fields.CFMModel(pw)
        c.impl = impl
        final = sampler.sample_rk4(c, x0, ts)
        _check_drift(final.cpu().numpy(), g["cfm_nfe5"], g["cfm_nfe5_f64"], f"scorehead CFM impl={impl}", med=1e-5, p999=1e-3)
    c = fields.CFMModel(pw)
    stepped = sampler.odeint(lambda t, x: c(t, x), x0, ts, method="rk4")[-1]
    r = drift_report(stepped.cpu().numpy(), g["cfm_nfe5"])
    assert r["median"] <= 1e-5 and r["p999"] <= 1e-3, r
    with pytest.raises(_lib.IdffError):      # a flow network of another shape is still rejected
        fields.CFMModel(packed("scorehead", cuda_device))(torch.tensor(0.0), torch.zeros(4, 3, device=cuda_device))


def test_scorehead_tensor_kernel_properties(cuda_device):
    """The score-head field on the tcgen05 kernel (k_tc16<SH>) at a ragged size: against the fp32 generic kernel per evaluation
    and over a solve with a trajectory; sub-batches reproduce the whole bit for bit; AUTO = raw kernel on clean particles and the
    fp32 re-solve where the fp16 split overflows."""
    pw = packed("scorehead256", cuda_device)
    N = 64 * 149 + 37
    x0 = (torch.randn(N, 2, generator=torch.Generator().manual_seed(5)) / 1.5).to(cuda_device)
    ts = torch.linspace(0, 1, 4)
    outs = {}
    for impl in (_lib.IDFF_IMPL_GENERIC, _lib.IDFF_IMPL_TENSOR16, _lib.IDFF_IMPL_AUTO):
        m = fields.CustomNetModelScoreHead(pw, 0.2, 0.2, 0.3)
        m.impl = impl
        for tv in (0.0, 0.4, 1.0):
            f = m(torch.tensor(tv), x0)
            if impl == _lib.IDFF_IMPL_GENERIC:
                outs[tv] = f
            else:
                scale = max(1.0, outs[tv].abs().max().item())
                assert (f - outs[tv]).abs().max().item() <= 1e-5 * scale, (impl, tv)
        traj = sampler.odeint(m, x0, ts, method="rk4")
        if impl == _lib.IDFF_IMPL_GENERIC:
            outs["traj"] = traj
        else:
            r = drift_report(traj.cpu().numpy(), outs["traj"].cpu().numpy())
            assert r["median"] <= 1e-5 and r["p999"] <= 1e-3, (impl, r)
            sub = sampler.odeint(m, x0[1000:3000].contiguous(), ts, method="rk4")
            assert torch.equal(sub, traj[:, 1000:3000])
    # BASELINE configs[1] scale (2^22 particles): finite, and any slice of the batch reproduces bit for bit on its own
    big = (torch.randn(1 << 22, 2, generator=torch.Generator().manual_seed(6)) / 1.5).to(cuda_device)
    mb = fields.CustomNetModelScoreHead(pw, 0.2, 0.2, 0.3)
    whole = sampler.sample_rk4(mb, big, ts)
    assert torch.isfinite(whole).all()
    for lo, hi in ((0, 64 * 500), (123457, 123457 + 70001), ((1 << 22) - 4099, 1 << 22)):
        assert torch.equal(sampler.sample_rk4(mb, big[lo:hi].contiguous(), ts), whole[lo:hi]), (lo, hi)
    # range guard: huge inputs overflow the fp16 split; AUTO re-solves exactly those rows in fp32
    xb = x0.clone()
    xb[::7] *= 3.0e4
    g = fields.CustomNetModelScoreHead(pw, 0.2, 0.2, 0.3)
    g.impl = _lib.IDFF_IMPL_GENERIC
    a = fields.CustomNetModelScoreHead(pw, 0.2, 0.2, 0.3)
    ref = sampler.sample_rk4(g, xb, ts)
    got = sampler.sample_rk4(a, xb, ts)
    fin = torch.isfinite(ref).all(dim=1)
    assert torch.isfinite(got[fin]).all()
    d = (got[fin] - ref[fin]).abs() / ref[fin].abs().clamp(min=1.0)
    assert d.median().item() <= 1e-5 and d.quantile(0.999).item() <= 1e-3


def test_host_stepped_odeint_equals_fused(cuda_device):
    """Any callable can be stepped on the host with forward(t, x); it must agree with the one-launch solve."""
    pw = packed("8gaussians", cuda_device)
    m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
    x0 = torch.tensor(load_golden("traj2d")["x0"][:300]).to(cuda_device)
    ts = torch.linspace(0, 1, 4)
    fused = sampler.odeint(m, x0, ts, method="rk4")
    stepped = sampler.odeint(lambda t, x: m(t, x), x0, ts, method="rk4")
    r = drift_report(fused.cpu().numpy(), stepped.cpu().numpy())
    assert r["median"] <= 2e-6 and r["p999"] <= 1e-3, r


def test_rk4_properties_at_scale(cuda_device):
    """N = 2^20 particles (BASELINE configs[1] lower end): particles are independent, so any sub-batch reproduces the
    whole bit-for-bit (this is what makes the multi-GPU sharding exact); a long grid crosses the per-launch interval
    limit and must equal the same grid solved in two calls."""
    pw = packed("checkerboard", cuda_device)
    m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
    g = torch.Generator(device=cuda_device).manual_seed(5)
    x0 = torch.randn(1 << 20, 2, device=cuda_device, generator=g) / 1.5
    ts = torch.linspace(0, 1, 3)
    whole = sampler.sample_rk4(m, x0, ts)
    assert torch.isfinite(whole).all()
    lo, hi = 777, 777 + 100003
    assert torch.equal(sampler.sample_rk4(m, x0[lo:hi], ts), whole[lo:hi])
    assert torch.equal(sampler.sample_rk4(m, x0[:5], ts), whole[:5])
    assert sampler.sample_rk4(m, x0[:0], ts).shape == (0, 2)
    # degenerate grids
    assert torch.equal(sampler.sample_rk4(m, x0[:64], torch.tensor([0.0])), x0[:64])
    long_ts = torch.linspace(0, 1, 131)                      # 130 intervals > 96 per launch
    a = sampler.sample_rk4(m, x0[:512], long_ts)
    mid = sampler.sample_rk4(m, x0[:512], long_ts[:97])
    b = sampler.sample_rk4(m, mid, long_ts[96:])
    assert torch.equal(a, b)
    with pytest.raises(_lib.IdffError):
        sampler.sample_rk4(m, x0[:8], torch.tensor([0.0, 0.5, 0.5, 1.0]))


@pytest.mark.parametrize("impl", [_lib.IDFF_IMPL_GENERIC, _lib.IDFF_IMPL_FUSED, _lib.IDFF_IMPL_TENSOR16, _lib.IDFF_IMPL_AUTO])
def test_sde_srk_and_euler_vs_golden(cuda_device, impl):
    g = load_golden("md")
    y = torch.tensor(g["y"]).to(cuda_device)
    pw = packed("polyala", cuda_device)
    sde = fields.SDE_cal(pw, input_dim=46, sigma=0.5, init_ind=3, max_length=200, dt=0.3, gamma1=0.2, gamma2=5.0 / 4)
    sde.impl = impl
    ts = torch.tensor(g["ts"])
    I_k, I_k0 = torch.tensor(g["I_k"]).to(cuda_device), torch.tensor(g["I_k0"]).to(cuda_device)
    assert sampler.sde_num_steps(ts, 0.3) == I_k.shape[0] == 4
    out = sampler.sdeint(sde, y, ts, dt=0.3, method="srk", bm=(I_k, I_k0))
    assert out.shape == (3, 64, 46) and torch.equal(out[0], y)
    eul = sampler.sdeint(sde, y, ts, dt=0.3, method="euler", bm=(I_k, None))
    # The golden start points are synthetic (not on the data manifold) and a quarter of the reference's own trajectories leave
    # it for good (|y| up to 3e8 after two steps, 3e18 after four).  The raw 3xFP16 tensor kernel (impl 5) holds values up
    # to 65504: it is compared on the trajectories the reference keeps inside that range and must flag the others non-finite;
    # the default route re-solves those in fp32 (safety net) and is compared on everything, like the fp32 kernels.
    for name, got, ref in (("srk", out, g["srk"]), ("euler", eul, g["euler"])):
        inr = np.abs(ref).max(axis=(0, 2)) < 1.0e3       # states that stay near the data range keep every activation below 65504
        assert inr.sum() >= 8
        for i in (1, 2):
            # values are O(100) degrees; 1e-5 relative per step over 4 steps x 3 drift evaluations
            if impl == _lib.IDFF_IMPL_TENSOR16:
                assert_field_close(got[i].cpu().numpy()[inr], ref[i][inr], 1e-4, f"{name} ts[{i}] (in-range trajectories)")
                far = np.abs(ref[2]).max(axis=1) > 1.0e7
                assert not torch.isfinite(got[2][torch.tensor(far)]).all(dim=1).any()     # loud, never silently wrong
            else:
                assert_field_close(got[i].cpu().numpy(), ref[i], 1e-4, f"{name} ts[{i}]")
    # noise drawn on the device when not injected: finite, right shape, differs between draws
    a = sampler.sdeint(sde, y, ts, dt=0.3)
    b = sampler.sdeint(sde, y, ts, dt=0.3)
    assert a.shape == (3, 64, 46) and not torch.equal(a, b)
    assert impl == _lib.IDFF_IMPL_TENSOR16 or torch.isfinite(a).all()   # (the raw fp16 kernel flags the runaway trajectories)


def test_fused_kernel_is_selected_and_agrees_with_generic(cuda_device):
    """IDFF_IMPL_AUTO must pick the persistent TMA-fed kernel for the benchmark shapes; forcing it on a shape it does
    not serve is an error (no silent fallback); both kernel families agree to fp32 round-off."""
    pw = packed("checkerboard", cuda_device)
    x0 = torch.tensor(load_golden("traj2d")["x0"]).to(cuda_device)
    ts = torch.linspace(0, 1, 3)
    outs = {}
    for impl in (_lib.IDFF_IMPL_GENERIC, _lib.IDFF_IMPL_FUSED, _lib.IDFF_IMPL_TENSOR, _lib.IDFF_IMPL_TENSOR16, _lib.IDFF_IMPL_AUTO):
        m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
        m.impl = impl
        outs[impl] = sampler.sample_rk4(m, x0, ts)
    assert torch.equal(outs[_lib.IDFF_IMPL_TENSOR16], outs[_lib.IDFF_IMPL_AUTO])   # AUTO = tcgen05 3xFP16 kernel on this shape
    r = drift_report(outs[_lib.IDFF_IMPL_TENSOR16].cpu().numpy(), outs[_lib.IDFF_IMPL_GENERIC].cpu().numpy())
    assert r["median"] <= 2e-5 and r["p999"] <= 1e-3, r                            # 3xFP16, truncating accumulation
    assert _lib.weights_nonfinite(pw.handle) == (0, 0)
    r = drift_report(outs[_lib.IDFF_IMPL_FUSED].cpu().numpy(), outs[_lib.IDFF_IMPL_GENERIC].cpu().numpy())
    assert r["median"] <= 5e-6 and r["p999"] <= 1e-3, r
    r = drift_report(outs[_lib.IDFF_IMPL_TENSOR].cpu().numpy(), outs[_lib.IDFF_IMPL_GENERIC].cpu().numpy())
    assert r["median"] <= 2e-5 and r["p999"] <= 1e-3, r                            # 3xTF32, truncating accumulation
    # ragged particle counts around the 32-particle tile / 148-CTA grid
    for impl in (_lib.IDFF_IMPL_FUSED, _lib.IDFF_IMPL_TENSOR, _lib.IDFF_IMPL_TENSOR16):
        m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
        m.impl = impl
        big = torch.randn(148 * 32 * 2 + 17, 2, device=cuda_device) / 1.5
        whole = sampler.sample_rk4(m, big, ts)
        for n in (1, 3, 31, 32, 33, 63, 64, 65, 4735, 4737, 9473):
            assert torch.equal(sampler.sample_rk4(m, big[:n], ts), whole[:n])
    sh = fields.CustomNetModelScoreHead(packed("scorehead", cuda_device), 0.2, 0.2, 0.2)
    sh.impl = _lib.IDFF_IMPL_FUSED
    with pytest.raises(_lib.IdffError):
        sampler.sample_rk4(sh, x0[:8], ts)


def test_tensor16_full_size_properties_and_range_guard(cuda_device):
    """BASELINE-size batch through the default (3xFP16 tensor) sampler: particles are independent, so any slice of a
    4 M-particle solve is reproduced bit-for-bit by solving the slice alone; trajectories end where sample_rk4 ends.
    Operands beyond the fp16 range make THAT particle's output non-finite in the raw kernel (IDFF_IMPL_TENSOR16; never
    silently wrong, counted); the default route (IDFF_IMPL_AUTO) re-solves exactly those particles in fp32."""
    from idff_b200.weights import PackedWeights
    pw = PackedWeights.from_npz(load_golden("weights_checkerboard"), device=cuda_device)   # own handle: fresh counters
    m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)                      # AUTO -> k_tc16
    n = 1 << 22
    x0 = torch.randn(n, 2, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(9)) / 1.5
    ts = torch.linspace(0, 1, 3)
    whole = sampler.sample_rk4(m, x0, ts)
    assert torch.isfinite(whole).all()
    for lo, hi in ((0, 4097), (1 << 21, (1 << 21) + 64), (n - 12345, n)):
        assert torch.equal(sampler.sample_rk4(m, x0[lo:hi], ts), whole[lo:hi])
    traj = sampler.odeint(m, x0[:70000], ts, method="rk4")
    assert torch.equal(traj[-1], whole[:70000]) and torch.equal(traj[0], x0[:70000])
    # the distribution moved toward the checkerboard support (|x| <= 4): a sanity property of the shipped model
    assert (whole.abs() < 6).float().mean().item() > 0.99
    # range guard
    assert _lib.weights_nonfinite(pw.handle) == (0, 0)
    bad = x0[:128].clone()
    bad[5] = 3.0e6                                                   # layer-1 pre-activations far beyond 65504
    raw = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
    raw.impl = _lib.IDFF_IMPL_TENSOR16                               # the raw kernel, no safety net
    out = sampler.sample_rk4(raw, bad, ts)
    assert not torch.isfinite(out[5]).all()
    keep = torch.ones(128, dtype=torch.bool, device=cuda_device)
    keep[5] = False
    assert torch.equal(out[keep], whole[:128][keep])                 # the other particles of the tile are untouched
    n_bad, n_fixed = _lib.weights_nonfinite(pw.handle)
    assert n_bad > 0 and n_fixed == 0
    mg = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
    mg.impl = _lib.IDFF_IMPL_GENERIC                                 # fp32 FMA arithmetic serves such inputs
    ref = sampler.sample_rk4(mg, bad, ts)
    assert torch.isfinite(ref).all()
    safe = sampler.sample_rk4(m, bad, ts)                            # AUTO: same kernel + the fp32 safety net
    assert torch.equal(safe[keep], whole[:128][keep]) and torch.equal(safe[5], ref[5])
    assert _lib.weights_nonfinite(pw.handle)[1] == 1
    inplace = bad.clone()                                            # in place: the safety net keeps its own copy of the input
    assert torch.equal(sampler.sample_rk4(m, inplace, ts, out=inplace), safe)


def _scaled_layers(name, s0, s12, dtype=torch.float32):
    """the shipped network with layer 0 scaled by s0 and layers 1, 2 by s12: activations, jets and adjoints leave the fp16
    range for part of the batch while every fp32 quantity stays finite"""
    from conftest import golden_layers
    layers, _ = golden_layers(name)
    sc = [s0, s12, s12, 1.0]
    return [((W * k).to(dtype), (b * k).to(dtype)) for (W, b), k in zip(layers, sc)]


def test_auto_route_is_range_safe_vs_oracle(cuda_device):
    """IDFF_IMPL_AUTO never returns a value the reference would not.  A network whose activations overflow
    fp16 (layer-0 weights x 8, every 7th input x 1e4): AUTO = 3xFP16 tensor kernel + fp32 re-solve of the overflowed
    particles (idff_fixup.cu) against the oracle (the reference's nested-autograd field, Autograd_example_order2.py:44-76)
    under the normal per-evaluation contract (1e-5 max|f|, or the reference's own fp32<->fp64 gap where that is larger --
    the x 8 network is stiffer than the shipped one), for single evaluations, whole solves, trajectories and the batched
    gamma sweep."""
    import idff_oracle as O
    from gpu_common import assert_field_vs_refs
    from idff_b200.weights import PackedWeights
    layers = _scaled_layers("checkerboard", 8.0, 1.0)
    layers64 = _scaled_layers("checkerboard", 8.0, 1.0, torch.float64)
    pw = PackedWeights([W.numpy() for W, _ in layers], [b.numpy() for _, b in layers], device=cuda_device)
    gen = torch.Generator().manual_seed(11)
    n = 4096 + 77
    x = torch.randn(n, 2, generator=gen) / 1.5
    x[::7] *= 1.0e4                                                  # every 7th particle far outside the data range
    xd = x.to(cuda_device)
    auto = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)                  # impl AUTO
    raw = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
    raw.impl = _lib.IDFF_IMPL_TENSOR16
    gen_m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
    gen_m.impl = _lib.IDFF_IMPL_GENERIC
    orc = O.Field2D(layers, 0.2, (0.2, 0.2), 2)
    orc64 = O.Field2D(layers64, 0.2, (0.2, 0.2), 2)
    for t in (0.0, 0.37, 5.0 / 6.0, 1.0):
        tt = torch.tensor(t)
        f_raw = raw(tt, xd)
        bad = ~torch.isfinite(f_raw).all(dim=1)
        assert bad.any() and not bad.all(), (t, int(bad.sum()))      # the test network does overflow fp16, for part of the batch
        f_auto = auto(tt, xd)
        assert torch.isfinite(f_auto).all()
        assert torch.equal(f_auto[~bad], f_raw[~bad])                # clean particles: the tensor kernel's own result
        assert torch.equal(f_auto[bad], gen_m(tt, xd)[bad])          # overflowed ones: fp32 FMA arithmetic
        ref = orc(tt, x)
        ref64 = orc64(tt.double(), x.double())
        assert torch.isfinite(ref).all()
        # per-evaluation contract per particle group: a 1e4-scale particle must not hide the O(1) ones behind max|f|
        for grp in (bad.cpu(), ~bad.cpu()):
            assert_field_vs_refs(f_auto.cpu()[grp].numpy(), ref[grp].numpy(), ref64[grp].numpy(), 1e-5, f"range-safe eval t={t}")
    assert _lib.weights_nonfinite(pw.handle)[1] > 0
    # whole solves: every 5th start point is out of the fp16 range from the first evaluation on
    x0 = (torch.randn(2048 + 5, 2, generator=gen) / 1.5)
    x0[::5] *= 1.0e4
    x0d = x0.to(cuda_device)
    ts = torch.linspace(0, 1, 3)
    out_raw = sampler.sample_rk4(raw, x0d, ts)
    bad = ~torch.isfinite(out_raw).all(dim=1)
    assert bad.any() and not bad.all()
    traj = sampler.odeint(auto, x0d, ts, method="rk4")
    out = sampler.sample_rk4(auto, x0d, ts)
    ref_traj = sampler.odeint(gen_m, x0d, ts, method="rk4")
    assert torch.equal(traj[-1], out) and torch.equal(traj[0], x0d)
    assert torch.equal(out[~bad], out_raw[~bad])
    assert torch.equal(traj[:, bad], ref_traj[:, bad])               # re-solved particles incl. their trajectory rows
    fin = torch.isfinite(ref_traj[-1]).all(dim=1)
    assert torch.equal(torch.isfinite(out).all(dim=1), fin)          # non-finite only where fp32 itself overflows
    # batched gamma sweep: same rows as per-tuple AUTO calls
    gam = [(0.2, 0.0), (0.2, 0.2), (1.0, 2.0), (0.5, 0.0)]
    models = [fields.CustomNetModel(pw, 0.2, a, b) for a, b in gam]
    outs = sampler.sample_rk4_sweep(models, x0d, ts)
    for mdl, o in zip(models, outs):
        assert torch.equal(o, sampler.sample_rk4(mdl, x0d, ts))


def test_cfm_solve_longer_than_one_launch(cuda_device):
    """CFMModel keeps the x captured at t == 0 for the whole solve (Autograd_example_order2.py:310-332); a grid of
    more than 96 intervals spans several launches and every launch must see that x0.  Against the oracle's CFM field under
    the restated rk4, every kernel family; a grid that does not start at t == 0 has no x0 in the reference either."""
    import idff_oracle as O
    from conftest import golden_layers
    layers, _ = golden_layers("checkerboard")
    pw = packed("checkerboard", cuda_device)
    x0 = torch.tensor(load_golden("traj2d")["x0"][:300])
    ts = torch.linspace(0, 1, 131)                                   # 130 intervals
    ref = O.odeint_rk4(O.FieldCFM(layers), x0, ts)
    for impl in IMPLS:
        c = fields.CFMModel(pw)
        c.impl = impl
        traj = sampler.odeint(c, x0.to(cuda_device), ts, method="rk4")
        assert drift_report(traj[-1].cpu().numpy(), ref[-1].numpy())["max"] <= 2e-4, impl
        assert drift_report(traj[100].cpu().numpy(), ref[100].numpy())["max"] <= 2e-4, impl
        with pytest.raises(_lib.IdffError):
            sampler.sample_rk4(c, x0.to(cuda_device), torch.linspace(0.25, 1, 4))


def test_order3_fused_three_jet_kernel(cuda_device):
    """The order-3 field (Autograd_example_order3.py:34-77) on its fused three-jet tcgen05 kernel (k_tc16o3, the default
    route for gamma3 != 0): independent particles (any slice reproduces the whole bit for bit, ragged tile counts included),
    trajectories end where sample_rk4 ends, agreement with the fp32 FMA kernels inside the accumulated-drift contract, and the
    fp16 range guard + fp32 safety net of the default route."""
    from idff_b200.weights import PackedWeights
    pw = PackedWeights.from_npz(load_golden("weights_checkerboard"), device=cuda_device)   # own handle: fresh counters
    m = fields.CustomNetModelOrder3(pw, 0.2, 0.2, 0.2, 0.2)                                # AUTO -> k_tc16o3
    n = (1 << 18) + 17
    x0 = torch.randn(n, 2, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(21)) / 1.5
    for nfe in (2, 5):
        ts = torch.linspace(0, 1, nfe + 1)
        before = _lib.launch_count()
        whole = sampler.sample_rk4(m, x0, ts)
        assert _lib.launch_count() - before == 2                       # ONE sampler launch + the safety net's scan
        assert torch.isfinite(whole).all()
        for lo, hi in ((0, 1), (0, 31), (5, 5 + 33), (1000, 1000 + 4737), (n - 9473, n)):
            assert torch.equal(sampler.sample_rk4(m, x0[lo:hi].contiguous(), ts), whole[lo:hi])
        traj = sampler.odeint(m, x0[:5000], ts, method="rk4")
        assert torch.equal(traj[-1], whole[:5000]) and torch.equal(traj[0], x0[:5000])
        for other in (_lib.IDFF_IMPL_GENERIC, _lib.IDFF_IMPL_FUSED, _lib.IDFF_IMPL_LAYERED):
            mo = fields.CustomNetModelOrder3(pw, 0.2, 0.2, 0.2, 0.2)
            mo.impl = other
            r = drift_report(whole[:8192].cpu().numpy(), sampler.sample_rk4(mo, x0[:8192], ts).cpu().numpy())
            assert r["median"] <= 5e-5 and r["p999"] <= 0.05, (nfe, other, r)   # order-3 kink outliers: SURVEY section 8(d)
    assert _lib.weights_nonfinite(pw.handle) == (0, 0)
    # single evaluations through forward(t, x): interior and boundary times
    mg = fields.CustomNetModelOrder3(pw, 0.2, 0.2, 0.2, 0.2)
    mg.impl = _lib.IDFF_IMPL_GENERIC
    for t in (0.0, 0.37, 1.0):
        a, b = m(torch.tensor(t), x0[:8192]), mg(torch.tensor(t), x0[:8192])
        # 8192 random points: a few sit on a SELU kink of some unit, where E (the second derivative data) jumps between two
        # fp32-equivalent evaluations (SURVEY section 8(d) "kink outliers"); the golden-point test holds the max, this the bulk
        r = drift_report(a.cpu().numpy(), b.cpu().numpy())
        scale = max(1.0, float(b.abs().max()))
        amp = 50.0 if t == 1.0 else 1.0      # the t == 1 branch divides by 1e-2: the MLP's fp32 round-off x 100 (gpu_common.py)
        assert r["median"] <= 2e-6 * scale * amp and r["p999"] <= 1e-4 * scale * amp, (t, r, scale)
    # range guard
    ts = torch.linspace(0, 1, 3)
    bad = x0[:100].clone()
    bad[7] = 3.0e6
    raw = fields.CustomNetModelOrder3(pw, 0.2, 0.2, 0.2, 0.2)
    raw.impl = _lib.IDFF_IMPL_TENSOR16
    out = sampler.sample_rk4(raw, bad, ts)
    assert not torch.isfinite(out[7]).all()
    keep = torch.ones(100, dtype=torch.bool, device=cuda_device)
    keep[7] = False
    good = sampler.sample_rk4(m, x0[:100].contiguous(), ts)
    assert torch.equal(out[keep], good[keep])
    safe = sampler.sample_rk4(m, bad, ts)
    assert torch.equal(safe[keep], good[keep]) and torch.equal(safe[7], sampler.sample_rk4(mg, bad, ts)[7])
    assert _lib.weights_nonfinite(pw.handle)[1] == 1


def _polyala_on_manifold(n, seed=0):
    """Data-like PolyALA states: fixed points of the shipped denoiser x -> x1hat(t = 0.5, x) (frame 3), found by iterating it
    from uniform angles and keeping the rows that converge (about half).  From such states the reference's SDE stays within
    +-190 degrees, as its generator does from real frames (MD_simulation.py:197-213)."""
    import idff_oracle as O
    from conftest import golden_layers
    layers, emb = golden_layers("polyala")
    gen = torch.Generator().manual_seed(seed)
    z = (torch.rand(4 * n, 46, generator=gen) * 2 - 1) * 180
    for _ in range(12):
        z = O.mlp_embedding_forward(layers, emb, torch.cat([z, torch.full((z.shape[0], 1), 0.5)], -1), torch.full((z.shape[0], 1), 3.0))
    ok = torch.isfinite(z).all(1) & (z.abs().max(1).values < 400)
    assert int(ok.sum()) >= n
    return z[ok][:n].contiguous(), layers, emb


@pytest.mark.parametrize("impl", [_lib.IDFF_IMPL_TENSOR16, _lib.IDFF_IMPL_AUTO])
def test_polyala_tensor_kernel_vs_oracle(cuda_device, impl):
    """the PolyALA drift + srk / euler on tcgen05 (k_md16) against the oracle -- the reference's SDE_cal.f
    (nested autograd, MD_simulation.py:497-526) under the restated torchsde schemes -- on data-like states, where the
    reference's solution stays on the manifold; plus agreement with the fp32 FFMA2 kernel on a larger batch with ragged
    tile counts, and the default route's safety net on trajectories that leave the fp16 range."""
    import idff_oracle as O
    y0, layers, emb = _polyala_on_manifold(300)
    pw = packed("polyala", cuda_device)
    ts = torch.linspace(0, 1, 3)
    grid = O.sde_step_grid(ts, 0.3)
    h = torch.tensor(np.diff(np.array(grid, dtype=np.float32)))
    gen = torch.Generator().manual_seed(5)
    ik = torch.randn(len(h), 300, 46, generator=gen) * h.sqrt()[:, None, None]
    ik0 = 0.5 * h[:, None, None] * ik + torch.randn(len(h), 300, 46, generator=gen) * (h ** 1.5)[:, None, None] / (2 * 3 ** 0.5)
    for init_ind, g2 in ((3, 1.25), (0, 5.0)):
        sde = fields.SDE_cal(pw, input_dim=46, sigma=0.5, init_ind=init_ind, max_length=200, dt=0.3, gamma1=0.2, gamma2=g2)
        sde.impl = impl
        orc = O.FieldMD(layers, emb, 46, sigma=0.5, init_ind=init_ind, dt=0.3, gamma1=0.2, gamma2=g2)
        yd, bm = y0.to(cuda_device), (ik.to(cuda_device), ik0.to(cuda_device))
        out = sampler.sdeint(sde, yd, ts, dt=0.3, method="srk", bm=bm)
        ref = O.sdeint_srk(orc, y0, ts, 0.3, ik, ik0)
        assert float(ref.abs().max()) < 400                              # the reference stays on the manifold
        for i in (1, 2):
            assert_field_close(out[i].cpu().numpy(), ref[i].numpy(), 1e-4, f"k_md16 srk ts[{i}] init_ind={init_ind}")
        eul = sampler.sdeint(sde, yd, ts, dt=0.3, method="euler", bm=(bm[0], None))
        refe = O.sdeint_euler(orc, y0, ts, 0.3, ik)
        for i in (1, 2):
            assert_field_close(eul[i].cpu().numpy(), refe[i].numpy(), 1e-4, f"k_md16 euler ts[{i}] init_ind={init_ind}")
        for t in (0.0, 0.3, 1.0):                                        # single drift evaluations through f(t, y)
            # on the manifold the drift (x1hat - y) / (1 - t) is a small difference of O(180) quantities: the 1e-5 budget is
            # relative to those (and divided by dt = 0.3 at t == 1), not to |f| ~ 0.1
            got_f, ref_f = sde.f(torch.tensor(t), yd).cpu().numpy(), orc.f(torch.tensor(t), y0).numpy()
            tol = 1e-5 * float(y0.abs().max()) / (0.3 if t == 1.0 else 1.0 - t)
            assert float(np.abs(got_f - ref_f).max()) <= tol, (t, float(np.abs(got_f - ref_f).max()), tol)
    # a larger batch with ragged tile counts: the tensor kernel against the fp32 FFMA2 kernel, trajectory by trajectory
    n = 148 * 32 + 45
    big = y0.repeat(17, 1)[:n].to(cuda_device) + 0.5 * torch.randn(n, 46, device=cuda_device)
    bmb = sampler.brownian_increments(len(h), ts, 0.3, (n, 46), cuda_device)
    sde = fields.SDE_cal(pw, input_dim=46, sigma=0.5, init_ind=3, max_length=200, dt=0.3, gamma1=0.2, gamma2=1.25)
    sde.impl = impl
    sf = fields.SDE_cal(pw, input_dim=46, sigma=0.5, init_ind=3, max_length=200, dt=0.3, gamma1=0.2, gamma2=1.25)
    sf.impl = _lib.IDFF_IMPL_FUSED
    a, b = sampler.sdeint(sde, big, ts, dt=0.3, bm=bmb), sampler.sdeint(sf, big, ts, dt=0.3, bm=bmb)
    assert torch.isfinite(b).all() and float(b.abs().max()) < 1000
    assert_field_close(a.cpu().numpy(), b.cpu().numpy(), 1e-4, "k_md16 vs FFMA2, 4781 trajectories")
    for lo, hi in ((0, 1), (31, 97), (n - 333, n)):                      # trajectories are independent
        if impl == _lib.IDFF_IMPL_AUTO and hi - lo < 256:
            continue                                                     # (the default route serves small batches on the FFMA2 kernel)
        part = sampler.sdeint(sde, big[lo:hi].contiguous(), ts, dt=0.3, bm=(bmb[0][:, lo:hi].contiguous(), bmb[1][:, lo:hi].contiguous()))
        assert torch.equal(part, a[:, lo:hi])
    # trajectories pushed off the manifold leave the fp16 range: raw kernel -> non-finite, default route -> fp32 values
    wild = big[:512].clone()
    wild[::8] = wild[::8] * 3000.0                                      # |y| ~ 5e5: beyond fp16 from the first operand on
    bmw = (bmb[0][:, :512].contiguous(), bmb[1][:, :512].contiguous())
    got, want = sampler.sdeint(sde, wild, ts, dt=0.3, bm=bmw), sampler.sdeint(sf, wild, ts, dt=0.3, bm=bmw)
    calm = torch.ones(512, dtype=torch.bool, device=cuda_device)
    calm[::8] = False
    assert_field_close(got[:, calm].cpu().numpy(), want[:, calm].cpu().numpy(), 1e-4, "calm trajectories next to wild ones")
    if impl == _lib.IDFF_IMPL_TENSOR16:
        assert not torch.isfinite(got[-1][~calm]).all(dim=1).all()       # some wild ones overflowed: loud
    else:
        fin = torch.isfinite(want[-1]).all(dim=1)
        assert torch.equal(torch.isfinite(got[-1]).all(dim=1), fin)      # non-finite only where fp32 itself overflows
        sel = fin & ~calm
        if sel.any():
            r = (got[-1][sel] - want[-1][sel]).abs().max().item()
            assert r <= 1e-3 * max(1.0, float(want[-1][sel].abs().max())), r


def test_small_n_route_is_opt_in_and_within_bounds(cuda_device):
    """idff_tune_small_n_route: small order-2 solves on the 32-particle-tile kernel.  Off by default (AUTO = k_tc16 for every N,
    so sub-batches reproduce the whole bit for bit); switched on, the golden bounds still hold and large batches are untouched."""
    g = load_golden("traj2d")
    x0 = torch.tensor(g["x0"]).to(cuda_device)
    pw = packed("checkerboard", cuda_device)
    m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
    raw = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
    raw.impl = _lib.IDFF_IMPL_TENSOR16
    ts = torch.linspace(0, 1, 3)
    base = sampler.sample_rk4(m, x0, ts)
    assert torch.equal(base, sampler.sample_rk4(raw, x0, ts))          # default: k_tc16
    try:
        sampler.set_small_n_route(4736)
        small = sampler.sample_rk4(m, x0, ts)
        _check_drift(small.cpu().numpy(), g["checkerboard_o2_nfe2_g2"][-1] if g["checkerboard_o2_nfe2_g2"].ndim == 3 else g["checkerboard_o2_nfe2_g2"],
                     g["checkerboard_o2_nfe2_g2_f64"], "small-N route")
        assert not torch.equal(small, base)                             # a different kernel rounds differently ...
        assert (small - base).abs().max().item() <= 1e-3
        big = torch.randn(6000, 2, device=cuda_device) / 1.5
        assert torch.equal(sampler.sample_rk4(m, big, ts), sampler.sample_rk4(raw, big, ts))   # ... and is not used above n_max
        m1 = fields.CustomNetModel(pw, 0.2, 0.2, 0.0)                   # one jet
        o1 = sampler.sample_rk4(m1, x0, ts)
    finally:
        sampler.set_small_n_route(0)
    r = drift_report(o1.cpu().numpy(), sampler.sample_rk4(m1, x0, ts).cpu().numpy())
    assert r["median"] <= 2e-5 and r["p999"] <= 2e-3, r


def test_bench_size_64M_particles_properties(cuda_device):
    """The upper end of BASELINE configs[1] and the bench workload (2^26 particles, nfe 2) through the default sampler, checked by
    size-independent properties: the solve is deterministic (two runs, same bits), equals the concatenation of four 2^24-particle
    solves (what four ranks of the particle-sharded launcher would return), any slice reproduces on its own, the output is finite
    and no particle needed the fp32 safety net."""
    from idff_b200.weights import PackedWeights
    pw = PackedWeights.from_npz(load_golden("weights_checkerboard"), device=cuda_device)
    m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2)
    n = 1 << 26
    x0 = torch.randn(n, 2, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(26)) / 1.5
    ts = torch.linspace(0, 1, 3)
    whole = sampler.sample_rk4(m, x0, ts)
    assert torch.isfinite(whole).all()
    again = sampler.sample_rk4(m, x0, ts)
    assert torch.equal(whole, again)
    del again
    q = n // 4
    for r in range(4):
        assert torch.equal(sampler.sample_rk4(m, x0[r * q:(r + 1) * q], ts), whole[r * q:(r + 1) * q])
    for lo, hi in ((n - 70001, n), (3 * q - 33, 3 * q + 31)):
        assert torch.equal(sampler.sample_rk4(m, x0[lo:hi], ts), whole[lo:hi])
    assert (whole.abs() < 6).float().mean().item() > 0.99
    assert _lib.weights_nonfinite(pw.handle) == (0, 0)
