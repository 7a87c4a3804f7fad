"""-m gpu: path sample (bit-exact), lambda, losses and their gradients, through the public Python API (ctypes ->
C-ABI -> sm_100a kernels), against the golden vectors produced by the reference's own code and the CPU oracle."""
import numpy as np
import pytest
import torch

import idff_oracle as O
from conftest import load_golden
from idff_b200 import _lib, losses
from idff_b200.torchcfm import conditional_flow_matching as cfm
from idff_b200.torchcfm import conditional_flow_matching_dynamic as cfmd

pytestmark = pytest.mark.gpu


def _matcher(key, sigma):
    dyn = key.startswith("dyn")
    mod = cfmd if dyn else cfm
    if "_Exact_" in key:
        return cfm.ExactOptimalTransportConditionalFlowMatcher(sigma)
    return mod.ConditionalFlowMatcher(sigma)


def test_path_sample_bit_exact_vs_reference_golden(cuda_device):
    g = load_golden("path")
    keys = sorted({k.rsplit("_", 1)[0] for k in g.files if k.endswith("_xt")})
    assert len(keys) >= 12
    for k in keys:
        sigma = float(k.rsplit("_s", 1)[1])
        FM = _matcher(k, sigma)
        x0, x1, t, eps = (torch.tensor(g[f"{k}_{n}"]).to(cuda_device) for n in ("x0", "x1", "t", "eps"))
        FM.sample_noise_like = lambda x, e=eps: e             # inject the reference's noise draw
        t2, xt, ut, eps2 = FM.sample_location_and_conditional_flow(x0, x1, t=t, return_noise=True)
        np.testing.assert_array_equal(xt.cpu().numpy(), g[k + "_xt"], err_msg=k)
        np.testing.assert_array_equal(ut.cpu().numpy(), g[k + "_ut"], err_msg=k)
        assert t2 is t and eps2 is eps
        np.testing.assert_array_equal(FM.sample_xt(x0, x1, t, eps).cpu().numpy(), g[k + "_xt"])
        lam = FM.compute_lambda(t).cpu().numpy()
        ref = np.broadcast_to(np.asarray(g[k + "_lambda"], dtype=np.float32), lam.shape)
        np.testing.assert_allclose(lam, ref, rtol=2e-7, atol=0)


def test_draw_order_t_then_eps(cuda_device):
    """t comes from the CPU generator (torch.rand) BEFORE eps = randn_like(x0) on x0's device (reference :184,:187)."""
    FM = cfm.ExactOptimalTransportConditionalFlowMatcher(sigma=0.5)
    x0 = torch.randn(64, 46, device=cuda_device)
    x1 = torch.randn(64, 46, device=cuda_device)
    torch.manual_seed(123)
    t, xt, ut, eps = FM.sample_location_and_conditional_flow(x0, x1, return_noise=True)
    torch.manual_seed(123)
    t_ref = torch.rand(64)
    eps_ref = torch.randn_like(x0)
    assert torch.equal(t.cpu(), t_ref) and torch.equal(eps, eps_ref)
    assert FM.x0 is x0 and FM.x1 is x1
    xt_ref, ut_ref = O.path_sample(x0.cpu(), x1.cpu(), t.cpu(), eps.cpu(), 0.5, True, ieee_sqrt=True)
    assert torch.equal(xt.cpu(), xt_ref) and torch.equal(ut.cpu(), ut_ref)
    out = FM.guided_sample_location_and_conditional_flow(x0, x1, y0="a", y1="b")
    assert len(out) == 5 and out[3] == "a" and out[4] == "b"
    with pytest.raises(AssertionError):
        FM.sample_location_and_conditional_flow(x0, x1, t=torch.rand(3, device=cuda_device))


@pytest.mark.parametrize("B,D", [(0, 2), (1, 1), (3, 5), (257, 3), (1000, 46), (4096, 2)])
@pytest.mark.parametrize("mode", [0, 1])
def test_path_sample_ragged_and_gather(cuda_device, B, D, mode):
    g = torch.Generator().manual_seed(B * 100 + D)
    x0, x1, eps = (torch.randn(B, D, generator=g) for _ in range(3))
    t = torch.rand(B, generator=g)
    sigma = 0.3
    xt, ut = cfm._path_sample(x0.to(cuda_device), x1.to(cuda_device), t.to(cuda_device), eps.to(cuda_device), sigma, mode)
    # IEEE sqrt: torch's CPU sqrt (MKL VML) is 1 ulp off in ~0.6% of entries above a few hundred elements
    xr, ur = O.path_sample(x0, x1, t, eps, sigma, bool(mode), ieee_sqrt=True)
    assert torch.equal(xt.cpu(), xr) and torch.equal(ut.cpu(), ur)
    if B > 1:
        i = torch.randint(0, B, (B,), generator=g)
        j = torch.randint(0, B, (B,), generator=g)
        xt, ut = cfm._path_sample(x0.to(cuda_device), x1.to(cuda_device), t.to(cuda_device), eps.to(cuda_device), sigma,
                                  mode, idx0=i.to(cuda_device), idx1=j.to(cuda_device))
        xr, ur = O.path_sample(x0[i], x1[j], t, eps, sigma, bool(mode), ieee_sqrt=True)
        assert torch.equal(xt.cpu(), xr) and torch.equal(ut.cpu(), ur)


def test_dynamic_matcher_repairs_with_plan(cuda_device):
    class FixedPlan:
        def sample_map_indices(self, a, b):
            n = a.shape[0]
            return torch.arange(n - 1, -1, -1, device=a.device), torch.arange(n, device=a.device).roll(1)
    FM = cfmd.ExactOptimalTransportConditionalFlowMatcher(sigma=0.1, ot_sampler=FixedPlan())
    x0 = torch.randn(33, 4, device=cuda_device)
    x1 = torch.randn(33, 4, device=cuda_device)
    t, xt, ut, eps = FM.sample_location_and_conditional_flow(x0, x1, return_noise=True)
    i, j = FixedPlan().sample_map_indices(x0, x1)
    xr, ur = O.path_sample(x0[i].cpu(), x1[j].cpu(), t.cpu(), eps.cpu(), 0.1, True, ieee_sqrt=True)
    assert torch.equal(xt.cpu(), xr) and torch.equal(ut.cpu(), ur)
    assert torch.equal(FM.x1, x1[j])
    lam = FM.compute_lambda(t)
    np.testing.assert_allclose(lam.cpu().numpy(), O.compute_lambda(t.cpu(), 0.1, True, True).numpy(), rtol=2e-7)


def test_path_sample_full_size_properties(cuda_device):
    """B = 2^24, D = 2 (BASELINE.md workload): ut is exactly x1 - x0; sigma = 0 gives the deterministic interpolant;
    the kernel is linear in eps; chunks of the batch reproduce the whole bit-for-bit."""
    B, D = 1 << 24, 2
    g = torch.Generator(device=cuda_device).manual_seed(7)
    x0 = torch.randn(B, D, device=cuda_device, generator=g)
    x1 = torch.randn(B, D, device=cuda_device, generator=g)
    eps = torch.randn(B, D, device=cuda_device, generator=g)
    t = torch.rand(B, device=cuda_device, generator=g)
    xt, ut = cfm._path_sample(x0, x1, t, eps, 0.0, 1)
    assert torch.equal(ut, x1 - x0)
    mu = t[:, None] * x1 + (1 - t[:, None]) * x0
    assert torch.equal(xt, mu + 0.0 * eps)
    xt2, _ = cfm._path_sample(x0, x1, t, eps, 0.5, 0)
    assert torch.equal(xt2, mu + 0.5 * eps)
    lo, hi = 12345, 12345 + (1 << 20) + 3
    xt3, _ = cfm._path_sample(x0[lo:hi], x1[lo:hi], t[lo:hi], eps[lo:hi], 0.5, 0)
    assert torch.equal(xt3, xt2[lo:hi])
    # ExactOT sigma_t against the same expression evaluated by torch on the GPU (IEEE sqrt, separate kernels)
    xt4, _ = cfm._path_sample(x0, x1, t, eps, 0.3, 1)
    assert torch.equal(xt4, mu + (0.3 * torch.sqrt(t * (1 - t)))[:, None] * eps)


def test_path_sample_rejects_cpu_tensors():
    FM = cfm.ConditionalFlowMatcher(0.1)
    with pytest.raises(_lib.IdffError):
        FM.sample_location_and_conditional_flow(torch.randn(4, 2), torch.randn(4, 2))


@pytest.mark.parametrize("k", ["B256_D2", "B10_D46", "B1000_D2"])
def test_losses_and_gradients_vs_golden(cuda_device, k):
    g = load_golden("loss")
    dev = cuda_device
    vt, st, x0, x1, t, xt = (torch.tensor(g[f"{k}_{n}"]).to(dev) for n in ("vt", "st", "x0", "x1", "t", "xt"))
    # tolerance: 1e-6 relative on the loss (fp64 block sums vs torch's fp32 pairwise mean), 2e-6 on gradients
    v = vt.clone().requires_grad_(True)
    l = losses.flow_loss(v, x1, t)
    l.backward()
    assert l.item() == pytest.approx(float(g[f"{k}_flow"]), rel=1e-6)
    np.testing.assert_allclose(v.grad.cpu().numpy(), g[f"{k}_flow_gvt"], rtol=2e-6, atol=1e-12)
    s = st.clone().requires_grad_(True)
    l = losses.score_loss(s, xt, x1, x0, t, 0.2)
    (l * 3.0).backward()
    assert l.item() == pytest.approx(float(g[f"{k}_score"]), rel=2e-6)
    np.testing.assert_allclose(s.grad.cpu().numpy() / 3.0, g[f"{k}_score_gst"], rtol=5e-6, atol=1e-9)
    v = (vt * 100).clone().requires_grad_(True)
    l = losses.md_periodic_loss(v, x1 * 90)
    l.backward()
    assert l.item() == pytest.approx(float(g[f"{k}_md"]), rel=1e-6)
    # the +-360 terms cancel in the gradient: absolute tolerance relative to the largest entry
    gm = g[f"{k}_md_gvt"]
    np.testing.assert_allclose(v.grad.cpu().numpy(), gm, rtol=2e-6, atol=2e-7 * np.abs(gm).max())
    # no-grad call: value only, repeated calls reuse the scratch deterministically
    with torch.no_grad():
        a = losses.flow_loss(vt, x1, t).item()
        b = losses.flow_loss(vt, x1, t).item()
    assert a == b == pytest.approx(float(g[f"{k}_flow"]), rel=1e-6)


def test_loss_large_batch_matches_fp64(cuda_device):
    B, D = 1 << 22, 2
    g = torch.Generator(device=cuda_device).manual_seed(3)
    vt = torch.randn(B, D, device=cuda_device, generator=g)
    x1 = torch.randn(B, D, device=cuda_device, generator=g)
    t = torch.rand(B, device=cuda_device, generator=g)
    got = losses.flow_loss(vt, x1, t).item()
    w = 1.0 / (np.float32(1.01) - t.double()[:, None])
    ref = ((w * (vt.double() - x1.double())) ** 2).mean().item()
    assert got == pytest.approx(ref, rel=2e-6)


def _fresh_mlp(device, seed=0):
    from idff_b200.torchcfm.models.models import MLP
    torch.manual_seed(seed)
    return MLP(dim=2, w=256, time_varying=True).to(device)


def test_train_step_matches_reference_formula(cuda_device):
    """One step of Autograd_example_order2.py:90-113 through the kernels (path sample + loss value/gradient) against
    the same step written with the reference's torch expressions, same weights and draws: parameters after AdamW agree."""
    from idff_b200 import train
    dev = cuda_device
    B = 256
    g = torch.Generator().manual_seed(5)
    x1 = train.checkerboard_device(B, dev)
    x0 = torch.randn(B, 2, generator=g).to(dev)
    t = torch.rand(B, generator=g).to(dev)
    ours, ref = _fresh_mlp(dev), _fresh_mlp(dev)
    FM = cfm.ExactOptimalTransportConditionalFlowMatcher(sigma=0.0)
    opt_o = torch.optim.AdamW(ours.parameters(), lr=2e-4)
    opt_r = torch.optim.AdamW(ref.parameters(), lr=2e-4)
    loss_o = train.flow_train_step(ours, opt_o, FM, x0, x1, t=t)
    # the reference's expressions (sigma = 0: xt is the interpolant)
    opt_r.zero_grad()
    xt = t[:, None] * x1 + (1 - t[:, None]) * x0
    vt = ref(torch.cat([xt, t[:, None]], dim=-1))
    loss_r = torch.mean(((1 / (1e-2 + 1 - t[:, None])) * (vt - x1)) ** 2)
    loss_r.backward()
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 1)
    opt_r.step()
    assert loss_o.item() == pytest.approx(loss_r.item(), rel=2e-6)
    for po, pr in zip(ours.parameters(), ref.parameters()):
        # AdamW's first step moves every weight by ~lr * sign(grad): compare the updates, not just the weights
        np.testing.assert_allclose(po.detach().cpu().numpy(), pr.detach().cpu().numpy(), rtol=0, atol=2e-7)


def test_graphed_trainer_replays_and_learns(cuda_device):
    """The captured step (data draw, path-sample kernel, MLP fwd/bwd, loss kernel, clip, AdamW) replays as one graph
    launch, launches our two kernels per step, and drives the loss down like the eager loop does."""
    from idff_b200 import train
    model = _fresh_mlp(cuda_device, seed=1)
    before = [p.detach().clone() for p in model.parameters()]
    tr = train.GraphedFlowTrainer(model, batch_size=256, lr=2e-3)
    first = []
    for _ in range(20):
        first.append(tr.step().item())
    for _ in range(400):
        tr.step()
    last = []
    for _ in range(20):
        last.append(tr.step().item())
    assert tr.steps == 440 and all(np.isfinite(first + last))
    assert np.mean(last) < 0.7 * np.mean(first), (np.mean(first), np.mean(last))
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, model.parameters()))


def test_md_train_step_matches_reference_formula(cuda_device):
    """One step of TrajectoryGenerator.train_model (MD_simulation.py:126-144) through the kernels (path sample, periodic loss
    value + gradient) against the same step written with the reference's torch expressions: loss and Adam-updated parameters."""
    from idff_b200 import train
    from idff_b200.torchcfm.models.models import MLP_Embedding
    dev = cuda_device
    T, D, B = 200, 46, 10
    g = torch.Generator().manual_seed(3)
    data = (180 * torch.sin(torch.cumsum(torch.randn(T, D, generator=g) * 0.05, 0))).to(dev)
    idx = torch.randint(1, T, (B,), generator=g).to(dev)
    t = torch.rand(B, generator=g).to(dev)
    torch.manual_seed(0)
    ours = MLP_Embedding(dim=D, w=128, time_embed=T, time_varying=True).to(dev)
    torch.manual_seed(0)
    ref = MLP_Embedding(dim=D, w=128, time_embed=T, time_varying=True).to(dev)
    FM = cfm.ExactOptimalTransportConditionalFlowMatcher(sigma=0.5)
    eps = torch.randn(B, D, generator=g).to(dev)
    FM.sample_noise_like = lambda x, e=eps: e
    opt_o, opt_r = torch.optim.Adam(ours.parameters()), torch.optim.Adam(ref.parameters())
    loss_o = train.md_train_step(ours, opt_o, FM, data, idx, t=t)
    opt_r.zero_grad()
    x0, x1 = data[idx - 1], data[idx]
    st = 0.5 * torch.sqrt(t * (1 - t))
    xt = t[:, None] * x1 + (1 - t[:, None]) * x0 + st[:, None] * eps
    vt = ref.net(torch.cat([xt, t[:, None], ref.time_index_embedding(idx)], dim=-1))
    loss_r = torch.mean((vt - x1) ** 2) + torch.mean((vt - (x1 + 360)) ** 2) + torch.mean((vt - (x1 - 360)) ** 2)
    loss_r.backward()
    opt_r.step()
    assert loss_o.item() == pytest.approx(loss_r.item(), rel=2e-6)
    for po, pr in zip(ours.parameters(), ref.parameters()):
        # Adam's first step moves a weight by lr * g / (|g| + 1e-8): fp32 noise in a near-zero gradient can flip it
        d = np.abs(po.detach().cpu().numpy() - pr.detach().cpu().numpy()).reshape(-1)
        assert np.quantile(d, 0.999) <= 2e-6 and d.max() <= 2.1e-3, (d.max(), np.quantile(d, 0.999))


def test_graphed_md_trainer_learns(cuda_device):
    from idff_b200 import train
    from idff_b200.torchcfm.models.models import MLP_Embedding
    dev = cuda_device
    T, D = 200, 46
    g = torch.Generator().manual_seed(4)
    data = (180 * torch.sin(torch.cumsum(torch.randn(T, D, generator=g) * 0.05, 0))).to(dev)
    torch.manual_seed(1)
    model = MLP_Embedding(dim=D, w=128, time_embed=T, time_varying=True).to(dev)
    tr = train.GraphedMDTrainer(model, data, batch_size=10, sigma=0.5)
    first = [tr.step().item() for _ in range(20)]
    for _ in range(600):
        tr.step()
    last = [tr.step().item() for _ in range(20)]
    assert np.isfinite(first + last).all() and np.mean(last) < np.mean(first), (np.mean(first), np.mean(last))


def test_scorehead_train_step_matches_reference_formula(cuda_device):
    """One step of example_order1_with_prediction_head.py:111-137 (flow loss on the first two outputs + score regression on the
    last two) through the kernels against the reference's torch expressions: loss and AdamW-updated parameters; and the same
    step captured as a CUDA graph learns."""
    from idff_b200 import train
    from idff_b200.torchcfm.models.models import MLP
    dev = cuda_device
    B, sig = 256, 0.2
    g = torch.Generator().manual_seed(6)
    x1 = (torch.randn(B, 2, generator=g) * 2).to(dev)
    x0 = torch.randn(B, 2, generator=g).to(dev)
    t = torch.rand(B, generator=g).to(dev)
    torch.manual_seed(3)
    ours = MLP(dim=4, w=256, time_varying=True).to(dev)
    torch.manual_seed(3)
    ref = MLP(dim=4, w=256, time_varying=True).to(dev)
    FM = cfm.ExactOptimalTransportConditionalFlowMatcher(sigma=0.0)
    opt_o, opt_r = torch.optim.AdamW(ours.parameters(), lr=1e-3), torch.optim.AdamW(ref.parameters(), lr=1e-3)
    loss_o = train.scorehead_train_step(ours, opt_o, FM, x0, x1, sig, t=t)
    opt_r.zero_grad()
    xt = t[:, None] * x1 + (1 - t[:, None]) * x0
    outs = ref.net(torch.cat([xt, xt, t[:, None]], dim=-1))
    vt, st = outs[:, :2], outs[:, 2:]
    grads = O.log_prob_xt_given_x1_dimwise(xt.cpu(), x1.cpu(), x0.cpu(), t.cpu(), sig).to(dev)
    loss_r = torch.mean(((1 / (1e-2 + 1 - t[:, None])) * (vt - x1)) ** 2) + torch.mean((st - grads) ** 2)
    loss_r.backward()
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 1)
    opt_r.step()
    assert loss_o.item() == pytest.approx(loss_r.item(), rel=5e-6)
    for po, pr in zip(ours.parameters(), ref.parameters()):
        d = np.abs(po.detach().cpu().numpy() - pr.detach().cpu().numpy()).reshape(-1)
        assert np.quantile(d, 0.999) <= 2e-6 and d.max() <= 2.1e-3, (d.max(), np.quantile(d, 0.999))
    torch.manual_seed(4)
    tr = train.GraphedFlowTrainer(MLP(dim=4, w=256, time_varying=True).to(dev), batch_size=256, lr=1e-3, score_sigma=sig)
    first = [tr.step().item() for _ in range(20)]
    for _ in range(300):
        tr.step()
    last = [tr.step().item() for _ in range(20)]
    assert np.isfinite(first + last).all() and np.mean(last) < np.mean(first)


def test_graph_trained_weights_reach_the_sampler(cuda_device):
    """A CUDA-graph replay updates the parameters without touching their version counters; the trainers bump them, so the
    sampler's packed-weights cache (keyed on data_ptr and _version) is rebuilt after further training."""
    from idff_b200 import fields, sampler, train
    from idff_b200.weights import packed_for
    model = _fresh_mlp(cuda_device, seed=2)
    tr = train.GraphedFlowTrainer(model, batch_size=256, lr=1e-2)
    x0 = torch.randn(512, 2, device=cuda_device)
    ts = torch.linspace(0, 1, 3)
    f = fields.CustomNetModel(model, 0.2, 0.2, 0.0)
    a = sampler.sample_rk4(f, x0, ts).clone()
    pw_a = packed_for(model, cuda_device)
    for _ in range(50):
        tr.step()
    b = sampler.sample_rk4(f, x0, ts)
    assert packed_for(model, cuda_device) is not pw_a
    assert not torch.equal(a, b)
    ref = O.odeint_rk4(O.Field2D([(m.weight.detach().cpu(), m.bias.detach().cpu()) for m in model.net
                                  if isinstance(m, torch.nn.Linear)], 0.2, (0.2, 0.0), 2), x0.cpu(), ts)[-1]
    assert (b.cpu() - ref).abs().median().item() < 1e-4


def test_aux_matchers_vs_reference_golden(cuda_device):
    """Target / SchrodingerBridge / VariancePreserving matchers (no IDFF configuration uses them; plain torch ops around the
    path-sample kernel) against outputs of the reference's own classes on the same inputs and injected noise."""
    g = load_golden("matchers")
    for key in ("B64_D2", "B10_D46"):
        x0, x1, eps, t = (torch.tensor(g[f"{key}_{n}"]).to(cuda_device) for n in ("x0", "x1", "eps", "t"))
        for name, FM in (("target", cfm.TargetConditionalFlowMatcher(0.1)), ("sb", cfm.SchrodingerBridgeConditionalFlowMatcher(0.5)),
                         ("vp", cfm.VariancePreservingConditionalFlowMatcher(0.2))):
            FM.sample_noise_like = lambda x, e=eps: e
            t2, xt, ut = FM.sample_location_and_conditional_flow(x0, x1, t=t)
            np.testing.assert_allclose(xt.cpu().numpy(), g[f"{key}_{name}_xt"], rtol=2e-6, atol=2e-6, err_msg=f"{key} {name} xt")
            # the SB flow divides by 2 t (1 - t) + 1e-8: near t = 0 / 1 it amplifies the last-ulp differences of xt - mu_t
            np.testing.assert_allclose(ut.cpu().numpy(), g[f"{key}_{name}_ut"], rtol=2e-4 if name == "sb" else 2e-6,
                                       atol=2e-4 if name == "sb" else 2e-6, err_msg=f"{key} {name} ut")
    with pytest.raises(ValueError):
        cfm.SchrodingerBridgeConditionalFlowMatcher(0.0)
    # compute_lambda and the path noise go through compute_sigma_t (reference :195-211), also for subclasses
    # that override only that method -- the kernel serves the two sigma_t forms it knows, everything else is torch
    t = torch.rand(64, device=cuda_device)
    FM = cfm.TargetConditionalFlowMatcher(0.1)
    np.testing.assert_allclose(FM.compute_lambda(t).cpu().numpy(),
                               (2 * (1 - (1 - 0.1) * t) ** 2 / (0.1 ** 2 + 1e-8)).cpu().numpy(), rtol=1e-6)

    class Halved(cfm.ConditionalFlowMatcher):
        def compute_sigma_t(self, t):
            return 0.5 * self.sigma * (1 - t)

    FM = Halved(0.3)
    assert FM._kernel_sigma_mode() is None and not FM._fusable()
    x0, x1, eps = (torch.randn(64, 2, device=cuda_device) for _ in range(3))
    FM.sample_noise_like = lambda x: eps
    _, xt, ut = FM.sample_location_and_conditional_flow(x0, x1, t=t)
    tp = t[:, None]
    np.testing.assert_allclose(xt.cpu().numpy(), (tp * x1 + (1 - tp) * x0 + 0.5 * 0.3 * (1 - tp) * eps).cpu().numpy(), rtol=1e-6, atol=1e-6)
    assert torch.equal(ut, x1 - x0)
    np.testing.assert_allclose(FM.compute_lambda(t).cpu().numpy(),
                               (2 * (0.15 * (1 - t)) ** 2 / (0.3 ** 2 + 1e-8)).cpu().numpy(), rtol=1e-6)
    assert cfm.ExactOptimalTransportConditionalFlowMatcher(0.5)._kernel_sigma_mode() == 1 and cfm.ConditionalFlowMatcher(0.5)._fusable()


@pytest.mark.parametrize("shape", [(256, 2), (10, 46), (1, 1), (4097, 3), (1 << 24, 2)])
def test_in_kernel_noise_is_randn_like_bit_for_bit(cuda_device, shape):
    """the fused path sample draws eps inside the kernel (Philox4x32-10 in ATen's element layout) -- the very
    tensor torch.randn_like(x0) returns for the same seed, the generator left where randn_like leaves it, xt / ut bit-exact
    against the reference's expression on that noise (torchcfm/conditional_flow_matching.py:184-189)."""
    B, D = shape
    g = torch.Generator(device=cuda_device).manual_seed(B + D)
    x0 = torch.randn(B, D, device=cuda_device, generator=g)
    x1 = torch.randn(B, D, device=cuda_device, generator=g)
    for sigma, FMc in ((0.5, cfm.ExactOptimalTransportConditionalFlowMatcher), (0.3, cfm.ConditionalFlowMatcher)):
        FM = FMc(sigma=sigma)
        torch.manual_seed(1234 + B)
        pre = torch.randn(7, device=cuda_device)                      # the generator is mid-stream, not at offset 0
        before = _lib.launch_count()
        t, xt, ut, eps = FM.sample_location_and_conditional_flow(x0, x1, return_noise=True)
        assert _lib.launch_count() == before + 1                      # ONE kernel: noise, xt and ut
        after = torch.randn(5, device=cuda_device)                    # what the next torch draw sees
        torch.manual_seed(1234 + B)
        assert torch.equal(torch.randn(7, device=cuda_device), pre)
        t_ref = torch.rand(B).to(cuda_device)
        eps_ref = torch.randn_like(x0)
        assert torch.equal(t, t_ref)
        assert torch.equal(eps, eps_ref), (shape, (eps != eps_ref).sum().item())
        assert torch.equal(torch.randn(5, device=cuda_device), after)  # same generator offset afterwards
        xt_ref, ut_ref = cfm._path_sample(x0, x1, t, eps_ref, sigma, FM._kernel_sigma_mode())   # the injected-eps kernel (bit-exact
        assert torch.equal(xt, xt_ref) and torch.equal(ut, ut_ref)                               # vs the oracle: tests above)
        # without return_noise the same xt comes back and eps never exists
        torch.manual_seed(1234 + B)
        torch.randn(7, device=cuda_device)
        t2, xt2, ut2 = FM.sample_location_and_conditional_flow(x0, x1)
        assert torch.equal(t2, t) and torch.equal(xt2, xt) and torch.equal(ut2, ut)
    if B <= 4097:
        xr, ur = O.path_sample(x0.cpu(), x1.cpu(), t.cpu(), eps.cpu(), 0.3, False, ieee_sqrt=True)
        assert torch.equal(xt.cpu(), xr) and torch.equal(ut.cpu(), ur)
