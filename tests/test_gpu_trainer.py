"""-m gpu: the fused training kernel (idff_train_flow_steps) against the reference's training loop
(Autograd_example_order2.py:89-112) run by the CPU oracle on the same weights and draws: per-step losses, gradient norms,
gradients, and the parameters / AdamW moments after a run of steps."""
import numpy as np
import pytest
import torch

import idff_oracle as O
from idff_b200 import _lib, train

pytestmark = pytest.mark.gpu

LR = 2e-4


def _fresh_mlp(device, seed=0):
    from idff_b200.torchcfm.models.models import MLP
    torch.manual_seed(seed)
    return MLP(dim=2, w=256, time_varying=True).to(device)


def _layers_of(model, dtype=torch.float32):
    ps = [p.detach().cpu().to(dtype).clone() for p in model.parameters()]
    return [(ps[2 * i], ps[2 * i + 1]) for i in range(4)]


def _draws(S, B, seed, sigma):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.as_tensor(O.checkerboard(S * B), dtype=torch.float32).view(S, B, 2) if hasattr(O, "checkerboard") else None
    x0 = torch.randn(S, B, 2, generator=g)
    t = torch.rand(S, B, generator=g)
    eps = torch.randn(S, B, 2, generator=g) if sigma != 0 else None
    return x0, x1.contiguous(), t, eps


@pytest.mark.parametrize("B,sigma", [(256, 0.0), (64, 0.5), (512, 0.0), (4, 0.0)])
def test_one_step_gradients_loss_and_update(cuda_device, B, sigma):
    """One step: unclipped gradients within 1e-5 * max|g| per tensor of autograd's, loss and gradient norm to 2e-6
    relative, AdamW's first update (~lr * sign(g)) to 1e-3 * lr wherever |g| is above the fp32 noise floor."""
    np.random.seed(3)
    dev = cuda_device
    model = _fresh_mlp(dev, seed=B)
    layers = _layers_of(model)
    x0, x1, t, eps = _draws(1, B, 11, sigma)
    losses_r, norms_r, after_r, grads_r = O.train_steps_2d(layers, x0, x1, t, eps, sigma=sigma, lr=LR, return_grads=True)
    tr = train.FusedFlowTrainer(model, batch_size=B, sigma=sigma, lr=LR)
    dump = torch.zeros(tr.n_params, device=dev)
    before = _lib.launch_count()
    loss, gnorm = tr.steps(x0.to(dev), x1.to(dev), t.to(dev), eps.to(dev) if eps is not None else None, return_gnorm=True,
                           grad_dump=dump)
    assert _lib.launch_count() - before == 1
    assert loss.item() == pytest.approx(losses_r[0], rel=2e-6)
    assert gnorm.item() == pytest.approx(norms_r[0], rel=2e-6)
    shapes = [tuple(p.shape) for p in model.parameters()]
    for o, sh, gr in zip(tr.offsets[:8], shapes, grads_r):
        got = dump[o: o + gr.numel()].view(sh).cpu().numpy()
        ref = gr.numpy()
        assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-12), (sh, np.abs(got - ref).max(), np.abs(ref).max())
    coef = min(1.0, 1.0 / (norms_r[0] + 1e-6))
    flat_r = torch.cat([torch.cat([W.reshape(-1), b.reshape(-1)]) for W, b in after_r]).numpy()
    flat_g = torch.cat([g.reshape(-1) for g in grads_r]).numpy() * coef
    flat_o = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().numpy()
    d = np.abs(flat_o - flat_r)
    solid = np.abs(flat_g) > 1e-6        # m/(sqrt(v)+1e-8) is insensitive to fp32 noise in g there
    assert d[solid].max() <= 1e-3 * LR, d[solid].max()
    assert d.max() <= 2 * LR


def test_many_steps_track_the_reference(cuda_device):
    """40 consecutive steps in ONE launch: losses within 1e-4 relative of the reference loop at every step (fp32 noise
    compounds through AdamW's normalised update), parameters within a few lr; median parameter difference << lr."""
    np.random.seed(4)
    dev = cuda_device
    S, B = 40, 256
    model = _fresh_mlp(dev, seed=7)
    layers = _layers_of(model)
    x0, x1, t, eps = _draws(S, B, 21, 0.0)
    losses_r, norms_r, after_r = O.train_steps_2d(layers, x0, x1, t, eps, sigma=0.0, lr=LR)
    tr = train.FusedFlowTrainer(model, batch_size=B, lr=LR)
    before = _lib.launch_count()
    loss, gnorm = tr.steps(x0.to(dev), x1.to(dev), t.to(dev), return_gnorm=True)
    assert _lib.launch_count() - before == 1 and tr.steps_done == S
    np.testing.assert_allclose(loss.cpu().numpy(), np.array(losses_r, dtype=np.float32), rtol=1e-4)
    np.testing.assert_allclose(gnorm.cpu().numpy(), np.array(norms_r, dtype=np.float32), rtol=1e-3)
    flat_r = torch.cat([torch.cat([W.reshape(-1), b.reshape(-1)]) for W, b in after_r]).numpy()
    flat_o = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().numpy()
    d = np.abs(flat_o - flat_r)
    moved = np.abs(flat_r - torch.cat([torch.cat([W.reshape(-1), b.reshape(-1)]) for W, b in layers]).numpy())
    print(f"after {S} steps: median|dp| {np.median(d):.2e} p99.9 {np.quantile(d, 0.999):.2e} max {d.max():.2e}; "
          f"median movement {np.median(moved):.2e}")
    assert np.median(d) <= 1e-2 * LR and np.quantile(d, 0.999) <= LR and d.max() <= 5 * LR
    assert np.median(moved) > 10 * np.median(d)


def test_split_launches_are_bit_identical_and_deterministic(cuda_device):
    """20 steps in one launch == 12 + 8 steps in two launches == a second run, bit for bit (fixed reduction orders;
    steps_done carries the AdamW bias correction across launches)."""
    np.random.seed(5)
    dev = cuda_device
    S, B = 20, 128
    x0, x1, t, eps = (v.to(dev) if v is not None else None for v in _draws(S, B, 31, 0.5))
    outs = []
    for split in ((S,), (12, 8), (S,)):
        model = _fresh_mlp(dev, seed=9)
        tr = train.FusedFlowTrainer(model, batch_size=B, sigma=0.5, lr=LR)
        ls, a = [], 0
        for n in split:
            ls.append(tr.steps(x0[a:a + n].contiguous(), x1[a:a + n].contiguous(), t[a:a + n].contiguous(),
                               eps[a:a + n].contiguous()))
            a += n
        outs.append((torch.cat(ls).cpu(), tr.state[: 3 * tr.offsets[10]].cpu().clone()))
    for l, st in outs[1:]:
        assert torch.equal(l, outs[0][0]) and torch.equal(st, outs[0][1])


@pytest.mark.parametrize("B,modes", [(256, (0, 1)), (64, (0, 1))])
def test_pair_and_single_cta_modes_agree(cuda_device, B, modes):
    """idff_debug_train_pair_mode: a row group on a CTA pair (cluster of 2, activation halves exchanged through distributed
    shared memory; mode 1, the default up to B = 296) against one CTA per 4 rows (mode 0): same algorithm, K-slices summed in a different fixed order -> losses agree to 1e-5
    relative, parameters to fp32 noise through 30 AdamW steps."""
    np.random.seed(6)
    dev = cuda_device
    S = 30
    x0, x1, t, eps = (v.to(dev) if v is not None else None for v in _draws(S, B, 41, 0.5))
    res = []
    for pair in modes:
        _lib.check(_lib.lib().idff_debug_train_pair_mode(pair), "pair_mode")
        try:
            model = _fresh_mlp(dev, seed=13)
            tr = train.FusedFlowTrainer(model, batch_size=B, sigma=0.5, lr=LR)
            loss = tr.steps(x0, x1, t, eps)
            res.append((loss.cpu().numpy(), torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().numpy()))
        finally:
            _lib.check(_lib.lib().idff_debug_train_pair_mode(1), "pair_mode")
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-5)
    d = np.abs(res[0][1] - res[1][1])
    assert np.median(d) <= 1e-2 * LR and d.max() <= 5 * LR, (np.median(d), d.max())


def test_model_aliases_state_and_feeds_the_sampler(cuda_device):
    """The torch module's parameters are views of the kernel state: after training, module forward, state_dict and the
    sampler's packed weights all see the trained network; the loss goes down on the checkerboard."""
    from idff_b200 import fields, sampler
    from idff_b200.weights import PackedWeights
    dev = cuda_device
    model = _fresh_mlp(dev, seed=2)
    w_before = model.net[0].weight.detach().clone()
    tr = train.FusedFlowTrainer(model, batch_size=256, lr=2e-3)
    first = tr.run(50).cpu().numpy()
    tr.run(1500)
    last = tr.run(50).cpu().numpy()
    assert np.isfinite(first).all() and np.isfinite(last).all()
    assert last.mean() < 0.7 * first.mean(), (first.mean(), last.mean())
    assert not torch.equal(w_before, model.net[0].weight.detach())
    assert model.net[0].weight.data_ptr() == tr.state.data_ptr() + 4 * (tr.offsets[9] + tr.offsets[0])
    x = torch.randn(64, 2, device=dev)
    tt = torch.full((64, 1), 0.3, device=dev)
    ref = O.mlp_forward(_layers_of(model), torch.cat([x, tt], 1).cpu())
    np.testing.assert_allclose(model(torch.cat([x, tt], 1)).detach().cpu().numpy(), ref.numpy(), rtol=1e-4, atol=1e-5)
    sd = tr.state_dict()
    assert set(sd) == set(model.state_dict()) and all(v.untyped_storage().nbytes() == v.numel() * 4 for v in sd.values())
    fresh = _fresh_mlp(dev, seed=99)
    fresh.load_state_dict(sd)                             # the checkpoint format of the reference (:116)
    assert torch.equal(fresh.net[4].weight, model.net[4].weight)
    pw = PackedWeights.from_module(model, device=dev)
    f = fields.CustomNetModel(pw, 0.2, 0.2, 0.0)
    out = sampler.sample_rk4(f, x, torch.linspace(0, 1, 3))
    assert torch.isfinite(out).all()


def test_external_parameter_writes_are_picked_up(cuda_device):
    """load_state_dict (an in-place copy into the aliased parameters) between launches: the next launch rebuilds the kernel's
    streaming weight copies by itself and trains the loaded network -- identical to a trainer built on that network."""
    np.random.seed(8)
    dev = cuda_device
    x0, x1, t, _ = (v.to(dev) if v is not None else None for v in _draws(3, 64, 51, 0.0))
    donor = _fresh_mlp(dev, seed=21)
    a = train.FusedFlowTrainer(_fresh_mlp(dev, seed=20), batch_size=64, lr=LR)
    a.steps(x0, x1, t)                                   # some history on other weights ...
    a.model.load_state_dict(donor.state_dict())          # ... then the weights are replaced from outside
    a.state[a.offsets[10]: a.offsets[11] + a.n_params].zero_()               # fresh optimizer state (exp_avg, exp_avg_sq)
    a.steps_done = 0
    la = a.steps(x0, x1, t)
    b = train.FusedFlowTrainer(_fresh_mlp(dev, seed=21), batch_size=64, lr=LR)
    lb = b.steps(x0, x1, t)
    assert torch.equal(la, lb)
    for pa, pb in zip(a.model.parameters(), b.model.parameters()):
        assert torch.equal(pa, pb)


def test_rejects_unsupported_shapes(cuda_device):
    dev = cuda_device
    tr = train.FusedFlowTrainer(_fresh_mlp(dev), batch_size=256)
    for B in (6, 1024, 2):
        with pytest.raises(_lib.IdffError):
            tr.steps(torch.zeros(1, B, 2, device=dev), torch.zeros(1, B, 2, device=dev), torch.zeros(1, B, device=dev))
    with pytest.raises(_lib.IdffError):
        tr.steps(torch.zeros(1, 8, 2), torch.zeros(1, 8, 2), torch.zeros(1, 8))          # CPU tensors: no fallback
    from idff_b200.torchcfm.models.models import MLP
    with pytest.raises(_lib.IdffError):
        train.FusedFlowTrainer(MLP(dim=2, w=128, time_varying=True).to(dev))


def test_eight_gaussians_device_sampler_and_training(cuda_device):
    """BASELINE configs[0] (8gaussians): the on-device data draw has the reference's law (8 modes at radius 5, std 0.1**0.25)
    and the fused trainer learns on it."""
    x = train.eight_gaussians_device(200000, cuda_device).cpu()
    rad = x.norm(dim=1)
    assert abs(rad.mean().item() - 5.0) < 0.1
    ang = torch.atan2(x[:, 1], x[:, 0])
    mode = torch.round(ang / (np.pi / 4)) % 8
    counts = torch.bincount(mode.long(), minlength=8).float() / len(x)
    assert (counts - 0.125).abs().max().item() < 0.01
    ref = O.eight_gaussians(200000)
    assert abs(x.std(0).mean().item() - ref.std(0).mean().item()) < 0.05
    tr = train.FusedFlowTrainer(_fresh_mlp(cuda_device, seed=5), batch_size=256, lr=2e-3, data=train.eight_gaussians_device)
    first = tr.run(50).cpu().numpy()
    tr.run(1500)
    last = tr.run(50).cpu().numpy()
    assert np.isfinite(last).all() and last.mean() < 0.7 * first.mean(), (first.mean(), last.mean())


def test_train_then_sample_end_to_end(cuda_device):
    """BASELINE configs[0] end to end on the device: train the 2-D MLP on 8gaussians with the fused kernel (10 000 steps, a
    quarter of a second), wrap it in the order-2 IDFF field, solve with rk4 and evaluate the MMD -- every stage one of this
    repo's kernels.  The untrained prior sits at MMD 0.6; the reference's shipped checkpoint at 0.012."""
    from idff_b200 import fields, sampler
    from idff_b200.mmd import compute_mmd_rbf
    dev = cuda_device
    model = _fresh_mlp(dev, seed=0)
    tr = train.FusedFlowTrainer(model, batch_size=256, lr=2e-4, data=train.eight_gaussians_device)
    for _ in range(10):
        loss = tr.run(1000)
    assert torch.isfinite(loss).all()
    g = torch.Generator(device=dev).manual_seed(1)
    true_s = train.eight_gaussians_device(2048, dev, generator=g)
    x0 = torch.randn(2048, 2, device=dev, generator=g) / 1.5
    out = sampler.sample_rk4(fields.CustomNetModel(model, 0.2, 0.2, 0.0), x0, torch.linspace(0, 1, 11))
    prior, got = compute_mmd_rbf(true_s, x0, 1.0), compute_mmd_rbf(true_s, out, 1.0)
    print(f"8gaussians, 10000 fused training steps, nfe 10: MMD {got:.4f} (prior {prior:.3f})")
    assert prior > 0.4 and got < 0.06


def _fresh_scorehead(device, seed=0):
    from idff_b200.torchcfm.models.models import MLP
    torch.manual_seed(seed)
    return MLP(dim=4, w=256, time_varying=True).to(device)     # example_order1_with_prediction_head.py:101


@pytest.mark.parametrize("B,sigma", [(256, 0.0), (64, 0.5), (512, 0.0), (4, 0.3)])
def test_scorehead_step_gradients_loss_and_update(cuda_device, B, sigma):
    """The order-1 script's step (example_order1_with_prediction_head.py:111-137: MLP(dim=4) on cat[xt, xt, t], flow loss +
    score regression, clip, AdamW lr 1e-3) in the persistent kernel against the oracle's loop on the same weights and draws --
    the bounds of the flow-network test."""
    np.random.seed(5)
    dev, lr, s0 = cuda_device, 1e-3, 0.2
    model = _fresh_scorehead(dev, seed=B)
    layers = _layers_of(model)
    x0, x1, t, eps = _draws(1, B, 13, sigma)
    losses_r, norms_r, after_r, grads_r = O.train_steps_2d(layers, x0, x1, t, eps, sigma=sigma, lr=lr, return_grads=True, score_sigma=s0)
    tr = train.FusedFlowTrainer(model, batch_size=B, sigma=sigma, lr=lr, score_sigma=s0)
    dump = torch.zeros(tr.n_params, device=dev)
    before = _lib.launch_count()
    loss, gnorm = tr.steps(x0.to(dev), x1.to(dev), t.to(dev), eps.to(dev) if eps is not None else None, return_gnorm=True,
                           grad_dump=dump)
    assert _lib.launch_count() - before == 1
    assert loss.item() == pytest.approx(losses_r[0], rel=2e-6)
    assert gnorm.item() == pytest.approx(norms_r[0], rel=2e-6)
    shapes = [tuple(p.shape) for p in model.parameters()]
    assert shapes[0] == (256, 5) and shapes[6] == (4, 256)
    for o, sh, gr in zip(tr.offsets[:8], shapes, grads_r):
        got = dump[o: o + gr.numel()].view(sh).cpu().numpy()
        ref = gr.numpy()
        assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-12), (sh, np.abs(got - ref).max(), np.abs(ref).max())
    coef = min(1.0, 1.0 / (norms_r[0] + 1e-6))
    flat_r = torch.cat([torch.cat([W.reshape(-1), b.reshape(-1)]) for W, b in after_r]).numpy()
    flat_g = torch.cat([g.reshape(-1) for g in grads_r]).numpy() * coef
    flat_o = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().numpy()
    d = np.abs(flat_o - flat_r)
    solid = np.abs(flat_g) > 1e-6
    assert d[solid].max() <= 1e-3 * lr, d[solid].max()
    assert d.max() <= 2 * lr


def test_scorehead_many_steps_and_field(cuda_device):
    """30 consecutive score-head steps in one launch track the oracle's loop; the trained module then drives the score-head field
    through the tcgen05 sampler (the packed weights follow the kernel-trained parameters)."""
    from idff_b200 import fields, sampler
    np.random.seed(6)
    dev, lr, s0, S, B = cuda_device, 1e-3, 0.2, 30, 256
    model = _fresh_scorehead(dev, seed=3)
    layers = _layers_of(model)
    x0, x1, t, eps = _draws(S, B, 23, 0.0)
    losses_r, norms_r, after_r = O.train_steps_2d(layers, x0, x1, t, eps, sigma=0.0, lr=lr, score_sigma=s0)
    tr = train.FusedFlowTrainer(model, batch_size=B, lr=lr, score_sigma=s0)
    loss = tr.steps(x0.to(dev), x1.to(dev), t.to(dev), None).cpu().numpy()
    np.testing.assert_allclose(loss, np.array(losses_r), rtol=2e-4)
    flat_r = torch.cat([torch.cat([W.reshape(-1), b.reshape(-1)]) for W, b in after_r]).numpy()
    flat_o = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().numpy()
    d = np.abs(flat_o - flat_r)
    assert np.median(d) <= 1e-6 and d.max() <= 5 * lr, (np.median(d), d.max())
    # bit-reproducible, independent of the launch split
    model2 = _fresh_scorehead(dev, seed=3)
    tr2 = train.FusedFlowTrainer(model2, batch_size=B, lr=lr, score_sigma=s0)
    la = tr2.steps(x0[:11].to(dev), x1[:11].to(dev), t[:11].to(dev), None)
    lb = tr2.steps(x0[11:].to(dev), x1[11:].to(dev), t[11:].to(dev), None)
    assert np.array_equal(torch.cat([la, lb]).cpu().numpy(), loss)
    for p, q in zip(model.parameters(), model2.parameters()):
        assert torch.equal(p, q)
    # the trained module as base_model of the score-head field
    m = fields.CustomNetModelScoreHead(model, 0.2, 0.2, 0.2)
    xs = torch.randn(4096, 2, device=dev) / 1.5
    out = sampler.sample_rk4(m, xs, torch.linspace(0, 1, 3))
    ref_layers = _layers_of(model)
    ref = O.odeint_rk4(O.FieldScoreHead(ref_layers, 0.2, (0.2, 0.2)), xs.cpu(), torch.linspace(0, 1, 3))[-1]
    dd = (out.cpu() - ref).abs()
    assert dd.median().item() <= 1e-5 and dd.quantile(0.999).item() <= 1e-3


def test_scorehead_train_then_sample_end_to_end(cuda_device):
    """The order-1 script end to end on the device (example_order1_with_prediction_head.py): its MLP(dim=4) trained on 8gaussians
    by the persistent kernel (flow loss + score regression, lr 1e-3, sigma 0.2; 10 000 steps, a quarter of a second), wrapped in
    its score-head field, solved by the tcgen05 sampler, judged by the MMD kernel."""
    from idff_b200 import fields, sampler
    from idff_b200.mmd import compute_mmd_rbf
    dev = cuda_device
    model = _fresh_scorehead(dev, seed=0)
    tr = train.FusedFlowTrainer(model, batch_size=256, lr=1e-3, score_sigma=0.2, data=train.eight_gaussians_device)
    for _ in range(10):
        loss = tr.run(1000)
    assert torch.isfinite(loss).all() and loss[-1] < 0.25 * tr_first_loss(model, dev)
    g = torch.Generator(device=dev).manual_seed(1)
    true_s = train.eight_gaussians_device(2048, dev, generator=g)
    x0 = torch.randn(2048, 2, device=dev, generator=g) / 1.5
    field = fields.CustomNetModelScoreHead(model, 0.2, 0.5, 0.0)      # gamma1 = 1/2: the score term vanishes, pure x1hat flow
    out = sampler.sample_rk4(field, x0, torch.linspace(0, 1, 11))
    prior, got = compute_mmd_rbf(true_s, x0, 1.0), compute_mmd_rbf(true_s, out, 1.0)
    print(f"8gaussians, score-head network, 10000 fused training steps, nfe 10: MMD {got:.4f} (prior {prior:.3f}), loss {loss[-1].item():.3f}")
    assert prior > 0.4 and got < 0.1


def tr_first_loss(model, dev):
    """loss of an untrained score-head network on the same task (a fresh copy, one step)"""
    fresh = _fresh_scorehead(dev, seed=0)
    t0 = train.FusedFlowTrainer(fresh, batch_size=256, lr=1e-3, score_sigma=0.2, data=train.eight_gaussians_device)
    return t0.run(1)[0]


def _polyala_case(device, T=200, D=46, seed=0):
    """A PolyALA-shaped data set (180 sin(cumsum(randn 0.05) + phase)) and a fresh MLP_Embedding (MD_simulation.py:118)."""
    from idff_b200.torchcfm.models.models import MLP_Embedding
    g = torch.Generator().manual_seed(seed)
    ds = 180.0 * torch.sin(torch.cumsum(torch.randn(T, D, generator=g) * 0.05, 0) + torch.rand(1, D, generator=g) * 6.28)
    torch.manual_seed(seed)
    model = MLP_Embedding(dim=D, w=128, time_embed=T, time_varying=True).to(device)
    return ds.float(), model


def _md_parts(model):
    ps = [p.detach().cpu().clone() for p in model.parameters()]
    return ps[0], [(ps[1 + 2 * i], ps[2 + 2 * i]) for i in range(4)]


@pytest.mark.parametrize("B,T,D", [(10, 200, 46), (16, 64, 47), (3, 200, 46), (7, 33, 5)])
def test_md_step_gradients_loss_and_update(cuda_device, B, T, D):
    """One PolyALA step in the cluster-resident kernel (idff_train_md_steps) vs the oracle's loop (MD_simulation.py:126-144 on torch
    CPU, same weights and draws): loss to 2e-6 relative, every gradient tensor (embedding included, duplicated indices too) within
    1e-5 * max|g|, Adam's first update to 1e-3 * lr where |g| is above the fp32 noise floor."""
    dev, lr = cuda_device, 1e-3
    ds, model = _polyala_case(dev, T, D, seed=B)
    emb, layers = _md_parts(model)
    g = torch.Generator().manual_seed(31 + B)
    idx = torch.randint(1, T, (1, B), generator=g)
    if B >= 3:
        idx[0, 2] = idx[0, 0]                      # a duplicated frame: its embedding gradient rows add up
    t = torch.rand(1, B, generator=g)
    eps = torch.randn(1, B, D, generator=g)
    losses_r, (emb_r, layers_r), grads_r = O.md_train_steps(emb, layers, ds, idx, t, eps, sigma=0.5, lr=lr, return_grads=True)
    tr = train.FusedMDTrainer(model, ds.to(dev), batch_size=B, sigma=0.5, lr=lr)
    dump = torch.zeros(tr.n_params, device=dev)
    before = _lib.launch_count()
    loss = tr.steps(idx.to(dev), t.to(dev), eps.to(dev), grad_dump=dump)
    assert _lib.launch_count() - before == 1
    assert loss.item() == pytest.approx(losses_r[0], rel=2e-6)
    for o, sh, gr in zip(tr.offsets[:9], tr.shapes, grads_r):
        got = dump[o: o + gr.numel()].view(sh).cpu().numpy()
        ref = gr.numpy()
        assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-12), (sh, np.abs(got - ref).max(), np.abs(ref).max())
    after_r = [emb_r] + [x for Wb in layers_r for x in Wb]
    for p, ref, gr in zip(model.parameters(), after_r, grads_r):
        d = (p.detach().cpu() - ref).abs().numpy()
        solid = np.abs(gr.numpy()) > 1e-4
        if solid.any():
            assert d[solid].max() <= 1e-3 * lr, (tuple(p.shape), d[solid].max())
        assert d.max() <= 2 * lr
    # moments live next to the parameters and follow torch's: exp_avg = (1 - beta1) g after one step
    m_blocks = tr.named_block("exp_avg")
    for mb, gr in zip(m_blocks, grads_r):
        np.testing.assert_allclose(mb.cpu().numpy(), 0.1 * gr.numpy(), rtol=1e-4, atol=1e-5 * max(float(gr.abs().max()), 1e-12))


def test_md_many_steps_track_the_reference(cuda_device):
    """50 consecutive PolyALA steps in ONE launch track the oracle's loop; bit-reproducible and independent of the launch split;
    the trained module drives SDE_cal through the sampler kernels."""
    dev, lr, S, B = cuda_device, 1e-3, 50, 10
    ds, model = _polyala_case(dev, seed=2)
    emb, layers = _md_parts(model)
    g = torch.Generator().manual_seed(77)
    idx = torch.randint(1, 200, (S, B), generator=g)
    t = torch.rand(S, B, generator=g)
    eps = torch.randn(S, B, 46, generator=g)
    losses_r, (emb_r, layers_r) = O.md_train_steps(emb, layers, ds, idx, t, eps, sigma=0.5, lr=lr)
    tr = train.FusedMDTrainer(model, ds.to(dev), batch_size=B, sigma=0.5, lr=lr)
    loss = tr.steps(idx.to(dev), t.to(dev), eps.to(dev)).cpu().numpy()
    np.testing.assert_allclose(loss, np.array(losses_r), rtol=2e-3)
    assert loss[-1] < loss[0]
    after_r = [emb_r] + [x for Wb in layers_r for x in Wb]
    dall = np.concatenate([(p.detach().cpu() - ref).abs().numpy().reshape(-1) for p, ref in zip(model.parameters(), after_r)])
    assert np.median(dall) <= 1e-5 and dall.max() <= 20 * lr, (np.median(dall), dall.max())
    ds2, model2 = _polyala_case(dev, seed=2)
    tr2 = train.FusedMDTrainer(model2, ds2.to(dev), batch_size=B, sigma=0.5, lr=lr)
    la = tr2.steps(idx[:17].to(dev).contiguous(), t[:17].to(dev).contiguous(), eps[:17].to(dev).contiguous())
    lb = tr2.steps(idx[17:].to(dev).contiguous(), t[17:].to(dev).contiguous(), eps[17:].to(dev).contiguous())
    assert np.array_equal(torch.cat([la, lb]).cpu().numpy(), loss)
    for p, q in zip(model.parameters(), model2.parameters()):
        assert torch.equal(p, q)
    # the kernel-trained module under the generator's field
    from idff_b200 import fields
    sde = fields.SDE_cal(model, input_dim=46, sigma=0.5, init_ind=3, max_length=200, dt=0.3, gamma1=0.2, gamma2=1.25)
    y = ds[5:9].to(dev).contiguous()
    f = sde.f(torch.tensor(0.3), y)
    emb2, layers2 = _md_parts(model)
    ref = O.FieldMD(layers2, emb2, 46, sigma=0.5, init_ind=3, dt=0.3, gamma1=0.2, gamma2=1.25).f(torch.tensor(0.3), y.cpu())
    assert (f.cpu() - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())


def test_md_trainer_is_deterministic_under_load(cuda_device):
    """The cluster kernel's hand-overs (st.async + mbarrier, no cluster barrier) leave no room for a race to hide: 6 repeats of
    400 steps at the largest batch, from the same state and draws, must agree bit for bit -- losses, parameters and both moments."""
    dev, S, B = cuda_device, 400, 16
    ds, _ = _polyala_case(dev, seed=9)
    g = torch.Generator().manual_seed(5)
    idx = torch.randint(1, 200, (S, B), generator=g).to(dev)
    t = torch.rand(S, B, generator=g).to(dev)
    eps = torch.randn(S, B, 46, generator=g).to(dev)
    ref = None
    for rep in range(6):
        _, model = _polyala_case(dev, seed=9)
        tr = train.FusedMDTrainer(model, ds.to(dev), batch_size=B, sigma=0.5)
        loss = tr.steps(idx, t, eps)
        snap = (loss.clone(), tr.state.clone())
        assert torch.isfinite(snap[0]).all() and torch.isfinite(snap[1]).all()
        if ref is None:
            ref = snap
        else:
            assert torch.equal(snap[0], ref[0]) and torch.equal(snap[1], ref[1]), rep
    assert ref[0][-1] < ref[0][0]


def test_md_trainer_argument_checks(cuda_device):
    dev = cuda_device
    ds, model = _polyala_case(dev, seed=1)
    tr = train.FusedMDTrainer(model, ds.to(dev))
    idx = torch.randint(1, 200, (2, 17), device=dev)
    with pytest.raises(_lib.IdffError):
        tr.steps(idx, torch.rand(2, 17, device=dev))           # batch above IDFF_TRAIN_MD_MAX_BATCH
    with pytest.raises(_lib.IdffError):
        tr.steps(idx[:, :10].int(), torch.rand(2, 10, device=dev))
    from idff_b200.torchcfm.models.models import MLP_Embedding
    with pytest.raises(_lib.IdffError):
        train.FusedMDTrainer(MLP_Embedding(dim=46, w=64, time_embed=200, time_varying=True).to(dev), ds.to(dev))


def test_md_steps_with_repaired_batches(cuda_device):
    """idff_train_md_steps_paired: the PolyALA loop with every batch re-paired before its path sample (the `_dynamic` matcher,
    conditional_flow_matching_dynamic.py:268) -- x0 / x1 from the given dataset rows, the embedding on index_sel in its original
    order -- tracks the oracle's loop with the same rows; FusedMDTrainer(ot_pairing=True).ot_rows produces rows that are pairs of
    each step's exact OT plan, and run() trains on them."""
    from scipy.optimize import linear_sum_assignment
    dev, lr, S, B, T, D = cuda_device, 1e-3, 12, 10, 200, 46
    ds, model = _polyala_case(dev, seed=5)
    emb, layers = _md_parts(model)
    g = torch.Generator().manual_seed(91)
    idx = torch.randint(1, T, (S, B), generator=g)
    t = torch.rand(S, B, generator=g)
    eps = torch.randn(S, B, D, generator=g)
    tr = train.FusedMDTrainer(model, ds.to(dev), batch_size=B, sigma=0.5, lr=lr, ot_pairing=True)
    torch.manual_seed(3)
    r0, r1 = tr.ot_rows(idx.to(dev))
    assert r0.shape == r1.shape == (S, B) and r0.dtype == torch.int64
    for s in range(S):                              # every (row0, row1) is a pair of step s's exact plan
        x0, x1 = ds[idx[s] - 1].double(), ds[idx[s]].double()
        _, col = linear_sum_assignment((torch.cdist(x0, x1) ** 2).numpy())
        plan = {(int(idx[s, i]) - 1, int(idx[s, col[i]])) for i in range(B)}
        assert all((int(a), int(b)) in plan for a, b in zip(r0[s].cpu(), r1[s].cpu()))
    losses_r, (emb_r, layers_r) = O.md_train_steps(emb, layers, ds, idx, t, eps, sigma=0.5, lr=lr, rows=(r0.cpu(), r1.cpu()))
    loss = tr.steps(idx.to(dev), t.to(dev), eps.to(dev), rows=(r0, r1)).cpu().numpy()
    np.testing.assert_allclose(loss, np.array(losses_r), rtol=1e-3)
    after_r = [emb_r] + [x for Wb in layers_r for x in Wb]
    dall = np.concatenate([(p.detach().cpu() - ref).abs().numpy().reshape(-1) for p, ref in zip(model.parameters(), after_r)])
    assert np.median(dall) <= 1e-5 and dall.max() <= 20 * lr, (np.median(dall), dall.max())
    out = tr.run(256)
    assert torch.isfinite(out).all() and out.shape == (256,)
