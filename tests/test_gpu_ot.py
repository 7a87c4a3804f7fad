"""-m gpu: the device OT coupling (idff_ot_assign) against the oracle (scipy's exact assignment on the fp64
cost, the restatement of pot.emd for uniform equal-size marginals, optimal_transport.py:59-94)."""
import numpy as np
import pytest
import torch

import idff_oracle as O
from idff_b200 import _lib
from idff_b200.torchcfm import optimal_transport as ot
from idff_b200.torchcfm import conditional_flow_matching_dynamic as cfmd

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,D", [(1, 2), (2, 2), (10, 46), (64, 2), (256, 2), (257, 3), (512, 2), (1024, 2)])
def test_assignment_is_the_exact_plan(cuda_device, n, D):
    g = torch.Generator().manual_seed(100 * n + D)
    x0 = torch.randn(n, D, generator=g)
    x1 = torch.randn(n, D, generator=g) * 1.5 + 0.3
    perm_ref, cost_ref = O.ot_exact_permutation(x0, x1)
    before = _lib.launch_count()
    perm, cost = ot.ot_assign(x0.to(cuda_device), x1.to(cuda_device), return_cost=True)
    assert _lib.launch_count() - before == 1
    perm = perm.cpu().numpy()
    assert sorted(perm.tolist()) == list(range(n))                        # a permutation: plan = P / n has uniform marginals
    assert cost.item() == pytest.approx(cost_ref, rel=1e-12, abs=1e-12)   # the same optimum ...
    np.testing.assert_array_equal(perm, perm_ref)                         # ... and (generic data: unique) the same plan


def test_auction_and_augmenting_path_kernels_agree(cuda_device):
    """n > 64 runs the epsilon-scaling auction, n <= 64 (or idff_debug_ot_sap) the shortest-augmenting-path kernel: same
    permutation and cost on generic data, for D = 2 and a wide D; non-finite inputs terminate and still return a permutation."""
    g = torch.Generator().manual_seed(77)
    for n, D in ((65, 2), (200, 2), (300, 7), (640, 2)):
        x0 = torch.randn(n, D, generator=g).to(cuda_device)
        x1 = (torch.randn(n, D, generator=g) * 0.7 - 0.4).to(cuda_device)
        pa, ca = ot.ot_assign(x0, x1, return_cost=True)
        _lib.check(_lib.lib().idff_debug_ot_sap(1), "ot_sap")
        try:
            ps, cs = ot.ot_assign(x0, x1, return_cost=True)
        finally:
            _lib.check(_lib.lib().idff_debug_ot_sap(0), "ot_sap")
        assert torch.equal(pa, ps), (n, D)
        assert ca.item() == pytest.approx(cs.item(), rel=1e-12)
    x0 = torch.randn(100, 2, generator=g)
    x0[3, 1] = float("nan")
    x0[40, 0] = float("inf")
    p = ot.ot_assign(x0.to(cuda_device), torch.randn(100, 2, generator=g).to(cuda_device))
    assert sorted(p.cpu().tolist()) == list(range(100))


def test_ties_and_degenerate_inputs(cuda_device):
    """Duplicate points (many optimal plans): any optimal permutation is acceptable, the cost must be the optimum."""
    for rep in (8, 48):                                  # 16 rows: augmenting paths; 96 rows: auction
        x0 = torch.tensor([[0.0, 0.0]] * rep + [[1.0, 1.0]] * rep)
        x1 = torch.tensor([[1.0, 1.0]] * rep + [[0.0, 0.0]] * rep)
        perm, cost = ot.ot_assign(x0.to(cuda_device), x1.to(cuda_device), return_cost=True)
        assert sorted(perm.cpu().tolist()) == list(range(2 * rep)) and cost.item() == 0.0
    z = torch.zeros(80, 2, device=cuda_device)           # every plan is optimal
    perm, cost = ot.ot_assign(z, z, return_cost=True)
    assert sorted(perm.cpu().tolist()) == list(range(80)) and cost.item() == 0.0
    for n in (33, 130):
        x = torch.randn(n, 2)
        perm, cost = ot.ot_assign(x.to(cuda_device), x.to(cuda_device), return_cost=True)
        assert perm.cpu().tolist() == list(range(n)) and cost.item() == 0.0
    with pytest.raises(_lib.IdffError):
        ot.ot_assign(torch.zeros(2000, 2, device=cuda_device), torch.zeros(2000, 2, device=cuda_device))
    with pytest.raises(_lib.IdffError):
        ot.ot_assign(torch.zeros(4, 2), torch.zeros(4, 2))                # CPU tensors: no fallback


def test_dynamic_matcher_couples_on_the_device(cuda_device):
    """conditional_flow_matching_dynamic.py:268: x0, x1 = ot_sampler.sample_plan(x0, x1) -- pairs follow the exact plan
    (every drawn pair (i, j) has j = perm[i]), i is uniform, and the path sample uses the re-paired rows."""
    dev = cuda_device
    B = 256
    g = torch.Generator().manual_seed(9)
    x0 = torch.randn(B, 2, generator=g).to(dev)
    x1 = (torch.randn(B, 2, generator=g) + 2.0).to(dev)
    perm_ref, _ = O.ot_exact_permutation(x0.cpu(), x1.cpu())
    i, j = ot.OTPlanSampler().sample_map_indices(x0, x1)
    assert i.is_cuda and j.is_cuda and i.dtype == torch.int64
    np.testing.assert_array_equal(j.cpu().numpy(), perm_ref[i.cpu().numpy()])
    assert 0.45 * B < len(set(i.cpu().tolist())) <= B                     # with replacement: ~63 % distinct
    i2, j2 = ot.OTPlanSampler().sample_map_indices(x0, x1, replace=False)  # without: every plan pair exactly once
    assert sorted(i2.cpu().tolist()) == list(range(B))
    np.testing.assert_array_equal(j2.cpu().numpy(), perm_ref[i2.cpu().numpy()])
    a0, b0 = ot.OTPlanSampler().sample_plan(x0, x1, replace=False)
    assert a0.shape == x0.shape and b0.shape == x1.shape
    FM = cfmd.ExactOptimalTransportConditionalFlowMatcher(sigma=0.0)
    t, xt, ut = FM.sample_location_and_conditional_flow(x0, x1)
    # sigma = 0: xt - t * ut = x0[i] and xt + (1 - t) * ut = x1[perm[i]] for the drawn i: every row of (FM.x0, FM.x1) is a plan pair
    a, b = FM.x0.cpu().numpy(), FM.x1.cpu().numpy()
    x0n, x1n = x0.cpu().numpy(), x1.cpu().numpy()
    for r in range(0, B, 17):
        src = np.where((x0n == a[r]).all(1))[0]
        assert len(src) >= 1 and (x1n[perm_ref[src[0]]] == b[r]).all()
    # labels follow their rows through the re-pairing (guided variant, :273-315)
    y0, y1 = torch.arange(B, device=dev), torch.arange(B, device=dev) + 1000
    out = FM.guided_sample_location_and_conditional_flow(x0, x1, y0=y0, y1=y1, return_noise=True)
    assert len(out) == 6
    la, lb = out[3].cpu().numpy(), out[4].cpu().numpy() - 1000
    np.testing.assert_array_equal(lb, perm_ref[la])
    assert torch.equal(FM.x0, x0[out[3]]) and torch.equal(FM.x1, x1[out[4] - 1000])
    xa, xb, ya, yb = ot.OTPlanSampler().sample_plan_with_labels(x0, x1, y0, None)
    assert yb is None and torch.equal(xa, x0[ya])
    # transport cost of the coupled pairs is lower than the independent pairing's
    assert ((FM.x0 - FM.x1) ** 2).sum(1).mean().item() < ((x0 - x1) ** 2).sum(1).mean().item()


def test_wasserstein_distance(cuda_device):
    """wasserstein(x0, x1, power) (optimal_transport.py:214-263): W2 from the assignment kernel, W1 on the host; both against
    the oracle's exact plan."""
    g = torch.Generator().manual_seed(12)
    x0, x1 = torch.randn(200, 2, generator=g), torch.randn(200, 2, generator=g) + 1.5
    _, cost = O.ot_exact_permutation(x0, x1)
    w2 = ot.wasserstein(x0.to(cuda_device), x1.to(cuda_device), power=2)
    assert w2 == pytest.approx(np.sqrt(cost / 200), rel=1e-9)
    assert ot.wasserstein(x0, x1, power=2) == pytest.approx(w2, rel=1e-9)            # host path, same optimum
    w1 = ot.wasserstein(x0.to(cuda_device), x1.to(cuda_device), power=1)
    assert 0 < w1 <= w2 + 1e-9                                                        # W1 <= W2
    with pytest.raises(ValueError):
        ot.wasserstein(x0, x1, method="sinkhorn")


@pytest.mark.gpu
@pytest.mark.parametrize("P,n,D", [(7, 10, 46), (40, 256, 2), (300, 64, 2), (5, 100, 3), (200, 128, 2)])
def test_batched_couplings_equal_single_calls(cuda_device, P, n, D):
    """idff_ot_assign_batch: one CTA per problem -> every row is exactly the single-problem result (same permutation, same
    fp64 cost bits), and a sample of them is scipy's optimum."""
    from scipy.optimize import linear_sum_assignment
    g = torch.Generator().manual_seed(100 + P)
    x0, x1 = torch.randn(P, n, D, generator=g), torch.randn(P, n, D, generator=g) * 1.5 + 0.3
    a, b = x0.to(cuda_device), x1.to(cuda_device)
    perm, cost = ot.ot_assign_batch(a, b, return_cost=True)
    assert perm.shape == (P, n) and cost.shape == (P,)
    for p in list(range(min(P, 6))) + [P - 1]:
        ps, cs = ot.ot_assign(a[p], b[p], return_cost=True)
        assert torch.equal(ps, perm[p]) and cs.item() == cost[p].item()
        M = (torch.cdist(x0[p].double(), x1[p].double()) ** 2).numpy()
        r, c = linear_sum_assignment(M)
        assert np.array_equal(perm[p].cpu().numpy(), c)
    assert all(sorted(row.tolist()) == list(range(n)) for row in perm.cpu())
    if n > 64:   # the auction's result does not depend on its block size (large batches run on smaller CTAs)
        try:
            for th in (64, 256):
                _lib.check(_lib.lib().idff_debug_ot_threads(th), "ot_threads")
                p2, c2 = ot.ot_assign_batch(a, b, return_cost=True)
                assert torch.equal(p2, perm) and torch.equal(c2, cost)
        finally:
            _lib.check(_lib.lib().idff_debug_ot_threads(0), "ot_threads")


@pytest.mark.gpu
def test_batched_plan_sampling_and_ot_paired_trainer(cuda_device):
    """OTPlanSampler.sample_plan_batch returns, per problem, rows of x0 paired with their OT partners (the reference's sample_plan
    per step, conditional_flow_matching_dynamic.py:268); FusedFlowTrainer(ot_pairing=True) trains on such pairs: the coupled
    pairs are shorter than the independent ones, and the loss goes down."""
    from idff_b200 import train
    from idff_b200.torchcfm.models.models import MLP
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(3)
    x0 = torch.randn(12, 64, 2, device=dev, generator=g)
    x1 = torch.randn(12, 64, 2, device=dev, generator=g) * 2.0
    s0, s1 = ot.OTPlanSampler().sample_plan_batch(x0, x1, generator=g)
    perm = ot.ot_assign_batch(x0, x1)
    for p in range(12):
        for r in range(0, 64, 7):
            i = (x0[p] == s0[p, r]).all(dim=1).nonzero()[0, 0]
            assert torch.equal(s1[p, r], x1[p, perm[p, i]])
    torch.manual_seed(0)
    tr = train.FusedFlowTrainer(MLP(dim=2, w=256, time_varying=True).to(dev), batch_size=64, ot_pairing=True)
    a0, a1, _, _ = tr.draw(16)
    tr.ot_pairing = False
    b0, b1, _, _ = tr.draw(16)
    tr.ot_pairing = True
    assert ((a1 - a0) ** 2).sum(-1).mean() < 0.8 * ((b1 - b0) ** 2).sum(-1).mean()
    first = tr.run(64).mean().item()
    tr.run(1024)
    assert tr.run(64).mean().item() < first


@pytest.mark.gpu
def test_dynamic_matcher_many_minibatches_at_once(cuda_device):
    """ExactOptimalTransportConditionalFlowMatcher.sample_location_and_conditional_flow_many: S minibatches coupled in one
    batched launch and path-sampled in one launch; every returned row is a pair of its own minibatch's exact plan and obeys
    xt = t x1 + (1 - t) x0 + sigma sqrt(t (1 - t)) eps,  ut = x1 - x0  (conditional_flow_matching_dynamic.py:268-271)."""
    dev = cuda_device
    torch.manual_seed(5)
    S, B, D = 9, 128, 2
    x0 = torch.randn(S, B, D, device=dev)
    x1 = torch.randn(S, B, D, device=dev) * 1.5 + 0.5
    FM = cfmd.ExactOptimalTransportConditionalFlowMatcher(sigma=0.3)
    t, xt, ut, eps = FM.sample_location_and_conditional_flow_many(x0, x1, return_noise=True)
    assert t.shape == (S, B) and xt.shape == (S, B, D) and ut.shape == (S, B, D) and eps.shape == (S, B, D)
    perm = ot.ot_assign_batch(x0, x1)
    for s in range(S):
        plan_ut = x1[s][perm[s]] - x0[s]                                    # the B pairs of minibatch s's plan
        d = torch.cdist(ut[s].double(), plan_ut.double())
        near, which = d.min(dim=1)
        assert near.max().item() < 1e-6                                     # every drawn pair is a pair of the plan
        xi, xj = x0[s][which], x1[s][perm[s][which]]
        tt = t[s][:, None]
        ref = tt * xj + (1 - tt) * xi + 0.3 * torch.sqrt(tt * (1 - tt)) * eps[s]
        assert (xt[s] - ref).abs().max().item() < 1e-5
    with pytest.raises(_lib.IdffError):      # host tensors: no CPU path
        FM.sample_location_and_conditional_flow_many(x0[:2, :16].cpu(), x1[:2, :16].cpu())
