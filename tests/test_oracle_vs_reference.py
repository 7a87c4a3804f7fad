"""Re-pins oracle/idff_oracle.py to the reference's OWN code objects whenever /root/reference is mounted
(build container only; skipped on the GPU box).  Bit-exact."""
import numpy as np
import pytest
import torch

import idff_oracle as O
import ref_loader as R

pytestmark = pytest.mark.skipif(not R.available(), reason="/root/reference not mounted")
torch.set_num_threads(4)


def _layers(net):
    return [(m.weight.detach(), m.bias.detach()) for m in net.net if isinstance(m, torch.nn.Linear)]


def test_field_order2_order3_bit_exact():
    net = R.load_2d_checkpoint("8gaussians")
    o2 = R.script_symbols(R.ORDER2, ["CustomNetModel"])["CustomNetModel"]
    o3 = R.script_symbols(R.ORDER3, ["CustomNetModel"])["CustomNetModel"]
    x = torch.randn(300, 2, generator=torch.Generator().manual_seed(4)) * 2
    for tv in (0.0, 0.25, 0.9, 1.0):
        t = torch.tensor(tv)
        a = o2(net, sigma=0.2, gamma1=0.3, gamma2=-0.7)(t, x.clone()).detach()
        b = O.Field2D(_layers(net), 0.2, (0.3, -0.7), 2)(t, x)
        assert torch.equal(a, b)
        a = o3(net, sigma=0.2, gamma1=0.3, gamma2=-0.7, gamma3=1.5)(t, x.clone()).detach()
        b = O.Field2D(_layers(net), 0.2, (0.3, -0.7, 1.5), 3)(t, x)
        assert torch.equal(a, b)


def test_md_field_bit_exact():
    net = R.load_md_checkpoint()
    S = R.script_symbols(R.MD, ["SDE_cal"])["SDE_cal"]
    y = torch.randn(17, 46, generator=torch.Generator().manual_seed(5)) * 50
    ref = S(net, input_dim=46, sigma=0.5, init_ind=7, max_length=200, dt=0.3, gamma1=0.2, gamma2=1.0)
    mine = O.FieldMD(_layers(net), net.time_index_embedding.weight.detach(), 46, 0.5, 7, 0.3, 0.2, 1.0)
    for tv in (0.0, 0.45, 1.0):
        t = torch.tensor(tv)
        assert torch.equal(ref.f(t, y.clone()).detach(), mine.f(t, y))
        assert torch.equal(ref.g(t, y.clone()).detach(), mine.g(t, y))


def test_matcher_bit_exact_and_draw_order():
    for dyn in (False, True):
        cfm = R.cfm(dynamic=dyn)
        FM = cfm.ConditionalFlowMatcher(sigma=0.3)
        x0, x1 = torch.randn(64, 3), torch.randn(64, 3)
        torch.manual_seed(9)
        t, xt, ut, eps = FM.sample_location_and_conditional_flow(x0, x1, return_noise=True)
        torch.manual_seed(9)
        t2, xt2, ut2, eps2 = O.sample_location_and_conditional_flow(x0, x1, 0.3, False, return_noise=True)
        for a, b in ((t, t2), (xt, xt2), (ut, ut2), (eps, eps2)):
            assert torch.equal(a, b)
        assert torch.equal(FM.compute_lambda(t) * torch.ones(1), O.compute_lambda(t, 0.3, False, dyn) * torch.ones(1))


def test_mmd_bit_exact():
    f = R.script_symbols(R.ORDER2, ["compute_mmd_rbf"])["compute_mmd_rbf"]
    a, b = torch.randn(500, 2), torch.randn(700, 2) + 0.5
    assert f(a, b, sigma=1.0) == pytest.approx(O.mmd_rbf(a, b, 1.0), rel=1e-6)


def test_md_metrics_equal_reference_method():
    import types
    fn = R.script_method(R.MD, "TrajectoryGenerator", "compute_metrics")
    g = torch.Generator().manual_seed(3)
    data = 100 * torch.randn(50, 7, generator=g)
    y = data[:, None, :] + 5 * torch.randn(50, 3, 7, generator=g)
    ref = fn(types.SimpleNamespace(dataset=data), y.numpy())
    got = O.md_compute_metrics(y.numpy(), data.numpy())
    assert set(ref) == set(got)
    for k in ref:
        assert float(ref[k]) == float(got[k]), k


def test_toy_samplers_equal_reference_functions():
    """The host-side toy generators of the mirror (idff_b200/torchcfm/utils.py) against the reference's functions executed from
    torchcfm/utils.py with the same RNG state."""
    import numpy as np
    from idff_b200.torchcfm import utils as U
    ref_rng = np.random.RandomState(5)
    f = R.script_symbols("torchcfm/utils.py", ["pinwheel", "spirals_2d", "checkerboard", "eight_normal_sample"],
                         extra_ns={"rng": ref_rng})
    U.rng = np.random.RandomState(5)
    assert torch.equal(f["pinwheel"](250), U.pinwheel(250))
    for name in ("spirals_2d", "checkerboard"):
        np.random.seed(9)
        a = f[name](300)
        np.random.seed(9)
        b = getattr(U, name)(300)
        assert torch.equal(a, b), name
    torch.manual_seed(4)
    a = f["eight_normal_sample"](128, 2, scale=5, var=0.1)
    torch.manual_seed(4)
    assert torch.equal(a, U.eight_normal_sample(128, 2, scale=5, var=0.1))


def test_mirror_module_surface_equals_reference():
    """The torchcfm mirror exports what the reference's modules define for the path (plots and torchdyn's moons aside), and the
    small GradModel wrapper computes what the reference's does (models.py:45-53)."""
    import ast
    import os
    from idff_b200.torchcfm.models import models as mirror
    ref = R.models()
    for name in ("MLP", "MLP_Embedding", "GradModel"):
        assert hasattr(mirror, name) and hasattr(ref, name)
    torch.manual_seed(0)
    act = torch.nn.Sequential(torch.nn.Linear(4, 16), torch.nn.Tanh(), torch.nn.Linear(16, 1))
    x = torch.randn(7, 4)
    a = ref.GradModel(act)(x.clone())
    b = mirror.GradModel(act)(x.clone())
    assert a.shape == (7, 3) and torch.equal(a, b) and b.requires_grad
    root = "/root/reference/torchcfm"
    skip = {"plot_trajectories", "plot_trajectories_m", "plt_flow_samples", "sample_moons"}     # presentation / torchdyn
    import importlib
    for mod in ("conditional_flow_matching", "conditional_flow_matching_dynamic", "optimal_transport", "utils"):
        tree = ast.parse(open(os.path.join(root, mod + ".py")).read())
        want = {n.name for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef))} - skip
        have = importlib.import_module("idff_b200.torchcfm." + mod)
        missing = [n for n in sorted(want) if not hasattr(have, n)]
        assert not missing, (mod, missing)
