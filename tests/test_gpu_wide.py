"""-m gpu: the layer-by-layer tcgen05 sampler for wide networks (BASELINE.json configs[2]: order-3 field, large-width
MLP) against the CPU oracle (the reference's nested-autograd field under the restated rk4) on seeded default-init
weights, and against the generic FMA kernel."""
import numpy as np
import pytest
import torch

import idff_oracle as O
from gpu_common import assert_field_close, drift_report
from idff_b200 import _lib, fields, sampler
from idff_b200.weights import PackedWeights

pytestmark = pytest.mark.gpu


def _net(width, seed=0):
    torch.manual_seed(seed)
    lin = [torch.nn.Linear(3, width), torch.nn.Linear(width, width), torch.nn.Linear(width, width), torch.nn.Linear(width, 2)]
    return [(m.weight.detach().clone(), m.bias.detach().clone()) for m in lin]


@pytest.mark.parametrize("width", [512, 1024])
def test_wide_field_and_rk4_vs_oracle(cuda_device, width):
    layers = _net(width)
    pw = PackedWeights([W.numpy() for W, _ in layers], [b.numpy() for _, b in layers], device=cuda_device)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(301, 2, generator=g) * 1.5
    xd = x.to(cuda_device)
    layers64 = [(W.double(), b.double()) for W, b in layers]
    for order, gam in ((2, (0.2, 0.2)), (3, (0.2, 0.3, 0.4)), (3, (0.0, 0.0, 0.0))):
        cls = fields.CustomNetModel if order == 2 else fields.CustomNetModelOrder3
        ref_f = O.Field2D(layers, 0.2, gam, order)
        ref64 = O.Field2D(layers64, 0.2, gam, order)
        for tv in (0.0, 0.37, 5.0 / 6.0, 1.0):
            outs = {}
            for impl in (_lib.IDFF_IMPL_GENERIC, _lib.IDFF_IMPL_TENSOR):
                m = cls(pw, 0.2, *gam)
                m.impl = impl
                outs[impl] = m(torch.tensor(tv, dtype=torch.float32), xd).cpu().numpy()
            r32 = ref_f(torch.tensor(tv, dtype=torch.float32), x).numpy()
            r64 = ref64(torch.tensor(np.float32(tv), dtype=torch.float64), x.double()).numpy()
            scale = max(1.0, np.abs(r32).max())
            gap = np.abs(r32 - r64).max()
            for impl, out in outs.items():
                e64 = np.abs(out - r64).max()
                # 1e-5 of max|f|, or the reference's own fp32 noise where that is larger (t = 1 divides by 1e-2)
                assert e64 <= max(1e-5 * scale, 1.5 * gap), (width, order, tv, impl, e64, scale, gap)
        # one solve, nfe = 2
        ts = torch.linspace(0, 1, 3)
        ref = O.odeint_rk4(ref_f, x, ts)[-1].numpy()
        m = cls(pw, 0.2, *gam)
        m.impl = _lib.IDFF_IMPL_TENSOR
        traj = sampler.odeint(m, xd, ts, method="rk4")
        assert torch.equal(traj[0], xd) and torch.equal(traj[-1], sampler.sample_rk4(m, xd, ts))
        r = drift_report(traj[-1].cpu().numpy(), ref)
        assert r["median"] <= 2e-5 and r["p999"] <= 5e-3, (width, order, r)


def test_wide_auto_route_and_chunk_independence(cuda_device):
    layers = _net(512, seed=3)
    pw = PackedWeights([W.numpy() for W, _ in layers], [b.numpy() for _, b in layers], device=cuda_device)
    m = fields.CustomNetModelOrder3(pw, 0.2, 0.2, 0.2, 0.2)        # AUTO
    x0 = torch.randn(5000, 2, device=cuda_device) / 1.5
    ts = torch.linspace(0, 1, 3)
    before = _lib.launch_count()
    whole = sampler.sample_rk4(m, x0, ts)
    assert _lib.launch_count() - before > 8                        # layer-by-layer path (many launches), not one kernel
    assert torch.isfinite(whole).all()
    mt = fields.CustomNetModelOrder3(pw, 0.2, 0.2, 0.2, 0.2)
    mt.impl = _lib.IDFF_IMPL_TENSOR
    for lo, hi in ((0, 17), (16, 48), (4000, 5000)):                # column-tile aligned starts: particles are independent
        assert torch.equal(sampler.sample_rk4(mt, x0[lo:hi], ts), whole[lo:hi])
