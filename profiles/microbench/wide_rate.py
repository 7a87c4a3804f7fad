"""Throughput of the wide (layer-by-layer tcgen05) sampler vs the generic FMA kernel: order-3 field, 3-w-w-w-2 MLP."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import torch
from idff_b200 import _lib, fields, sampler
from idff_b200.torchcfm.models.models import MLP
from idff_b200.weights import PackedWeights
dev = torch.device("cuda:0")
for width, n in ((1024, 1 << 17), (2048, 1 << 16), (512, 1 << 18)):
    torch.manual_seed(0)
    pw = PackedWeights.from_module(MLP(dim=2, w=width, time_varying=True), device=dev)
    x = torch.randn(n, 2, device=dev) / 1.5
    ts = torch.linspace(0, 1, 3)
    F = 3 * width + 2 * width * width + 2 * width
    for order, gam in ((3, (0.2, 0.2, 0.2)), (2, (0.2, 0.2))):
        cls = fields.CustomNetModelOrder3 if order == 3 else fields.CustomNetModel
        for impl, name, nn in ((3, "tensor-wide", n), (3, "tensor-wide, 32-particle tiles", n), (1, "generic", n // 8)):
            if "32-particle" in name and order != 3:
                continue
            _lib.lib().idff_debug_wide_dense3(0 if "32-particle" in name else 1)
            m = cls(pw, 0.2, *gam); m.impl = impl
            xx = x[:nn]
            sampler.sample_rk4(m, xx, ts); torch.cuda.synchronize()
            t0 = time.time(); sampler.sample_rk4(m, xx, ts); torch.cuda.synchronize(); dt = time.time() - t0
            tf = nn * 2.0 * F * (6 * 2 * order + 2) / dt / 1e12
            print(f"w={width} order {order} {name:30s} N={nn:7d}: {nn * 8 / dt / 1e6:8.3f} M particle-steps/s  {tf:7.1f} algorithmic TFLOP/s", flush=True)
