"""One launch of the PolyALA chain kernel (idff_md_generate: 200 frames, 5 trajectories) for timing / an ncu capture."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, torch
from idff_b200 import md
from idff_b200.weights import PackedWeights
dev = torch.device("cuda:0")
pw = PackedWeights.from_npz(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests", "golden", "weights_polyala.npz")), device=dev)
K, D, T, NFE, dt = int(os.environ.get("K", 5)), 46, int(os.environ.get("T", 200)), 3, 0.3
g = torch.Generator(device=dev).manual_seed(1)
x_init = torch.randn(K, D, device=dev, generator=g)
from idff_b200 import sampler
ik, ik0 = sampler.brownian_increments(4, torch.linspace(0, 1, NFE), dt, (T, K, D), dev, g)     # (n_steps, T, K, D), the right law
ik, ik0 = ik.transpose(0, 1).contiguous(), ik0.transpose(0, 1).contiguous()
kw = dict(sigma=0.5, gamma1=0.2, gamma2=5.0, dt=dt, device=dev, x_init=x_init, noise=(ik, ik0))
md.generate_trajectories_chain(pw, K, D, T, NFE, **kw)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    y = md.generate_trajectories_chain(pw, K, D, T, NFE, **kw)
b.record(); torch.cuda.synchronize()
print(f"idff_md_generate T={T} K={K}: {a.elapsed_time(b) / 3:.2f} ms per chain = {a.elapsed_time(b) / 3 / T * 1e3:.1f} us per frame; finite={bool(torch.isfinite(y).all())}")
