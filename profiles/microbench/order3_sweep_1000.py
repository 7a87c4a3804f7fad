import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import numpy as np, torch
from gpu_common import packed
from conftest import load_golden
from idff_b200 import sweep, _lib
dev = torch.device("cuda:0")
pw = packed("checkerboard", dev)
x0 = torch.randn(2048, 2, device=dev) / 1.5
true_s = torch.tensor(load_golden("mmd")["true_checkerboard"]).to(dev)
grid = [0.0, 0.1, 0.2, 0.3, 0.5, 0.7, 1.0, 1.5, 2.0, 3.0]
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); b = _lib.launch_count()
    best, res = sweep.find_best_gammas_mmd(pw, x0, true_s, gamma_range=grid, nfe=2, sigma=0.2, sigma_kernel=1.0, order=3)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"order-3 sweep, {len(res)} tuples x 2048 particles, nfe 2: {1e3 * (t1 - t0):.1f} ms, {_lib.launch_count() - b} launches, best {best}, finite {np.isfinite([v for _, v in res]).all()}")
