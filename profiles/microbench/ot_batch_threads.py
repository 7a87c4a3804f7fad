import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from idff_b200 import _lib
from idff_b200.torchcfm import optimal_transport as ot
dev = torch.device("cuda:0")
def timeit(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
g = torch.Generator(device=dev).manual_seed(0)
for n, D, P in ((256, 2, 1024), (256, 2, 4096), (512, 2, 1024), (128, 2, 4096), (256, 46, 1024), (1024, 2, 592)):
    x0 = torch.randn(P, n, D, device=dev, generator=g); x1 = torch.randn(P, n, D, device=dev, generator=g)
    ref = None
    for th in (512, 384, 256, 192, 128, 64):
        _lib.check(_lib.lib().idff_debug_ot_threads(th), "t")
        perm, cost = ot.ot_assign_batch(x0, x1, return_cost=True)
        if ref is None: ref = (perm.clone(), cost.clone())
        same = torch.equal(perm, ref[0]) and torch.equal(cost, ref[1])
        s = timeit(lambda: ot.ot_assign_batch(x0, x1))
        print(f"n={n} D={D} P={P} threads={th}: {s * 1e3:.2f} ms, {s / P * 1e6:.1f} us per problem, identical={same}", flush=True)
_lib.lib().idff_debug_ot_threads(0)
