import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from idff_b200.torchcfm import conditional_flow_matching as cfm
dev = torch.device("cuda:0")
def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
for n, D in ((1 << 24, 2), (1 << 22, 2), (1 << 20, 46), (1 << 24, 1)):
    x0, x1, t = torch.randn(n, D, device=dev), torch.randn(n, D, device=dev), torch.rand(n, device=dev)
    for want_eps in (False, True):
        s = timeit(lambda: cfm._path_sample_philox(x0, x1, t, 0.2, 1, want_eps))
        byt = n * (4 * D + 1 + (1 if want_eps else 0) * D) * 4
        print(f"n={n} D={D} return_noise={want_eps}: {s*1e6:.1f} us, {n/s/1e9:.1f} G rows/s, {byt/s/1e9:.0f} GB/s", flush=True)
