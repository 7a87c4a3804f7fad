import torch, time, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from idff_b200.torchcfm import optimal_transport as ot
from scipy.optimize import linear_sum_assignment
for n, D in ((10, 46), (256, 2), (512, 2), (1024, 2)):
    x0 = torch.randn(n, D, device="cuda"); x1 = torch.randn(n, D, device="cuda") + 1
    ot.ot_assign(x0, x1); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): ot.ot_assign(x0, x1)
    b.record(); torch.cuda.synchronize()
    M = (torch.cdist(x0.double(), x1.double()) ** 2).cpu().numpy(); linear_sum_assignment(M)
    t0 = time.perf_counter()
    for _ in range(5):
        M = (torch.cdist(x0.double(), x1.double()) ** 2).cpu().numpy()
    t1 = time.perf_counter()
    for _ in range(5):
        linear_sum_assignment(M)
    t2 = time.perf_counter()
    print(f"n={n} D={D}: idff_ot_assign {a.elapsed_time(b) / 5 * 1e3:.0f} us; host: cdist->cpu {(t1 - t0) / 5 * 1e6:.0f} us + scipy {(t2 - t1) / 5 * 1e6:.0f} us")
