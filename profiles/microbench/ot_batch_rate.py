"""idff_ot_assign_batch: time per coupled minibatch when P problems share one launch (one CTA per problem), against P = 1."""
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
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
for n, D in ((256, 2), (10, 46), (64, 2), (512, 2), (1024, 2)):
    for P in (1, 148, 296, 592, 1024, 4096):
        if n >= 512 and P > 1024: continue
        x0 = torch.randn(P, n, D, device=dev, generator=g); x1 = torch.randn(P, n, D, device=dev, generator=g)
        s = timeit(lambda: ot.ot_assign_batch(x0, x1))
        print(f"n={n} D={D} P={P}: {s * 1e3:.3f} ms per launch, {s / P * 1e6:.1f} us per problem", flush=True)
# the whole OT-paired training loop of the 2-D flow network: draws + batched coupling + 1024 fused steps
from idff_b200 import train
from idff_b200.torchcfm.models.models import MLP
for pairing in (False, True):
    torch.manual_seed(0)
    tr = train.FusedFlowTrainer(MLP(dim=2, w=256, time_varying=True).to(dev), batch_size=256, ot_pairing=pairing)
    s = timeit(lambda: tr.run(1024))
    print(f"FusedFlowTrainer B=256 ot_pairing={pairing}: {s / 1024 * 1e6:.1f} us/step including draws{' and couplings' if pairing else ''}", flush=True)
