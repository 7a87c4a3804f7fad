import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from idff_b200 import train, fields, sampler
from idff_b200.mmd import compute_mmd_rbf
from idff_b200.torchcfm.models.models import MLP
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MLP(dim=2, w=256, time_varying=True).to(dev)
tr = train.FusedFlowTrainer(model, batch_size=256, lr=float(os.environ.get("LR", 2e-4)), data=train.eight_gaussians_device)
true_s = train.eight_gaussians_device(2048, dev)
x0 = torch.randn(2048, 2, device=dev) / 1.5
print("prior mmd", compute_mmd_rbf(true_s, x0, 1.0))
done = 0
for chunk in (2000, 8000, 10000, 20000, 40000):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(chunk // 1000):
        l = tr.run(1000)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    done += chunk
    f = fields.CustomNetModel(model, 0.2, 0.2, 0.0)
    for nfe in (2, 10):
        out = sampler.sample_rk4(f, x0, torch.linspace(0, 1, nfe + 1))
        print(f"steps {done} ({dt:.2f}s for {chunk}): loss {l[-100:].mean().item():.3f} nfe {nfe} mmd {compute_mmd_rbf(true_s, out, 1.0):.4f}")
