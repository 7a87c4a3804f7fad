"""ncu target: idff_path_sample_philox on (2^24, 2) fp32 rows (the bench aux shape), a few launches."""
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from idff_b200.torchcfm import conditional_flow_matching as cfm
dev = torch.device("cuda:0")
n = 1 << 24
x0, x1, t = torch.randn(n, 2, device=dev), torch.randn(n, 2, device=dev), torch.rand(n, device=dev)
for _ in range(4):
    out = cfm._path_sample_philox(x0, x1, t, 0.2, 1, False)
torch.cuda.synchronize()
print("ok", [tuple(o.shape) for o in out if o is not None])
