"""ncu target: idff_ot_assign_batch, 1024 problems of n = 256 (D = 2), one CTA of 128 threads per problem."""
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from idff_b200.torchcfm import optimal_transport as ot
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
x0 = torch.randn(1024, 256, 2, device=dev, generator=g); x1 = torch.randn(1024, 256, 2, device=dev, generator=g)
for _ in range(3):
    perm = ot.ot_assign_batch(x0, x1)
torch.cuda.synchronize()
print("ok", perm.shape)
