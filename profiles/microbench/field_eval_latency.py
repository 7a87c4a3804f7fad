import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from gpu_common import packed
from idff_b200 import _lib, fields
dev = torch.device("cuda:0")
def timeit(fn, reps=50, warm=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for name, mk in (("order2", lambda pw: fields.CustomNetModel(pw, 0.2, 0.2, 0.2)), ("order3", lambda pw: fields.CustomNetModelOrder3(pw, 0.2, 0.2, 0.2, 0.2))):
    pw = packed("checkerboard", dev)
    for n in (256, 1024, 2048, 4096, 8192, 16384):
        x = torch.randn(n, 2, device=dev) / 1.5
        t = torch.tensor(0.37)
        row = []
        for impl in (_lib.IDFF_IMPL_AUTO, _lib.IDFF_IMPL_FUSED, _lib.IDFF_IMPL_TENSOR16):
            m = mk(pw); m.impl = impl
            try:
                row.append("%.1f" % timeit(lambda: m(t, x)))
            except Exception as e:
                row.append("n/a")
        print(name, "N=%d" % n, "auto/fused/tensor16 us:", row, flush=True)
pw = packed("scorehead256", dev)
for n in (256, 2048, 4096, 16384):
    x = torch.randn(n, 2, device=dev) / 1.5
    row = []
    for impl in (_lib.IDFF_IMPL_AUTO, _lib.IDFF_IMPL_GENERIC, _lib.IDFF_IMPL_TENSOR16):
        m = fields.CustomNetModelScoreHead(pw, 0.2, 0.2, 0.2); m.impl = impl
        row.append("%.1f" % timeit(lambda: m(torch.tensor(0.37), x)))
    print("scorehead N=%d auto/generic/tensor16 us:" % n, row, flush=True)
