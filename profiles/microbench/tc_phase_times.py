import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests"]
import numpy as np, torch
from gpu_common import packed
from idff_b200 import fields, sampler, _lib
dev = torch.device("cuda:0")
pw = packed("checkerboard", dev)
xx = torch.randn(1 << 21, 2, device=dev) / 1.5
IMPL = int(sys.argv[1]) if len(sys.argv) > 1 else 5
m = fields.CustomNetModel(pw, 0.2, 0.2, 0.2); m.impl = IMPL
for _ in range(2):
    sampler.sample_rk4(m, xx, torch.linspace(0, 1, 3)); torch.cuda.synchronize()
t0 = time.time(); sampler.sample_rk4(m, xx, torch.linspace(0, 1, 3)); torch.cuda.synchronize(); dt = time.time() - t0
print("impl", IMPL, ": 2M particles nfe2:", xx.shape[0] * 8 / dt / 1e6, "M particle-steps/s")
buf = (C.c_int64 * 32)(); _lib.check((_lib.lib().idff_debug_tensor16_stamps if IMPL == 5 else _lib.lib().idff_debug_tensor_stamps)(buf), "stamps"); s = np.array(buf[:])
names = ["L1 start", "L1 done+publish", "acc0 ready", "epi0 done+publish", "acc1 ready", "epi1 done+publish", "reduce1 done", "acc2 ready", "epi2 done+publish", "acc3 ready", "epi3+reduce done", "ebar", "update+ebar"]
for i in range(1, 13): print(f"  E: {names[i]:22s} +{s[i]-s[i-1]:7d}   (t={s[i]-s[0]})")
for g in range(4): print(f"  MMA gemm {g}: start t={s[16+2*g]-s[0]:7d}  issue-done t={s[17+2*g]-s[0]:7d}  issue {s[17+2*g]-s[16+2*g]}")

if IMPL == 5: print(f"  MMA thread waits over the evaluation: acc_free(mt0) {s[24]}  acc_free(mt1) {s[25]}  b_ready[0] {s[26]}  b_ready[1] {s[27]}  weight ring {s[28]}")
