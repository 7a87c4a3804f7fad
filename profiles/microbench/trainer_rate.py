"""Rate of the fused training kernel (idff_train_flow_steps): us/step and samples/s at the reference's batch (256) and others,
S steps per launch, inputs pre-drawn on the device (and, separately, including the device draws)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from idff_b200 import train
from idff_b200.torchcfm.models.models import MLP

import ctypes as C
from idff_b200 import _lib
dev = torch.device("cuda:0")
for B, pair in ((256, 0), (256, 1), (64, 0), (512, 0)):
    _lib.lib().idff_debug_train_pair_mode(pair)
    print("mode:", "CTA pair per 4 rows (default where B <= 296)" if pair and B <= 296 else "one CTA per 4 rows", flush=True)
    torch.manual_seed(0)
    tr = train.FusedFlowTrainer(MLP(dim=2, w=256, time_varying=True).to(dev), batch_size=B)
    for S in (1, 64, 1024):
        x0, x1, t, eps = tr.draw(S)
        tr.steps(x0, x1, t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            tr.steps(x0, x1, t)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * S)
        e0.record()
        for _ in range(reps):
            tr.run(S)
        e1.record(); torch.cuda.synchronize()
        us2 = e0.elapsed_time(e1) * 1e3 / (reps * S)
        print(f"B={B} S={S}: {us:.2f} us/step = {B / us:.2f} M samples/s (pre-drawn); {us2:.2f} us/step incl. draws", flush=True)

# phase stamps (clock64 per CTA) of the last step of a 64-step launch at B = 256, both modes
import numpy as np
for pair in (0, 1):
    _lib.lib().idff_debug_train_pair_mode(pair)
    torch.manual_seed(0)
    tr = train.FusedFlowTrainer(MLP(dim=2, w=256, time_varying=True).to(dev), batch_size=256)
    x0, x1, t, eps = tr.draw(64)
    tr.steps(x0, x1, t); torch.cuda.synchronize()
    buf = (C.c_int64 * (148 * 8))()
    _lib.check(_lib.lib().idff_debug_train_stamps(buf, 148), "stamps")
    st = np.array(buf[:], dtype=np.int64).reshape(148, 8)
    d = np.diff(st[:, :7], axis=1)
    names = ["rows phase", "barrier 1 wait", "wgrad job", "barrier 2 wait", "adamw", "barrier 3 wait"]
    nrow = 128 if pair else 64
    print("phase stamps, pair mode" if pair else "phase stamps, one CTA per row group")
    for lo, hi, what in ((0, nrow, "row CTAs (also tiles)"), (nrow, 128, "tile-only CTAs"), (128, 144, "thin jobs"), (144, 145, "db3/loss"), (145, 148, "idle")):
        if hi > lo:
            print(" ", what, {n: int(np.median(d[lo:hi, i])) for i, n in enumerate(names)}, "total", int(np.median(st[lo:hi, 6] - st[lo:hi, 0])))
_lib.lib().idff_debug_train_pair_mode(1)
