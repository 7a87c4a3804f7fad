"""CPU-only checks of the C-ABI: the library loads, exports every symbol include/idff_b200.h declares, validates
arguments without touching a GPU, and its host-side scalars follow the reference's fp32 operation order."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from idff_b200 import _lib


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "idff_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(idff_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 15
    L = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/idff_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert _lib.lib().idff_version() >= 200


def test_loaded_library_was_built_from_this_checkout():
    """build provenance: the Makefile stamps the sha256 of csrc/ + the header into the library"""
    assert _lib.source_hash() == _lib.expected_source_hash()


def test_argument_validation_needs_no_gpu():
    L = _lib.lib()
    assert L.idff_path_sample(None, None, None, None, None, None, 0.0, 0, None, None, -1, 2, None) == -1
    assert b"bad shape" in L.idff_last_error()
    assert L.idff_path_sample(None, None, None, None, None, None, 0.0, 0, None, None, 0, 2, None) == 0   # empty batch
    assert L.idff_path_sample(None, None, None, None, None, None, 0.0, 0, None, None, 4, 2, None) == -1  # null pointers
    assert L.idff_loss(7, None, None, None, None, None, 0.0, 1.0, None, None, 1, 1, None, None) == -1
    assert L.idff_loss(0, None, None, None, None, None, 0.0, 1.0, None, None, 0, 2, None, None) == -1    # mean over nothing
    assert L.idff_mmd_rbf(None, 1, None, 1, 2, 1.0, None, 0, 1, 0, 1, None) == -1
    assert L.idff_sde_num_steps(None, 0, 0.1) == -1
    ts = np.array([0.0, 0.5, 1.0], dtype=np.float32)
    assert L.idff_sde_num_steps(ts.ctypes.data_as(C.c_void_p), 3, 0.3) == 4        # 0,.3,.6,.9,1
    assert L.idff_sde_num_steps(ts.ctypes.data_as(C.c_void_p), 3, 0.25) == 4
    with pytest.raises(_lib.IdffError):
        _lib.require_cuda(torch.zeros(2), "x")                                      # no CPU path


def _stage(variant, order, t, sigma, gam, end_div=1e-2):
    p = _lib.FieldParams()
    p.variant, p.order, p.dim, p.sigma, p.end_div = variant, order, 2, sigma, end_div
    for i, g in enumerate(gam):
        p.gamma[i] = g
    out = (C.c_float * 8)()
    assert _lib.lib().idff_stage_scalars(C.byref(p), t, out) == 0
    return np.array(out[:], dtype=np.float32)


def test_stage_scalars_follow_reference_op_order():
    for tv in (0.0, 1.0 / 6.0, 0.37, float(np.float32(1) / np.float32(3)), 0.9, 1.0):
        t = torch.tensor(tv, dtype=torch.float32)
        for gam in ((0.2, 0.3), (1.0, 2.0), (-0.5, 0.0)):
            s = _stage(_lib.IDFF_FIELD_MOMENTUM, 2, tv, 0.2, gam)
            kind = 0 if tv == 0 else (1 if tv == 1 else 2)
            assert s[1] == kind and s[0] == t.item()
            st2 = (0.2 * torch.sqrt(t * (1 - t))) ** 2                         # Autograd_example_order2.py:47
            a = torch.sqrt(1 - (gam[0] + gam[1]) * st2)                        # :73
            assert s[2] == a.item() and s[3] == (1 - t).item()
            one = torch.ones(())
            c1 = 0.5 * (2 * gam[0] - 1) * one * st2 / (0.2 ** 2)               # :74 with grad == 1
            c2 = gam[1] * one * st2 / (0.2 ** 2)                               # :75
            assert s[4] == c1.item() and s[5] == c2.item() and s[6] == 0.0
        s = _stage(_lib.IDFF_FIELD_MD, 2, tv, 0.5, (0.2, 1.25), end_div=0.3)
        st2 = (0.5 ** 2) * t * (1 - t)                                          # MD_simulation.py:502
        assert s[2] == torch.sqrt(1 - st2).item()                               # :523
        assert s[4] == (-(0.5 * (2 * 0.2 - 1) * torch.ones(()) * st2 / 0.5 ** 2)).item()
        assert s[5] == (-(1.25 * torch.ones(()) * st2 / 0.5 ** 2)).item()
        assert s[7] == np.float32(0.3)
        s = _stage(_lib.IDFF_FIELD_SCOREHEAD, 1, tv, 0.2, (0.3, 0.7))
        st2 = (0.2 * torch.sqrt(t * (1 - t))) ** 2
        assert s[2] == torch.sqrt(1 - (0.3 + 0.7) * st2).item()
        assert s[4] == (0.5 * (2 * 0.3 - 1) * torch.ones(()) * st2).item()
        assert s[5] == (0.7 * torch.ones(()) * st2).item()


def test_new_entry_points_validate_without_a_gpu():
    """The trainer's state layout matches nn.Linear's parameter shapes; the later entry points reject bad arguments and treat
    empty work as a no-op before touching a device."""
    L = _lib.lib()
    off = (C.c_int64 * 12)()
    assert L.idff_train_flow_offsets(off) == 0
    o = list(off)
    torch.manual_seed(0)
    from idff_b200.torchcfm.models.models import MLP
    sizes = [p.numel() for p in MLP(dim=2, w=256, time_varying=True).parameters()]
    assert o[0] == 0 and [o[i + 1] - o[i] for i in range(7)] == sizes[:7] and o[8] == sum(sizes)
    assert o[9] == 0 and o[10] >= o[8] and o[11] - o[10] == o[10] - o[9] and (o[10] * 4) % 16 == 0
    assert L.idff_train_flow_state_floats() > o[11] + o[8]
    assert L.idff_train_flow_offsets(None) == -1
    opt = _lib.AdamW(2e-4, 0.9, 0.999, 1e-8, 1e-2, 1.0)
    assert L.idff_train_flow_steps(None, None, None, None, None, 1, 256, 0.0, 1, C.byref(opt), 0, None, None, None, None) == -1
    fake = C.c_void_p(4096)     # aligned, never dereferenced: validation happens first
    assert L.idff_train_flow_steps(fake, None, None, None, None, 0, 256, 0.0, 1, C.byref(opt), 0, None, None, None, None) == 0
    assert L.idff_train_flow_steps(fake, fake, fake, fake, None, 1, 6, 0.0, 1, C.byref(opt), 0, None, None, None, None) == -1
    assert b"multiple of 4" in L.idff_last_error()
    assert L.idff_train_flow_steps(fake, fake, fake, fake, None, 1, 1024, 0.0, 1, C.byref(opt), 0, None, None, None, None) == -1
    assert L.idff_ot_assign(None, None, 0, 2, None, None, None) == 0
    assert L.idff_ot_assign(None, None, 5000, 2, None, None, None) == -1
    assert L.idff_ot_assign(None, None, 8, 2, None, None, None) == -1
    assert L.idff_ot_assign_batch(None, None, 0, 8, 2, None, None, None) == 0
    assert L.idff_ot_assign_batch(None, None, 4, 8, 2, None, None, None) == -1
    assert L.idff_ot_assign_batch(None, None, -1, 8, 2, None, None, None) == -1
    assert L.idff_md_generate_sweep(None, None, 0, 4, None, None, 2, 0.3, 0, None, None, None, 5, None) == -1
    assert L.idff_debug_train_stamps(None, 4) == -1


def test_round2_trainer_entry_points_validate_without_a_gpu():
    """Score-head and PolyALA trainers: state layouts follow the modules' parameters() order and shapes; bad arguments are rejected
    (and empty work accepted) before any device is touched."""
    L = _lib.lib()
    from idff_b200.torchcfm.models.models import MLP, MLP_Embedding
    off = (C.c_int64 * 12)()
    assert L.idff_train_offsets(1, off) == 0                                         # IDFF_TRAIN_NET_SCOREHEAD
    o = list(off)
    sizes = [p.numel() for p in MLP(dim=4, w=256, time_varying=True).parameters()]   # example_order1_with_prediction_head.py:101
    assert sizes[0] == 256 * 5 and sizes[6] == 4 * 256
    assert o[0] == 0 and [o[i + 1] - o[i] for i in range(7)] == sizes[:7] and o[8] == sum(sizes)
    assert L.idff_train_state_floats(1) > o[11] + o[8] and L.idff_train_state_floats(0) == L.idff_train_flow_state_floats()
    assert L.idff_train_state_floats(7) == -1 and L.idff_train_offsets(7, off) == -1
    opt = _lib.AdamW(1e-3, 0.9, 0.999, 1e-8, 1e-2, 1.0)
    fake = C.c_void_p(4096)
    assert L.idff_train_steps(1, fake, None, None, None, None, 0, 256, 0.0, 1, 0.2, C.byref(opt), 0, None, None, None, None) == 0
    assert L.idff_train_steps(2, fake, fake, fake, fake, None, 1, 256, 0.0, 1, 0.2, C.byref(opt), 0, None, None, None, None) == -1
    assert L.idff_train_steps(1, fake, fake, fake, fake, None, 1, 258, 0.0, 1, 0.2, C.byref(opt), 0, None, None, None, None) == -1
    # PolyALA: MLP_Embedding(dim=46, w=128, time_embed=200) -- embedding first, as module.parameters() yields it
    off13 = (C.c_int64 * 13)()
    assert L.idff_train_md_offsets(46, 200, off13) == 0
    o = list(off13)
    sizes = [p.numel() for p in MLP_Embedding(dim=46, w=128, time_embed=200, time_varying=True).parameters()]
    assert sizes[0] == 200 * 128 and sizes[1] == 128 * 175
    assert o[0] == 0 and [o[i + 1] - o[i] for i in range(8)] == sizes[:8] and o[9] == sum(sizes) == 87086
    assert o[10] == 0 and o[11] >= o[9] and o[12] == 2 * o[11] and L.idff_train_md_state_floats(46, 200) == 3 * o[11]
    assert L.idff_train_md_state_floats(48, 200) == -1 and L.idff_train_md_state_floats(46, 500) == -1
    adam = _lib.AdamW(1e-3, 0.9, 0.999, 1e-8, 0.0, 0.0)
    assert L.idff_train_md_steps(fake, None, 200, 46, None, None, None, 0, 10, 0.5, 1, C.byref(adam), 0, None, None, None) == 0
    assert L.idff_train_md_steps(fake, fake, 200, 46, fake, fake, None, 1, 17, 0.5, 1, C.byref(adam), 0, None, None, None) == -1
    assert b"batch" in L.idff_last_error()
    decay = _lib.AdamW(1e-3, 0.9, 0.999, 1e-8, 1e-2, 0.0)                            # torch.optim.Adam's weight_decay is L2, not served
    assert L.idff_train_md_steps(fake, fake, 200, 46, fake, fake, None, 1, 10, 0.5, 1, C.byref(decay), 0, None, None, None) == -1
    assert L.idff_debug_train_md_stamps(None) == -1
    assert L.idff_tune_small_n_route(0) == 0


def test_header_is_valid_c_and_links_from_a_c_program(tmp_path):
    """include/idff_b200.h is the boundary a non-Python host binds: it must compile as plain C (gcc -std=c99 -pedantic) and a C
    program that takes the address of every declared entry point must link against libidff_b200.so and run (no GPU touched)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    names = _declared_symbols()
    src = tmp_path / "abi.c"
    lines = ['#include "idff_b200.h"', "#include <stdio.h>", "int main(void) {", "  const void* fn[] = {"]
    lines += [f"    (const void*)&{n}," for n in names]
    lines += ["  };", "  unsigned i, n = 0;", "  for (i = 0; i < sizeof(fn) / sizeof(fn[0]); ++i) n += fn[i] != 0;",
              '  printf("%u %d %s\\n", n, idff_version(), idff_source_hash());',
              "  return idff_sde_num_steps(0, 0, 0.1f) == IDFF_EINVAL ? 0 : 1;", "}"]
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-Wno-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-lidff_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    n, ver, h = r.stdout.split()
    assert int(n) == len(names) and int(ver) >= 200 and h == _lib.expected_source_hash()


def test_sde_step_count_matches_the_oracle_grid_on_random_grids():
    """idff_sde_num_steps (torchsde's fixed-step grid: curr_t += dt in fp32, the last step clipped to ts[-1]) against the oracle's
    restatement on random output grids and step sizes -- host logic only."""
    import idff_oracle as O
    rng = np.random.default_rng(0)
    L = _lib.lib()
    for _ in range(300):
        n = int(rng.integers(2, 9))
        ts = np.sort(rng.uniform(0.0, 1.0, n)).astype(np.float32)
        ts[0], ts[-1] = 0.0, np.float32(rng.choice([1.0, 0.75, float(ts[-1]) if ts[-1] > ts[0] else 1.0]))
        ts = np.sort(ts)
        dt = float(np.float32(rng.choice([0.3, 0.1, 1.0 / 3.0, 0.25, float(rng.uniform(0.05, 0.6))])))
        got = L.idff_sde_num_steps(ts.ctypes.data_as(C.c_void_p), n, dt)
        want = len(O.sde_step_grid(torch.tensor(ts), dt)) - 1
        assert got == want, (ts, dt, got, want)
