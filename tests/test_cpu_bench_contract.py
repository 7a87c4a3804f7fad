"""CPU-only checks of the bench contract and of the constants shared between the header and the Python binding."""
import json
import os
import re
import subprocess
import sys

from conftest import ROOT
from idff_b200 import _lib


def test_reference_arm_prints_one_contract_line():
    """bench.py --impl reference times the oracle on the host cores and prints ONE JSON line with the contract keys."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-particles", "512"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "idff_particle_steps_per_sec" and d["unit"] == "particle-steps/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["vs_baseline"] is None and d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1", "--cpu-particles", "512"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_header_constants_match_binding():
    src = open(os.path.join(ROOT, "include", "idff_b200.h")).read()
    consts = dict(re.findall(r"#define\s+(IDFF_(?:FIELD|IMPL)_[A-Z]+)\s+(-?\d+)", src))
    for name, val in consts.items():
        assert getattr(_lib, name) == int(val), name
    assert {"IDFF_IMPL_AUTO", "IDFF_IMPL_GENERIC", "IDFF_IMPL_FUSED", "IDFF_IMPL_TENSOR", "IDFF_IMPL_LAYERED"} <= set(consts)


def test_flop_model_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.MAC_PER_PASS == 132352                                   # F(3,256,2)
    assert abs(bench.flops_per_particle_solve(2) - 6.88e6) < 0.01e6        # 2F[(4 nfe - 2) 2n + 2], nfe = 2, n = 2
