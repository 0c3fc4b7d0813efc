"""CPU: the reference arm of bench.py (the unmodified reference program on a bounded sample) - no GPU needed."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _ref_binary():
    sys.path.insert(0, str(ROOT))
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle.ref_binary()


def test_reference_arm_runs_the_unmodified_program_on_all_cores():
    """Under torch.distributed.run OMP_NUM_THREADS is 1 (the N >= 2 reference numbers ran on one thread);
    the arm must set the thread count itself, keep the workload's config and say what it sampled."""
    if _ref_binary() is None:
        pytest.skip("oracle/_ref/WDPMCL_ref not built (needs /root/reference at build time)")
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                          "--sample-size-ref", "96"], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    cores = len(os.sched_getaffinity(0))
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "reference"
    assert line["cpu_baseline"]["cores"] == cores
    assert line["config"]["rows"] == 32768 and "2 row stripes" in line["config"]["partition"]  # the workload's config, as our arm prints it
    assert "96x96" in line["sample"] and "UNMODIFIED" in line["sample"]
    assert line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""
