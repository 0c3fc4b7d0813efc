"""GPU (-m gpu): the time-to-converge driver (scripts/converge_dist.py) and its binary checkpoints."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
SCRIPT = ROOT / "scripts" / "converge_dist.py"


def _run(*args):
    res = subprocess.run([sys.executable, str(SCRIPT), *map(str, args)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


def test_checkpoint_resume_is_bit_exact(cuda_lib, tmp_path):
    """Six blocks in one go == three blocks, a checkpoint, three more after --resume: same grid checksum, same
    last max_diff (the binary side format keeps every bit; the reference's %f scratch file would not)."""
    common = ["--size", 700, "--block-iters", 40, "--tol-mm", 0.0001]
    whole = _run(*common, "--max-blocks", 6)
    ck = tmp_path / "ck"
    first = _run(*common, "--max-blocks", 3, "--checkpoint-dir", ck)
    assert (ck / "meta.json").exists() and (ck / "water_rank0.bin").stat().st_size == 700 * 700 * 8
    second = _run(*common, "--max-blocks", 6, "--checkpoint-dir", ck, "--resume", "--csv", tmp_path / "c.csv")
    assert first["iterations"] == 120 and second["iterations"] == whole["iterations"] == 240
    assert second["checksum"] == whole["checksum"] and second["checksum"] != first["checksum"]
    assert second["last_max_diff_m"] == whole["last_max_diff_m"]
    assert len((tmp_path / "c.csv").read_text().strip().splitlines()) == 7  # header + all six blocks, also the resumed ones


def test_stops_at_the_reference_criterion(cuda_lib):
    """A generous tolerance is met after the first block: converged, one block, not the budget."""
    r = _run("--size", 300, "--block-iters", 200, "--tol-mm", 5000.0, "--max-blocks", 50)
    assert r["converged"] and r["blocks"] == 1
