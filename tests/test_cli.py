"""The command-line host (wdpm_b200/host/wdpm_host.c) as a drop-in for WDPMCL.

CPU: usage texts and exit codes equal the unmodified reference binary's (tests/golden/usage_*.txt).
GPU: validate_WDPM.sh's Add -> Drain -> Subtract sequence on basin5 through the binary; output files
byte-identical to the reference's OpenCL-branch files, report lines equal apart from run times, the
backend lines and file paths."""
import re
import subprocess
from pathlib import Path

import pytest

from conftest import GOLDEN, golden_text, gunzip_to

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def host_bin(cuda_lib):
    from wdpm_b200 import build
    return build.build_host()


@pytest.mark.parametrize("name,argv", [("usage_all", []), ("usage_add", ["add"]), ("usage_subtract", ["subtract"]),
                                       ("usage_drain", ["drain"]), ("usage_badcount", ["add", "a", "b", "c"])])
def test_usage_matches_reference(host_bin, name, argv):
    gold = (GOLDEN / f"{name}.txt").read_text()
    code, text = gold.split("\n", 1)
    res = subprocess.run([str(host_bin)] + argv, capture_output=True, text=True)
    assert res.returncode == int(code.split("=")[1]) == 42
    assert res.stdout == text


def test_drain_without_water_file_exits_42(host_bin, tmp_path):
    dem = gunzip_to("basin5.asc.gz", tmp_path / "basin5.asc")
    res = subprocess.run([str(host_bin), "drain", str(dem), str(tmp_path / "missing.asc"), str(tmp_path / "o.asc"), "NULL",
                          "0.1", "1.0", "1", "1", "0.005", "0"], capture_output=True, text=True)
    assert res.returncode == 42 and "Error water file missing" in res.stdout


def _normalise(report: str) -> list[str]:
    out = []
    for ln in report.split("\n"):
        if re.search(r"(DEM|Water|Output|Scratch) file:", ln) or "for Computation" in ln or "backend flags" in ln or "Run Time" in ln:
            continue
        toks = ln.split()
        if toks and toks[0].isdigit() and len(toks) in (3, 5):  # per-block line: drop the run-time column
            ln = " ".join(toks[:-1])
        out.append(ln.rstrip())
    return out


@pytest.mark.gpu
def test_validation_sequence_through_the_binary(host_bin, tmp_path):
    from test_oracle import _check_goldens
    dem = gunzip_to("basin5.asc.gz", tmp_path / "basin5.asc")
    add, drain, sub = tmp_path / "add.asc", tmp_path / "drain.asc", tmp_path / "sub.asc"
    runs = [
        ("add10", ["add", dem, "NULL", add, "NULL", "10", "1.0", "1.0", "1", "1", "0.005", "0"], add),
        ("drain", ["drain", dem, add, drain, "NULL", "0.1", "1.0", "1", "1", "0.005", "0"], drain),
        ("sub10", ["subtract", dem, drain, sub, "NULL", "10", "1.0", "1", "1", "0.005", "0"], sub),
    ]
    for name, argv, out in runs:
        res = subprocess.run([str(host_bin)] + [str(a) for a in argv], capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stdout[-500:] + res.stderr
        text = out.read_text()
        assert text == golden_text(f"ref_opencl_{name}.asc.gz"), name
        _check_goldens(name, text)
        assert _normalise(res.stdout) == _normalise((GOLDEN / f"ref_opencl_{name}.txt").read_text()), name


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", ["1", "2"])
def test_chained_modules_write_the_same_files_as_separate_runs(host_bin, tmp_path, gpus):
    """`wdpmcl_b200 chain`: Add -> Drain -> Subtract in one process, grids handed on in HBM (one GPU) or in
    host memory (stripes) with the "%f" quantisation a file gives: every output file and report must equal
    the unmodified reference's, exactly as for three separate runs."""
    import os
    from test_oracle import _check_goldens
    dem = gunzip_to("basin5.asc.gz", tmp_path / "basin5.asc")
    add, drain, sub = tmp_path / "add.asc", tmp_path / "drain.asc", tmp_path / "sub.asc"
    pfs = []
    for name, body in (("add10", f"add {dem} NULL {add} NULL 10 1.0 1.0 1 1 0.005 0"),
                       ("drain", f"drain {dem} {add} {drain} NULL 0.1 1.0 1 1 0.005 0"),
                       ("sub10", f"subtract {dem} {drain} {sub} NULL 10 1.0 1 1 0.005 0")):
        pf = tmp_path / f"{name}.txt"
        pf.write_text(body + "\n")
        pfs.append(str(pf))
    res = subprocess.run([str(host_bin), "chain"] + pfs, capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, WDPM_B200_GPUS=gpus))
    assert res.returncode == 0, res.stdout[-800:] + res.stderr
    for name, out in (("add10", add), ("drain", drain), ("sub10", sub)):
        text = out.read_text()
        assert text == golden_text(f"ref_opencl_{name}.asc.gz"), name
        _check_goldens(name, text)
    # the chained stdout is the three reports one after the other
    reports = res.stdout.split("WDPM run summary")
    assert len(reports) == 4
    want = []
    for name in ("add10", "drain", "sub10"):
        want += _normalise((GOLDEN / f"ref_opencl_{name}.txt").read_text())
    assert [ln for ln in _normalise(res.stdout) if ln] == [ln for ln in want if ln]  # blank lines at the seams aside


@pytest.mark.gpu
def test_parameter_file_and_scratch_resume(host_bin, tmp_path):
    """Parameter-file form (WDPMCL.c:334-342) with an iteration limit and a scratch file, then a resume
    from that scratch file: the two-leg run must end where the uninterrupted run ends."""
    dem = gunzip_to("basin5.asc.gz", tmp_path / "basin5.asc")
    full, part, scratch = tmp_path / "full.asc", tmp_path / "part.asc", tmp_path / "scratch.asc"
    pf = tmp_path / "params.txt"
    pf.write_text(f"add {dem} NULL {part} {scratch}\n10 1.0 1.0\n1 1 0.005 3000\n")
    r1 = subprocess.run([str(host_bin), str(pf)], capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0 and "No Scratch file found" in r1.stdout and scratch.exists()
    assert "Maximum number of iterations: 3000" in r1.stdout
    r2 = subprocess.run([str(host_bin), "add", str(dem), "NULL", str(part), str(scratch), "10", "1.0", "1.0", "1", "1", "0.005", "0"],
                        capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0 and "Scratch file found" in r2.stdout
    # the scratch file holds "%f" text, so the resumed leg starts from values rounded to 1e-6 m:
    # it converges to the same ponds within that rounding, not bit for bit
    import numpy as np
    from wdpm_b200 import ascgrid
    _, w = ascgrid.read_asc(part)
    _, ref = ascgrid.read_asc(GOLDEN / "ref_opencl_add10.asc.gz")
    valid = ref >= 0
    assert abs(w[valid].sum() - ref[valid].sum()) / ref[valid].sum() < 1e-4
    assert np.abs(w - ref)[valid].max() < 1e-2


@pytest.mark.parametrize("name", ["ref_opencl_add10", "ref_serial_drain", "ref_opencl_sub10", "ref_opencl_add300"])
def test_asc_reader_and_writer_are_byte_exact(host_bin, tmp_path, name):
    """The host's parallel parser + formatter (wdpm_host.c read_grid / write_grid): reading a file the
    reference wrote (fprintf "%f ", WDPMCL.c:1549) and writing it back must give the same bytes."""
    src = gunzip_to(f"{name}.asc.gz", tmp_path / "in.asc")
    out = tmp_path / "out.asc"
    res = subprocess.run([str(host_bin), "--asc-roundtrip", str(src), str(out)], capture_output=True, text=True)
    assert res.returncode == 0
    assert out.read_bytes() == src.read_bytes()


def test_asc_reader_parses_the_dem_like_strtod(host_bin, tmp_path):
    """basin5 has 4-decimal values and a differently formatted header; after a round trip through the
    host every value must equal numpy's (correctly rounded) parse of the original text."""
    import numpy as np
    from wdpm_b200 import ascgrid
    src = gunzip_to("basin5.asc.gz", tmp_path / "basin5.asc")
    out = tmp_path / "out.asc"
    assert subprocess.run([str(host_bin), "--asc-roundtrip", str(src), str(out)]).returncode == 0
    h0, a = ascgrid.read_asc(src)
    h1, b = ascgrid.read_asc(out)
    assert (h0.ncols, h0.nrows, h0.cellsize, h0.nodata) == (h1.ncols, h1.nrows, h1.cellsize, h1.nodata)
    assert np.array_equal(np.round(a, 6), b)  # "%f" keeps six decimals; the DEM has four


@pytest.mark.gpu
def test_binary_on_several_gpus_gives_the_same_files(host_bin, tmp_path):
    """WDPM_B200_GPUS=N: one host thread drives N stripe solvers (begin / enqueue / end); the validation
    sequence must still reproduce the reference's files byte for byte."""
    import os
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run under gpurun --gpus 2)")
    env = dict(os.environ, WDPM_B200_GPUS=str(min(n, 4)))
    dem = gunzip_to("basin5.asc.gz", tmp_path / "basin5.asc")
    add, drain, sub = tmp_path / "add.asc", tmp_path / "drain.asc", tmp_path / "sub.asc"
    runs = [
        ("add10", ["add", dem, "NULL", add, "NULL", "10", "1.0", "1.0", "1", "1", "0.005", "0"], add),
        ("drain", ["drain", dem, add, drain, "NULL", "0.1", "1.0", "1", "1", "0.005", "0"], drain),
        ("sub10", ["subtract", dem, drain, sub, "NULL", "10", "1.0", "1", "1", "0.005", "0"], sub),
    ]
    for name, argv, out in runs:
        res = subprocess.run([str(host_bin)] + [str(a) for a in argv], capture_output=True, text=True, timeout=900, env=env)
        assert res.returncode == 0, res.stdout[-800:] + res.stderr
        assert out.read_text() == golden_text(f"ref_opencl_{name}.asc.gz"), name
        assert _normalise(res.stdout) == _normalise((GOLDEN / f"ref_opencl_{name}.txt").read_text()), name
