"""The "%f" round trip without the text (include/wdpm_quantize.h): the host function against Python's own
correctly rounded formatting and parsing (CPU), and the device kernel + the HBM hand-over against the host
function (GPU)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import random_case

ROOT = Path(__file__).resolve().parent.parent
SRC = """
#include "wdpm_quantize.h"
void quantize_array(double *a, long n) { for (long i = 0; i < n; i++) a[i] = wdpm_quantize6(a[i]); }
"""


def _host_lib():
    out = Path(__file__).resolve().parent / "emul" / "_build" / "libquantize.so"
    out.parent.mkdir(exist_ok=True)
    hdr = ROOT / "include" / "wdpm_quantize.h"
    if not out.exists() or hdr.stat().st_mtime > out.stat().st_mtime:
        subprocess.run(["/usr/bin/gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-I", str(ROOT / "include"), "-x", "c", "-", "-o", str(out), "-lm"],
                       input=SRC.encode(), check=True)
    return C.CDLL(str(out))


def _quantize_host(a: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(a, dtype=np.float64).copy()
    _host_lib().quantize_array(b.ctypes.data_as(C.c_void_p), C.c_long(b.size))
    return b


def _cases(rng, n):
    u = rng.uniform(size=n)
    parts = [u * 0.5, u * 1e-5, -99999.0 * u,
             np.ldexp(np.floor(u * 2 ** 20), -7 - rng.integers(0, 8, n)),                 # dyadic values: exact decimal ties occur
             np.floor(u * 1e6) / 1e6 + (rng.uniform(size=n) - 0.5) * 1e-15,              # next to 6-decimal numbers
             (np.floor(u * 2e6) + 0.5) / 1e6 + rng.integers(-1, 2, n) * 1.1e-16,         # next to decimal ties
             np.array([0.0, -0.0, 1e-9, -1e-9, 5e-7, 0.0078125, 0.0234375, 123456.7890625, 4.4e9, 1e300, -4.9e-7])]
    return np.concatenate(parts)


def test_quantize6_is_the_text_round_trip():
    rng = np.random.default_rng(5)
    a = _cases(rng, 60_000)
    got = _quantize_host(a)
    want = np.array([float("%f" % v) if abs(v) < 4.5e9 else v for v in a])
    assert np.array_equal(got.view(np.int64), want.view(np.int64))  # bit for bit, sign of zero included


@pytest.mark.gpu
def test_device_quantize_and_hand_over_match_the_host_function(cuda_lib):
    from wdpm_b200 import F64, Solver, ascgrid
    rng = np.random.default_rng(6)
    D, W = random_case(rng, 120, 500, np.float64)
    W[1:-1, 1:-1] = np.where(D[1:-1, 1:-1] > -99999.0, rng.permutation(_cases(rng, 60_000))[: 120 * 500].reshape(120, 500).clip(0, 4e9), 0)
    dem, w = D[1:-1, 1:-1], W[1:-1, 1:-1]
    with Solver(120, 500, -99999.0, 0, dtype=F64, kernel=2, fused_variant=2) as a, Solver(120, 500, -99999.0, 2, dtype=F64) as b:
        a.upload(dem, w)
        b.copy_state_from(a)            # different module, different tiling (pitch): grids travel inside HBM
        assert np.array_equal(b.download_water(), w)
        b.quantize_water()
        want = np.where(dem > -99999.0, _quantize_host(w).reshape(w.shape), w)
        assert np.array_equal(b.download_water().view(np.int64), want.view(np.int64))
        r, c, e = b.find_outlet()       # the elevations arrived too
        assert e == dem[r - 1, c - 1]
