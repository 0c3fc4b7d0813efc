"""CPU: the fused kernel's schedule (wdpm_b200/csrc/mw_schedule.h) proven against the oracle.

tests/emul/mw_emul.cpp executes the CUDA kernel's exact sequence of bulk loads, tile relaxations
and write-backs on a model of the shared-memory row ring, with the kernel's own index arithmetic and
relax functions compiled for the host. Bit-equality with the oracle shows the halo widths, phase lags
and multi-iteration pipelining are right; its hazard counters show the ring is large enough."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import random_case
from oracle import pyoracle as po

HERE = Path(__file__).resolve().parent
SRC = HERE / "emul" / "mw_emul.cpp"
NCFG = 8


def _build(extra=(), name="libmw_emul.so"):
    out = HERE / "emul" / "_build" / name
    out.parent.mkdir(exist_ok=True)
    deps = [SRC, HERE.parent / "wdpm_b200" / "csrc" / "mw_schedule.h", HERE.parent / "wdpm_b200" / "csrc" / "relax.cuh"]
    if not out.exists() or any(d.stat().st_mtime > out.stat().st_mtime for d in deps):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                        *extra, str(SRC), "-o", str(out)], check=True)
    return C.CDLL(str(out))


@pytest.fixture(scope="module")
def emul():
    return _build()


def run_emul(lib, cfg, module, w, d, nodata, n_launches, chunk_triples, outlet=(-10, -10), td=0.0):
    sfx, ct = ("_f64", C.c_double) if w.dtype == np.float64 else ("_f32", C.c_float)
    R, Cc = w.shape[0] - 2, w.shape[1] - 2
    tdv = np.array([td], dtype=w.dtype)
    err = np.zeros(5, dtype=np.int64)
    rc = getattr(lib, "mw_emul_run" + sfx)(cfg, module, w.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), R, Cc,
                                          ct(nodata), n_launches, chunk_triples, outlet[0], outlet[1],
                                          tdv.ctypes.data_as(C.c_void_p), err.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return float(tdv[0]), err


def cfg_info(lib, cfg):
    a = [C.c_int() for _ in range(4)]
    lib.mw_emul_cfg_info(cfg, *[C.byref(x) for x in a])
    return dict(zip(("W", "TWV", "K", "NRING"), (x.value for x in a)))


@pytest.mark.parametrize("cfg", range(1, NCFG))
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_schedule_is_bit_exact(oracle, emul, cfg, dt):
    K = cfg_info(emul, cfg)["K"]
    rng = np.random.default_rng(100 + cfg)
    for rows, cols, chunk_triples in ((50, 70, 0), (61, 130, 5), (97, 45, 7), (1, 1, 0), (3, 200, 1)):
        for mod in (po.ADD, po.SUBTRACT, po.DRAIN):
            D, W = random_case(rng, rows, cols, dt)
            outlet = (oracle.find_outlet(D) or (1, 1)) if mod == po.DRAIN else (-10, -10)
            a, b = W.copy(), W.copy()
            ta = oracle.iterate(a, D, -99999.0, mod, 3 * K, outlet=outlet if mod == po.DRAIN else (0, 0), totaldrain=0.25)
            tb, err = run_emul(emul, cfg, mod, b, D, -99999.0, 3, chunk_triples, outlet, 0.25)
            assert not err.any(), (cfg, rows, cols, mod, err)
            assert np.array_equal(a, b), (cfg, rows, cols, mod)
            assert ta == tb


def test_production_window_one_iteration(oracle, emul):
    """cfg 0 = the 512-column production window, on a grid wider than one strip."""
    rng = np.random.default_rng(7)
    D, W = random_case(rng, 40, 1100, np.float64)
    a, b = W.copy(), W.copy()
    oracle.iterate(a, D, -99999.0, po.ADD, 2)
    _, err = run_emul(emul, 0, po.ADD, b, D, -99999.0, 2, 6)
    assert not err.any() and np.array_equal(a, b)


def test_drain_outlet_on_ownership_boundaries(oracle, emul):
    """The outlet's neighbours straddle strips and chunks: events must still fold in sub-pass order."""
    rng = np.random.default_rng(9)
    info = cfg_info(emul, 1)
    twv = info["TWV"]
    for orow, ocol in ((15, twv), (15, twv - 1), (16, twv + 1), (14, 2 * twv), (1, 1), (30, 60)):
        D, W = random_case(rng, 30, 60, np.float64, nodata_fraction=0.0, wet_fraction=1.0)
        a, b = W.copy(), W.copy()
        ta = oracle.iterate(a, D, -99999.0, po.DRAIN, 4, outlet=(orow, ocol), totaldrain=1.0)
        tb, err = run_emul(emul, 1, po.DRAIN, b, D, -99999.0, 4, 5, (orow, ocol), 1.0)
        assert not err.any() and np.array_equal(a, b) and ta == tb, (orow, ocol)


def test_hazard_checks_fire_when_ring_is_too_small():
    """Negative control: with three ring rows fewer the emulator must report a load landing on
    rows whose write-back is still in flight."""
    lib = _build(extra=("-DWDPM_NRING_DELTA=-3",), name="libmw_emul_shrunk.so")
    rng = np.random.default_rng(5)
    D, W = random_case(rng, 50, 70, np.float64)
    _, err = run_emul(lib, 7, po.ADD, W.copy(), D, -99999.0, 1, 0)
    assert err[0] > 0 or err[1] > 0


# ---- warp-autonomous schedule (k_fused_wa): two tiles per lane, windows sliding by warp shuffle ----

def run_wa(lib, cfg, module, mode, w, d, nodata, n_launches, chunk_triples):
    sfx, ct = ("_f64", C.c_double) if w.dtype == np.float64 else ("_f32", C.c_float)
    R, Cc = w.shape[0] - 2, w.shape[1] - 2
    err = np.zeros(5, dtype=np.int64)
    rc = getattr(lib, "wa_emul_run" + sfx)(cfg, module, mode, w.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), R, Cc,
                                          ct(nodata), n_launches, chunk_triples, err.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return err


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 5, 6])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_wa_schedule_is_bit_exact(oracle, emul, cfg, dt):
    """Grids narrower and wider than one strip (several warps per row triple, several strips), chunked rows."""
    rng = np.random.default_rng(300 + cfg)
    W_, TWV = C.c_int(), C.c_int()
    emul.wa_emul_cfg_info(cfg, C.byref(W_), C.byref(TWV))
    for rows, cols, chunk_triples in ((40, 70, 0), (31, TWV.value + 40, 4), (25, 2 * TWV.value + 7, 3), (1, 1, 0), (3, 200, 1)):
        for mod, mode in ((po.ADD, 0), (po.ADD, 1), (po.SUBTRACT, 0)):
            D, Wt = random_case(rng, rows, cols, dt)
            a, b = Wt.copy(), Wt.copy()
            oracle.iterate(a, D, -99999.0, mod, 3)
            err = run_wa(emul, cfg, mod, mode, b, D, -99999.0, 3, chunk_triples)
            assert not err.any(), (cfg, rows, cols, mod, err)
            assert np.array_equal(a, b), (cfg, rows, cols, mod, mode, int((a != b).sum()))


def test_wa_unguarded_add_on_a_clean_grid(oracle, emul):
    """No activity test at all (Add with the cap-free steps, kOptNoGuard): exact as long as the water is +0 wherever the reference
    skips the centre - dry cells, NODATA cells and the halo ring included - which is what the solver checks."""
    rng = np.random.default_rng(41)
    for cfg, rows, cols, ct, dt in ((0, 40, 300, 5, np.float64), (1, 33, 420, 0, np.float64), (0, 40, 300, 5, np.float32), (3, 20, 700, 0, np.float32)):
        D, Wt = random_case(rng, rows, cols, dt, wet_fraction=0.5, nodata_fraction=0.15)
        assert not np.signbit(Wt).any() and not Wt[D <= -99999.0].any()
        a, b = Wt.copy(), Wt.copy()
        oracle.iterate(a, D, -99999.0, po.ADD, 4)
        err = run_wa(emul, cfg, po.ADD, 3, b, D, -99999.0, 4, ct)
        assert not err.any() and np.array_equal(a, b)
        assert not np.signbit(b).any()  # the invariant is preserved


def test_wa_guard_handles_what_the_unguarded_step_must_not_see(oracle, emul):
    """Negative water, -0.0 and water on NODATA cells: the guarded step treats them as the reference does."""
    rng = np.random.default_rng(43)
    D, Wt = random_case(rng, 30, 260, np.float64, wet_fraction=0.8, nodata_fraction=0.1)
    Wt[rng.uniform(size=Wt.shape) < 0.05] = -0.25
    Wt[rng.uniform(size=Wt.shape) < 0.05] = -0.0
    Wt[D <= -99999.0] = 0.125
    a, b = Wt.copy(), Wt.copy()
    oracle.iterate(a, D, -99999.0, po.ADD, 3)
    err = run_wa(emul, 0, po.ADD, 0, b, D, -99999.0, 3, 4)
    assert not err.any() and np.array_equal(a, b) and np.array_equal(np.signbit(a), np.signbit(b))


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_wa_staggered_schedule_is_bit_exact(oracle, emul, dt):
    """kOptStagger: triple slot 1 runs half a step behind slot 0, windows stay in registers across the CTA barrier in
    between, rows go home half a step after they are finished. Same ring (no extra rows): the emulator's hazard and
    same-half-step race counters must stay at zero."""
    rng = np.random.default_rng(77)
    W_, TWV = C.c_int(), C.c_int()
    for cfg in (1, 4):
        emul.wa_emul_cfg_info(cfg, C.byref(W_), C.byref(TWV))
        for rows, cols, chunk_triples in ((40, 70, 0), (31, TWV.value + 40, 4), (25, 2 * TWV.value + 7, 3), (1, 1, 0), (3, 200, 1), (64, 90, 7)):
            for mod, mode in ((po.ADD, 4), (po.ADD, 5), (po.SUBTRACT, 4)):
                D, Wt = random_case(rng, rows, cols, dt)
                a, b = Wt.copy(), Wt.copy()
                oracle.iterate(a, D, -99999.0, mod, 3)
                err = run_wa(emul, cfg, mod, mode, b, D, -99999.0, 3, chunk_triples)
                assert not err.any(), (cfg, rows, cols, mod, err)
                assert np.array_equal(a, b), (cfg, rows, cols, mod, mode, int((a != b).sum()))
    # unguarded on a clean grid
    D, Wt = random_case(rng, 33, 420, dt, wet_fraction=0.5, nodata_fraction=0.15)
    a, b = Wt.copy(), Wt.copy()
    oracle.iterate(a, D, -99999.0, po.ADD, 4)
    err = run_wa(emul, 1, po.ADD, 7, b, D, -99999.0, 4, 0)
    assert not err.any() and np.array_equal(a, b)


def test_wa_staggered_hazard_checks_fire_when_ring_is_too_small():
    lib = _build(extra=("-DWDPM_NRING_DELTA=-3",), name="libmw_emul_shrunk.so")
    rng = np.random.default_rng(5)
    D, Wt = random_case(rng, 50, 400, np.float64)
    err = run_wa(lib, 1, po.ADD, 4, Wt.copy(), D, -99999.0, 1, 0)
    assert err[0] > 0 or err[1] > 0


def test_wa_deep_prefetch_needs_the_strict_order():
    """cfg 5 loads two steps ahead on the ring of a one-step prefetch. That only works because the loads of a step go
    out after its write-backs were read out; a ring three rows shorter must trip the hazard counters."""
    lib = _build(extra=("-DWDPM_NRING_DELTA=-3",), name="libmw_emul_shrunk.so")
    rng = np.random.default_rng(6)
    D, Wt = random_case(rng, 60, 400, np.float64)
    err = run_wa(lib, 5, po.ADD, 0, Wt.copy(), D, -99999.0, 1, 0)
    assert err[0] > 0 or err[1] > 0


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_wa_drain_is_bit_exact(oracle, emul, dt):
    """Drain in the warp-autonomous schedule: warps with an outlet mark in reach take the outlet walk (runoff.cl:104-111),
    contacts are recorded once - by a storing lane of the CTA that owns the centre - and fold in sub-pass order.
    Outlets on strip borders, warp borders (window column 186 of a two-warp row triple), chunk borders and the rim."""
    rng = np.random.default_rng(91)
    W_, TWV = C.c_int(), C.c_int()
    for cfg in (0, 1):
        emul.wa_emul_cfg_info(cfg, C.byref(W_), C.byref(TWV))
        twv = TWV.value
        rows, cols = 36, twv + 60
        spots = [(15, twv), (15, twv - 1), (16, twv + 1), (1, 1), (rows, cols), (12, 186 - 12 + 1), (13, 186 - 12), (18, 7), (21, twv + 30)]
        for orow, ocol in spots:
            for mode in (0, 1):
                D, Wt = random_case(rng, rows, cols, dt, nodata_fraction=0.02, wet_fraction=0.95)
                D[orow, ocol] = dt(480.0)
                a, b = Wt.copy(), Wt.copy()
                ta = oracle.iterate(a, D, -99999.0, po.DRAIN, 4, outlet=(orow, ocol), totaldrain=1.0)
                sfx, ct = ("_f64", C.c_double) if dt == np.float64 else ("_f32", C.c_float)
                tdv = np.array([1.0], dtype=dt)
                err = np.zeros(5, dtype=np.int64)
                rc = getattr(emul, "wa_emul_drain" + sfx)(cfg, mode, b.ctypes.data_as(C.c_void_p), D.ctypes.data_as(C.c_void_p), rows, cols, ct(-99999.0),
                                                          4, 5, orow, ocol, tdv.ctypes.data_as(C.c_void_p), err.ctypes.data_as(C.c_void_p))
                assert rc == 0 and not err.any(), (cfg, orow, ocol, mode, err)
                assert np.array_equal(a, b), (cfg, orow, ocol, mode, int((a != b).sum()))
                assert dt(ta) == tdv[0], (cfg, orow, ocol, mode)
