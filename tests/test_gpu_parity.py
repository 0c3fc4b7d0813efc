"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle - bit for bit."""
import numpy as np
import pytest

from conftest import GOLDEN, golden_text, random_case
from oracle import pyoracle as po
from wdpm_b200 import ascgrid

pytestmark = pytest.mark.gpu

MODS = [("add", 0), ("subtract", 1), ("drain", 2)]
NODATA = -99999.0


def _solver(cuda_lib, D, W, module, dtype, **kw):
    from wdpm_b200 import F32, F64, Solver
    s = Solver(D.shape[0] - 2, D.shape[1] - 2, NODATA, module, dtype=F64 if dtype == np.float64 else F32, **kw)
    s.upload(D[1:-1, 1:-1], W[1:-1, 1:-1])
    return s


def _cuda_iterate(cuda_lib, D, W, module, dtype, n, outlet=None, td=0.0, **kw):
    s = _solver(cuda_lib, D, W, module, dtype, **kw)
    if module == 2:
        s.set_outlet(*outlet)
        s.set_total_drain(td)
    s.iterate(n)
    out = s.download_water()
    tdo = s.get_total_drain()
    info = s.info()
    s.close()
    return ascgrid.pad_grid(out, dtype(0)), tdo, info


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("mn,mod", MODS)
def test_colour_kernel_matches_oracle(cuda_lib, oracle, dt, mn, mod):
    from wdpm_b200 import KERNEL_COLOUR
    rng = np.random.default_rng(21)
    for rows, cols in ((1, 1), (2, 7), (33, 65), (100, 131), (257, 190)):
        D, W = random_case(rng, rows, cols, dt)
        outlet = oracle.find_outlet(D) or (1, 1)
        a = W.copy()
        ta = oracle.iterate(a, D, NODATA, mod, 25, outlet=outlet, totaldrain=0.5)
        b, tb, _ = _cuda_iterate(cuda_lib, D, W, mod, dt, 25, outlet=outlet, td=0.5, kernel=KERNEL_COLOUR)
        assert np.array_equal(a, b), (rows, cols)
        if mod == 2:
            assert dt(ta) == dt(tb)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("mn,mod", MODS)
def test_resident_kernel_matches_oracle(cuda_lib, oracle, dt, mn, mod):
    """The small-grid kernel: one cooperative launch for all iterations, tiles resident in shared memory."""
    from wdpm_b200 import KERNEL_RESIDENT
    rng = np.random.default_rng(23)
    for rows, cols, n in ((1, 1, 3), (2, 7, 4), (33, 65, 25), (100, 131, 25), (257, 190, 12), (482, 471, 7), (40, 900, 8)):
        D, W = random_case(rng, rows, cols, dt)
        outlet = oracle.find_outlet(D) or (1, 1)
        a = W.copy()
        ta = oracle.iterate(a, D, NODATA, mod, n, outlet=outlet, totaldrain=0.5)
        b, tb, info = _cuda_iterate(cuda_lib, D, W, mod, dt, n, outlet=outlet, td=0.5, kernel=KERNEL_RESIDENT)
        assert info["kernel"] == KERNEL_RESIDENT
        assert np.array_equal(a, b), (rows, cols, int((a != b).sum()))
        if mod == 2:
            assert dt(ta) == dt(tb)


def _variants(dtype_code):
    from wdpm_b200 import solver
    v, out = 1, []
    while solver.fused_variant_info(v, dtype_code) is not None:
        out.append(v)
        v += 1
    return out


def _implements(variant, code, mod):
    from wdpm_b200 import Solver
    from wdpm_b200.solver import WdpmError
    try:
        Solver(8, 8, NODATA, mod, dtype=code, kernel=2, fused_variant=variant).close()
        return True
    except WdpmError as e:
        assert e.code == -5
        return False


WA_VARIANTS = {np.float64: [17, 18, 19, 20, 21, 22, 23, 24, 25], np.float32: [14, 15, 16, 17, 18, 19, 20, 21, 22, 23]}


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("mod", [0, 1])
def test_warp_autonomous_blocks_match_oracle(cuda_lib, oracle, dt, mod):
    """k_fused_wa through wdpm_run_block: after the first prologue with a zero threshold > 0 fp64 Add switches to the
    unguarded step (water is then known to be +0 wherever the reference skips a centre); grids wider than a strip,
    chunked rows, NODATA values in the water file, half the cells dry."""
    from wdpm_b200 import F32, F64, Solver
    rng = np.random.default_rng(61)
    for variant in WA_VARIANTS[dt]:
        for rows, cols, chunk_rows in ((70, 450, 0), (64, 1200, 12)):
            D, W = random_case(rng, rows, cols, dt, depth=0.05, wet_fraction=0.5, nodata_fraction=0.1)
            W[D <= NODATA] = dt(NODATA)  # as a water file written by the Add module has it
            thres = 0.004
            a = W.copy()
            s = Solver(rows, cols, NODATA, mod, dtype=F64 if dt == np.float64 else F32, zero_threshold=thres, kernel=2,
                       fused_variant=variant, fused_chunk_rows=chunk_rows)
            s.upload(D[1:-1, 1:-1], W[1:-1, 1:-1])
            for _ in range(3):
                md, ms, _ = oracle.block(a, D, NODATA, mod, dt(thres), 9)
                r = s.run_block(9)
                assert r.max_diff == md, (variant, rows, cols)
                assert abs(r.masked_sum - ms) <= 1e-12 * abs(ms)
            b = ascgrid.pad_grid(s.download_water(), dt(0))
            s.close()
            # NODATA cells of the water file: the reference's threshold pass zeroes them (WDPMCL.c:1055-1065)
            assert np.array_equal(a, b), (variant, rows, cols, int((a != b).sum()))


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("mn,mod", MODS)
def test_fused_kernel_matches_oracle_all_variants(cuda_lib, oracle, dt, mn, mod):
    """Every tiling variant (window width, triples per step, iterations per launch), several chunkings."""
    from wdpm_b200 import F32, F64, KERNEL_FUSED, solver
    code = F64 if dt == np.float64 else F32
    rng = np.random.default_rng(31)
    for variant in _variants(code):
        K = solver.fused_variant_info(variant, code)["iters_per_launch"]
        if mod == 2 and not _implements(variant, code, mod):
            continue  # the warp-autonomous variants implement Add and Subtract (Drain keeps k_fused)
        for rows, cols, chunk_rows in ((50, 70, 0), (61, 130, 15), (97, 700, 21), (1, 1, 0), (200, 45, 48)):
            D, W = random_case(rng, rows, cols, dt)
            outlet = oracle.find_outlet(D) or (1, 1)
            n = 3 * K + (1 if K > 1 else 0)  # also exercise the remainder path
            a = W.copy()
            ta = oracle.iterate(a, D, NODATA, mod, n, outlet=outlet, totaldrain=0.25)
            b, tb, info = _cuda_iterate(cuda_lib, D, W, mod, dt, n, outlet=outlet, td=0.25, kernel=KERNEL_FUSED,
                                        fused_variant=variant, fused_chunk_rows=chunk_rows)
            assert info["kernel"] == KERNEL_FUSED
            assert np.array_equal(a, b), (variant, rows, cols, chunk_rows, int((a != b).sum()))
            if mod == 2:
                assert dt(ta) == dt(tb), (variant, rows, cols)


def test_fused_drain_outlet_on_ownership_boundaries(cuda_lib, oracle):
    from wdpm_b200 import F64, KERNEL_FUSED, solver
    twv = solver.fused_variant_info(2, F64)["strip_cols"]
    rng = np.random.default_rng(9)
    for orow, ocol in ((15, twv), (15, twv - 1), (16, twv + 1), (14, 2 * twv), (1, 1), (30, 60)):
        D, W = random_case(rng, 30, 60, np.float64, nodata_fraction=0.0, wet_fraction=1.0)
        a = W.copy()
        ta = oracle.iterate(a, D, NODATA, po.DRAIN, 4, outlet=(orow, ocol), totaldrain=1.0)
        b, tb, _ = _cuda_iterate(cuda_lib, D, W, 2, np.float64, 4, outlet=(orow, ocol), td=1.0, kernel=KERNEL_FUSED,
                                 fused_variant=2, fused_chunk_rows=15)
        assert np.array_equal(a, b) and ta == tb, (orow, ocol)


def _adversarial_case(rng, rows, cols, dt, kind):
    """Inputs built to sit on the arithmetic's edges rather than in its middle."""
    nodata = NODATA
    if kind == "flat":            # exactly flat terrain: every comparison is a tie broken by water alone
        d = np.full((rows, cols), 500.0)
        w = rng.uniform(0, 0.3, d.shape)
    elif kind == "ulp":           # elevations a few ulps apart, water of the order of one ulp of the elevation
        d = 500.0 + np.float64(np.spacing(dt(500.0))) * rng.integers(-3, 4, (rows, cols))
        w = np.float64(np.spacing(dt(500.0))) * rng.uniform(0, 4, d.shape)
    elif kind == "tiny":          # water from the subnormal range up to millimetres
        d = (500 + 3 * rng.standard_normal((rows, cols))).round(4)
        lo = -44.0 if dt == np.float32 else -320.0
        w = 10.0 ** rng.uniform(lo, -3, d.shape)
    elif kind == "negative":      # negative and zero elevations (Drain's outlet rule wants dem > 0)
        d = (-50 + 30 * rng.standard_normal((rows, cols))).round(3)
        d[rng.uniform(size=d.shape) < 0.1] = 0.0
        w = rng.uniform(0, 2.0, d.shape)
    elif kind == "all_nodata":
        d = np.full((rows, cols), nodata)
        w = rng.uniform(0, 0.3, d.shape)
    else:                         # "dry": nothing to move
        d = (500 + 3 * rng.standard_normal((rows, cols))).round(4)
        w = np.zeros_like(d)
    if kind not in ("all_nodata", "flat"):
        d[rng.uniform(size=d.shape) < 0.05] = nodata
    w[rng.uniform(size=w.shape) < 0.2] = 0
    D = ascgrid.pad_grid(d.astype(dt), dt(nodata))
    W = ascgrid.pad_grid(np.where(d > nodata, w, 0).astype(dt), dt(0))
    return D, W


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["flat", "ulp", "tiny", "negative", "all_nodata", "dry"])
def test_adversarial_inputs_match_oracle(cuda_lib, oracle, dt, kind):
    """Ties, water at the rounding error of the surface sums, subnormal water, negative elevations, grids
    without a valid cell, dry grids, single rows and columns - all kernels, all modules, zero threshold 0."""
    from wdpm_b200 import F32, F64
    from wdpm_b200.solver import PRODUCTION_FUSED_VARIANT_F64
    code = F64 if dt == np.float64 else F32
    rng = np.random.default_rng(500 + ["flat", "ulp", "tiny", "negative", "all_nodata", "dry"].index(kind))
    variants = [(1, 0), (2, 2), (3, 0)] + ([(2, PRODUCTION_FUSED_VARIANT_F64), (2, 13), (2, 18)] if dt == np.float64 else [(2, 7), (2, 15)])
    for rows, cols in ((40, 70), (1, 90), (75, 1)):
        D, W = _adversarial_case(rng, rows, cols, dt, kind)
        for mod in (0, 1, 2):
            outlet = oracle.find_outlet(D)
            if mod == 2 and outlet is None:
                continue  # no cell with dem > 0: the reference has no outlet either
            a = W.copy()
            ta = oracle.iterate(a, D, NODATA, mod, 12, outlet=outlet or (0, 0), totaldrain=0.0)
            for kernel, variant in variants:
                if mod == 2 and kernel == 2 and variant and not _implements(variant, code, mod):
                    continue
                b, tb, _ = _cuda_iterate(cuda_lib, D, W, mod, dt, 12, outlet=outlet, td=0.0, kernel=kernel, fused_variant=variant)
                assert np.array_equal(a, b), (kind, rows, cols, mod, kernel, variant, int((a != b).sum()))
                if mod == 2:
                    assert dt(ta) == dt(tb), (kind, rows, cols, kernel, variant)


def test_drain_without_a_positive_elevation_has_no_outlet(cuda_lib):
    from wdpm_b200 import F64, Solver
    from wdpm_b200.solver import WdpmError
    dem = np.full((20, 30), -3.0)
    with Solver(20, 30, NODATA, 2, dtype=F64) as s:
        s.upload(dem, np.full_like(dem, 0.1))
        with pytest.raises(WdpmError):
            s.find_outlet()
        with pytest.raises(WdpmError):
            s.iterate(1)  # Drain refuses to run without an outlet


def _outlet_set(D, rng, twv):
    """Outlets that stress the bookkeeping: neighbours of each other, on strip / chunk borders, on the rim."""
    R, Cc = D.shape[0] - 2, D.shape[1] - 2
    want = [(1, 1), (1, 2), (2, 2), (R, Cc), (15, min(twv, Cc)), (15, min(twv + 1, Cc)), (16, min(twv - 1, Cc)), (30, 7), (32, 9), (R // 2, 1)]
    for _ in range(6):
        want.append((int(rng.integers(1, R + 1)), int(rng.integers(1, Cc + 1))))
    out = []
    for rc in want:
        if D[rc] > NODATA and rc not in out:
            out.append(rc)
    return out


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("kernel,variant", [(1, 0), (2, 1), (2, 2), (2, 3), (3, 0), (2, 17), (2, 18)])
def test_drain_outlet_set_matches_oracle(cuda_lib, oracle, dt, kernel, variant):
    """Extension (BASELINE configs[4]): a set of outlets. Water grid and every outlet's total, bit for bit
    (variants 17 / 18: the warp-autonomous kernel's outlet path)."""
    from wdpm_b200 import F32, F64, Solver, solver
    code = F64 if dt == np.float64 else F32
    twv = solver.fused_variant_info(2, code)["strip_cols"]
    rng = np.random.default_rng(77)
    for rows, cols, chunk_rows, n in ((60, 90, 15, 7), (45, 200, 0, 5)):
        D, W = random_case(rng, rows, cols, dt, nodata_fraction=0.04, wet_fraction=0.9)
        outlets = _outlet_set(D, rng, twv)
        assert len(outlets) >= 10
        a = W.copy()
        t0 = np.zeros(len(outlets), dtype=dt)
        t0[0] = dt(0.5)
        ta = oracle.iterate_outlets(a, D, NODATA, n, outlets, totals=t0)
        s = Solver(rows, cols, NODATA, 2, dtype=code, kernel=kernel, fused_variant=variant, fused_chunk_rows=chunk_rows if kernel == 2 else 0)
        s.upload(D[1:-1, 1:-1], W[1:-1, 1:-1])
        s.set_outlets(outlets)
        s.set_total_drain(0.5)
        s.iterate(n)
        b = ascgrid.pad_grid(s.download_water(), dt(0))
        tb = s.get_outlet_drains(len(outlets)).astype(dt)
        total = s.get_total_drain()
        s.close()
        assert np.array_equal(a, b), (rows, cols, int((a != b).sum()))
        assert np.array_equal(ta, tb), (ta, tb)
        acc = ta[0]
        for v in ta[1:]:
            acc = dt(acc + v)
        assert dt(total) == acc


def test_outlet_set_can_be_replaced_and_survives_upload(cuda_lib, oracle):
    """Marks are taken out when the set changes, and put back after a fresh DEM upload."""
    from wdpm_b200 import F64, Solver
    rng = np.random.default_rng(78)
    D, W = random_case(rng, 40, 50, np.float64, nodata_fraction=0.0, wet_fraction=1.0)
    first, second = [(5, 5), (6, 6), (20, 30)], [(10, 10), (33, 41)]
    s = Solver(40, 50, NODATA, 2, dtype=F64, kernel=2, fused_variant=2)
    s.upload(D[1:-1, 1:-1], W[1:-1, 1:-1])
    s.set_outlets(first)
    s.set_outlets(second)          # the cells of `first` must get their elevations back
    s.upload(D[1:-1, 1:-1], W[1:-1, 1:-1])   # and `second` must be marked again in the new grid
    s.set_total_drain(0.0)
    s.iterate(4)
    b = ascgrid.pad_grid(s.download_water(), 0.0)
    tb = s.get_outlet_drains(2)
    r, c, _ = s.find_outlet()      # searches true elevations, then installs the reference's single outlet
    s.close()
    a = W.copy()
    ta = oracle.iterate_outlets(a, D, NODATA, 4, second)
    assert np.array_equal(a, b) and np.array_equal(ta, tb)
    assert (r, c) == oracle.find_outlet(D)


@pytest.mark.parametrize("dn,dt", [("f64", np.float64), ("f32", np.float32)])
@pytest.mark.parametrize("mn,mod", MODS)
@pytest.mark.parametrize("kernel", [1, 2])
def test_golden_vectors_from_verbatim_kernel(cuda_lib, dn, dt, mn, mod, kernel):
    """tests/golden/crop_vectors.npz was produced by the verbatim runoff.cl in the build container."""
    g = np.load(GOLDEN / "crop_vectors.npz")
    D = ascgrid.pad_grid(g["dem"].astype(dt), dt(g["nodata"]))
    W = g[f"w0_{dn}"]
    b, tb, _ = _cuda_iterate(cuda_lib, D, W, mod, dt, int(g["iters"]), outlet=tuple(int(x) for x in g["outlet"]), td=0.0,
                             kernel=kernel, fused_variant=2 if kernel == 2 else 0, fused_chunk_rows=24)
    assert np.array_equal(b, g[f"w_{mn}_{dn}"])
    if mod == 2:
        assert dt(tb) == g[f"td_{mn}_{dn}"]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("kernel", [1, 2])
def test_run_block_matches_oracle_block(cuda_lib, oracle, dt, kernel):
    """Threshold + snapshot + iterations + masked max-diff / sum / totaldrain (WDPMCL.c:1055-1268)."""
    from wdpm_b200 import F32, F64, Solver
    rng = np.random.default_rng(41)
    for mod in (0, 1, 2):
        D, W = random_case(rng, 120, 150, dt, depth=0.02)
        W[5:20, 5:20] = dt(-99999.0)  # water files carry NODATA values; the threshold pass must clear them
        outlet = oracle.find_outlet(D)
        thres = 0.004
        a = W.copy()
        td = 0.125
        s = Solver(120, 150, NODATA, mod, dtype=F64 if dt == np.float64 else F32, zero_threshold=thres, kernel=kernel,
                   fused_variant=2 if kernel == 2 else 0, fused_chunk_rows=30)
        s.upload(D[1:-1, 1:-1], W[1:-1, 1:-1])
        if mod == 2:
            s.set_outlet(*outlet)
            s.set_total_drain(td)
        for _ in range(2):
            md, ms, td = oracle.block(a, D, NODATA, mod, dt(thres), 30, outlet=outlet, totaldrain=td)
            r = s.run_block(30)
            assert r.max_diff == md
            assert abs(r.masked_sum - ms) <= 1e-12 * abs(ms)
            if mod == 2:
                assert dt(r.total_drain) == dt(td)
            assert r.wet_cells == int(np.count_nonzero((a > 0) & (D > NODATA)))
        assert np.array_equal(ascgrid.pad_grid(s.download_water(), dt(0)), a)
        s.close()


def test_initial_conditions_and_outlet(cuda_lib, oracle):
    from oracle_backend import OracleBackend
    from wdpm_b200 import F64, Solver
    rng = np.random.default_rng(51)
    D, W = random_case(rng, 90, 110, np.float64, depth=0.05)
    dem, w = D[1:-1, 1:-1], W[1:-1, 1:-1]
    for what in ("add", "subtract"):
        s = Solver(90, 110, NODATA, 0, dtype=F64)
        s.upload(dem, w)
        ob = OracleBackend(90, 110, NODATA, 0, 0.0)
        ob.upload(dem, w)
        if what == "add":
            s.apply_add(0.3, 0.8)
            ob.apply_add(0.3, 0.8)
        else:
            s.apply_subtract(0.02)
            ob.apply_subtract(0.02)
        assert np.array_equal(s.download_water(), ob.download_water())
        r, c, e = s.find_outlet()
        assert (r, c) == oracle.find_outlet(D) and e == D[r, c]
        assert s.get_cell_water(r, c) == ob.get_cell_water(r, c)
        s.close()


def test_basin5_validation_sequence_on_gpu(cuda_lib, basin5, tmp_path):
    """BASELINE.json configs[1] and validate_WDPM.sh: Add 10 mm -> Drain -> Subtract 10 mm on basin5,
    through the CUDA solver; output files must equal the unmodified reference's OpenCL-branch files."""
    from test_oracle import _check_goldens
    from wdpm_b200.wdpmcl import ModuleParams, default_backend, run_module
    hdr, dem = basin5
    steps = [
        ("add10", ModuleParams("add", depth_mm=10, runoff_fraction=1.0, elevation_tol_mm=1.0, zero_threshold_mm=0.005), 30000),
        ("drain", ModuleParams("drain", elevation_tol_mm=0.1, drain_tol_m3=1.0, zero_threshold_mm=0.005), 11000),
        ("sub10", ModuleParams("subtract", depth_mm=10, elevation_tol_mm=1.0, zero_threshold_mm=0.005), 1000),
    ]
    water = None
    for name, params, iters in steps:
        rep = run_module(dem, hdr.nodata, hdr.cellsize, params, water=water, backend=default_backend())
        assert rep.iterations == iters
        path = tmp_path / f"{name}.asc"
        ascgrid.write_asc(path, hdr, rep.water)
        text = path.read_text()
        assert text == golden_text(f"ref_opencl_{name}.asc.gz"), name
        _check_goldens(name, text)
        _, water = ascgrid.read_asc(path)


def test_basin5_fused_kernel_equals_colour_kernel(cuda_lib, basin5):
    """Same DEM through both kernels for two blocks of Add: identical grids and reductions."""
    from wdpm_b200 import F64, KERNEL_COLOUR, KERNEL_FUSED, Solver
    hdr, dem = basin5
    res = []
    for kernel, variant in ((KERNEL_COLOUR, 0), (KERNEL_FUSED, 1), (KERNEL_FUSED, 5), (KERNEL_FUSED, 6), (KERNEL_FUSED, 0), (3, 0), (0, 0)):
        s = Solver(hdr.nrows, hdr.ncols, hdr.nodata, 0, dtype=F64, zero_threshold=5e-6, kernel=kernel, fused_variant=variant)
        s.upload(dem, None)
        s.apply_add(0.3, 1.0)
        r = [s.run_block(200) for _ in range(2)]
        res.append((s.download_water(), [(x.max_diff, x.masked_sum, x.wet_cells) for x in r]))
        s.close()
    for w, r in res[1:]:
        assert np.array_equal(w, res[0][0])
        assert r == res[0][1]


@pytest.mark.parametrize("dt,size", [(np.float32, 8192), (np.float64, 4096), (np.float64, 32768)])
def test_large_grid_properties(cuda_lib, dt, size):
    """Sizes the oracle cannot reach in seconds: the fused kernel must equal the colour kernel bit for
    bit (two independent CUDA code paths), conserve mass to rounding, keep margins dry and be a no-op
    on a dry grid. Also checks an oracle-computed crop far from the crop's edges."""
    import torch
    from oracle import pyoracle
    from wdpm_b200 import F32, F64, KERNEL_COLOUR, KERNEL_FUSED, Solver, synth
    if size >= 32768:  # BASELINE.json's full size: 35 GB on the device, ~30 GB of host arrays
        import psutil
        if torch.cuda.get_device_properties(0).total_memory < 100e9 or psutil.virtual_memory().available < 60e9:
            pytest.skip("not enough memory for the 32768^2 case")
    code = F64 if dt == np.float64 else F32
    dem = synth.fractal_dem(size, size, seed=size, device="cuda", dtype=torch.float64)
    dem = (dem - dem.min()).to(torch.float32 if dt == np.float32 else torch.float64).cpu().numpy()
    n_it = 6 if size < 32768 else 3
    outs = []
    for kernel in (KERNEL_COLOUR, KERNEL_FUSED):
        s = Solver(size, size, NODATA, 0, dtype=code, kernel=kernel)
        s.upload(dem, None)
        s.apply_add(0.1, 1.0)
        s.iterate(n_it)
        outs.append(s.download_water())
        if kernel == KERNEL_FUSED:
            s.upload_water(None)          # dry grid: nothing may change
            s.iterate(2)
            assert not s.download_water().any()
        s.close()
    assert np.array_equal(outs[0], outs[1])
    total = outs[1].astype(np.float64).sum()
    assert abs(total - 0.1 * size * size) / (0.1 * size * size) < (1e-12 if dt == np.float64 else 2e-6)
    # oracle on a crop: a sub-pass moves information at most 2 cells, so cells further than
    # 2*9*n_it from the crop edge cannot see the cut. The crop origin must be a multiple of 3 so the
    # colour of every cell (row, col mod 3) is the same in the crop as in the full grid.
    r0, c0, n = 3 * ((size // 2 - 100) // 3), 3 * (size // 9), 320
    D = ascgrid.pad_grid(dem[r0:r0 + n, c0:c0 + n], dt(NODATA))
    W = np.where(D > NODATA, dt(0.1), dt(0)).astype(dt)
    pyoracle.Oracle().iterate(W, D, NODATA, 0, n_it)
    m = 18 * n_it + 4
    assert np.array_equal(W[1 + m:-1 - m, 1 + m:-1 - m], outs[1][r0 + m:r0 + n - m, c0 + m:c0 + n - m])


def test_basin5_add_300mm_to_convergence(cuda_lib, basin5, tmp_path):
    """BASELINE.json configs[0]: basin5 Add 300 mm, rof 1.0, tol 1 mm, threshold 0.005 mm. The reference needs
    320 000 iterations (352 s on 8 host threads through its OpenCL branch, 996 s serial); the output file
    must equal its file byte for byte (tests/golden/ref_opencl_add300.asc.gz, md5 equal to the serial run's)."""
    from wdpm_b200.wdpmcl import ModuleParams, default_backend, run_module
    hdr, dem = basin5
    rep = run_module(dem, hdr.nodata, hdr.cellsize, ModuleParams("add", depth_mm=300, runoff_fraction=1.0, elevation_tol_mm=1.0,
                                                                   zero_threshold_mm=0.005), backend=default_backend())
    assert rep.iterations == 320000
    path = tmp_path / "add300.asc"
    ascgrid.write_asc(path, hdr, rep.water)
    assert path.read_text() == golden_text("ref_opencl_add300.asc.gz")
    assert "%10.2f" % rep.final_vol == "3301066.70" and "%10.4f" % rep.water_frac == "    0.3062"
    print(f"cfg1 on GPU: {rep.solver_ms/1e3:.2f} s of device time, {rep.launches} launches")


def test_auto_picks_the_production_tilings(cuda_lib):
    """AUTO on a grid above the small-grid limit: every module gets the warp-autonomous kernel (fp64: 12 warps on a
    380-column window; fp32: 24 warps on 752 columns)."""
    from wdpm_b200 import F32, F64, Solver, solver
    from wdpm_b200.solver import PRODUCTION_FUSED_VARIANT_F32, PRODUCTION_FUSED_VARIANT_F64
    rows = cols = 2100  # 4.4 M cells
    def tiling(module, dtype, thres):
        with Solver(rows, cols, NODATA, module, dtype=dtype, zero_threshold=thres) as s:
            i = s.info()
        return {k: i[k] for k in ("window_cols", "strip_cols", "cta_threads", "smem_bytes")}, i["warp_autonomous"]
    def variant(v, dtype):
        i = solver.fused_variant_info(v, dtype)
        return {k: i[k] for k in ("window_cols", "strip_cols", "cta_threads", "smem_bytes")}
    for module in (0, 1, 2):
        assert tiling(module, F64, 5e-6) == (variant(PRODUCTION_FUSED_VARIANT_F64, F64), 1)
        assert tiling(module, F32, 5e-6) == (variant(PRODUCTION_FUSED_VARIANT_F32, F32), 1)


def test_tier2_parity_against_the_serial_path(cuda_lib, basin5):
    """BASELINE.json north_star, second parity bullet: against the reference's SERIAL CPU path
    (WDPMCL.c:1074-1125 with runoffs :1934-1964, runoffd :1967-2006 and drain() :1859-1897 - here the oracle's
    SCHED_SERIAL, itself pinned to the unmodified program's serial output files in test_oracle.py) the total water
    volume agrees to 1e-9 relative and every cell's depth within the run's elevation tolerance. basin5, the
    validate_WDPM.sh chain, each module from the reference's own hand-over file, full fp64 precision on both sides;
    Add is cut at 4 000 iterations to keep the CPU side to seconds (it is bit-identical anyway), Drain and Subtract
    run to the reference's stop criterion. (The serial and OpenCL branches differ by design: Subtract
    uses the Add formula there, Drain wipes the outlet's 3x3 after every iteration - SURVEY.md 8a, rows a8/a9.)"""
    from oracle_backend import factory
    from oracle import pyoracle as po
    from wdpm_b200.wdpmcl import ModuleParams, default_backend, run_module
    hdr, dem = basin5

    def water_of(name):
        import tempfile
        with tempfile.NamedTemporaryFile("w", suffix=".asc", delete=False) as f:
            f.write(golden_text(name))
            path = f.name
        return ascgrid.read_asc(path)[1]

    steps = [
        ("add", ModuleParams("add", depth_mm=10, runoff_fraction=1.0, elevation_tol_mm=1.0, zero_threshold_mm=0.005, iteration_limit=4000), None),
        ("drain", ModuleParams("drain", elevation_tol_mm=0.1, drain_tol_m3=1.0, zero_threshold_mm=0.005), "ref_opencl_add10.asc.gz"),
        ("subtract", ModuleParams("subtract", depth_mm=10, elevation_tol_mm=1.0, zero_threshold_mm=0.005, iteration_limit=1000), "ref_opencl_drain.asc.gz"),
    ]
    valid = dem > hdr.nodata
    for name, params, water_file in steps:
        water = None if water_file is None else water_of(water_file)
        gpu = run_module(dem, hdr.nodata, hdr.cellsize, params, water=water, backend=default_backend())
        cpu = run_module(dem, hdr.nodata, hdr.cellsize, params, water=water, backend=factory(schedule=po.SCHED_SERIAL))
        # Drain runs to its stop criterion (11 000 iterations on both paths): the serial branch empties the outlet's
        # 3x3 after every iteration, so the two transients differ and only the drained states are comparable
        assert gpu.iterations == cpu.iterations == (params.iteration_limit or 11000)
        vg, vc = float(np.sum(gpu.water[valid])), float(np.sum(cpu.water[valid]))
        assert abs(vg - vc) <= 1e-9 * abs(vc), (name, vg, vc)
        worst = float(np.max(np.abs(gpu.water[valid] - cpu.water[valid])))
        assert worst <= params.elevation_tol_mm / 1000.0, (name, worst)
        if name == "add":  # serial Add is runoffs == runoffadd bit for bit
            assert np.array_equal(gpu.water, cpu.water)


def test_final_statistics_on_the_device(cuda_lib, oracle):
    """the order-free part of the final report (WDPMCL.c:1394-1459) - valid cells, cells with more than
    1 mm of water, deepest water - from one device pass; single solver and stripes (counts add, maxima combine)."""
    from wdpm_b200 import F32, F64, Solver
    from wdpm_b200.stripes import StripeSolver, connect_in_process, plan_stripes
    rng = np.random.default_rng(91)
    for dt, code in ((np.float64, F64), (np.float32, F32)):
        D, W = random_case(rng, 140, 310, dt, depth=0.004, wet_fraction=0.6)
        W[3, 5] = dt(7.25)  # the deepest cell, on a valid cell
        D[3, 5] = dt(500.0)
        dem, w = D[1:-1, 1:-1], W[1:-1, 1:-1]
        valid = dem > NODATA
        want = (int(valid.sum()), int(((w > dt(0.001)) & valid).sum()), float(w[valid].max()))
        with Solver(140, 310, NODATA, 2, dtype=code) as s:   # Drain: the outlet mark must still count as a valid cell
            s.upload(dem, w)
            s.find_outlet()
            assert s.final_statistics() == want
        plan = plan_stripes(140, 3)
        ss = [StripeSolver(140, 310, NODATA, 0, st, dtype=code, fused_variant=2) for st in plan]
        connect_in_process(ss)
        for s, st in zip(ss, plan):
            s.upload_band(dem[st.band_row0:st.band_row0 + st.band_rows], w[st.band_row0:st.band_row0 + st.band_rows])
        parts = [s.final_statistics() for s in ss]
        for s in ss:
            s.close()
        assert (sum(p[0] for p in parts), sum(p[1] for p in parts), max(p[2] for p in parts)) == want
