"""GPU (-m gpu): row-stripe solvers.

* one GPU: several in-process stripes driven in lockstep (compute phase for all, then halo push for
  all) must reproduce the single-solver grid bit for bit - checks stripe geometry, band upload,
  halo rows and the push kernel;
* two or more GPUs: the real thing (one process per GPU, CUDA IPC, NVLink peer stores, device-side
  arrival flags) via tests/dist_stripes_check.py under torch.distributed.run; skipped on a 1-GPU box.
"""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import random_case
from wdpm_b200 import ascgrid

pytestmark = pytest.mark.gpu
NODATA = -99999.0


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("module", [0, 1, 2])
@pytest.mark.parametrize("n_stripes,variant", [(2, 2), (3, 2), (2, 18)])
def test_in_process_stripes_match_single_solver(cuda_lib, oracle, dt, module, n_stripes, variant):
    from wdpm_b200 import F32, F64, KERNEL_FUSED, Solver
    from wdpm_b200.stripes import StripeSolver, connect_in_process, plan_stripes
    code = F64 if dt == np.float64 else F32
    rng = np.random.default_rng(77)
    rows, cols = 130, 420
    D, W = random_case(rng, rows, cols, dt)
    dem, w0 = D[1:-1, 1:-1], W[1:-1, 1:-1]
    outlet = oracle.find_outlet(D)
    n_it = 6
    ref = W.copy()
    td_ref = oracle.iterate(ref, D, NODATA, module, n_it, outlet=outlet, totaldrain=0.5)

    plan = plan_stripes(rows, n_stripes)
    ss = [StripeSolver(rows, cols, NODATA, module, st, dtype=code, fused_variant=variant, fused_chunk_rows=21) for st in plan]
    connect_in_process(ss)
    for s, st in zip(ss, plan):
        s.upload_band(dem[st.band_row0:st.band_row0 + st.band_rows], w0[st.band_row0:st.band_row0 + st.band_rows])
        if module == 2:
            s.set_outlet(*outlet)
            s.set_total_drain(0.5 if st.row0 <= outlet[0] < st.row0 + st.rows else 0.0)
    for _ in range(n_it):
        for s in ss:
            s.phase(0)
        for s in ss:
            s.synchronize()
        for s in ss:
            s.phase(1)
        for s in ss:
            s.synchronize()
    full = np.concatenate([s.download_owned() for s in ss], axis=0)
    td = sum(s.get_total_drain() for s in ss)
    for s in ss:
        s.close()
    assert np.array_equal(full, ref[1:-1, 1:-1]), int((full != ref[1:-1, 1:-1]).sum())
    if module == 2:
        assert dt(td) == dt(td_ref)


@pytest.mark.parametrize("n_stripes,variant", [(2, 2), (3, 2), (2, 18), (3, 18)])
def test_in_process_stripes_with_an_outlet_set(cuda_lib, oracle, n_stripes, variant):
    """Outlet sets across stripes: outlets on and next to stripe borders; the water grid and every outlet's total
    are bit-exact (an outlet's total lives on the stripe that owns its row)."""
    from wdpm_b200 import F64
    from wdpm_b200.stripes import StripeSolver, connect_in_process, plan_stripes
    rng = np.random.default_rng(79)
    rows, cols = 130, 300
    D, W = random_case(rng, rows, cols, np.float64, nodata_fraction=0.02, wet_fraction=0.95)
    dem, w0 = D[1:-1, 1:-1], W[1:-1, 1:-1]
    plan = plan_stripes(rows, n_stripes)
    border = plan[1].row0
    want = [(border, 10), (border - 1, 11), (border + 1, 40), (border, 41), (1, 1), (rows, cols), (60, 150), (61, 151)]
    outlets = [rc for rc in want if D[rc] > NODATA]
    n_it = 6
    ref = W.copy()
    td_ref = oracle.iterate_outlets(ref, D, NODATA, n_it, outlets)
    ss = [StripeSolver(rows, cols, NODATA, 2, st, dtype=F64, fused_variant=variant, fused_chunk_rows=21) for st in plan]
    connect_in_process(ss)
    for s, st in zip(ss, plan):
        s.upload_band(dem[st.band_row0:st.band_row0 + st.band_rows], w0[st.band_row0:st.band_row0 + st.band_rows])
        s.set_outlets(outlets)
        s.set_total_drain(0.0)
    for _ in range(n_it):
        for s in ss:
            s.phase(0)
        for s in ss:
            s.synchronize()
        for s in ss:
            s.phase(1)
        for s in ss:
            s.synchronize()
    full = np.concatenate([s.download_owned() for s in ss], axis=0)
    td = sum(s.get_outlet_drains(len(outlets)) for s in ss)
    for s in ss:
        s.close()
    assert np.array_equal(full, ref[1:-1, 1:-1]), int((full != ref[1:-1, 1:-1]).sum())
    # an outlet on a stripe border is drained by centres of two stripes; both record their contacts with the stripe that
    # owns the outlet, in sub-pass order, so every total is the single-solver one bit for bit (the other stripes hold 0)
    assert np.array_equal(td, td_ref), (td, td_ref)


def test_multi_gpu_stripes_over_nvlink(cuda_lib):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run under gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    script = Path(__file__).resolve().parent / "dist_stripes_check.py"
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, timeout=900)
    assert "STRIPES_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-3000:]
