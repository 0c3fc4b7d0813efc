"""Launched by tests/test_gpu_halo_race.py (one process, one GPU).

N in-process stripe solvers on the SAME device run FREE (wdpm_block_begin / enqueue / end on their own
streams - the host never waits between iterations), so what orders their iterations is only the device-side
halo protocol of the iteration kernel: peer stores + arrival flags + k_halo_wait. The assembled grid must
equal the single-solver grid bit for bit.

With WDPM_B200_LIB pointing at the -DWDPM_TEST_HOOKS build, WDPM_TEST_HALO_READER_DELAY_NS makes the CTAs
that only READ a stripe's bottom-halo rows late, and WDPM_TEST_HALO_OLD_COUNT=1 restores the round-1 flag rule
(count only the CTAs that own exported rows): the stripe below then overwrites halo rows that are still to be
read - the write-after-read race of VERDICT r1. (It cannot change a grid: a cell depends on at most two rows
below it within one iteration - a centre pushes to its upper neighbours first, runoff.cl:28-30 - and the late
readers own rows at least four above the halo; but it is a race, and the counters make it visible.)

HOOK_COUNTERS a b: a = reader CTAs that were delayed, b = how many of them woke up AFTER the stripe below had
already finished the next iteration (whose export overwrites the rows they were about to read).

usage: stripes_freerun_check.py N_STRIPES MODULE CHUNK_ROWS ITERS [DTYPE]   -> prints FREERUN_EQUAL or FREERUN_DIFFER <n>
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))

from conftest import random_case  # noqa: E402
from wdpm_b200 import F32, F64, Solver  # noqa: E402
from wdpm_b200.stripes import StripeSolver, connect_in_process, plan_stripes  # noqa: E402

NODATA = -99999.0


def main():
    n_stripes, module, chunk_rows, iters = (int(x) for x in sys.argv[1:5])
    dt = np.float32 if len(sys.argv) > 5 and sys.argv[5] == "f32" else np.float64
    code = F64 if dt == np.float64 else F32
    rng = np.random.default_rng(4242)
    rows, cols = 250, 330
    D, W = random_case(rng, rows, cols, dt, nodata_fraction=0.03, wet_fraction=0.9)
    dem, w0 = D[1:-1, 1:-1], W[1:-1, 1:-1]

    with Solver(rows, cols, NODATA, module, dtype=code, zero_threshold=1e-4, kernel=2, fused_variant=2) as s:
        s.upload(dem, w0)
        if module == 2:
            outlet = s.find_outlet()[:2]
            s.set_total_drain(0.0)
        ref_res = s.run_block(iters)
        ref = s.download_water()

    plan = plan_stripes(rows, n_stripes)
    ss = [StripeSolver(rows, cols, NODATA, module, st, dtype=code, zero_threshold=1e-4, fused_variant=2, fused_chunk_rows=chunk_rows)
          for st in plan]
    connect_in_process(ss)
    for s, st in zip(ss, plan):
        s.upload_band(dem[st.band_row0:st.band_row0 + st.band_rows], w0[st.band_row0:st.band_row0 + st.band_rows])
        if module == 2:
            s.set_outlet(*outlet)
            s.set_total_drain(0.0)
    for s in ss:
        s.block_begin()
    done = 0
    while done < iters:  # small slices, round-robin: nothing here blocks
        n = min(5, iters - done)
        for s in ss:
            s.block_enqueue(n)
        done += n
    res = [s.block_end() for s in ss]
    full = np.concatenate([s.download_owned() for s in ss], axis=0)
    lib = ss[0]._lib
    if hasattr(lib, "wdpm_debug_counters"):  # hooks build: how many reader CTAs were actually delayed
        import ctypes as C
        cnt = (C.c_int32 * 4)()
        lib.wdpm_debug_counters(cnt)
        print("HOOK_COUNTERS", list(cnt))
    for s in ss:
        s.close()
    ndiff = int((full != ref).sum())
    md = max(r.max_diff for r in res)
    if ndiff == 0 and md == ref_res.max_diff and sum(r.wet_cells for r in res) == ref_res.wet_cells:
        print("FREERUN_EQUAL")
    else:
        print(f"FREERUN_DIFFER {ndiff}")


if __name__ == "__main__":
    main()
