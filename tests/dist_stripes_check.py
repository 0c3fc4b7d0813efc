"""Launched by tests/test_gpu_stripes.py under torch.distributed.run (one rank per GPU).

Every rank builds the same DEM, runs its stripe through DistributedSolver (CUDA IPC + NVLink halo
pushes + device-side arrival flags), rank 0 also runs the whole DEM on its GPU alone; the assembled
stripes and the combined block results must match the single-GPU run bit for bit."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))

from conftest import random_case  # noqa: E402
from wdpm_b200 import F32, F64, Solver  # noqa: E402
from wdpm_b200.stripes import DistributedSolver  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    rows, cols = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (700, 900)
    failures = []
    # chunk_rows 0 = the solver's own chunking (one wave of CTAs at this size); 3 = one row triple per CTA, i.e.
    # several waves: a stripe's upward halo is then exported, and its flag raised, long before its kernel ends
    cases = [(dt, code, module, 0) for dt, code in ((np.float64, F64), (np.float32, F32)) for module in (0, 1, 2)]
    cases += [(np.float64, F64, 0, 3), (np.float64, F64, 2, 3)]
    for dt, code, module, chunk_rows in cases:
        if True:
            rng = np.random.default_rng(99)
            D, W = random_case(rng, rows, cols, dt, depth=0.05)
            dem, w0 = D[1:-1, 1:-1], W[1:-1, 1:-1]
            ds = DistributedSolver(rows, cols, -99999.0, module, device=local, dtype=code, zero_threshold=1e-3, fused_chunk_rows=chunk_rows)
            st = ds.stripe
            ds.upload_band(dem[st.band_row0:st.band_row0 + st.band_rows], w0[st.band_row0:st.band_row0 + st.band_rows])
            if module == 0:
                ds.solver.apply_add(0.01, 1.0)
            elif module == 1:
                ds.solver.apply_subtract(0.01)
            else:
                # outlet: global minimum, found per stripe and combined on the host
                try:
                    cand = ds.solver.find_outlet()
                except Exception:
                    cand = None
                allc = [None] * world
                dist.all_gather_object(allc, cand)
                best = min((c for c in allc if c is not None), key=lambda c: (c[2], c[0], c[1]))
                ds.solver.set_outlet(best[0], best[1])
                ds.solver.set_total_drain(0.0)
            res = [ds.run_block(25) for _ in range(2)]
            owned = ds.solver.download_owned()
            parts = [None] * world
            dist.all_gather_object(parts, owned)
            if rank == 0:
                full = np.concatenate(parts, axis=0)
                s = Solver(rows, cols, -99999.0, module, dtype=code, zero_threshold=1e-3, device=local, kernel=2)
                s.upload(dem, w0)
                if module == 0:
                    s.apply_add(0.01, 1.0)
                elif module == 1:
                    s.apply_subtract(0.01)
                else:
                    s.find_outlet()
                    s.set_total_drain(0.0)
                ref = [s.run_block(25) for _ in range(2)]
                ok = np.array_equal(full, s.download_water())
                for a, b in zip(res, ref):
                    ok = ok and a.max_diff == b.max_diff and a.wet_cells == b.wet_cells and abs(a.masked_sum - b.masked_sum) <= 1e-12 * abs(b.masked_sum)
                    if module == 2:
                        ok = ok and dt(a.total_drain) == dt(b.total_drain)
                if not ok:
                    failures.append((dt.__name__, module, chunk_rows))
                s.close()
            ds.close()
    if rank == 0:
        print("STRIPES_OK" if not failures else f"STRIPES_FAIL {failures}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
