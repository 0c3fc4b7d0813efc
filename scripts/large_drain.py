#!/usr/bin/env python
"""BASELINE.json configs[4] at the reference's semantics: Drain on a synthetic DEM of --size^2 cells in fp64,
row stripes over all ranks (one process per GPU), starting from a uniform --water-mm layer.

The reference has exactly one outlet (lowest cell with dem > 0, src/WDPMCL.c:1005-1017): outlet 0 here.
--outlets K > 1 adds the "many drain outlets" of configs[4] through wdpm_set_outlets (an extension,
include/wdpm_b200.h): the K-1 lowest cells of the DEM's rim (first/last row and column), ties by
row-major position. Checks the size-independent property a Drain run offers: water left + water
drained = water put in, to rounding, and that the outlet found by the stripes is the global minimum.

python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/large_drain.py --size 65536 --blocks 2
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wdpm_b200 import DRAIN, F64, synth  # noqa: E402
from wdpm_b200.stripes import DistributedSolver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=65536)
ap.add_argument("--blocks", type=int, default=2)
ap.add_argument("--water-mm", type=float, default=300.0)
ap.add_argument("--thres-mm", type=float, default=0.0, help="zero-depth threshold; > 0 destroys mass (SURVEY appendix A quirk 6), so the balance check needs 0")
ap.add_argument("--outlets", type=int, default=64, help="size of the outlet set (1 = the reference's single outlet)")
ap.add_argument("--out", default="")
a = ap.parse_args()

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = a.size
t0 = time.time()
dem = synth.fractal_dem(n, n, seed=n, device=f"cuda:{local}", dtype=torch.float64)
gmin = float(dem.min().item())
gpos = int(torch.argmin(dem).item())  # first occurrence in row-major order = the reference's tie rule
rim = []
if a.outlets > 1:  # every rank holds the whole DEM here, so every rank picks the same rim cells
    idx = torch.arange(n, device=dem.device)
    zeros, last = torch.zeros_like(idx), torch.full_like(idx, n - 1)
    rr = torch.cat([zeros, last, idx[1:-1], idx[1:-1]])
    cc = torch.cat([idx, idx, zeros[1:-1], last[1:-1]])
    ev = dem[rr, cc]
    key = rr * n + cc
    order = torch.argsort(key)          # row-major first, then a stable sort by elevation: ties -> first in row-major order
    ev, key = ev[order], key[order]
    pick = torch.argsort(ev, stable=True)[: a.outlets + 1]
    rim = [(int(k) // n + 1, int(k) % n + 1) for k in key[pick].tolist()]  # padded coordinates
from wdpm_b200.stripes import plan_stripes  # noqa: E402
st = plan_stripes(n, world)[rank]
band = dem[st.band_row0:st.band_row0 + st.band_rows].cpu().numpy()
del dem
torch.cuda.empty_cache()  # hand the FFT workspace back before the solver allocates its grids
ds = DistributedSolver(n, n, -99999.0, DRAIN, device=local, dtype=F64, zero_threshold=a.thres_mm / 1000)
assert ds.stripe == st
water = np.full_like(band, a.water_mm / 1000.0)
ds.upload_band(band, water)
t_setup = time.time() - t0

# outlet: each stripe finds its own minimum, the host combines (elevation, then row-major position)
try:
    cand = ds.solver.find_outlet()
except Exception:
    cand = None
allc = [None] * world
dist.all_gather_object(allc, cand)
best = min((c for c in allc if c is not None), key=lambda c: (c[2], c[0], c[1]))
assert best[2] == gmin and (best[0] - 1) * n + (best[1] - 1) == gpos, (best, gmin, gpos)
# rim cells ON the stripe borders of an 8-way partition (left column at a stripe's first row, right column at the row
# above it): their neighbouring centres belong to two stripes at 8 GPUs and to one at 2 - the case that used to make a
# total depend on the partition
forced = []
if a.outlets >= 32:
    for p8 in plan_stripes(n, 8)[1:]:
        forced += [(p8.row0, 1), (p8.row0 - 1, n)]
outlets = [(best[0], best[1])]
for rc in forced + rim:
    if rc not in outlets and len(outlets) < a.outlets:
        outlets.append(rc)
ds.solver.set_outlets(outlets)
owner = st.row0 <= best[0] < st.row0 + st.rows
w_out = ds.solver.get_cell_water(best[0], best[1]) if owner else 0.0
ds.solver.set_total_drain(max(w_out, 0.0) if owner else 0.0)  # src/WDPMCL.c:1029

put_in = a.water_mm / 1000.0 * n * n
lines = []
for b in range(a.blocks):
    r = ds.run_block(1000)
    # totaldrain counts the outlet's initial water twice: remove it once for the balance
    drained = r.total_drain - max(a.water_mm / 1000.0, 0.0)
    bal = (r.masked_sum + drained - put_in) / put_in
    lines.append({"block": b + 1, "ms_per_iteration": r.iterate_ms / 1000, "cell_updates_per_s": n * n * 1000 / (r.block_ms / 1e3),
                  "max_diff": r.max_diff, "water_left_m": r.masked_sum, "total_drain_m": r.total_drain, "balance_rel_err": bal,
                  "wet_fraction": r.wet_cells / (n * n)})
    if rank == 0:
        print(json.dumps(lines[-1]), flush=True)
    assert abs(bal) < (1e-9 if a.thres_mm == 0 else 1e-5), bal
# what must not depend on the number of GPUs: the water grid (order-free checksum) and every outlet's total (bit patterns)
per_outlet = ds.outlet_drains(len(outlets))
checksum = ds.solver.water_checksum()
parts = [None] * world
dist.all_gather_object(parts, checksum)
checksum = sum(parts) % (1 << 64)
if rank == 0:
    info = ds.solver.info()
    summary = {"size": n, "gpus": world, "dtype": "f64", "module": "drain", "outlet": [best[0], best[1]], "min_elevation": best[2],
               "n_outlets": len(outlets), "outlets_head": outlets[:8],
               "setup_s": t_setup, "device_bytes_per_gpu": info["device_bytes"], "blocks": lines,
               "water_checksum": f"{checksum:016x}", "outlet_totals_bits": [f"{int(np.float64(v).view(np.uint64)):016x}" for v in per_outlet],
               "outlet_totals_m": [float(v) for v in per_outlet],
               "total_drain_outlet_order_m": float(np.add.reduce(np.asarray(per_outlet, dtype=np.float64))) if len(per_outlet) < 8 else float(__import__("functools").reduce(lambda x, y: x + y, [float(v) for v in per_outlet])),
               "outlets_on_stripe_borders": [list(o) for o in outlets if any(abs(o[0] - p.row0) <= 1 for p in plan_stripes(n, world)[1:])]}
    print(json.dumps(summary))
    if a.out:
        Path(a.out).write_text(json.dumps(summary, indent=1))
ds.close()
dist.destroy_process_group()
