#!/usr/bin/env python
"""Time several fused-kernel variants on one synthetic DEM in one process (developer tool).

python scripts/variant_sweep.py --size 8192 --cases f64:0:13,f64:0:17,f32:0:12 [--iters 200] [--blocks 3]
case = dtype:module:variant[:chunk_rows]   (module 0 add, 1 subtract, 2 drain)
Prints ms per iteration of the last block (device time of the iteration kernels only) and the roofline fraction.
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from wdpm_b200 import ADD, F32, F64, Solver, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=8192)
ap.add_argument("--cases", default="f64:0:13,f64:0:17")
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--blocks", type=int, default=3)
ap.add_argument("--add-mm", type=float, default=300.0)
ap.add_argument("--out", default="")
a = ap.parse_args()
try:
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:
    peak = 6650.0

dem64 = synth.fractal_dem(a.size, a.size, seed=a.size, device="cuda", dtype=torch.float64)
dem_f64 = dem64.cpu().numpy()
dem_f32 = (dem64 - dem64.min()).to(torch.float32).cpu().numpy()
del dem64
torch.cuda.empty_cache()
rows = []
for case in a.cases.split(","):
    parts = case.split(":")
    dt, mod, var = parts[0], int(parts[1]), int(parts[2])
    chunk = int(parts[3]) if len(parts) > 3 else 0
    code, dem, es = (F64, dem_f64, 8) if dt == "f64" else (F32, dem_f32, 4)
    try:
        s = Solver(a.size, a.size, -99999.0, mod, dtype=code, zero_threshold=5e-6, kernel=2, fused_variant=var, fused_chunk_rows=chunk)
    except Exception as e:
        print(case, "create failed:", e)
        continue
    if mod == ADD:
        s.upload(dem, None)
        s.apply_add(a.add_mm / 1000, 1.0)
    else:
        s.upload(dem, np.full_like(dem, a.add_mm / 1000))
        if mod == 2:
            s.find_outlet()
            s.set_total_drain(0.0)
    ms = []
    for _ in range(a.blocks):
        r = s.run_block(a.iters)
        ms.append(r.iterate_ms / a.iters)
    info = s.info()
    s.close()
    best = min(ms[1:]) if len(ms) > 1 else ms[0]
    cups = a.size * a.size / (best / 1e3)
    frac = cups * 3 * es / 1e9 / peak
    row = dict(case=case, ms_per_iter=[round(x, 4) for x in ms], best=best, cell_updates_per_s=cups, roofline_frac=frac,
               max_diff=r.max_diff, wet=r.wet_cells, grid=info["grid_ctas"], chunk_rows=info["chunk_rows"], window=info["window_cols"],
               strip=info["strip_cols"], threads=info["cta_threads"], smem=info["smem_bytes"])
    rows.append(row)
    print(json.dumps(row), flush=True)
if a.out:
    Path(a.out).write_text("\n".join(json.dumps(r) for r in rows) + "\n")
