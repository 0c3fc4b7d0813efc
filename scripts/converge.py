#!/usr/bin/env python
"""Run Add on a synthetic DEM block by block until the reference's stop test fires (or --max-blocks),
logging per block: iterations, max_diff, wet fraction, ms per iteration, cumulative device seconds.

python scripts/converge.py --size 8192 --dtype f64 --add-mm 300 --max-blocks 200 [--out profiles/x.csv]
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wdpm_b200 import ADD, F32, F64, Solver, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=8192)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--add-mm", type=float, default=300.0)
ap.add_argument("--tol-mm", type=float, default=1.0)
ap.add_argument("--thres-mm", type=float, default=0.005)
ap.add_argument("--max-blocks", type=int, default=200)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--out", default="")
a = ap.parse_args()
code, tdt = (F64, torch.float64) if a.dtype == "f64" else (F32, torch.float32)
dem = synth.fractal_dem(a.size, a.size, seed=a.size, device="cuda", dtype=torch.float64)
if a.dtype == "f32":
    dem = dem - dem.min()
dem = dem.to(tdt).cpu().numpy()
s = Solver(a.size, a.size, -99999.0, ADD, dtype=code, zero_threshold=a.thres_mm / 1000, fused_variant=a.variant)
s.upload(dem, None)
s.apply_add(a.add_mm / 1000, 1.0)
cells = a.size * a.size
rows = ["block,iterations,max_diff_m,wet_fraction,ms_per_iteration,cum_device_s"]
cum = 0.0
for b in range(1, a.max_blocks + 1):
    r = s.run_block(1000)
    cum += r.block_ms / 1e3
    rows.append(f"{b},{b*1000},{r.max_diff:.6g},{r.wet_cells/cells:.5f},{r.iterate_ms/1000:.5f},{cum:.3f}")
    if b <= 5 or b % 10 == 0:
        print(rows[-1], flush=True)
    if r.max_diff <= a.tol_mm / 1000:
        print("converged:", rows[-1])
        break
if a.out:
    Path(a.out).write_text("\n".join(rows) + "\n")
print(rows[-1])
s.close()
