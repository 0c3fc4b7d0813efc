#!/usr/bin/env python
"""Run a few iterations of the solver on a synthetic DEM - the command line ncu profiles.

python scripts/profile_iterate.py --size 8192 --dtype f64 --iters 12 [--variant V] [--warm-blocks B]
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402,F811
from wdpm_b200 import ADD, F32, F64, Solver, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=8192)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--kernel", type=int, default=2)
ap.add_argument("--warm-blocks", type=int, default=0, help="1000-iteration blocks to run first (ages the water state)")
ap.add_argument("--add-mm", type=float, default=300.0)
ap.add_argument("--module", type=int, default=0, help="0 add, 1 subtract, 2 drain (water = uniform add-mm layer)")
a = ap.parse_args()
code, tdt = (F64, torch.float64) if a.dtype == "f64" else (F32, torch.float32)
dem = synth.fractal_dem(a.size, a.size, seed=a.size, device="cuda", dtype=torch.float64)
if a.dtype == "f32":
    dem = dem - dem.min()
dem = dem.to(tdt).cpu().numpy()
s = Solver(a.size, a.size, -99999.0, a.module, dtype=code, zero_threshold=5e-6, kernel=a.kernel, fused_variant=a.variant)
if a.module == ADD:
    s.upload(dem, None)
    s.apply_add(a.add_mm / 1000, 1.0)
else:
    s.upload(dem, np.full_like(dem, a.add_mm / 1000))
    if a.module == 2:
        s.find_outlet()
        s.set_total_drain(0.0)
for _ in range(a.warm_blocks):
    r = s.run_block(1000)
    print("warm block", r)
r = s.run_block(a.iters)
print(r, s.info())
print("ms per iteration", r.iterate_ms / a.iters, "cell-updates/s", a.size * a.size * a.iters / (r.iterate_ms / 1e3))
s.close()
