#!/usr/bin/env python
"""Per-warp timeline of one CTA of the fused kernel (developer probe).

Build (here, no GPU needed):   python scripts/timeline.py --build
Run (GPU box):                 python scripts/timeline.py --size 8192 [--dtype f32] [--module 2] [--variant V] --cta 70 --out gpurun_out/tl.npy

The instrumented library is the product source compiled with -DWDPM_TIMELINE into
wdpm_b200/libwdpm_b200_tl.so; each warp's lane 0 stamps clock64 at fixed points of eight steps:
0 step start, 1 loads landed (compute) / copies issued (data-movement warp), 2 window in registers issued,
3/5/7 end of colour sub-step 1/2/3, 4/6 past the row-group barrier, 8 past the step barrier.
"""
import argparse
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
TL_LIB = ROOT / "wdpm_b200" / "libwdpm_b200_tl.so"

ap = argparse.ArgumentParser()
ap.add_argument("--build", action="store_true")
ap.add_argument("--size", type=int, default=8192)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
ap.add_argument("--module", type=int, default=0, help="0 add, 1 subtract, 2 drain (uniform 300 mm water layer)")
ap.add_argument("--cta", type=int, default=300)
ap.add_argument("--step0", type=int, default=100)
ap.add_argument("--out", default="gpurun_out/timeline.npy")
a = ap.parse_args()

if a.build:
    sys.path.insert(0, str(ROOT))
    from wdpm_b200 import build
    cmd = [build._nvcc(), *build.NVCC_FLAGS, "-DWDPM_TIMELINE", "-ccbin", "/usr/bin/g++", "-o", str(TL_LIB), *map(str, build.SOURCES)]
    subprocess.run(cmd, check=True)
    print(TL_LIB)
    sys.exit(0)

os.environ["WDPM_B200_LIB"] = str(TL_LIB)
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from wdpm_b200 import ADD, F32, F64, Solver, synth  # noqa: E402
from wdpm_b200.solver import load_library  # noqa: E402

lib = load_library()
dem = synth.fractal_dem(a.size, a.size, seed=a.size, device="cuda", dtype=torch.float64)
if a.dtype == "f32":
    dem = dem - dem.min()
dem = dem.to(torch.float64 if a.dtype == "f64" else torch.float32).cpu().numpy()
s = Solver(a.size, a.size, -99999.0, a.module, dtype=F64 if a.dtype == "f64" else F32, zero_threshold=5e-6, kernel=2, fused_variant=a.variant)
if a.module == ADD:
    s.upload(dem, None)
    s.apply_add(0.3, 1.0)
else:
    s.upload(dem, np.full_like(dem, 0.3))
    if a.module == 2:
        s.find_outlet()
        s.set_total_drain(0.0)
s.run_block(20)
assert lib.wdpm_debug_timeline(a.cta, a.step0, None, 0) == 0
r = s.run_block(10)
n = 8 * 32 * 10
buf = np.zeros(n, dtype=np.int64)
assert lib.wdpm_debug_timeline(0, 0, buf.ctypes.data_as(C.POINTER(C.c_longlong)), n) == 0
tl = buf.reshape(8, 32, 10)
Path(a.out).parent.mkdir(exist_ok=True)
np.save(a.out, tl)
print("ms/iteration", r.iterate_ms / 10, s.info())
t0 = tl[:, :, 0][tl[:, :, 0] > 0].min()
for step in range(1, 4):
    print("step", step)
    for w in range(32):
        if not tl[step, w, 0]:
            continue
        row = tl[step, w]
        print(f"  warp {w:2d}: " + " ".join(f"{(x - t0) if x > 0 else -1:7d}" for x in row[:9]))
s.close()
