#!/usr/bin/env python
"""basin5-sized grids: iterations/s of the colour kernel vs fused variants."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wdpm_b200 import ADD, F64, Solver, ascgrid  # noqa: E402

hdr, dem = ascgrid.read_asc(Path(__file__).resolve().parent.parent / "tests" / "golden" / "basin5.asc.gz")
for kernel, variant, chunk in ((1, 0, 0), (2, 5, 9), (3, 0, 0), (0, 0, 0)):
    s = Solver(hdr.nrows, hdr.ncols, hdr.nodata, ADD, dtype=F64, zero_threshold=5e-6, kernel=kernel, fused_variant=variant,
               fused_chunk_rows=chunk)
    s.upload(dem, None)
    s.apply_add(0.3, 1.0)
    s.run_block(1000)
    t = time.perf_counter()
    r = s.run_block(5000)
    dt = time.perf_counter() - t
    print(f"kernel {kernel} variant {variant} chunk {chunk}: {r.iterate_ms/5000*1000:.2f} us/iteration (wall {dt/5000*1e6:.2f}), ctas {s.info()['grid_ctas']}")
    s.close()
