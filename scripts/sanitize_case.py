#!/usr/bin/env python
"""Smallest end-to-end case for compute-sanitizer: every kernel of the library once, both fused forms
of chunking, all three modules, checked against the oracle."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from conftest import random_case  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from wdpm_b200 import F64, Solver, ascgrid  # noqa: E402

o = po.Oracle()
rng = np.random.default_rng(3)
for module in (0, 1, 2):
    for kernel, variant, chunk in ((1, 0, 0), (2, 2, 15), (2, 5, 0), (2, 10, 0)):
        D, W = random_case(rng, 70, 420, np.float64)
        outlet = o.find_outlet(D)
        a = W.copy()
        md, ms, td = o.block(a, D, -99999.0, module, 1e-3, 6, outlet=outlet, totaldrain=0.0)
        s = Solver(70, 420, -99999.0, module, dtype=F64, zero_threshold=1e-3, kernel=kernel, fused_variant=variant, fused_chunk_rows=chunk)
        s.upload(D[1:-1, 1:-1], W[1:-1, 1:-1])
        if module == 2:
            s.find_outlet()
            s.set_total_drain(0.0)
        r = s.run_block(6)
        ok = np.array_equal(ascgrid.pad_grid(s.download_water(), 0.0), a) and r.max_diff == md
        print("module", module, "kernel", kernel, "variant", variant, "ok" if ok else "MISMATCH")
        s.close()
