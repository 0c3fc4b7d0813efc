"""TEST INFRASTRUCTURE: an object with wdpm_b200.Solver's interface backed by the CPU oracle,
so the host logic in wdpm_b200/wdpmcl.py can be exercised (and pinned against the reference's
output files) on a machine without a GPU. Never imported by the product."""
from __future__ import annotations

import numpy as np

from oracle import pyoracle as po
from wdpm_b200 import ascgrid
from wdpm_b200.solver import BlockResult


class OracleBackend:
    def __init__(self, rows, cols, nodata, module, zero_threshold, dtype=np.float64, schedule=po.SCHED_OPENCL):
        self.o = po.Oracle()
        self.rows, self.cols, self.nodata, self.module = rows, cols, nodata, module
        self.np_dtype = dtype
        self.thres = zero_threshold
        self.schedule = schedule
        self.outlet = (0, 0)
        self.td = 0.0

    def upload(self, dem, water=None):
        self.D = ascgrid.pad_grid(np.asarray(dem, dtype=self.np_dtype), self.np_dtype(self.nodata))
        w = np.zeros((self.rows, self.cols), self.np_dtype) if water is None else np.asarray(water, dtype=self.np_dtype)
        self.W = ascgrid.pad_grid(w, self.np_dtype(0))

    def apply_add(self, depth, rof):
        w, valid = self.W, self.D > self.nodata
        m = valid & (w > 0)
        w[m] += self.np_dtype(depth)
        m2 = valid & (w <= 0)
        w[m2] = self.np_dtype(depth * rof)

    def apply_subtract(self, depth):
        w, valid = self.W, self.D > self.nodata
        v = w - self.np_dtype(depth)
        w[valid] = np.where(v > 0, v, 0)[valid]

    def find_outlet(self):
        r, c = self.o.find_outlet(self.D)
        self.outlet = (r, c)
        return r, c, float(self.D[r, c])

    def get_cell_water(self, r, c):
        return float(self.W[r, c])

    def set_total_drain(self, v):
        self.td = float(self.np_dtype(v))

    def get_total_drain(self):
        return self.td

    def run_block(self, n):
        md, ms, td = self.o.block(self.W, self.D, self.nodata, self.module, self.thres, n, schedule=self.schedule,
                                  outlet=self.outlet, totaldrain=self.td)
        self.td = td
        wet = int(np.count_nonzero((self.W > 0) & (self.D > self.nodata)))
        return BlockResult(md, ms, td, wet, n, 0, 0.0, 0.0)

    def download_water(self):
        return self.W[1:-1, 1:-1].copy()


def factory(dtype=np.float64, schedule=po.SCHED_OPENCL):
    def make(rows, cols, nodata, module, zero_threshold):
        return OracleBackend(rows, cols, nodata, module, zero_threshold, dtype=dtype, schedule=schedule)
    return make
