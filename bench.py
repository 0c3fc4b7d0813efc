#!/usr/bin/env python
"""Benchmark of the WDPM redistribution path (BASELINE.json: cell-updates/s, 32768^2 DEM).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
    python bench.py --impl reference [...]                         the reference's CPU path, bounded sample

One STEP = one convergence block of the reference's solver loop
(/root/reference/src/WDPMCL.c:1054-1268): zero-threshold + snapshot, 1000 iterations of the nine
colour sub-passes, masked max-difference / water-balance reductions - over the whole DEM.
1 cell-update = one interior cell carried through one iteration, so a step is rows*cols*1000
cell-updates. Workload at every N: BASELINE.json configs[3], the synthetic 32768 x 32768 fractal DEM
(wdpm_b200/synth.py), Add 300 mm, runoff fraction 1.0, zero threshold 0.005 mm, fp64 (the
reference's precision); N > 1 partitions the same DEM into row stripes (strong scaling).

`value`  : device-timed (CUDA events on the solver's stream, max over ranks), grids resident in HBM.
`e2e`    : the same block through the public C-ABI call sequence a host makes per block when it
           round-trips like the reference does (WDPMCL.c:1129-1153, :1217-1233): upload DEM + water
           from pinned host memory, run the block, download the water grid - all inside the timed region.
`roofline`: dominant kernel = the fused iteration kernel; algorithmic bytes per launch =
           cells * 3 * sizeof(T) * iterations_per_launch (read dem, read water, write water once per
           iteration) over its mean launch time, measured with CUDA events around the
           launch sequence inside the library.
`cpu_baseline`: the verbatim reference kernels (oracle/_ref/librunoffcl_ref.so, built from
           /root/reference/src/runoff.cl; kind "reference") or, if that binary is absent, the C oracle
           (kind "port"), all host threads, on a window of the SAME DEM in the SAME state the timed GPU
           steps start from.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BLOCK_ITERS = 1000
ADD_MM = 300.0
THRES_MM = 0.005
NODATA = -99999.0


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks and throttle reasons with NVML while the timed region runs."""

    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None
            return self
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.2)

    def stop(self) -> dict:
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# --------------------------------------------------------------------------- reference arm


def cpu_kernel():
    """(callable iterate(w, d, n_iters), kind, threads) - the reference kernels if built, else the port."""
    from oracle import pyoracle as po
    po_o = po.Oracle()
    if po.RefCL.available():
        ref = po.RefCL()
        # keep the column-major layout across calls (the reference flattens once per block)
        import ctypes as C

        class Runner:
            kind = "reference"
            threads = po_o.threads

            def __init__(self, w, d):
                self.wf = np.ascontiguousarray(w.T)
                self.df = np.ascontiguousarray(d.T)
                self.R, self.Cc = w.shape[0] - 2, w.shape[1] - 2
                self.td = np.zeros(1)

            def iterate(self, n):
                ref.lib.refcl_iterate_f64(C.c_int(0), self.wf.ctypes.data_as(C.c_void_p), self.df.ctypes.data_as(C.c_void_p),
                                          C.c_double(NODATA), C.c_int(self.R), C.c_int(self.Cc), C.c_int(n),
                                          self.td.ctypes.data_as(C.c_void_p), C.c_int(0), C.c_int(0))
        return Runner

    class Runner:  # noqa: F811
        kind = "port"
        threads = po_o.threads

        def __init__(self, w, d):
            self.w, self.d = w, d

        def iterate(self, n):
            po_o.iterate(self.w, self.d, NODATA, po.ADD, n)
    return Runner


def time_cpu_sample(D: np.ndarray, W: np.ndarray, budget_s: float = 12.0):
    """Cell-updates/s of the CPU path on padded window (D, W); iterations chosen to fill ~budget_s."""
    Runner = cpu_kernel()
    r = Runner(W.copy(), D)
    cells = (D.shape[0] - 2) * (D.shape[1] - 2)
    t = time.perf_counter()
    r.iterate(4)
    per_it = (time.perf_counter() - t) / 4
    n = int(max(8, min(2000, budget_s / max(per_it, 1e-6))))
    t = time.perf_counter()
    r.iterate(n)
    dt = time.perf_counter() - t
    return cells * n / dt, Runner.kind, Runner.threads, n, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from wdpm_b200 import ascgrid, synth
    size = min(args.sample_size, args.size)
    dem = synth.fractal_dem(size, size, seed=args.size, device="cpu").numpy()
    D = ascgrid.pad_grid(dem, NODATA)
    W = np.where(D > NODATA, ADD_MM / 1000.0, 0.0)
    Runner = cpu_kernel()
    r = Runner(W, D)
    iters = args.ref_iters_per_step
    cells = size * size
    for _ in range(args.warmup):
        r.iterate(iters)
    t = time.perf_counter()
    for _ in range(args.steps):
        r.iterate(iters)
    dt = time.perf_counter() - t
    value = cells * iters * args.steps / dt
    sample = f"{size}x{size} window of the synthetic DEM (seed {args.size}), Add {ADD_MM:g} mm from the dry start, {iters} iterations per step"
    line = {
        "impl": "reference", "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": Runner.threads, "kind": Runner.kind, "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- our arm


def workload_config(args, world):
    return {"workload": f"synthetic {args.size}x{args.size} fractal DEM (H=0.7, sigma 3.34 m), Add {ADD_MM:g} mm, rof 1.0, "
                        f"zero-threshold {THRES_MM} mm, one step = one {BLOCK_ITERS}-iteration convergence block",
            "rows": args.size, "cols": args.size, "block_iters": args.block_iters,
            "partition": "single GPU" if world == 1 else f"{world} row stripes, halo exchange over NVLink",
            "l2_policy": "inputs larger than L2 (3 grids x rows x cols x 8 B per iteration)"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from wdpm_b200 import ADD, F32, F64, KERNEL_AUTO, Solver, ascgrid, synth
    from wdpm_b200.stripes import DistributedSolver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dtype_code, np_dt, torch_dt, esize = (F64, np.float64, torch.float64, 8) if args.dtype == "f64" else (F32, np.float32, torch.float32, 4)
    size = args.size
    cells = size * size
    dem_dev = synth.fractal_dem(size, size, seed=size, device=f"cuda:{local_rank}", dtype=torch.float64)
    if args.dtype == "f32":
        dem_dev = dem_dev - dem_dev.min()  # fp32 mode: base elevation removed before rounding

    common = dict(dtype=dtype_code, zero_threshold=THRES_MM / 1000, kernel=KERNEL_AUTO, fused_variant=args.variant)
    if world == 1:
        s = Solver(size, size, NODATA, ADD, device=local_rank, iters_per_launch=args.iters_per_launch, **common)
        r0, nrows, o0, orows = 0, size, 0, size
    else:
        ds = DistributedSolver(size, size, NODATA, ADD, device=local_rank, **common)
        s = ds.solver
        st = ds.stripe
        r0, nrows, o0, orows = st.band_row0, st.band_rows, st.owned_row0, st.owned_rows
    # this rank's rows (owned + halos) in pinned host memory: what a host application would hand over
    dem_host = torch.empty((nrows, size), dtype=torch_dt, pin_memory=True)
    dem_host.copy_(dem_dev[r0:r0 + nrows].to(torch_dt))
    del dem_dev
    torch.cuda.empty_cache()
    water_host = torch.zeros((nrows, size), dtype=torch_dt, pin_memory=True)
    owned_host = water_host[o0 - r0:o0 - r0 + orows]  # contiguous view: the owned rows inside the band

    def upload(with_water: bool):
        if world == 1:
            s.upload_ptr(dem_host.data_ptr(), water_host.data_ptr() if with_water else None)
        else:
            ds.upload_band_ptr(dem_host.data_ptr(), water_host.data_ptr() if with_water else None)

    def run_block():
        return s.run_block(args.block_iters) if world == 1 else ds.run_block(args.block_iters)

    stream = torch.cuda.Stream()
    s.set_stream(stream.cuda_stream)
    upload(False)
    s.apply_add(ADD_MM / 1000.0, 1.0)

    for _ in range(args.warmup):
        run_block()

    # state the timed steps start from, for the CPU baseline sample (N=1, rank 0)
    do_cpu = args.cpu_baseline and rank == 0 and world == 1
    if do_cpu:
        s.download_water_ptr(water_host.data_ptr())
        n = min(args.sample_size, size)
        c0 = (size - n) // 2
        Dw = ascgrid.pad_grid(dem_host.numpy()[c0:c0 + n, c0:c0 + n].astype(np.float64), NODATA)
        Ww = ascgrid.pad_grid(water_host.numpy()[c0:c0 + n, c0:c0 + n].astype(np.float64), 0.0)
        wet = float(np.count_nonzero(Ww > 0)) / (n * n)

    # ---- timed region: K steps, device time on the solver's stream, max over ranks
    sampler = ClockSampler(physical_gpu_index(local_rank)).start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iter_ms, klaunch, last = [], 0, None
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        r = run_block()
        iter_ms.append(r.iterate_ms)
        klaunch += r.launches
        last = r
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    total_ms = max_over_ranks(ev0.elapsed_time(ev1))
    value = cells * args.block_iters * args.steps / (total_ms / 1e3)

    # ---- end to end: host buffers in, host buffers out, every step (max over ranks, wall clock
    # bracketed by barriers since the host<->device copies are synchronous calls)
    if world == 1:
        s.download_water_ptr(water_host.data_ptr())
    else:
        # refresh the whole band (halos too) so the re-upload resumes the same state
        s.download_water_ptr(owned_host.data_ptr())
        parts = [None] * world
        dist.all_gather_object(parts, (o0, owned_host.numpy()[[0, 1, 2, 3, 4, 5, -3, -2, -1]].copy()))
        if rank > 0:
            water_host[0:o0 - r0] = torch.from_numpy(parts[rank - 1][1][-(o0 - r0):])
        if rank + 1 < world:
            nb = nrows - (o0 - r0) - orows
            water_host[o0 - r0 + orows:] = torch.from_numpy(parts[rank + 1][1][:nb])
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        upload(True)
        r2 = run_block()
        s.download_water_ptr(owned_host.data_ptr() if world > 1 else water_host.data_ptr())
        klaunch += r2.launches
        if world > 1 and e2e_steps > 1:
            parts = [None] * world
            dist.all_gather_object(parts, owned_host.numpy()[[0, 1, 2, 3, 4, 5, -3, -2, -1]].copy())
            if rank > 0:
                water_host[0:o0 - r0] = torch.from_numpy(parts[rank - 1][-(o0 - r0):])
            if rank + 1 < world:
                nb = nrows - (o0 - r0) - orows
                water_host[o0 - r0 + orows:] = torch.from_numpy(parts[rank + 1][:nb])
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = cells * args.block_iters * e2e_steps / t_e2e

    info2 = s.info()
    K = info2["iters_per_launch"]
    iter_launches = args.block_iters // K
    my_cells = orows * size
    peak, peak_src = measured_peak_gbs()
    launch_ms = float(np.mean(iter_ms)) / iter_launches
    achieved = my_cells * 3 * esize * K / (launch_ms / 1e3) / 1e9
    kernel_name = "k_fused" if info2["kernel"] == 2 else "k_colour"
    traffic = None
    try:  # DRAM bytes per launch from the committed ncu capture of this very configuration, if there is one
        tj = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        traffic = tj[f"{kernel_name}:{args.dtype}:{size}x{size}:{world}gpu"]["traffic_bytes"]
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name, "launch_ms": launch_ms,
                "algorithmic_bytes_per_launch": my_cells * 3 * esize * K, "peak_source": peak_src, "per": "GPU (slowest rank's launch time)"}

    line = {
        "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": 2 * nrows * size * esize * world if world > 1 else 2 * cells * esize,
                "d2h_bytes_per_step": cells * esize + 64 * world, "steps": e2e_steps},
        "gpu_launches": int(klaunch), "clocks": clocks, "roofline": roofline,
        "state": {"max_diff": last.max_diff, "wet_fraction": last.wet_cells / cells,
                  "iterations_done": (args.warmup + args.steps + e2e_steps) * args.block_iters},
        "tiling": {k: info2[k] for k in ("kernel", "strip_cols", "window_cols", "chunk_rows", "grid_ctas", "cta_threads", "smem_bytes", "iters_per_launch", "sm_count")},
    }
    if do_cpu:
        v, kind, threads, n_it, dt = time_cpu_sample(Dw, Ww, args.cpu_budget)
        line["cpu_baseline"] = {"value": v, "unit": "cell-updates/s", "cores": threads, "kind": kind,
                                "sample": f"{Dw.shape[0]-2}x{Dw.shape[1]-2} centre window of the same DEM in the state after warm-up "
                                          f"(wet fraction {wet:.3f}), {n_it} iterations in {dt:.1f} s"}
    if world == 1:
        s.close()
    else:
        ds.close()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=32768)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--block-iters", type=int, default=BLOCK_ITERS)
    ap.add_argument("--iters-per-launch", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--sample-size", type=int, default=2048)
    ap.add_argument("--ref-iters-per-step", type=int, default=100)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
