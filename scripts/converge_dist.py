#!/usr/bin/env python
"""Time-to-converge of the Add module on a synthetic DEM (BASELINE.json: "cell-updates/s + time-to-converge,
32k^2 DEM"): blocks of 1000 iterations until the reference's stop test fires (max |w - w_block_start| <= the
elevation tolerance, /root/reference/src/WDPMCL.c:1283-1376), or a time / block budget runs out. One process
per GPU (row stripes, halo exchange inside the iteration kernel); also runs on a single GPU without torchrun.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/converge_dist.py \
        --size 32768 --budget-s 540 --csv profiles/converge_32768_f64_8gpu.csv --json profiles/converge_32768_f64_8gpu.json

Per block the CSV holds: block, iterations, max_diff (m), wet fraction, ms per iteration (device time of the
iteration kernels, slowest rank), cumulative device seconds, cumulative wall seconds.

Checkpoints (a binary side format): --checkpoint-dir DIR writes, every --checkpoint-every blocks
and at the end, each rank's owned rows as a raw little-endian array (water_rank<r>.bin) plus meta.json; --resume
continues from them (same --size / dtype / rank count). The reference's own checkpoint is the %f scratch file
(WDPMCL.c:1290-1299), which quantises to 1e-6 m; the binary form resumes bit for bit. (On this project's GPU pool
nothing but 64 MiB of gpurun_out/ survives a call, so a 32768^2 run cannot span calls; the feature is exercised at
small sizes by tests/test_gpu_converge.py.)
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from wdpm_b200 import ADD, F32, F64, Solver, synth  # noqa: E402
from wdpm_b200.stripes import HALO_ABOVE, HALO_BELOW, DistributedSolver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=32768)
ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
ap.add_argument("--add-mm", type=float, default=300.0)
ap.add_argument("--tol-mm", type=float, default=1.0)
ap.add_argument("--thres-mm", type=float, default=0.005)
ap.add_argument("--max-blocks", type=int, default=10 ** 9)
ap.add_argument("--budget-s", type=float, default=0.0, help="stop after this many wall seconds of iterating (0 = none)")
ap.add_argument("--block-iters", type=int, default=1000)
ap.add_argument("--csv", default="")
ap.add_argument("--json", default="")
ap.add_argument("--checkpoint-dir", default="")
ap.add_argument("--checkpoint-every", type=int, default=0, help="blocks between checkpoints (0 = only at the end)")
ap.add_argument("--resume", action="store_true")
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

code, tdt, npdt = (F64, torch.float64, np.float64) if a.dtype == "f64" else (F32, torch.float32, np.float32)
n = a.size
cells = n * n
t_start = time.time()
dem = synth.fractal_dem(n, n, seed=n, device=f"cuda:{local}", dtype=torch.float64)
if a.dtype == "f32":
    dem = dem - dem.min()
common = dict(dtype=code, zero_threshold=a.thres_mm / 1000)
if world == 1:
    s = Solver(n, n, -99999.0, ADD, device=local, **common)
    ds = None
    band_row0, band_rows, owned_row0, owned_rows = 0, n, 0, n
else:
    ds = DistributedSolver(n, n, -99999.0, ADD, device=local, **common)
    s = ds.solver
    st = ds.stripe
    band_row0, band_rows, owned_row0, owned_rows = st.band_row0, st.band_rows, st.owned_row0, st.owned_rows
band = dem[band_row0:band_row0 + band_rows].to(tdt).cpu().numpy()
del dem
torch.cuda.empty_cache()


def upload(water_band):
    if world == 1:
        s.upload(band, water_band)
    else:
        ds.upload_band(band, water_band)


ckdir = Path(a.checkpoint_dir) if a.checkpoint_dir else None
meta_key = dict(size=n, dtype=a.dtype, world=world, add_mm=a.add_mm, thres_mm=a.thres_mm)
blocks_done, iters_done, cum_dev, cum_wall, history = 0, 0, 0.0, 0.0, []
if a.resume:
    meta = json.loads((ckdir / "meta.json").read_text())
    assert {k: meta[k] for k in meta_key} == meta_key, "checkpoint belongs to another run"
    owned = np.fromfile(ckdir / f"water_rank{rank}.bin", dtype=npdt).reshape(owned_rows, n)
    water_band = np.zeros((band_rows, n), dtype=npdt)
    o = owned_row0 - band_row0
    water_band[o:o + owned_rows] = owned
    if world > 1:  # halo rows come from the neighbours' owned rows
        parts = [None] * world
        dist.all_gather_object(parts, (owned[:HALO_BELOW].copy(), owned[-HALO_ABOVE:].copy()))
        if rank > 0:
            water_band[:o] = parts[rank - 1][1][-o:]
        if rank + 1 < world:
            nb = band_rows - o - owned_rows
            water_band[o + owned_rows:] = parts[rank + 1][0][:nb]
    upload(water_band)
    blocks_done, iters_done, cum_dev, cum_wall, history = meta["blocks"], meta["iterations"], meta["cum_device_s"], meta["cum_wall_s"], meta["history"]
    del water_band, owned
else:
    upload(None)
    s.apply_add(a.add_mm / 1000, 1.0)
t_setup = time.time() - t_start


def checkpoint():
    if ckdir is None:
        return
    ckdir.mkdir(parents=True, exist_ok=True)
    out = np.empty((owned_rows, n), dtype=npdt)
    if world == 1:
        s.download_water(out)
    else:
        s.download_owned(out)
    out.tofile(ckdir / f"water_rank{rank}.bin")
    if world > 1:
        dist.barrier()
    if rank == 0:
        (ckdir / "meta.json").write_text(json.dumps(dict(meta_key, blocks=blocks_done, iterations=iters_done, cum_device_s=cum_dev,
                                                         cum_wall_s=cum_wall, history=history)))


converged = False
t_loop = time.time()
while blocks_done < a.max_blocks:
    r = s.run_block(a.block_iters) if world == 1 else ds.run_block(a.block_iters)
    blocks_done += 1
    iters_done += a.block_iters
    cum_dev += r.block_ms / 1e3
    wall = cum_wall + (time.time() - t_loop)
    history.append([blocks_done, iters_done, r.max_diff, r.wet_cells / cells, r.iterate_ms / a.block_iters, round(cum_dev, 3), round(wall, 3)])
    if rank == 0 and (blocks_done <= 5 or blocks_done % 20 == 0):
        print(",".join(f"{x:.6g}" if isinstance(x, float) else str(x) for x in history[-1]), flush=True)
    if r.max_diff <= a.tol_mm / 1000:
        converged = True
        break
    stop = a.budget_s > 0 and time.time() - t_loop > a.budget_s
    if world > 1:  # every rank must take the same decision
        flag = torch.tensor([1 if stop else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        stop = bool(flag.item())
    if stop:
        break
    if ckdir is not None and a.checkpoint_every > 0 and blocks_done % a.checkpoint_every == 0:
        cum_wall_save = cum_wall
        cum_wall = wall
        checkpoint()
        cum_wall = cum_wall_save
cum_wall += time.time() - t_loop
checkpoint()
checksum = s.water_checksum()
if world > 1:
    parts = [None] * world
    dist.all_gather_object(parts, checksum)
    checksum = sum(parts) % (1 << 64)

if rank == 0:
    if a.csv:
        Path(a.csv).parent.mkdir(parents=True, exist_ok=True)
        Path(a.csv).write_text("block,iterations,max_diff_m,wet_fraction,ms_per_iteration,cum_device_s,cum_wall_s\n" +
                               "\n".join(",".join(f"{x:.9g}" if isinstance(x, float) else str(x) for x in h) for h in history) + "\n")
    # where the run is heading: fit log(max_diff) against log(iterations) over the last third of the blocks
    est = None
    if not converged and len(history) >= 30:
        tail = history[-max(10, len(history) // 3):]
        x = np.log([h[1] for h in tail])
        y = np.log([max(h[2], 1e-300) for h in tail])
        slope, icpt = np.polyfit(x, y, 1)
        if slope < 0:
            it_needed = float(np.exp((np.log(a.tol_mm / 1000) - icpt) / slope))
            ms = float(np.mean([h[4] for h in tail]))
            est = {"fit": "max_diff ~ iterations^slope over the last third of the blocks", "slope": float(slope),
                   "iterations_to_tolerance": it_needed, "seconds_to_tolerance_at_current_rate": it_needed * ms / 1e3}
    info = s.info()
    summary = {"metric": "time_to_converge", "workload": f"synthetic {n}x{n} fractal DEM, Add {a.add_mm:g} mm, tolerance {a.tol_mm:g} mm, zero threshold {a.thres_mm:g} mm, {a.dtype}",
               "n_gpus": world, "converged": converged, "blocks": blocks_done, "iterations": iters_done, "device_seconds": cum_dev,
               "wall_seconds": cum_wall, "setup_seconds": t_setup, "last_max_diff_m": history[-1][2], "last_wet_fraction": history[-1][3],
               "cell_updates_per_s": cells * iters_done / cum_dev, "mean_ms_per_iteration": cum_dev / iters_done * 1e3,
               "checksum": f"{checksum:016x}", "extrapolation": est, "warp_autonomous": info.get("warp_autonomous"), "resumed": a.resume}
    print(json.dumps(summary))
    if a.json:
        Path(a.json).write_text(json.dumps(summary, indent=1) + "\n")
if world == 1:
    s.close()
else:
    ds.close()
    dist.destroy_process_group()
