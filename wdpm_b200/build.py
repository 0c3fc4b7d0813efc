"""Build the CUDA library in-tree: wdpm_b200/libwdpm_b200.so (sm_100a only).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU
box with the working tree. `python -m wdpm_b200.build` or build_library().
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
LIB = PKG / "libwdpm_b200.so"
HOST_BIN = PKG / "host" / "wdpmcl_b200"
SOURCES = [PKG / "csrc" / "solver.cu"]
HEADERS = [PKG / "csrc" / "kernels.cuh", PKG / "csrc" / "relax.cuh", PKG / "csrc" / "mw_schedule.h",
           ROOT / "include" / "wdpm_b200.h", ROOT / "include" / "wdpm_quantize.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",          # no FMA contraction: results must match the reference bit for bit
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    if force or _stale(LIB, SOURCES + HEADERS):
        cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", "-o", str(LIB), *map(str, SOURCES)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
    return LIB


HOOKS_LIB = ROOT / "tests" / "_build" / "libwdpm_b200_hooks.so"


def build_hooks_library(force: bool = False) -> Path:
    """TEST build of the same sources with -DWDPM_TEST_HOOKS (tests/test_gpu_halo_race.py): lets a test delay
    chosen CTAs and switch the halo protocol back to its pre-fix CTA count, to show that the race it guards
    against is real. Lives under tests/_build; the product never loads it."""
    if force or _stale(HOOKS_LIB, SOURCES + HEADERS):
        HOOKS_LIB.parent.mkdir(parents=True, exist_ok=True)
        cmd = [_nvcc(), *NVCC_FLAGS, "-DWDPM_TEST_HOOKS", "-ccbin", "/usr/bin/g++", "-o", str(HOOKS_LIB), *map(str, SOURCES)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc (test hooks) failed:\n" + res.stdout + res.stderr)
    return HOOKS_LIB


def build_host(force: bool = False) -> Path | None:
    """The drop-in command-line host (C) linked against the library."""
    src = PKG / "host" / "wdpm_host.c"
    if not src.exists():
        return None
    if force or _stale(HOST_BIN, [src, ROOT / "include" / "wdpm_b200.h", ROOT / "include" / "wdpm_quantize.h", LIB]):
        cmd = ["/usr/bin/gcc", "-O2", "-std=c11", "-fopenmp", "-Wall", "-Wextra", "-I", str(ROOT / "include"), str(src), "-o", str(HOST_BIN),
               "-L", str(PKG), "-lwdpm_b200", "-Wl,-rpath,$ORIGIN/..", "-lm", "-lpthread"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("host build failed:\n" + res.stdout + res.stderr)
    return HOST_BIN


if __name__ == "__main__":
    print(build_library(force=True, verbose=False))
    print(build_host(force=True))
