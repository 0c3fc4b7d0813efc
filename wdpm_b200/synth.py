"""Synthetic fractal DEMs for the large benchmark configurations.

The reference ships one DEM (dem/basin5.asc); BASELINE.json's 8192^2 / 32768^2 /
65536^2 cases are this repository's definition: spectral
synthesis - white Gaussian noise, FFT, amplitude proportional to k^-(H+1) with
H = 0.7, inverse FFT, rescaled to mean 500 m and standard deviation 3.34 m
(basin5's), quantised to 0.0001 m like basin5, cell size 10 m, NODATA -99999,
every cell valid. The noise comes from numpy's default_rng(seed) on the CPU and
from torch's Philox generator on a CUDA device, so a DEM is reproducible per
device type, not across them; it is an input, and every parity check feeds the
same generated array to both sides.
"""
from __future__ import annotations

import numpy as np
import torch

MEAN_ELEV = 500.0
SIGMA_ELEV = 3.34
HURST = 0.7
CELLSIZE = 10.0
NODATA = -99999.0
QUANTUM = 1e-4


def fractal_dem(rows: int, cols: int, seed: int, device: str | torch.device = "cpu",
                dtype: torch.dtype = torch.float64) -> torch.Tensor:
    """Return a (rows, cols) DEM tensor on `device` in `dtype` (values in metres)."""
    device = torch.device(device)
    if device.type == "cuda":
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        noise = torch.randn(rows, cols, generator=gen, device=device, dtype=torch.float32)
    else:
        rng = np.random.default_rng(seed)
        noise = torch.from_numpy(rng.standard_normal((rows, cols), dtype=np.float32))
    spec = torch.fft.rfft2(noise)
    del noise
    ky = torch.fft.fftfreq(rows, device=device, dtype=torch.float32)[:, None]
    kx = torch.fft.rfftfreq(cols, device=device, dtype=torch.float32)[None, :]
    k = torch.sqrt(ky * ky + kx * kx)
    k[0, 0] = 1.0
    amp = k.pow_(-(HURST + 1.0))
    amp[0, 0] = 0.0
    spec *= amp
    del amp, k
    field = torch.fft.irfft2(spec, s=(rows, cols))
    del spec
    field = field.to(torch.float64)
    field -= field.mean()
    field *= SIGMA_ELEV / field.std()
    field += MEAN_ELEV
    field = torch.round(field / QUANTUM) * QUANTUM
    return field.to(dtype)
