"""Synthetic inputs of the shapes BASELINE.json names (3DEP/WTK are unreachable offline).

Deterministic, seeded, float32-representable (SURVEY.md §8d "Synthetic DEM"): base 1500 m, analytic
N-S and oblique ridges (200-400 m relief, 5-15 km wavelength) plus spectral fBm roughness.  Arrays are
`[row=north, col=east]` like the reference's flipped rasters (`ssrs/raster.py:49`).
"""
from __future__ import annotations

import numpy as np


def synthetic_dem(rows: int, cols: int, resolution: float, seed: int = 20211018,
                  rough_rms: float = 25.0) -> np.ndarray:
    """float32 elevation [rows, cols] in metres."""
    y = (np.arange(rows, dtype=np.float64) * resolution)[:, None]
    x = (np.arange(cols, dtype=np.float64) * resolution)[None, :]
    km = 1000.0
    z = 1500.0 + 0.0 * (x + y)
    # N-S ridges (crest lines along y), wavelength 9 km, modulated along-crest
    z = z + 160.0 * np.sin(2 * np.pi * x / (9.0 * km)) * (1.0 + 0.35 * np.sin(2 * np.pi * y / (23.0 * km)))
    # oblique ridges
    z = z + 110.0 * np.sin(2 * np.pi * (0.8 * x + 0.6 * y) / (13.0 * km) + 0.7)
    # a few gaussian massifs
    rng = np.random.RandomState(seed)
    lx, ly = cols * resolution, rows * resolution
    for _ in range(6):
        cx, cy = rng.uniform(0, lx), rng.uniform(0, ly)
        sx, sy = rng.uniform(2.5, 6.0) * km, rng.uniform(2.5, 6.0) * km
        z = z + rng.uniform(120.0, 300.0) * np.exp(-0.5 * (((x - cx) / sx) ** 2 + ((y - cy) / sy) ** 2))
    # fBm roughness by spectral synthesis, H = 0.7  (amplitude ~ k^-(H+1))
    if rough_rms > 0:
        ky = np.fft.fftfreq(rows, d=resolution)[:, None]
        kx = np.fft.rfftfreq(cols, d=resolution)[None, :]
        kk = np.sqrt(kx * kx + ky * ky)
        kk[0, 0] = np.inf
        kcut = 1.0 / (4.0 * resolution)   # keep the smallest scales smooth: slopes stay O(0.1)
        amp = kk ** (-1.7) * np.exp(-(kk / kcut) ** 2)
        # band-limit to wavelengths below ~20 km so the spectrum is grid-size independent
        amp = amp * (kk > 1.0 / (20.0 * km))
        phase = rng.uniform(0, 2 * np.pi, size=amp.shape)
        f = np.fft.irfft2(amp * np.exp(1j * phase), s=(rows, cols))
        f *= rough_rms / max(f.std(), 1e-30)
        z = z + f
    return np.ascontiguousarray(z.astype(np.float32))


def synthetic_wind_lattice(rows: int, cols: int, resolution: float, spacing_m: float = 2000.0,
                           seed: int = 7, speed0: float = 8.0, dirn0: float = 270.0):
    """Jittered ~2 km lattice of (x, y, speed, direction) points covering the padded region
    (BASELINE config 4; stands in for WTK points, `ssrs/simulator.py:765-792`)."""
    rng = np.random.RandomState(seed)
    lx, ly = (cols - 1) * resolution, (rows - 1) * resolution
    pad = 1.5 * spacing_m
    gx = np.arange(-pad, lx + pad + 1, spacing_m)
    gy = np.arange(-pad, ly + pad + 1, spacing_m)
    xx, yy = np.meshgrid(gx, gy)
    xx = xx + rng.uniform(-0.25, 0.25, xx.shape) * spacing_m
    yy = yy + rng.uniform(-0.25, 0.25, yy.shape) * spacing_m
    speed = speed0 + 3.0 * np.sin(2 * np.pi * xx / 40e3) * np.cos(2 * np.pi * yy / 35e3)
    dirn = dirn0 + 30.0 * np.cos(2 * np.pi * yy / 50e3)
    return xx.ravel(), yy.ravel(), speed.ravel(), np.mod(dirn.ravel(), 360.0)


def seasonal_wind_conditions(count: int, seed: int = 11):
    """`count` seeded (speed, direction) pairs: speed ~ Weibull(k=2, lambda=8), direction ~ von Mises(270, kappa=2)."""
    rng = np.random.RandomState(seed)
    speed = 8.0 * rng.weibull(2.0, size=count)
    dirn = np.mod(270.0 + np.degrees(rng.vonmises(0.0, 2.0, size=count)), 360.0)
    return np.maximum(speed, 1.0), dirn
