"""Synthetic inputs for sizes the reference ships no data for: phantoms, sampling
masks (unshifted, DC at [0,0], DC always sampled, ~30 % density) and complex Gaussian noise with the
image-domain SNR of ``CS_MRI/noises.mat * 3`` (sigma = 15 * N / 256 per component)."""
from __future__ import annotations

import numpy as np


def phantom(N: int, seed: int = 0, n_ellipses: int = 10) -> np.ndarray:
    """Random-ellipse phantom in [0,1], quantised to uint8/255 like S1:85-90. float32 (N,N)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:N, 0:N].astype(np.float64)
    yy = (yy - N / 2) / (N / 2)
    xx = (xx - N / 2) / (N / 2)
    img = np.zeros((N, N))
    # outer body
    img[(xx / 0.85) ** 2 + (yy / 0.92) ** 2 <= 1] = 0.35
    for _ in range(n_ellipses):
        cx, cy = rng.uniform(-0.5, 0.5, 2)
        a, b = rng.uniform(0.05, 0.35, 2)
        th = rng.uniform(0, np.pi)
        val = rng.uniform(0.1, 1.0)
        xr = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        yr = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        img[(xr / a) ** 2 + (yr / b) ** 2 <= 1] += val * 0.5
    img = np.clip(img, 0, 1)
    return np.float32(np.uint8((img * 255.0).round()) / 255.)


def phantoms(B: int, N: int, seed0: int = 0) -> np.ndarray:
    return np.stack([phantom(N, seed0 + i) for i in range(B)])


def make_mask(kind: str, N: int, seed: int = 0, density: float = 0.30) -> np.ndarray:
    """uint8 (N,N) sampling mask, DC at [0,0]. kind in {'random','radial','cartesian'}."""
    rng = np.random.default_rng(seed)
    f = np.fft.fftfreq(N) * N                       # unshifted frequency index
    fy, fx = np.meshgrid(f, f, indexing='ij')
    r = np.sqrt(fx ** 2 + fy ** 2) / (N / 2)
    if kind == 'random':
        # variable-density Bernoulli + fully sampled low-frequency disc
        core = r <= 0.08
        pdf = (1 - np.minimum(r, 1.0)) ** 3 + 0.02
        lo, hi = 0.0, 1e3                          # bisection: expected density == target
        for _ in range(60):
            sc = 0.5 * (lo + hi)
            if np.where(core, 1.0, np.minimum(pdf * sc, 1.0)).mean() < density:
                lo = sc
            else:
                hi = sc
        m = (rng.uniform(size=(N, N)) < np.minimum(pdf * hi, 1.0)) | core
    elif kind == 'cartesian':
        # full columns (constant along rows, like the shipped Q_Cartesian30); dense low frequencies
        ncol = int(round(density * N))
        core = np.abs(f) <= max(2, int(0.08 * N) // 2)
        cols = core.copy()
        rest = np.flatnonzero(~core)
        p = (1 - np.abs(f[rest]) / (N / 2)) ** 2 + 1e-3
        extra = max(ncol - int(core.sum()), 0)
        pick = rng.choice(rest, size=min(extra, rest.size), replace=False, p=p / p.sum())
        cols[pick] = True
        m = np.broadcast_to(cols[None, :], (N, N)).copy()
    elif kind == 'radial':
        t = np.arange(-N // 2, N // 2)
        phase = rng.uniform(0, 1)
        nlines = max(2, int(density * N))
        while True:                                 # add lines until the target density is reached
            m = np.zeros((N, N), dtype=bool)
            for ang in (np.arange(nlines) + phase) * np.pi / nlines:
                xs = np.round(t * np.cos(ang)).astype(int) % N
                ys = np.round(t * np.sin(ang)).astype(int) % N
                m[ys, xs] = True
            if m.mean() >= density - 0.004 or nlines > 4 * N:
                break
            nlines += 1
    else:
        raise ValueError(kind)
    m[0, 0] = True
    return m.astype(np.uint8)


def make_noise(N: int, seed: int = 1234, B: int | None = None, sigma: float | None = None) -> np.ndarray:
    """complex128 noise, sigma = 15*N/256 per component (already 'x3')."""
    rng = np.random.default_rng(seed)
    s = 15.0 * N / 256.0 if sigma is None else sigma
    shape = (N, N) if B is None else (B, N, N)
    return rng.normal(0, s, shape) + 1j * rng.normal(0, s, shape)
