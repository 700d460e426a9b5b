"""Plug-and-play ADMM loops (reference S3:255-296, S6:262-308, S6:491-525), batched and GPU-resident.

Per iteration the reference crosses the CPU<->GPU boundary six times (x-update in NumPy on the host,
denoiser on the device).  Here the whole iteration stays on the device: the x-update is the CUDA
kernel behind ``pnpadmm_xupdate_f32`` (the cluster kernel for 256x256), the denoiser is a PyTorch
module (bf16 tensor cores), and the combine / dual / clamp steps are the pointwise C-ABI kernels.
The float32 round trips and the [0,1] clamps of x, z AND the dual w are kept.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from .solver import AdmmSolver, ArrayLike, cnc_combine, dual_update_


def _setup(images, mask, noises, reo, device):
    img = torch.as_tensor(images)
    single = img.ndim == 2
    if single:
        img = img[None]
    B, N = int(img.shape[0]), int(img.shape[1])
    m = torch.as_tensor(mask)
    solver = AdmmSolver(B, N, dtype='float32', mask_batched=(m.ndim == 3), device=device)
    y = solver.acquire(img, m, noises)
    z = solver.zero_filled(y)                     # x = |ifft2(y)|; z = x; w = 0   (S3:246-249)
    w = torch.zeros_like(z)
    solver.prepare(y, m, reo)
    return solver, z, w, single


def _finish(x, single, as_numpy):
    x = x[0] if single else x
    return x.cpu().numpy() if as_numpy else x


@torch.no_grad()
def pnp_admm_l1(images: ArrayLike, mask: ArrayLike, noises: ArrayLike, denoiser: Callable, *, iter_num: int = 50,
                reo: float = 0.26, device=None, callback: Optional[Callable] = None):
    """PnP-ADMM-L1 (S3:255-296): z = D(x + w); w = w + x - z; clamp(x, z, w)."""
    as_numpy = not isinstance(images, torch.Tensor)
    solver, z, w, single = _setup(images, mask, noises, reo, device)
    x = z
    for i in range(iter_num):
        x, xpw = solver.xupdate(z, w, want_xpw=True)                # S3:259-264 (+ x + w, S3:290)
        z = denoiser(xpw[:, None], i)[:, 0].contiguous()            # S3:290
        dual_update_(x, z, w, clamp01=True)                         # S3:293-296
        if callback is not None:
            callback(i, x, z, w)
    return _finish(x, single, as_numpy)


@torch.no_grad()
def pnp_admm_cnc(images: ArrayLike, mask: ArrayLike, noises: ArrayLike, denoiser1: Callable,
                 denoiser2: Optional[Callable] = None, *, alpha: float = 1.2, iter_num: int = 50, lambda1: float = 4.0,
                 reo: float = 0.45, b: float = 0.3, device=None, callback: Optional[Callable] = None):
    """PnP-ADMM-CNC (S6:262-308; with two denoisers S6:491-525):
    s = D1(z); t = (1-a) z + a (x+w) + a reo lambda b (z - s); z = D2(t); w = w + x - z; clamps."""
    if denoiser2 is None:
        denoiser2 = denoiser1
    as_numpy = not isinstance(images, torch.Tensor)
    solver, z, w, single = _setup(images, mask, noises, reo, device)
    coef = alpha * reo * lambda1 * b
    x = z
    for i in range(iter_num):
        x = solver.xupdate(z, w)                                    # S6:266-271
        s = denoiser1(z[:, None], i)[:, 0].contiguous()             # S6:300
        t = cnc_combine(z, x, w, s, alpha, coef)                    # S6:301
        z = denoiser2(t[:, None], i)[:, 0].contiguous()             # S6:302
        dual_update_(x, z, w, clamp01=True)                         # S6:305-308
        if callback is not None:
            callback(i, x, z, w)
    return _finish(x, single, as_numpy)
