"""CPU oracle: NumPy fp64 restatement of the reference ADMM reconstruction loop.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` leg may import it, and there only as the checker / CPU baseline.
The product path (``pnp_admm_cnc_mri_b200``) never routes through this file.

Parity status: PINNED.  ``oracle/run_reference.py`` executes the unmodified
reference scripts (via runpy) in the build container and
``oracle/make_golden.py`` asserts this restatement is bit-identical
(``np.array_equal``) to their ``out[0]`` for ADMM-L1 and ADMM-CNC; the 90-row
PSNR table mined from the reference's ``results/*.log`` is reproduced to the
printed decimals (``tests/test_oracle_golden.py``).

Reference citations are ``file:line`` under the upstream repo, with the script
aliases of SURVEY.md (S1 = ADMM_L1, S3 = PNP_ADMM_L1_D, S4 = ADMM_CNC,
S6 = PNP_ADMM_CNC_D).

Algorithm contract::

    y  = fft2(img) * mask + noises          # noise on ALL bins        S1:99
    x  = |ifft2(y)| ; z = x ; w = 0                                    S1:100-105
    loop: X = fft2(z - w)                                              S1:115-116
          X[idx] = (La2*X[idx] + y[idx]) / (1 + La2), La2 = 1/(2 reo)  S1:117-118
          x = |Re(ifft2(X))|                                           S1:119-120
          L1 : z = soft(x + w, reo*lambda1)                            S1:123
          CNC: s = soft(z, 1/b)                                        S4:127
               t = (1-alpha) z + alpha (x+w) + alpha reo lambda1 b (z-s)  S4:128
               z = soft(t, alpha*reo*lambda1)                          S4:129
          w = w + x - z                                                S1:126
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np


# --------------------------------------------------------------------------
# a1  soft-threshold                                             S1:18-19
# --------------------------------------------------------------------------
def soft(x, c):
    """``fmax(|x| - c, 0) * sign(x)`` with ``sign(0) == 0`` (S1:18-19, S4:18-19)."""
    return np.fmax(np.fabs(x) - c, 0) * np.sign(x)


# --------------------------------------------------------------------------
# a2  acquisition / zero-filled initialisation                   S1:97-105
# --------------------------------------------------------------------------
def acquire(img, mask, noises):
    """``y = fft2(img) * mask + noises`` (S1:99).  Noise lands on every bin."""
    return np.fft.fft2(img) * mask + noises


def zero_filled(y):
    """``x0 = |ifft2(y)|`` — complex magnitude (S1:100,104)."""
    return np.absolute(np.fft.ifft2(y))


# --------------------------------------------------------------------------
# a3  closed-form x-update                                       S1:115-120
# --------------------------------------------------------------------------
def x_update(z, w, y, index, reo):
    xtilde = np.copy(z - w)                                   # S1:115
    xf = np.fft.fft2(xtilde)                                  # S1:116
    La2 = 1.0 / 2.0 / reo                                     # S1:117
    xf[index] = (La2 * xf[index] + y[index]) / (1.0 + La2)    # S1:118
    x = np.real(np.fft.ifft2(xf))                             # S1:119
    return np.absolute(x)                                     # S1:120


# --------------------------------------------------------------------------
# S1:97-126  ADMM-L1 for one image
# --------------------------------------------------------------------------
def admm_l1(img, mask, noises, iter_num=50, lambda1=0.1, reo=0.015,
            return_state=False):
    img = np.asarray(img).squeeze()
    y = acquire(img, mask, noises)
    index = np.nonzero(mask)
    x = zero_filled(y)
    z = np.copy(x)
    w = np.zeros(img.shape, dtype=np.float64)
    for _ in range(iter_num):
        x = x_update(z, w, y, index, reo)
        z = soft(x + w, reo * lambda1)                        # S1:123
        w = w + x - z                                         # S1:126
    return (x, z, w, y) if return_state else x


# --------------------------------------------------------------------------
# S4:101-132  ADMM-CNC for one image
# --------------------------------------------------------------------------
def admm_cnc(img, mask, noises, alpha=0.45, iter_num=50, lambda1=0.5, reo=0.05, b=64,
             return_state=False):
    img = np.asarray(img).squeeze()
    y = acquire(img, mask, noises)
    index = np.nonzero(mask)
    x = zero_filled(y)
    z = np.copy(x)
    w = np.zeros(img.shape, dtype=np.float64)
    for _ in range(iter_num):
        x = x_update(z, w, y, index, reo)
        s = soft(z, 1 / b)                                                          # S4:127
        t = (1 - alpha) * z + alpha * (x + w) + alpha * reo * lambda1 * b * (z - s)  # S4:128
        z = soft(t, alpha * reo * lambda1)                                          # S4:129
        w = w + x - z                                                               # S4:132
    return (x, z, w, y) if return_state else x


def admm_from_y(y, mask, prox, iter_num, lambda1, reo, alpha=0.0, b=1.0, z0=None, w0=None):
    """Same loop starting from a given measurement ``y`` (used to test the
    solve-from-y C-ABI entry point).  ``prox`` is 'l1' or 'cnc'."""
    index = np.nonzero(mask)
    x = zero_filled(y)
    z = np.copy(x) if z0 is None else np.array(z0, dtype=np.float64)
    w = np.zeros(x.shape, dtype=np.float64) if w0 is None else np.array(w0, dtype=np.float64)
    for _ in range(iter_num):
        x = x_update(z, w, y, index, reo)
        if prox == 'l1':
            z = soft(x + w, reo * lambda1)
        else:
            s = soft(z, 1 / b)
            t = (1 - alpha) * z + alpha * (x + w) + alpha * reo * lambda1 * b * (z - s)
            z = soft(t, alpha * reo * lambda1)
        w = w + x - z
    return x, z, w


# --------------------------------------------------------------------------
# PnP variants with an injectable denoiser.
#   S3:241-296 (PnP-L1), S6:250-308 / S6:477-525 (PnP-CNC)
# The denoiser is any callable ``D(arr_f32[N,N], i) -> arr_f32[N,N]`` so the
# CUDA path and the oracle can be driven by the *same* network.
# float32 round trips and the [0,1] clamps of x, z AND the dual w are kept.
# --------------------------------------------------------------------------
def _f32(a):
    return np.asarray(a).astype(np.float32)


def pnp_admm_l1(img, mask, noises, denoiser: Callable, iter_num=50, reo=0.26, fft_dtype=None):
    """S3:241-296.  ``fft_dtype=np.complex128`` reproduces the author's
    NumPy-1.x behaviour (float32 state upcast to complex128 in fft2);
    ``None`` keeps whatever the installed NumPy does (>=2.0: complex64)."""
    img = np.asarray(img).squeeze()
    y = acquire(img, mask, noises)
    index = np.nonzero(mask)
    x = zero_filled(y)
    z = np.copy(x)
    w = np.zeros(img.shape, dtype=np.float64)
    for i in range(iter_num):
        zz, ww = (z, w) if fft_dtype is None else (z.astype(np.float64), w.astype(np.float64))
        x = x_update(zz, ww, y, index, reo)                   # S3:259-264
        x = _f32(x)                                           # S3:266
        z = _f32(np.absolute(z))                              # S3:269-270
        w = _f32(np.absolute(w))                              # S3:273-274
        z = _f32(denoiser(x + w, i))                          # S3:290
        w = w + x - z                                         # S3:293
        x = np.clip(x, 0, 1)                                  # S3:294
        z = np.clip(z, 0, 1)                                  # S3:295
        w = np.clip(w, 0, 1)                                  # S3:296
    return x


def pnp_admm_cnc(img, mask, noises, denoiser1: Callable, denoiser2: Optional[Callable] = None,
                 alpha=1.2, iter_num=50, lambda1=4.0, reo=0.45, b=0.3, fft_dtype=None):
    """S6:250-308 (one model, ``denoiser2 is None``) and S6:477-525 (model pair)."""
    if denoiser2 is None:
        denoiser2 = denoiser1
    img = np.asarray(img).squeeze()
    y = acquire(img, mask, noises)
    index = np.nonzero(mask)
    x = zero_filled(y)
    z = np.copy(x)
    w = np.zeros(img.shape, dtype=np.float64)
    for i in range(iter_num):
        zz, ww = (z, w) if fft_dtype is None else (z.astype(np.float64), w.astype(np.float64))
        x = x_update(zz, ww, y, index, reo)                   # S6:266-271
        x = _f32(x)
        z = _f32(np.absolute(z))                              # S6:277
        w = _f32(np.absolute(w))                              # S6:282
        s = _f32(denoiser1(z, i))                             # S6:300
        c = np.float32(alpha * reo * lambda1 * b)
        t = np.float32(1 - alpha) * z + np.float32(alpha) * (x + w) + c * (z - s)   # S6:301
        z = _f32(denoiser2(_f32(t), i))                       # S6:302
        w = w + x - z                                         # S6:305
        x = np.clip(x, 0, 1)                                  # S6:306
        z = np.clip(z, 0, 1)
        w = np.clip(w, 0, 1)
    return x


# --------------------------------------------------------------------------
# a10  DPIR sigma schedule                              utils/utils_pnp.py:14-23
# --------------------------------------------------------------------------
def get_rho_sigma(sigma=2.55 / 255, iter_num=15, modelSigma1=49.0, modelSigma2=2.55, w=1.0):
    modelSigmaS = np.logspace(np.log10(modelSigma1), np.log10(modelSigma2), iter_num).astype(np.float32)
    modelSigmaS_lin = np.linspace(modelSigma1, modelSigma2, iter_num).astype(np.float32)
    sigmas = (modelSigmaS * w + modelSigmaS_lin * (1 - w)) / 255.
    rhos = list(map(lambda x: 0.23 * (sigma ** 2) / (x ** 2), sigmas))
    return rhos, sigmas


# --------------------------------------------------------------------------
# metrics                                       utils/utils_image.py:543-636
# --------------------------------------------------------------------------
def calculate_psnr(img1, img2, border=0):
    """utils_image.py:543-556 — images in [0,255]."""
    h, w = img1.shape[:2]
    a = img1[border:h - border, border:w - border].astype(np.float64)
    b = img2[border:h - border, border:w - border].astype(np.float64)
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return float('inf')
    return 20 * math.log10(255.0 / math.sqrt(mse))


def psnr_zero_fill(x, im_orig):
    """utils_image.py:559-564 (used for the zero-filling print, S1:101)."""
    M, N = np.shape(x)
    mse = (np.sum((np.absolute(x - im_orig)) ** 2)) / (M * N)
    return 10 * np.log10(255 * 255 / mse)


def calculate_re(img1, img2, border=0):
    """utils_image.py:622-636 — ||H - E||_2 / ||H||_2."""
    h, w = img1.shape[:2]
    a = img1[border:h - border, border:w - border].astype(np.float64)
    b = img2[border:h - border, border:w - border].astype(np.float64)
    return float(np.linalg.norm(b - a) / np.linalg.norm(b))


def _gauss_window_11():
    # cv2.getGaussianKernel(11, 1.5): exp(-(i-5)^2 / (2 sigma^2)), normalised.
    i = np.arange(11, dtype=np.float64) - 5.0
    k = np.exp(-(i * i) / (2 * 1.5 * 1.5))
    k /= k.sum()
    return k


def calculate_ssim(img1, img2, border=0):
    """utils_image.py:570-615 — 11x11 sigma 1.5 Gaussian, 'valid' region.
    Restated with separable correlation (window is an outer product), no cv2."""
    h, w = img1.shape[:2]
    a = np.squeeze(img1[border:h - border, border:w - border]).astype(np.float64)
    b = np.squeeze(img2[border:h - border, border:w - border]).astype(np.float64)
    k = _gauss_window_11()

    def filt(im):
        # valid-region separable filter == filter2D(...)[5:-5, 5:-5]
        t = np.zeros((im.shape[0] - 10, im.shape[1]), dtype=np.float64)
        for i in range(11):
            t += k[i] * im[i:i + t.shape[0], :]
        o = np.zeros((t.shape[0], im.shape[1] - 10), dtype=np.float64)
        for j in range(11):
            o += k[j] * t[:, j:j + o.shape[1]]
        return o

    C1 = (0.01 * 255) ** 2
    C2 = (0.03 * 255) ** 2
    mu1, mu2 = filt(a), filt(b)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 ** 2, mu2 ** 2, mu1 * mu2
    s1 = filt(a ** 2) - mu1_sq
    s2 = filt(b ** 2) - mu2_sq
    s12 = filt(a * b) - mu1_mu2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return float(m.mean())


# --------------------------------------------------------------------------
# preprocessing                        utils_image.py:145-157,181-186,495-508
# --------------------------------------------------------------------------
def uint2single(img_u8):
    """utils_image.py:181-182."""
    return np.float32(img_u8 / 255.)


def single2uint(img):
    """utils_image.py:185-186: np.uint8((img.clip(0, 1) * 255.).round()) — the img_E of S6:315 / S6:531."""
    return np.uint8((img.clip(0, 1) * 255.).round())


def preprocess_uint8(img_u8):
    """S1:85-90: gray uint8 (H,W) -> modcrop(8) -> /255 float32 (the clip
    round trip S1:89-90 is the identity for uint8 input)."""
    H, W = img_u8.shape[:2]
    img_u8 = img_u8[:H - H % 8, :W - W % 8]
    return uint2single(img_u8)
