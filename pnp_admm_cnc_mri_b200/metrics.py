"""Image-quality metrics of the reference harness (utils/utils_image.py:543-636): PSNR, SSIM, RE on
[0,255] images.  Host-side (NumPy) like the reference's; they are evaluated once per image after the
loop (S1:144-146) and are not part of the accelerated path."""
from __future__ import annotations

import math

import numpy as np


def _crop(a, border):
    h, w = a.shape[:2]
    return a[border:h - border, border:w - border]


def calculate_psnr(img1, img2, border=0):
    """20 log10(255 / sqrt(MSE)) in float64 (utils_image.py:543-556)."""
    if img1.shape != img2.shape:
        raise ValueError('Input images must have the same dimensions.')
    a = _crop(img1, border).astype(np.float64)
    b = _crop(img2, border).astype(np.float64)
    mse = np.mean((a - b) ** 2)
    return float('inf') if mse == 0 else 20 * math.log10(255.0 / math.sqrt(mse))


def psnr(x, im_orig):
    """utils_image.py:559-564 (the zero-filling print; accepts complex x)."""
    M, N = np.shape(x)
    mse = np.sum(np.absolute(x - im_orig) ** 2) / (M * N)
    return 10 * np.log10(255 * 255 / mse)


def calculate_re(img1, img2, border=0):
    """||H - E||_2 / ||H||_2 (utils_image.py:622-636)."""
    if img1.shape != img2.shape:
        raise ValueError('Input images must have the same dimensions.')
    a = _crop(img1, border).astype(np.float64)
    b = _crop(img2, border).astype(np.float64)
    return float(np.linalg.norm(b - a) / np.linalg.norm(b))


def _gaussian_11():
    i = np.arange(11, dtype=np.float64) - 5.0
    k = np.exp(-(i * i) / (2 * 1.5 * 1.5))
    return k / k.sum()


def _filter_valid(im, k):
    rows = im.shape[0] - 10
    t = sum(k[i] * im[i:i + rows, :] for i in range(11))
    cols = im.shape[1] - 10
    return sum(k[j] * t[:, j:j + cols] for j in range(11))


def calculate_ssim(img1, img2, border=0):
    """SSIM with the 11x11 sigma-1.5 Gaussian window on the valid region (utils_image.py:570-615)."""
    if img1.shape != img2.shape:
        raise ValueError('Input images must have the same dimensions.')
    a = np.squeeze(_crop(img1, border)).astype(np.float64)
    b = np.squeeze(_crop(img2, border)).astype(np.float64)
    if a.ndim != 2:
        raise ValueError('grayscale images only on this path')
    k = _gaussian_11()
    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    mu1, mu2 = _filter_valid(a, k), _filter_valid(b, k)
    s1 = _filter_valid(a * a, k) - mu1 * mu1
    s2 = _filter_valid(b * b, k) - mu2 * mu2
    s12 = _filter_valid(a * b, k) - mu1 * mu2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    return float(m.mean())
