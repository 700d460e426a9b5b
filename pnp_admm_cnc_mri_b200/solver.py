"""Batched host API of the ADMM reconstruction path (PyTorch is plumbing only: device memory and
streams; all arithmetic runs in libpnpadmm.so through the C ABI of include/pnpadmm.h).

    x = admm_solve(images[B,N,N], mask[N,N] | [B,N,N], noises[N,N] | [B,N,N],
                   prox='l1' | 'cnc', iter_num=, lambda1=, reo=, alpha=, b=, dtype='float32')

mirrors what the reference does per image inside ``ADMM_L1`` (S1:97-132) / ``ADMM_CNC``
(S4:101-138): acquisition ``y = fft2(img)*mask + noises``, zero-filled start, ``iter_num`` ADMM
iterations, returns the last x-update.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _abi

ArrayLike = Union[np.ndarray, torch.Tensor]

_PROX = {'l1': _abi.PROX_L1, 'cnc': _abi.PROX_CNC}
_KERNEL = {'auto': _abi.KERNEL_AUTO, 'cluster': _abi.KERNEL_CLUSTER, 'streaming': _abi.KERNEL_STREAMING, 'rowsep': _abi.KERNEL_ROWSEP}
_ROWSEP_CACHE = {}


def mask_is_row_separable(mask) -> bool:
    """True if the sampling mask consists of full k-space lines, mask[kr, kc] == mask[0, kc] (Cartesian undersampling such as
    CS_MRI/Q_Cartesian30): the reconstruction then runs on the row-separable kernel (kernel='rowsep').  Host arrays are tested
    directly; a CUDA tensor is tested once per (storage, version) with one device reduction and remembered."""
    if os.environ.get('PNPADMM_NO_ROWSEP'):
        return False
    if isinstance(mask, torch.Tensor) and mask.is_cuda:
        key = (mask.data_ptr(), mask._version, tuple(mask.shape))
        hit = _ROWSEP_CACHE.get(key)
        if hit is None:
            if len(_ROWSEP_CACHE) > 64:
                _ROWSEP_CACHE.clear()
            hit = _ROWSEP_CACHE[key] = bool(((mask != 0) == (mask[:1] != 0)).all().item()) if mask.ndim == 2 else False
        return hit
    m = np.asarray(mask.cpu() if isinstance(mask, torch.Tensor) else mask)
    return bool(m.ndim == 2 and ((m != 0) == (m[:1] != 0)).all())


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _abi.PnpAdmmError('no CUDA device visible: this package has no CPU path')
    return torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class AdmmSolver:
    """Owns the device workspace for a fixed (B, N, dtype, mask layout) and exposes the C-ABI calls
    on torch CUDA tensors.  All calls are asynchronous on the current torch stream."""

    def __init__(self, B: int, N: int, dtype: str = 'float32', mask_batched: bool = False, device=None):
        if dtype not in ('float32', 'float64'):
            raise ValueError("dtype must be 'float32' or 'float64'")
        self.lib = _abi.load()
        self.device = _require_cuda(device)
        self.B, self.N = int(B), int(N)
        self.f64 = dtype == 'float64'
        self.sfx = 'f64' if self.f64 else 'f32'
        self.rdtype = torch.float64 if self.f64 else torch.float32
        self.cdtype = torch.complex128 if self.f64 else torch.complex64
        self.mask_batched = bool(mask_batched)
        nbytes = self.lib.pnpadmm_workspace_bytes(self.B, self.N, int(self.f64), int(self.mask_batched))
        if nbytes == 0:
            raise ValueError(f'bad problem size B={B}, N={N}')
        with torch.cuda.device(self.device):
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.ws_bytes = nbytes
        self._prepared_reo = None

    # -- helpers ---------------------------------------------------------------------------
    def _fn(self, name):
        return getattr(self.lib, f'pnpadmm_{name}_{self.sfx}')

    def _real(self, a: ArrayLike, name: str) -> torch.Tensor:
        t = torch.as_tensor(a)
        t = t.to(device=self.device, dtype=self.rdtype).contiguous()
        if t.shape != (self.B, self.N, self.N):
            raise ValueError(f'{name} must have shape {(self.B, self.N, self.N)}, got {tuple(t.shape)}')
        return t

    def _mask(self, mask: ArrayLike) -> torch.Tensor:
        m = torch.as_tensor(mask)
        m = (m != 0).to(device=self.device, dtype=torch.uint8).contiguous()
        want = (self.B, self.N, self.N) if self.mask_batched else (self.N, self.N)
        if m.shape != want:
            raise ValueError(f'mask must have shape {want}, got {tuple(m.shape)}')
        return m

    def new(self) -> torch.Tensor:
        return torch.empty((self.B, self.N, self.N), dtype=self.rdtype, device=self.device)

    # -- a2 ----------------------------------------------------------------------------------
    def acquire(self, img: ArrayLike, mask: ArrayLike, noises: ArrayLike, spectrum: str = 'numpy') -> torch.Tensor:
        """y = fft2(img) * mask + noises (S1:99).

        spectrum='numpy' follows NumPy's dtype rule in the float64 build: a float32 image (what the
        reference feeds, utils_image.uint2single) gets a complex64-rounded spectrum (NumPy >= 2), a
        float64 image a full double one.  'double' never rounds (NumPy 1.x, the author's setup)."""
        if spectrum not in ('numpy', 'double'):
            raise ValueError("spectrum must be 'numpy' or 'double'")
        src_dtype = torch.as_tensor(img).dtype
        round_f32 = int(spectrum == 'numpy' and src_dtype in (torch.float32, torch.float16, torch.uint8))
        img = self._real(img, 'img')
        m = self._mask(mask)
        nz = torch.as_tensor(noises).to(device=self.device, dtype=self.cdtype).contiguous()
        if nz.shape == (self.N, self.N):
            nb = 0
        elif nz.shape == (self.B, self.N, self.N):
            nb = 1
        else:
            raise ValueError(f'noises must be (N,N) or (B,N,N), got {tuple(nz.shape)}')
        y = torch.empty((self.B, self.N, self.N), dtype=self.cdtype, device=self.device)
        with torch.cuda.device(self.device):
            extra = (round_f32,) if self.f64 else ()
            _abi.check(self._fn('acquire')(img.data_ptr(), m.data_ptr(), nz.data_ptr(), y.data_ptr(), self.B, self.N,
                                           int(self.mask_batched), nb, *extra, self.ws.data_ptr(), self.ws_bytes,
                                           _stream_ptr()))
        return y

    def zero_filled(self, y: torch.Tensor) -> torch.Tensor:
        """x0 = |ifft2(y)| (S1:100,104)."""
        y = self._cplx(y)
        x0 = self.new()
        with torch.cuda.device(self.device):
            _abi.check(self._fn('zero_filled')(y.data_ptr(), x0.data_ptr(), self.B, self.N, self.ws.data_ptr(),
                                               self.ws_bytes, _stream_ptr()))
        return x0

    def _cplx(self, y: ArrayLike) -> torch.Tensor:
        y = torch.as_tensor(y).to(device=self.device, dtype=self.cdtype).contiguous()
        if y.shape != (self.B, self.N, self.N):
            raise ValueError(f'y must have shape {(self.B, self.N, self.N)}, got {tuple(y.shape)}')
        return y

    # -- data term ----------------------------------------------------------------------------
    def prepare(self, y: torch.Tensor, mask: ArrayLike, reo: float) -> None:
        y = self._cplx(y)
        m = self._mask(mask)
        with torch.cuda.device(self.device):
            _abi.check(self._fn('prepare')(y.data_ptr(), m.data_ptr(), self.B, self.N, int(self.mask_batched), float(reo),
                                           self.ws.data_ptr(), self.ws_bytes, _stream_ptr()))
        self._prepared_reo = float(reo)

    # -- a3 ----------------------------------------------------------------------------------
    def xupdate(self, z: torch.Tensor, w: torch.Tensor, want_xpw: bool = False, kernel: str = 'auto'):
        """x = |Re(ifft2(blend(fft2(z - w))))| (S1:115-120); optionally also x + w."""
        if self._prepared_reo is None:
            raise _abi.PnpAdmmError('xupdate called before prepare()')
        z, w = self._real(z, 'z'), self._real(w, 'w')
        x = self.new()
        xpw = self.new() if want_xpw else None
        with torch.cuda.device(self.device):
            _abi.check(self._fn('xupdate')(z.data_ptr(), w.data_ptr(), x.data_ptr(), _ptr(xpw), self.B, self.N,
                                           int(self.mask_batched), _KERNEL[kernel], self.ws.data_ptr(), self.ws_bytes,
                                           _stream_ptr()))
        return (x, xpw) if want_xpw else x

    # -- the loop -----------------------------------------------------------------------------
    def iterate(self, x: torch.Tensor, z: torch.Tensor, w: torch.Tensor, prox: str, iter_num: int, lambda1: float,
                reo: float, alpha: float = 0.0, b: float = 1.0, kernel: str = 'auto') -> None:
        """`iter_num` ADMM iterations in place on (z, w); x receives the last x-update."""
        if self._prepared_reo is None or self._prepared_reo != float(reo):
            raise _abi.PnpAdmmError('iterate: call prepare(y, mask, reo) with the same reo first')
        for t, n in ((x, 'x'), (z, 'z'), (w, 'w')):
            if not (t.is_cuda and t.dtype == self.rdtype and t.is_contiguous() and t.shape == (self.B, self.N, self.N)):
                raise ValueError(f'{n} must be a contiguous CUDA {self.rdtype} tensor of shape {(self.B, self.N, self.N)}')
        with torch.cuda.device(self.device):
            _abi.check(self._fn('iterate')(x.data_ptr(), z.data_ptr(), w.data_ptr(), self.B, self.N,
                                           int(self.mask_batched), _PROX[prox], int(iter_num), float(lambda1), float(reo),
                                           float(alpha), float(b), _KERNEL[kernel], self.ws.data_ptr(), self.ws_bytes,
                                           _stream_ptr()))

    def solve(self, y: torch.Tensor, mask: ArrayLike, prox: str, iter_num: int, lambda1: float, reo: float,
              alpha: float = 0.0, b: float = 1.0, kernel: str = 'auto') -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """x0 = |ifft2(y)|, z = x0, w = 0, then the loop (S1:100-126 / S4:104-132). Returns (x, z, w)."""
        if prox not in _PROX:
            raise ValueError("prox must be 'l1' or 'cnc'")
        y = self._cplx(y)
        m = self._mask(mask)
        x, z, w = self.new(), self.new(), self.new()
        with torch.cuda.device(self.device):
            _abi.check(self._fn('solve')(y.data_ptr(), m.data_ptr(), x.data_ptr(), z.data_ptr(), w.data_ptr(), self.B,
                                         self.N, int(self.mask_batched), _PROX[prox], int(iter_num), float(lambda1),
                                         float(reo), float(alpha), float(b), _KERNEL[kernel], self.ws.data_ptr(),
                                         self.ws_bytes, _stream_ptr()))
        self._prepared_reo = float(reo)
        return x, z, w

    def reconstruct(self, img: ArrayLike, mask: ArrayLike, noises: ArrayLike, prox: str, iter_num: int, lambda1: float, reo: float,
                    alpha: float = 0.0, b: float = 1.0, kernel: str = 'auto') -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """From images to reconstructions in one call (S1:97-132 / S4:101-138 per image): acquisition, zero-filled start,
        data term and `iter_num` iterations.  `img`: (B,N,N) real in [0,1], or uint8 gray levels (divided by 255 on the device).
        At N = 256 in float32 with one mask / noise for the batch all of it runs inside the cluster kernel.  Returns (x, z, w)."""
        if prox not in _PROX:
            raise ValueError("prox must be 'l1' or 'cnc'")
        t = torch.as_tensor(img)
        if t.shape != (self.B, self.N, self.N):
            raise ValueError(f'img must have shape {(self.B, self.N, self.N)}, got {tuple(t.shape)}')
        if t.dtype == torch.uint8:
            img8, imgf = t.to(self.device).contiguous(), None
        else:
            img8, imgf = None, t.to(device=self.device, dtype=self.rdtype).contiguous()
        if (kernel == 'auto' and self.N in (256, 512, 1024) and not self.f64 and not self.mask_batched and iter_num >= 1
                and mask_is_row_separable(mask)):
            kernel = 'rowsep'                  # full k-space lines: every image row is solved on its own (K3)
        m = self._mask(mask)
        nz, nb = _noise_arg(self, noises)
        if kernel == 'rowsep' and nb:
            kernel = 'auto'
        x, z, w = self.new(), self.new(), self.new()
        with torch.cuda.device(self.device):
            _abi.check(self._fn('reconstruct')(_ptr(imgf), _ptr(img8), m.data_ptr(), nz.data_ptr(), x.data_ptr(), z.data_ptr(),
                                               w.data_ptr(), self.B, self.N, int(self.mask_batched), nb, _PROX[prox], int(iter_num),
                                               float(lambda1), float(reo), float(alpha), float(b), _KERNEL[kernel],
                                               self.ws.data_ptr(), self.ws_bytes, _stream_ptr()))
        self._prepared_reo = None          # the fused path does not leave a prepared data term for xupdate / iterate
        return x, z, w


class HostPipeline:
    """Back-to-back reconstructions from HOST buffers (the drop-in functions' situation: uint8 images on the host in,
    float32 reconstructions on the host out), pipelined over `n_slots` device slots: H2D copies, kernels and the D2H
    copy run on three streams so the copies of neighbouring batches overlap the kernels.  Owns the C-ABI pipeline
    object (events + captured compute graphs), its streams, the device scratch and the workspace; several
    HostPipelines can coexist on one device.

        pipe = HostPipeline(B, N, n_slots=3)
        pipe.submit(slot, h_img_u8, h_mask_u8, h_noise_c64_as_float, h_x, prox='cnc', iter_num=50, ...)
        pipe.wait(slot)          # h_x of that slot is valid

    Host tensors should be pinned (otherwise the copies are synchronous)."""

    def __init__(self, B: int, N: int, n_slots: int = 2, device=None, output: str = 'float32'):
        import ctypes
        if output not in ('float32', 'uint8'):
            raise ValueError("output must be 'float32' (the reconstructions x) or 'uint8' (img_E as the reference saves it)")
        self.output = output
        self.lib = _abi.load()
        self.device = _require_cuda(device)
        self.B, self.N, self.n_slots = int(B), int(N), int(n_slots)
        with torch.cuda.device(self.device):
            nbytes = self.lib.pnpadmm_host_pipeline_scratch_bytes(self.B, self.N, self.n_slots)
            if nbytes == 0:
                raise ValueError(f'bad pipeline shape B={B}, N={N}, n_slots={n_slots} (2..4 slots)')
            h = ctypes.c_void_p()
            _abi.check(self.lib.pnpadmm_pipeline_create(ctypes.byref(h), self.n_slots))
            self._h = h
            if output == 'uint8':
                _abi.check(self.lib.pnpadmm_pipeline_set_output(h, _abi.OUT_U8))
            self.scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.ws_bytes = self.lib.pnpadmm_workspace_bytes(self.B, self.N, 0, 0)
            self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
            self.s_compute, self.s_h2d, self.s_d2h = (torch.cuda.Stream(self.device) for _ in range(3))

    def submit(self, slot: int, h_img: torch.Tensor, h_mask: torch.Tensor, h_noise: torch.Tensor, h_x: torch.Tensor, *,
               prox: str = 'cnc', iter_num: int = 50, lambda1: float = 0.5, reo: float = 0.05, alpha: float = 0.45,
               b: float = 64.0, kernel: str = 'auto') -> None:
        B, N = self.B, self.N
        if not (h_img.dtype == torch.uint8 and tuple(h_img.shape) == (B, N, N) and h_img.is_contiguous() and not h_img.is_cuda):
            raise ValueError(f'h_img must be a contiguous host uint8 tensor of shape {(B, N, N)}')
        if not (h_mask.dtype == torch.uint8 and tuple(h_mask.shape) == (N, N) and h_mask.is_contiguous() and not h_mask.is_cuda):
            raise ValueError(f'h_mask must be a contiguous host uint8 tensor of shape {(N, N)}')
        if not (h_noise.dtype == torch.float32 and h_noise.numel() == 2 * N * N and h_noise.is_contiguous() and not h_noise.is_cuda):
            raise ValueError('h_noise must be a contiguous host float32 tensor holding (N,N) complex64 as (re, im) pairs')
        want = torch.uint8 if self.output == 'uint8' else torch.float32
        if not (h_x.dtype == want and tuple(h_x.shape) == (B, N, N) and h_x.is_contiguous() and not h_x.is_cuda):
            raise ValueError(f'h_x must be a contiguous host {want} tensor of shape {(B, N, N)}')
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pnpadmm_reconstruct_host_pipelined_f32(
                self._h, h_img.data_ptr(), h_mask.data_ptr(), h_noise.data_ptr(), h_x.data_ptr(), B, N, _PROX[prox],
                int(iter_num), float(lambda1), float(reo), float(alpha), float(b), _KERNEL[kernel],
                self.scratch.data_ptr(), self.scratch.numel(), self.ws.data_ptr(), self.ws_bytes, int(slot),
                self.s_compute.cuda_stream, self.s_h2d.cuda_stream, self.s_d2h.cuda_stream))

    def wait(self, slot: int) -> None:
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pnpadmm_reconstruct_host_wait(self._h, int(slot)))

    def close(self) -> None:
        if getattr(self, '_h', None) is not None and self._h:
            with torch.cuda.device(self.device):
                for s in (self.s_compute, self.s_h2d, self.s_d2h):
                    s.synchronize()
                self.lib.pnpadmm_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _noise_arg(solver, noises):
    nz = torch.as_tensor(noises).to(device=solver.device, dtype=solver.cdtype).contiguous()
    if nz.shape == (solver.N, solver.N):
        return nz, 0
    if nz.shape == (solver.B, solver.N, solver.N):
        return nz, 1
    raise ValueError(f'noises must be (N,N) or (B,N,N), got {tuple(nz.shape)}')


def admm_solve(images: ArrayLike, mask: ArrayLike, noises: ArrayLike, *, prox: str = 'l1', iter_num: int = 50,
               lambda1: float = 0.1, reo: float = 0.015, alpha: float = 0.45, b: float = 64.0,
               dtype: str = 'float32', kernel: str = 'auto', return_state: bool = False, device=None):
    """Batched ADMM-L1 / ADMM-CNC reconstruction.

    images : (B,N,N) or (N,N) real in [0,1] (the reference's ``img_L``, S1:85-90)
    mask   : (N,N) or (B,N,N) 0/1, DC at [0,0] (``Q_*.mat['Q1']``)
    noises : (N,N) or (B,N,N) complex, already scaled (S1:186)
    Returns x with the input's leading shape; numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.
    """
    if prox not in _PROX:
        raise ValueError("prox must be 'l1' or 'cnc' (PnP denoisers go through pnp_admm_cnc_mri_b200.pnp)")
    as_numpy = not isinstance(images, torch.Tensor)
    img = torch.as_tensor(images)
    single = img.ndim == 2
    if single:
        img = img[None]
    if img.ndim != 3 or img.shape[1] != img.shape[2]:
        raise ValueError(f'images must be (B,N,N) or (N,N), got {tuple(img.shape)}')
    B, N = int(img.shape[0]), int(img.shape[1])
    m = torch.as_tensor(mask)
    solver = AdmmSolver(B, N, dtype=dtype, mask_batched=(m.ndim == 3), device=device)
    if dtype == 'float32' and not return_state:
        # one call from images to reconstructions (acquisition fused into the cluster kernel where it applies)
        x, z, w = solver.reconstruct(img, m, noises, prox, iter_num, lambda1, reo, alpha, b, kernel=kernel)
        y = None
    else:
        y = solver.acquire(img, m, noises)
        x, z, w = solver.solve(y, m, prox, iter_num, lambda1, reo, alpha, b, kernel=kernel)

    def out(t):
        t = t[0] if single else t
        return t.cpu().numpy() if as_numpy else t

    if return_state:
        return out(x), out(z), out(w), (y[0] if single else y).cpu().numpy() if as_numpy else (y[0] if single else y)
    return out(x)


# ---- pointwise pieces (a1, a8, a6) on CUDA tensors --------------------------------------------
def _pw_args(*ts):
    t0 = ts[0]
    if t0.dtype not in (torch.float32, torch.float64):
        raise ValueError('float32 / float64 CUDA tensors only')
    for t in ts:
        if not (t.is_cuda and t.is_contiguous() and t.dtype == t0.dtype and t.shape == t0.shape):
            raise ValueError('operands must be contiguous CUDA tensors of identical shape and dtype')
    return 'f64' if t0.dtype == torch.float64 else 'f32'


def soft(x: torch.Tensor, c: float) -> torch.Tensor:
    """soft(x, c) = fmax(|x| - c, 0) * sign(x) (S1:18-19) on a CUDA tensor."""
    sfx = _pw_args(x)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _abi.check(getattr(_abi.load(), 'pnpadmm_soft_' + sfx)(x.data_ptr(), out.data_ptr(), float(c), x.numel(), _stream_ptr()))
    return out


def cnc_combine(z, x, w, s, alpha: float, coef: float) -> torch.Tensor:
    """t = (1 - alpha) z + alpha (x + w) + coef (z - s) (S6:301)."""
    sfx = _pw_args(z, x, w, s)
    t = torch.empty_like(z)
    with torch.cuda.device(z.device):
        _abi.check(getattr(_abi.load(), 'pnpadmm_cnc_combine_' + sfx)(z.data_ptr(), x.data_ptr(), w.data_ptr(), s.data_ptr(),
                                                                      t.data_ptr(), float(alpha), float(coef), z.numel(),
                                                                      _stream_ptr()))
    return t


def dual_update_(x, z, w, clamp01: bool) -> None:
    """In place: w = w + x - z; optionally clamp x, z, w to [0,1] (S3:293-296)."""
    sfx = _pw_args(x, z, w)
    with torch.cuda.device(x.device):
        _abi.check(getattr(_abi.load(), 'pnpadmm_dual_update_' + sfx)(x.data_ptr(), z.data_ptr(), w.data_ptr(), int(clamp01),
                                                                      x.numel(), _stream_ptr()))


def image_metrics(x: torch.Tensor, ref_u8: torch.Tensor, quantize: bool = False) -> torch.Tensor:
    """Per-image (PSNR dB, SSIM, RE) on the device: ``x`` (B,N,N) float32/float64 reconstructions on unit scale,
    ``ref_u8`` (B,N,N) uint8 ground truth.  Same definitions as the reference's ``calculate_psnr`` /
    ``calculate_ssim`` / ``calculate_re`` (utils/utils_image.py:543-636, border 0), double precision.
    ``quantize=False`` scores ``255 x`` (S1:133), ``True`` scores ``uint8(round(255 clip(x)))`` (S6:315).
    Returns a (B,3) float64 CUDA tensor; nothing is copied to the host."""
    dev = _require_cuda(x.device)
    if x.dim() == 2:
        x, ref_u8 = x[None], ref_u8[None]
    if x.dtype not in (torch.float32, torch.float64) or ref_u8.dtype != torch.uint8 or x.shape != ref_u8.shape \
            or x.dim() != 3 or x.shape[1] != x.shape[2]:
        raise ValueError('image_metrics: x (B,N,N) float32/float64 and ref_u8 (B,N,N) uint8 of the same shape')
    x, ref_u8 = x.contiguous(), ref_u8.to(dev).contiguous()
    B, N = int(x.shape[0]), int(x.shape[1])
    lib = _abi.load()
    out = torch.empty((B, 3), dtype=torch.float64, device=dev)
    scratch = torch.empty(lib.pnpadmm_metrics_scratch_bytes(B), dtype=torch.uint8, device=dev)
    sfx = 'f64' if x.dtype == torch.float64 else 'f32'
    with torch.cuda.device(dev):
        _abi.check(getattr(lib, 'pnpadmm_metrics_' + sfx)(x.data_ptr(), ref_u8.data_ptr(), B, N, int(bool(quantize)),
                                                          out.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream_ptr()))
    return out
