"""DnCNN / FDnCNN / IRCNN / FFDNet forward through the tensor-core kernels of the C ABI (``pnpadmm_dncnn_forward_bf16``,
``pnpadmm_dncnn_forward_dilated_bf16``, ``pnpadmm_ffdnet_forward_bf16``).

The reference runs its denoisers as stock ``torch.nn`` modules (models/network_dncnn.py:36-67, 120-141, called
from S6:353-359 / S3:20-35).  On B200 the 64->64 layers are hand-written tcgen05 implicit GEMMs
(csrc/dncnn_tc.cuh); this module only packs the weights of an ``nn.Module`` into the layouts the kernels read
and owns the two bf16 activation buffers.  Host logic only: there is no fallback when the library is missing.
"""
from __future__ import annotations

import ctypes
from typing import List, Tuple

import torch
import torch.nn as nn

from . import _abi


MAX_DILATION = 4


def conv_layers(net: nn.Module) -> List[nn.Conv2d]:
    """The convolutions of a DnCNN / FDnCNN / IRCNN in forward order (``net.model`` is conv, ReLU, conv, ...).  Middle layers
    may be dilated (dilation = padding <= 4, IRCNN); the first and the last layer may not."""
    convs = [m for m in net.model if isinstance(m, nn.Conv2d)]
    if len(convs) < 2:
        raise ValueError('need at least a first and a last convolution')
    for k, c in enumerate(convs):
        d = c.dilation[0]
        ok_d = d == 1 or (0 < k < len(convs) - 1 and 1 <= d <= MAX_DILATION)
        if c.kernel_size != (3, 3) or c.dilation != (d, d) or c.padding != (d, d) or not ok_d or c.stride != (1, 1) or c.bias is None:
            raise ValueError(f'layer {k}: only conv3x3, stride 1, padding = dilation (<= {MAX_DILATION} in middle layers, 1 in the '
                             f'first and last), with bias is supported')
    if convs[0].out_channels != 64 or convs[-1].in_channels != 64 or convs[-1].out_channels != 1:
        raise ValueError('expected 64 feature channels and one output channel')
    for c in convs[1:-1]:
        if c.in_channels != 64 or c.out_channels != 64:
            raise ValueError('middle layers must be 64 -> 64')
    return convs


def pack_conv64(weight: torch.Tensor, n_out: int = 64) -> torch.Tensor:
    """[c_out][64][3][3] float -> [kx][c_in // 8][ky][n_out][c_in % 8] bf16 (rows c_out.. n_out-1 zero): the
    shared-memory image of the B operand (no-swizzle K-major core matrices of 8 rows x 8 c_in).  For one (kx, k-chunk)
    the three ky taps are stacked along N (row = ky * n_out + c_out), so one N = 3 n_out instruction feeds the three
    output rows an input row contributes to (csrc/dncnn_tc.cuh)."""
    co = weight.shape[0]
    if weight.shape[1:] != (64, 3, 3) or co > n_out:
        raise ValueError(f'bad weight shape {tuple(weight.shape)}')
    w = torch.zeros((n_out, 64, 3, 3), dtype=torch.float32, device=weight.device)
    w[:co] = weight.float()
    w = w.permute(3, 1, 2, 0).reshape(3, 8, 8, 3, n_out)       # kx, c_in // 8, c_in % 8, ky, c_out
    return w.permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16)


def pack_dncnn(net: nn.Module, device) -> Tuple[dict, int, int]:
    convs = conv_layers(net)
    head, mids, tail = convs[0], convs[1:-1], convs[-1]
    dev = torch.device(device)
    bf = lambda t: t.detach().to(dev, torch.float32).to(torch.bfloat16)      # noqa: E731  (weights live in bf16, like net.to(bf16))
    packed = {
        'w_head': bf(head.weight).float().contiguous(),
        'b_head': bf(head.bias).float().contiguous(),
        'w_mid': (torch.stack([pack_conv64(bf(c.weight)) for c in mids]).contiguous() if mids
                  else torch.zeros(0, dtype=torch.bfloat16, device=dev)),
        'b_mid': (torch.stack([bf(c.bias).float() for c in mids]).contiguous() if mids
                  else torch.zeros(0, dtype=torch.float32, device=dev)),
        'w_tail': pack_conv64(bf(tail.weight), 16),
        'b_tail': bf(tail.bias).float().contiguous(),
        'dilations': [int(c.dilation[0]) for c in mids],
    }
    return packed, head.in_channels, len(mids)


class FusedDnCNN:
    """``y = D(x)`` for a DnCNN / IRCNN (residual, 1 input channel; IRCNN's middle layers are dilated) or FDnCNN (non-residual,
    2 input channels).  ``repack(net)`` re-reads the weights of a module with the same architecture (IRCNN's sigma-indexed
    weight switch, S3:280-288).

    x: (B, cin, H, W) float32 CUDA tensor -> (B, 1, H, W) float32.  Work is enqueued on the current stream.
    The instance owns the two bf16 activation buffers (2 x B*H*W*128 bytes, re-allocated when the shape changes), so calls on
    one instance must be stream-ordered; use one instance per stream for concurrent forwards.
    """

    def __init__(self, net: nn.Module, residual: bool, device='cuda'):
        if not torch.cuda.is_available():
            raise _abi.PnpAdmmError('no CUDA device visible: the tensor-core denoiser has no CPU path')
        self.lib = _abi.load()
        self.device = torch.device(device)
        if self.device.type == 'cuda' and self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.residual = bool(residual)
        self.repack(net)
        self._act = None
        self._act_key = None

    def repack(self, net: nn.Module) -> None:
        self.w, self.cin, self.n_mid = pack_dncnn(net, self.device)
        if self.cin not in (1, 2):
            raise ValueError(f'{self.cin} input channels: only DnCNN / IRCNN (1) and FDnCNN (2) are on the path')
        dil = self.w['dilations']
        self._dil = (ctypes.c_int * len(dil))(*dil) if any(d != 1 for d in dil) else None

    def _buffers(self, B, H, W):
        key = (B, H, W)
        if self._act_key != key:
            n = self.lib.pnpadmm_dncnn_activation_bytes(B, H, W)
            self._act = (torch.empty(n, dtype=torch.uint8, device=self.device), torch.empty(n, dtype=torch.uint8, device=self.device))
            self._act_key = key
        return self._act

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.ndim != 4 or x.shape[1] != self.cin or not x.is_cuda:
            raise ValueError(f'expected a CUDA tensor of shape (B, {self.cin}, H, W), got {tuple(x.shape)} on {x.device}')
        if x.device != self.w['w_head'].device:
            raise ValueError(f'input on {x.device} but the packed weights live on {self.w["w_head"].device}')
        x = x.float().contiguous()
        B, _, H, W = (int(v) for v in x.shape)
        w = self.w
        # the library launches on the CURRENT device and stream: make them the tensors' own
        with torch.cuda.device(x.device):
            a0, a1 = self._buffers(B, H, W)
            out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
            tail_args = (w['w_head'].data_ptr(), w['b_head'].data_ptr(),
                         w['w_mid'].data_ptr() if self.n_mid else None, w['b_mid'].data_ptr() if self.n_mid else None,
                         w['w_tail'].data_ptr(), w['b_tail'].data_ptr(), int(self.residual), a0.data_ptr(), a1.data_ptr(),
                         torch.cuda.current_stream().cuda_stream)
            if self._dil is None:
                _abi.check(self.lib.pnpadmm_dncnn_forward_bf16(x.data_ptr(), out.data_ptr(), B, self.cin, H, W, self.n_mid, *tail_args))
            else:
                _abi.check(self.lib.pnpadmm_dncnn_forward_dilated_bf16(x.data_ptr(), out.data_ptr(), B, self.cin, H, W, self.n_mid,
                                                                       self._dil, *tail_args))
        return out


def pack_ffdnet(net: nn.Module, device) -> Tuple[dict, int]:
    """Weights of an FFDNet (models/network_ffdnet.py:31-54: conv 5->64, (nb - 2) x conv 64->64, conv 64->4, all with bias)
    in the kernels' layouts.  The first layer becomes a 64->64 image whose input channels 5..63 are zero."""
    convs = [m for m in net.model if isinstance(m, nn.Conv2d)]
    for k, c in enumerate(convs):
        if c.kernel_size != (3, 3) or c.padding != (1, 1) or c.dilation != (1, 1) or c.stride != (1, 1) or c.bias is None:
            raise ValueError(f'layer {k}: only conv3x3, stride 1, padding 1, dilation 1, with bias is supported')
    head, mids, tail = convs[0], convs[1:-1], convs[-1]
    if (head.in_channels, head.out_channels, tail.in_channels, tail.out_channels) != (5, 64, 64, 4) or getattr(net, 'sf', 2) != 2:
        raise ValueError('expected the gray FFDNet: 4 sub-pixels + noise map -> 64 -> ... -> 4 sub-pixels, scale factor 2')
    for c in mids:
        if c.in_channels != 64 or c.out_channels != 64:
            raise ValueError('middle layers must be 64 -> 64')
    dev = torch.device(device)
    bf = lambda t: t.detach().to(dev, torch.float32).to(torch.bfloat16)      # noqa: E731
    w_head = torch.zeros((64, 64, 3, 3), dtype=torch.bfloat16, device=dev)
    w_head[:, :5] = bf(head.weight)
    packed = {
        'w_head': pack_conv64(w_head),
        'b_head': bf(head.bias).float().contiguous(),
        'w_mid': (torch.stack([pack_conv64(bf(c.weight)) for c in mids]).contiguous() if mids
                  else torch.zeros(0, dtype=torch.bfloat16, device=dev)),
        'b_mid': (torch.stack([bf(c.bias).float() for c in mids]).contiguous() if mids
                  else torch.zeros(0, dtype=torch.float32, device=dev)),
        'w_tail': pack_conv64(bf(tail.weight), 16),
        'b_tail': bf(tail.bias).float().contiguous(),
    }
    return packed, len(mids)


class FusedFFDNet:
    """``y = FFDNet(x, sigma)`` (reference models/network_ffdnet.py:56-73) on the tensor-core kernels: x (B, 1, H, W) float32
    CUDA -> (B, 1, H, W) float32; ``sigma`` is the scalar level of the noise map (S3:64: 15 / 255).  Buffer ownership and
    stream ordering as for :class:`FusedDnCNN`."""

    def __init__(self, net: nn.Module, device='cuda'):
        if not torch.cuda.is_available():
            raise _abi.PnpAdmmError('no CUDA device visible: the tensor-core denoiser has no CPU path')
        self.lib = _abi.load()
        self.device = torch.device(device)
        if self.device.type == 'cuda' and self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.w, self.n_mid = pack_ffdnet(net, self.device)
        self._act = None
        self._act_key = None

    def _buffers(self, B, H2, W2):
        key = (B, H2, W2)
        if self._act_key != key:
            n = self.lib.pnpadmm_dncnn_activation_bytes(B, H2, W2)
            self._act = (torch.empty(n, dtype=torch.uint8, device=self.device), torch.empty(n, dtype=torch.uint8, device=self.device))
            self._act_key = key
        return self._act

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, sigma: float) -> torch.Tensor:
        if x.ndim != 4 or x.shape[1] != 1 or not x.is_cuda:
            raise ValueError(f'expected a CUDA tensor of shape (B, 1, H, W), got {tuple(x.shape)} on {x.device}')
        if x.device != self.w['w_head'].device:
            raise ValueError(f'input on {x.device} but the packed weights live on {self.w["w_head"].device}')
        x = x.float().contiguous()
        B, _, H, W = (int(v) for v in x.shape)
        w = self.w
        with torch.cuda.device(x.device):
            a0, a1 = self._buffers(B, (H + 1) // 2, (W + 1) // 2)
            out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
            _abi.check(self.lib.pnpadmm_ffdnet_forward_bf16(
                x.data_ptr(), out.data_ptr(), B, H, W, float(sigma), self.n_mid, w['w_head'].data_ptr(), w['b_head'].data_ptr(),
                w['w_mid'].data_ptr() if self.n_mid else None, w['b_mid'].data_ptr() if self.n_mid else None,
                w['w_tail'].data_ptr(), w['b_tail'].data_ptr(), a0.data_ptr(), a1.data_ptr(),
                torch.cuda.current_stream().cuda_stream))
        return out


def to_chunk_planar(x_nhwc: torch.Tensor) -> torch.Tensor:
    """Dense (B, H, W, 64) -> the kernels' inter-layer layout (B, H, 8, W, 8): chunk-planar rows."""
    B, H, W, C = x_nhwc.shape
    return x_nhwc.reshape(B, H, W, 8, 8).permute(0, 1, 3, 2, 4).contiguous()


def from_chunk_planar(x_cp: torch.Tensor) -> torch.Tensor:
    B, H, _, W, _ = x_cp.shape
    return x_cp.permute(0, 1, 3, 2, 4).reshape(B, H, W, 64).contiguous()


def conv64(x_nhwc: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, relu: bool = True, dilation: int = 1) -> torch.Tensor:
    """One 64 -> 64 layer through ``pnpadmm_conv64_bf16`` / ``pnpadmm_conv64_dilated_bf16``: x (B, H, W, 64) bf16 CUDA -> same
    shape (parity tests)."""
    lib = _abi.load()
    if x_nhwc.dtype != torch.bfloat16 or x_nhwc.ndim != 4 or x_nhwc.shape[-1] != 64 or not x_nhwc.is_cuda:
        raise ValueError('expected a (B, H, W, 64) bf16 CUDA tensor')
    x = to_chunk_planar(x_nhwc)
    B, H, W = int(x.shape[0]), int(x.shape[1]), int(x.shape[3])
    wp = pack_conv64(weight.to(x.device))
    bs = bias.to(x.device, torch.float32).contiguous()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        if dilation == 1:
            _abi.check(lib.pnpadmm_conv64_bf16(x.data_ptr(), out.data_ptr(), wp.data_ptr(), bs.data_ptr(), B, H, W, int(relu), st))
        else:
            _abi.check(lib.pnpadmm_conv64_dilated_bf16(x.data_ptr(), out.data_ptr(), wp.data_ptr(), bs.data_ptr(), B, H, W, int(relu),
                                                       int(dilation), st))
    return from_chunk_planar(out)
