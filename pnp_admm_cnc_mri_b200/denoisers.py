"""CNN denoisers of the PnP variants (the only dense contraction on the path; PyTorch, bf16 tensor cores).

Architectures follow the reference's model zoo so that KAIR checkpoints load with ``strict=True``
(same module / parameter names), but are written from scratch:

  DnCNN   (models/network_dncnn.py:36-67)    17 or 20 conv 3x3 x 64, ReLU, residual  x - n(x)
  FDnCNN  (models/network_dncnn.py:120-141)  20 conv, noise-level map as 2nd channel, non-residual
  IRCNN   (models/network_dncnn.py:70-109)   7 dilated conv (1,2,3,4,3,2,1), residual
  FFDNet  (models/network_ffdnet.py:31-73)   pixel-unshuffle x2 + sigma map -> 15 conv -> pixel-shuffle
  UNetRes (models/network_unet.py:76-136)    DRUNet: 4 scales x 4 ResBlocks, bias-free, stride-conv /
                                             conv-transpose resampling

The checkpoints are not shipped (model_zoo/README.md) so weights default to PyTorch's random init
under a fixed seed; ``weights=`` takes a state_dict / path for real checkpoints.

``build_denoiser(model_name, ...)`` returns a callable ``D(x[B,1,H,W] float32, i) -> float32`` that
mirrors the reference dispatch ``denoising_step1`` (S3:19-68) / ``denoising_step2`` (S6:18-67) /
``denoising_step`` (S6:353-359), batch-first.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def _conv_relu_chain(channels: Sequence[int], bias: bool = True, dilations: Optional[Sequence[int]] = None) -> nn.Sequential:
    """conv3x3, ReLU, conv3x3, ReLU, ..., conv3x3 — convs sit at even indices (KAIR key layout)."""
    layers = []
    n = len(channels) - 1
    for i in range(n):
        d = 1 if dilations is None else dilations[i]
        layers.append(nn.Conv2d(channels[i], channels[i + 1], 3, 1, d, dilation=d, bias=bias))
        if i < n - 1:
            layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


class DnCNN(nn.Module):
    def __init__(self, in_nc=1, out_nc=1, nc=64, nb=17, act_mode='R'):
        super().__init__()
        if act_mode != 'R':
            raise NotImplementedError("only act_mode='R' (no batch norm) is on the reference path")
        self.model = _conv_relu_chain([in_nc] + [nc] * (nb - 1) + [out_nc])

    def forward(self, x):
        return x - self.model(x)


class FDnCNN(nn.Module):
    def __init__(self, in_nc=2, out_nc=1, nc=64, nb=20, act_mode='R'):
        super().__init__()
        if act_mode != 'R':
            raise NotImplementedError("only act_mode='R'")
        self.model = _conv_relu_chain([in_nc] + [nc] * (nb - 1) + [out_nc])

    def forward(self, x):
        return self.model(x)


class IRCNN(nn.Module):
    def __init__(self, in_nc=1, out_nc=1, nc=64):
        super().__init__()
        self.model = _conv_relu_chain([in_nc] + [nc] * 6 + [out_nc], dilations=(1, 2, 3, 4, 3, 2, 1))

    def forward(self, x):
        return x - self.model(x)


class FFDNet(nn.Module):
    def __init__(self, in_nc=1, out_nc=1, nc=64, nb=15, act_mode='R'):
        super().__init__()
        if act_mode != 'R':
            raise NotImplementedError("only act_mode='R'")
        self.sf = 2
        self.model = _conv_relu_chain([in_nc * 4 + 1] + [nc] * (nb - 1) + [out_nc * 4])

    def forward(self, x, sigma):
        h, w = x.shape[-2:]
        x = F.pad(x, (0, (-w) % 2, 0, (-h) % 2), mode='replicate')
        x = F.pixel_unshuffle(x, self.sf)
        m = sigma.to(x.dtype).expand(x.shape[0], 1, x.shape[-2], x.shape[-1])
        x = self.model(torch.cat((x, m), 1))
        return F.pixel_shuffle(x, self.sf)[..., :h, :w]


class _ResBlock(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.res = nn.Sequential(nn.Conv2d(ch, ch, 3, 1, 1, bias=False), nn.ReLU(inplace=True),
                                 nn.Conv2d(ch, ch, 3, 1, 1, bias=False))

    def forward(self, x):
        return x + self.res(x)


class UNetRes(nn.Module):
    """DRUNet backbone (bias-free)."""

    def __init__(self, in_nc=2, out_nc=1, nc=(64, 128, 256, 512), nb=4, act_mode='R', downsample_mode='strideconv',
                 upsample_mode='convtranspose'):
        super().__init__()
        if (act_mode, downsample_mode, upsample_mode) != ('R', 'strideconv', 'convtranspose'):
            raise NotImplementedError('only the DRUNet configuration used by the reference is implemented')
        c = list(nc)
        blocks = lambda ch: [_ResBlock(ch) for _ in range(nb)]
        self.m_head = nn.Conv2d(in_nc, c[0], 3, 1, 1, bias=False)
        self.m_down1 = nn.Sequential(*blocks(c[0]), nn.Conv2d(c[0], c[1], 2, 2, 0, bias=False))
        self.m_down2 = nn.Sequential(*blocks(c[1]), nn.Conv2d(c[1], c[2], 2, 2, 0, bias=False))
        self.m_down3 = nn.Sequential(*blocks(c[2]), nn.Conv2d(c[2], c[3], 2, 2, 0, bias=False))
        self.m_body = nn.Sequential(*blocks(c[3]))
        self.m_up3 = nn.Sequential(nn.ConvTranspose2d(c[3], c[2], 2, 2, 0, bias=False), *blocks(c[2]))
        self.m_up2 = nn.Sequential(nn.ConvTranspose2d(c[2], c[1], 2, 2, 0, bias=False), *blocks(c[1]))
        self.m_up1 = nn.Sequential(nn.ConvTranspose2d(c[1], c[0], 2, 2, 0, bias=False), *blocks(c[0]))
        self.m_tail = nn.Conv2d(c[0], out_nc, 3, 1, 1, bias=False)

    def forward(self, x0):
        x1 = self.m_head(x0)
        x2 = self.m_down1(x1)
        x3 = self.m_down2(x2)
        x4 = self.m_down3(x3)
        x = self.m_body(x4)
        x = self.m_up3(x + x4)
        x = self.m_up2(x + x3)
        x = self.m_up1(x + x2)
        return self.m_tail(x + x1)


def count_params(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())


# ----------------------------------------------------------------------------------------------
# test-time wrappers on the DRUNet path
# ----------------------------------------------------------------------------------------------
def augment(x: torch.Tensor, mode: int) -> torch.Tensor:
    """The 8 dihedral transforms of utils_image.augment_img_tensor4 (utils_image.py:333-349)."""
    if mode == 0:
        return x
    k, flip = {1: (1, True), 2: (0, True), 3: (3, False), 4: (2, True), 5: (1, False), 6: (2, False), 7: (3, True)}[mode]
    y = x.rot90(k, [2, 3]) if k else x
    return y.flip([2]) if flip else y


def augment_inverse_mode(mode: int) -> int:
    """S3:46-50: modes 3 and 5 (the two pure 90-degree rotations) are undone by 8 - mode, the rest are involutions."""
    return 8 - mode if mode in (3, 5) else mode


def split_forward(model: Callable, x: torch.Tensor, refield=32, min_size=256, modulo=16) -> torch.Tensor:
    """utils_model.test_split_fn (utils_model.py:76-109), sf = 1: one padded pass up to min_size^2 pixels,
    otherwise four overlapping quadrants of (h//2//refield + 1) * refield pixels stitched at h//2, w//2."""
    h, w = x.shape[-2:]
    if h * w <= min_size ** 2:
        xp = F.pad(x, (0, (-w) % modulo, 0, (-h) % modulo), mode='replicate')
        return model(xp)[..., :h, :w]
    ph = (h // 2 // refield + 1) * refield
    pw = (w // 2 // refield + 1) * refield
    tiles = [x[..., :ph, :pw], x[..., :ph, w - pw:], x[..., h - ph:, :pw], x[..., h - ph:, w - pw:]]
    if h * w <= 4 * min_size ** 2:
        # the four quadrants share a shape: one batched forward instead of four
        b = x.shape[0]
        e = model(torch.cat(tiles, 0))
        es = [e[i * b:(i + 1) * b] for i in range(4)]
    else:
        es = [split_forward(model, t, refield, min_size, modulo) for t in tiles]
    out = x.new_zeros((x.shape[0], es[0].shape[1], h, w))
    out[..., :h // 2, :w // 2] = es[0][..., :h // 2, :w // 2]
    out[..., :h // 2, w // 2:] = es[1][..., :h // 2, (-w + w // 2):]
    out[..., h // 2:, :w // 2] = es[2][..., (-h + h // 2):, :w // 2]
    out[..., h // 2:, w // 2:] = es[3][..., (-h + h // 2):, (-w + w // 2):]
    return out


def get_rho_sigma(sigma=2.55 / 255, iter_num=15, modelSigma1=49.0, modelSigma2=2.55, w=1.0):
    """DPIR schedule (utils/utils_pnp.py:14-23): log-space from modelSigma1 to modelSigma2, / 255."""
    s_log = np.logspace(np.log10(modelSigma1), np.log10(modelSigma2), iter_num).astype(np.float32)
    s_lin = np.linspace(modelSigma1, modelSigma2, iter_num).astype(np.float32)
    sigmas = (s_log * w + s_lin * (1 - w)) / 255.
    rhos = [0.23 * (sigma ** 2) / (float(s) ** 2) for s in sigmas]
    return rhos, sigmas


# ----------------------------------------------------------------------------------------------
# model construction + reference dispatch
# ----------------------------------------------------------------------------------------------
def _arch(model_name: str) -> str:
    if 'fdncnn' in model_name:
        return 'fdncnn'
    if 'dncnn' in model_name:
        return 'dncnn'
    for k in ('drunet', 'ircnn', 'ffdnet'):
        if k in model_name:
            return k
    raise ValueError(f'unknown denoiser {model_name!r} (dncnn_*, fdncnn_*, drunet_*, ircnn_*, ffdnet_* are on the path; '
                     f'BM3D is out of scope)')


def build_model(model_name: str, seed: int = 0, weights=None) -> nn.Module:
    """Instantiate the architecture the reference builds for `model_name` (S3:122-209 / S6:128-218)."""
    arch = _arch(model_name)
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        if arch == 'dncnn':
            nb = 20 if model_name in ('dncnn_gray_blind', 'dncnn_color_blind', 'dncnn3') else 17      # S3:127-130
            m = DnCNN(1, 1, 64, nb, 'R')
        elif arch == 'fdncnn':
            m = FDnCNN(2, 1, 64, 20, 'R')                                                             # S3:150
        elif arch == 'drunet':
            m = UNetRes(2, 1, (64, 128, 256, 512), 4, 'R', 'strideconv', 'convtranspose')             # S3:168-169
        elif arch == 'ircnn':
            m = IRCNN(1, 1, 64)                                                                       # S3:187
        else:
            m = FFDNet(1, 1, 64, 15, 'R')                                                             # S3:203
    finally:
        torch.random.set_rng_state(gen_state)
    if weights is not None:
        sd = torch.load(weights, map_location='cpu') if isinstance(weights, (str, bytes)) else weights
        m.load_state_dict(sd, strict=True)
    m.eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m


class Denoiser:
    """Batched callable D(x, i) with the reference's per-model pre/post-processing.

    x : (B,1,H,W) float32 CUDA tensor; returns float32.  The network itself runs in `dtype`
    (bf16 channels_last by default; float32 for parity tests).
    """

    def __init__(self, model_name: str, iter_num: int = 50, x8: bool = False, noises=None, dtype=torch.bfloat16,
                 device='cuda', seed: int = 0, weights=None, ircnn_weights: Optional[Sequence] = None, model: Optional[nn.Module] = None,
                 fused: Optional[bool] = None):
        self.name = model_name
        self.arch = _arch(model_name)
        self.dtype = dtype
        self.device = torch.device(device)
        self.x8 = bool(x8) and self.arch in ('drunet', 'ircnn')       # only these branches look at x8 (S3:39-62)
        if self.arch == 'ircnn' and ircnn_weights is None and weights is not None:
            # KAIR's ircnn_gray.pth is ONE file holding 25 state-dicts keyed '0'..'24' (S3:187-189 `model25`)
            sd = torch.load(weights, map_location='cpu') if isinstance(weights, (str, bytes)) else weights
            if isinstance(sd, dict) and '0' in sd and isinstance(sd['0'], dict):
                ircnn_weights, weights = sd, None
        net = model if model is not None else build_model(model_name, seed, weights)
        # DnCNN / FDnCNN / IRCNN / FFDNet (all 64-channel conv3x3 chains; IRCNN's are dilated) in bf16 on a GPU run on the
        # hand-written tensor-core kernels (csrc/dncnn_tc.cuh) unless fused=False asks for the stock PyTorch module (the A/B
        # baseline); DRUNet stays in PyTorch.
        can_fuse = self.arch in ('dncnn', 'fdncnn', 'ircnn', 'ffdnet') and dtype == torch.bfloat16 and self.device.type == 'cuda'
        if fused and not can_fuse:
            raise ValueError('fused=True needs a DnCNN / FDnCNN / IRCNN / FFDNet in bf16 on a CUDA device')
        self.fused = None
        if can_fuse and fused is not False:
            from .dncnn_fused import FusedDnCNN, FusedFFDNet
            self.fused = (FusedFFDNet(net, device=self.device) if self.arch == 'ffdnet'
                          else FusedDnCNN(net, residual=(self.arch in ('dncnn', 'ircnn')), device=self.device))
        self.net = net.to(self.device, dtype).to(memory_format=torch.channels_last)
        self.sigmas = None
        self.noise_map = None
        self.ircnn_weights = ircnn_weights
        self._ircnn_idx = 0
        if self.arch in ('drunet', 'ircnn'):
            # S3:162-165: sigma = max(0.255/255, 15/255), modelSigma1 = 49, modelSigma2 = 15
            _, s = get_rho_sigma(sigma=max(0.255 / 255., 15 / 255.), iter_num=iter_num, modelSigma1=49, modelSigma2=15.0, w=1.0)
            self.sigmas = torch.tensor(s, device=self.device)
        if self.arch == 'fdncnn':
            if noises is None:
                raise ValueError('fdncnn needs the k-space noise array (its noise-level map is |noises| / 255, S3:27-30)')
            nm = torch.as_tensor(np.absolute(np.asarray(noises))).float() / 255.
            self.noise_map = nm.to(self.device)[None, None]
        self.ffdnet_sigma = torch.full((1, 1, 1, 1), 15 / 255., device=self.device)                 # S3:64

    def _run(self, x, *extra):
        xin = x.to(self.dtype).contiguous(memory_format=torch.channels_last)
        if self.dtype == torch.float32 and xin.is_cuda:
            # a float32 denoiser is the parity configuration: keep cuDNN off the TF32 path
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                return self.net(xin, *extra).float()
        return self.net(xin, *extra).float()

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, i: int = 0) -> torch.Tensor:
        if self.arch == 'dncnn':
            return self.fused(x) if self.fused is not None else self._run(x)                        # S3:20-22
        if self.arch == 'fdncnn':
            nm = self.noise_map.expand(x.shape[0], 1, *x.shape[-2:])
            xin = torch.cat((x, nm), 1)
            return self.fused(xin) if self.fused is not None else self._run(xin)                    # S3:26-35
        if self.arch == 'ffdnet':
            if self.fused is not None:
                return self.fused(x, 15 / 255.)                                                     # S3:64-66
            return self._run(x, self.ffdnet_sigma.to(self.dtype))                                   # S3:64-66
        mode = i % 8 if self.x8 else 0
        if mode:
            x = augment(x, mode)                                                                    # S3:40-41
        if self.arch == 'drunet':
            smap = self.sigmas[i].float().expand(x.shape[0], 1, *x.shape[-2:])
            y = split_forward(lambda t: self._run(t), torch.cat((x, smap), 1), 32, 256, 16)         # S3:43-44
        else:
            if self.ircnn_weights is not None:                                                      # S3:280-288
                idx = int(math.ceil(float(self.sigmas[i]) * 255. / 2.) - 1)
                if idx != self._ircnn_idx:                                                          # `former_idx`, starts at 0
                    sets = self.ircnn_weights
                    sd = sets[str(idx)] if isinstance(sets, dict) and str(idx) in sets else sets[idx]
                    self.net.load_state_dict(sd, strict=True)
                    if self.fused is not None:
                        self.fused.repack(self.net)
                    self._ircnn_idx = idx
            y = self.fused(x) if self.fused is not None else self._run(x)                           # S3:56
        if mode:
            y = augment(y, augment_inverse_mode(mode))                                              # S3:46-50
        return y


def build_denoiser(model_name: str, **kw) -> Denoiser:
    return Denoiser(model_name, **kw)
