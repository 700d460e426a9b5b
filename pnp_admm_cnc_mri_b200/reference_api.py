"""Drop-in entry points with the reference's own names and call signatures.

    out        = ADMM_L1(mask, noises, iter_num=, lambda1=, reo=)                      S1:29
    out        = ADMM_CNC(mask, noises, alpha=, iter_num=, lambda1=, reo=, b=)         S4:31
    out        = PNP_ADMM_L1_D(model_name, mask, noises, iter_num=, reo=)              S3:77
    out, psnr1 = PNP_ADMM_CNC_D(model_name, mask, noises, alpha=, ..., b=)             S6:79
    out, psnr1 = PNP_ADMM_CNC_DnCNN(model_name1, model_name2, mask, noises, **opts)    S6:372
    soft(x, c);  analyze_parse_*(defaults...)   (CLI flags --iter_num --lambda1 --reo --alpha --b)

Like the reference, each function walks ``testsets/<testset_name>`` (default ``Set1``) under the current
directory, reconstructs every image with the given mask / noise, writes ``results/<name>/*.png`` and an
append-mode log, and returns the 22-slot list ``out`` whose first n entries are the reconstructions
(S1:44-45,132).  Differences, all additive: the images are reconstructed as ONE batch on the GPU;
keyword-only extras (``testset_name``, ``testsets``, ``results``, ``save_E``, ``images``, ``image_names``,
``model_zoo``, ``denoiser_dtype``, ``dtype``, ``device``) override what the reference hard-codes; when
``model_zoo/<name>.pth`` is absent the network is random-initialised (the KAIR checkpoints are not redistributable).
``dtype='float64'`` runs ADMM_L1 / ADMM_CNC on the fp64 validation build (the reference loop is float64; the default
float32 kernels agree with it to <= 1e-4 relative).  PSNR / SSIM / RE are evaluated on the device
(``pnpadmm_metrics_*``, the reference's formulas in double precision), one call per batch.
"""
from __future__ import annotations

import argparse
import logging
import os
from collections import OrderedDict
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import metrics
from .denoisers import Denoiser, build_model, count_params
from .pnp import pnp_admm_cnc, pnp_admm_l1
from .solver import admm_solve, image_metrics
from .solver import soft as _soft_cuda

IMG_EXTENSIONS = ('.jpg', '.JPG', '.jpeg', '.JPEG', '.png', '.PNG', '.ppm', '.PPM', '.bmp', '.BMP', '.tif')
_EXTRA = ('testset_name', 'testsets', 'results', 'save_E', 'images', 'image_names', 'model_zoo', 'denoiser_dtype',
          'fix_model2', 'seed', 'dtype', 'device')


# ----------------------------------------------------------------------------------------------
# a1
# ----------------------------------------------------------------------------------------------
def soft(x, c):
    """``np.fmax(np.fabs(x) - c, 0) * np.sign(x)`` (S1:18-19), evaluated by the CUDA kernel."""
    if isinstance(x, torch.Tensor):
        return _soft_cuda(x.contiguous(), c)
    a = np.asarray(x)
    dt = torch.float64 if a.dtype == np.float64 else torch.float32
    return _soft_cuda(torch.as_tensor(a).to('cuda', dt).contiguous(), c).cpu().numpy()


# ----------------------------------------------------------------------------------------------
# CLI parsers (S1:21-27, S4:21-29, S3:70-75, S6:69-77, S6:361-370)
# ----------------------------------------------------------------------------------------------
def _parser(alpha=None, iter_num=None, lambda1=None, reo=None, b=None):
    p = argparse.ArgumentParser()
    if alpha is not None:
        p.add_argument('--alpha', type=float, default=alpha, help='Step size in Plug-and Play')
    p.add_argument('--iter_num', type=int, default=iter_num, help='Number of iterations')
    if lambda1 is not None:
        p.add_argument('--lambda1', type=float, default=lambda1, help='regularization parameter')
    p.add_argument('--reo', type=float, default=reo, help='Lagrange parameter')
    if b is not None:
        p.add_argument('--b', type=float, default=b, help='convex parameter')
    return p


def analyze_parse_ADMM_L1(default_iter_num, default_lambda1, default_reo, argv=None):
    return _parser(None, default_iter_num, default_lambda1, default_reo).parse_args(argv)


def analyze_parse_ADMM_CNC(default_alpha, default_iter_num, default_lambda1, default_reo, default_b, argv=None):
    return _parser(default_alpha, default_iter_num, default_lambda1, default_reo, default_b).parse_args(argv)


def analyze_parse_PNP_ADMM_L1_D(default_iter_num, default_reo, argv=None):
    return _parser(None, default_iter_num, None, default_reo).parse_args(argv)


def analyze_parse_PNP_ADMM_CNC_D(default_alpha, default_iter_num, default_lambda1, default_reo, default_b, argv=None):
    return _parser(default_alpha, default_iter_num, default_lambda1, default_reo, default_b).parse_args(argv)


analyze_parse_PNP_ADMM_CNC_DnCNN = analyze_parse_PNP_ADMM_CNC_D

# call-site presets of the reference drivers (S1:171, S4:176, S3:339-347, S6:569-577)
PRESETS = {
    'ADMM_L1': dict(iter_num=50, lambda1=0.1, reo=0.015),
    'ADMM_CNC': dict(alpha=0.45, iter_num=50, lambda1=0.5, reo=0.05, b=64),
    'PNP_ADMM_L1_D': {'fdncnn_gray': dict(iter_num=50, reo=0.25), 'dncnn_15': dict(iter_num=50, reo=0.15),
                      'ffdnet_gray': dict(iter_num=50, reo=0.25), 'ircnn_gray': dict(iter_num=50, reo=0.145),
                      'drunet_gray': dict(iter_num=50, reo=0.26)},
    'PNP_ADMM_CNC_D': {'fdncnn_gray': dict(alpha=0.9, iter_num=50, lambda1=0.2, reo=0.45, b=0.3),
                       'ffdnet_gray': dict(alpha=0.9, iter_num=50, lambda1=1.35, reo=0.45, b=0.3),
                       'ircnn_gray': dict(alpha=0.5, iter_num=50, lambda1=1.3, reo=0.45, b=2),
                       'drunet_gray': dict(alpha=1, iter_num=50, lambda1=0.8, reo=0.8, b=0.45)},
    'PNP_ADMM_CNC_DnCNN': dict(alpha=1.2, iter_num=50, lambda1=4, reo=0.45, b=0.3),
}


# ----------------------------------------------------------------------------------------------
# harness pieces shared by the five entry points
# ----------------------------------------------------------------------------------------------
def _logger(name: str, log_path: str) -> logging.Logger:
    """File (append) + stream handlers, format of utils_logger.logger_info (utils_logger.py:25-44)."""
    log = logging.getLogger(name)
    if not log.handlers:
        fmt = logging.Formatter('%(asctime)s.%(msecs)03d : %(message)s', datefmt='%y-%m-%d %H:%M:%S')
        fh = logging.FileHandler(log_path, mode='a')
        fh.setFormatter(fmt)
        sh = logging.StreamHandler()
        sh.setFormatter(fmt)
        log.setLevel(logging.INFO)
        log.addHandler(fh)
        log.addHandler(sh)
    return log


def get_image_paths(dataroot: str) -> List[str]:
    """Sorted recursive listing (utils_image.py:66-82)."""
    if not os.path.isdir(dataroot):
        raise AssertionError('{:s} is not a valid directory'.format(dataroot))
    out = []
    for dp, _, fns in sorted(os.walk(dataroot)):
        for fn in sorted(fns):
            if fn.endswith(IMG_EXTENSIONS):
                out.append(os.path.join(dp, fn))
    if not out:
        raise AssertionError('{:s} has no valid image file'.format(dataroot))
    return sorted(out)


def _load_images(extras):
    """-> (img_H list of uint8 (H,W), names with extension). S1:83-86: imread gray, modcrop(8)."""
    if extras.get('images') is not None:
        imgs = [np.asarray(a) for a in extras['images']]
        names = list(extras.get('image_names') or ['%02d.png' % (i + 1) for i in range(len(imgs))])
        L_path = '<arrays>'
    else:
        import cv2
        L_path = os.path.join(extras.get('testsets', 'testsets'), extras.get('testset_name', 'Set1'))
        paths = get_image_paths(L_path)
        imgs = [cv2.imread(p, 0) for p in paths]                         # utils_image.py:149
        names = [os.path.basename(p) for p in paths]
    out = []
    for a in imgs:
        if a.dtype != np.uint8 or a.ndim != 2:
            raise ValueError('images must be 2-D uint8 gray-level arrays')
        H, W = a.shape
        out.append(a[:H - H % 8, :W - W % 8])                           # modcrop(img, 8)
    return out, names, L_path


def _split_opts(opts):
    extras = {k: opts[k] for k in _EXTRA if k in opts}
    return extras


class _Run:
    """Bookkeeping the five reference functions share: paths, logger, metrics, out list."""

    def __init__(self, result_tag: str, extras: dict, slots: int = 22):
        self.testset_name = extras.get('testset_name', 'Set1')
        self.result_name = self.testset_name + '_dn_' + result_tag
        self.E_path = os.path.join(extras.get('results', 'results'), self.result_name)
        os.makedirs(self.E_path, exist_ok=True)
        self.logger = _logger(self.result_name, os.path.join(self.E_path, self.result_name + '.log'))
        self.save_E = extras.get('save_E', True)
        A = np.zeros((256, 256), dtype='uint8')
        self.out = [A] * slots                                           # S1:44-45
        self.psnr1 = [0] * 22
        self.res = OrderedDict(psnr=[], ssim=[], re=[])
        self.scores = None

    def check_slots(self, n_images: int):
        """The reference raises IndexError at `out[n] = x` for the image past the last slot (S1:132), after having
        reconstructed the ones before; a batch is reconstructed at once, so fail before launching it."""
        if n_images > len(self.out):
            raise IndexError(f'list assignment index out of range: {n_images} images but the reference\'s `out` list has '
                             f'{len(self.out)} slots (S1:44-45)')

    def score(self, x: torch.Tensor, img_H, as_uint8: bool):
        """PSNR / SSIM / RE of the whole batch on the device (utils_image.py:543-636 in double precision;
        img_E = 255 x, S1:133, or uint8(round(255 x)), S6:315): one (B,3) array comes back."""
        ref = torch.as_tensor(np.stack(img_H)).to(x.device)
        self.scores = image_metrics(x, ref, quantize=as_uint8).cpu().numpy()

    def record(self, n, x, img_H, name, suffix, as_uint8: bool, psnr_fmt: str):
        self.out[n] = x
        img_E = np.uint8((x * 255.0).round()) if as_uint8 else x * 255   # S6:315 | S1:133
        if self.save_E:
            import cv2
            cv2.imwrite(os.path.join(self.E_path, os.path.splitext(name)[0] + suffix), np.squeeze(img_E))
        p, s, r = (float(v) for v in self.scores[n])
        for k, v in zip(('psnr', 'ssim', 're'), (p, s, r)):
            self.res[k].append(v)
        self.logger.info(('{:s} - PSNR: ' + psnr_fmt + ' dB; SSIM: {:.4f} ; RE: {:.4f}.').format(name, p, s, r))
        self.psnr1[n] = p

    def averages(self, alpha=None):
        ap, as_, ar = (sum(v) / len(v) for v in self.res.values())
        if alpha is None:
            self.logger.info('------> testset_name: ({}), Average PSNR:({:.3f})dB, Average ssim : ({:.3f}), Average re : ({:.3f}) )'
                             .format(self.testset_name, ap, as_, ar))
        else:
            self.logger.info('------> testset_name: ({}), alpha: ({:.3f}), Average PSNR:({:.3f})dB, Average ssim : ({:.3f}), '
                             'Average re : ({:.3f}) )'.format(self.testset_name, alpha, ap, as_, ar))
        return ap, as_, ar


def _stack(img_H):
    shapes = {a.shape for a in img_H}
    if len(shapes) != 1:
        raise ValueError('all images of a test set must share one shape to be reconstructed as a batch')
    return np.stack([np.float32(a / 255.) for a in img_H])              # uint2single (utils_image.py:181)


def _zero_fill_print(img_L, mask, noises):
    """S1:99-101 prints the zero-filling PSNR of each image (host NumPy, outside the timed path)."""
    for a in img_L:
        y = np.fft.fft2(a) * mask + noises
        print('zero-filling psnr = %.4f' % metrics.psnr(np.fft.ifft2(y) * 255, a * 255))


# ----------------------------------------------------------------------------------------------
# S1 / S4
# ----------------------------------------------------------------------------------------------
def ADMM_L1(mask, noises, **ADMM_L1_opts):
    iter_num = ADMM_L1_opts.get('iter_num', 20)                          # S1:35-37 fallbacks
    lambda1 = ADMM_L1_opts.get('lambda1', 0.04)
    reo = ADMM_L1_opts.get('reo', 0.04)
    extras = _split_opts(ADMM_L1_opts)
    run = _Run('ADMM_L1', extras)
    img_H, names, L_path = _load_images(extras)
    run.logger.info(L_path)
    run.check_slots(len(names))
    img_L = _stack(img_H)
    _zero_fill_print(img_L, mask, noises)                                # S1:99-101
    xd = admm_solve(torch.as_tensor(img_L), mask, noises, prox='l1', iter_num=iter_num, lambda1=lambda1, reo=reo,
                    dtype=extras.get('dtype', 'float32'), device=extras.get('device'))
    run.score(xd, img_H, False)
    x = xd.cpu().numpy()
    for n, name in enumerate(names):
        run.record(n, x[n].astype(np.float64), img_H[n], name, '_PDG L1.png', False, '{:.2f}')
    run.averages()
    return run.out


def ADMM_CNC(mask, noises, **ADMM_CNC_opts):
    iter_num = ADMM_CNC_opts.get('iter_num', 4)                          # S4:37-41 fallbacks
    alpha = ADMM_CNC_opts.get('alpha', 0.4)
    lambda1 = ADMM_CNC_opts.get('lambda1', 0.04)
    reo = ADMM_CNC_opts.get('reo', 2.75)
    b = ADMM_CNC_opts.get('b', 1)
    extras = _split_opts(ADMM_CNC_opts)
    run = _Run('ADMM_CNC', extras)
    img_H, names, L_path = _load_images(extras)
    run.logger.info(L_path)
    run.check_slots(len(names))
    img_L = _stack(img_H)
    _zero_fill_print(img_L, mask, noises)                                # S4:103-105
    xd = admm_solve(torch.as_tensor(img_L), mask, noises, prox='cnc', iter_num=iter_num, lambda1=lambda1, reo=reo, alpha=alpha,
                    b=b, dtype=extras.get('dtype', 'float32'), device=extras.get('device'))
    run.score(xd, img_H, False)
    x = xd.cpu().numpy()
    for n, name in enumerate(names):
        run.record(n, x[n].astype(np.float64), img_H[n], name, '_ADMM CNC.png', False, '{:.4f}')
    run.averages()
    return run.out


# ----------------------------------------------------------------------------------------------
# S3 / S6
# ----------------------------------------------------------------------------------------------
def _denoiser(model_name, iter_num, x8, noises, extras, logger, weights_name=None):
    zoo = extras.get('model_zoo', 'model_zoo')
    path = os.path.join(zoo, (weights_name or model_name) + '.pth')
    weights = path if os.path.exists(path) else None
    ircnn_weights = None
    if 'ircnn' in model_name and weights is not None:
        sd25 = torch.load(path, map_location='cpu')                      # 25 state dicts keyed '0'..'24' (S3:188)
        ircnn_weights = [sd25[str(i)] for i in range(len(sd25))]
        weights = ircnn_weights[0]
    model = build_model(model_name, seed=extras.get('seed', 0), weights=weights)
    logger.info('Model path: {:s}{}'.format(path, '' if os.path.exists(path) else '  (absent: random-init weights)'))
    logger.info('Params number: {}'.format(count_params(model)))
    return Denoiser(model_name, iter_num=iter_num, x8=x8, noises=noises, model=model, ircnn_weights=ircnn_weights,
                    dtype=extras.get('denoiser_dtype', torch.bfloat16), device=extras.get('device') or 'cuda')


def PNP_ADMM_L1_D(model_name, mask, noises, **PNP_ADMM_L1_D_opts):
    iter_num = PNP_ADMM_L1_D_opts.get('iter_num', 20)                    # S3:83-84
    reo = PNP_ADMM_L1_D_opts.get('reo', 0.04)
    extras = _split_opts(PNP_ADMM_L1_D_opts)
    run = _Run(model_name, extras)
    x8 = 'drunet' in model_name          # S3:87 sets True; the dncnn / fdncnn / ircnn branches reset it (S3:130,142,181)
    D = _denoiser(model_name, iter_num, x8, noises, extras, run.logger)
    img_H, names, L_path = _load_images(extras)
    run.check_slots(len(names))
    img_L = _stack(img_H)
    _zero_fill_print(img_L, mask, noises)                                # S3:241-243
    xd = pnp_admm_l1(torch.as_tensor(img_L), mask, noises, D, iter_num=iter_num, reo=reo, device=extras.get('device'))
    run.score(xd, img_H, False)
    x = xd.cpu().numpy()
    for n, name in enumerate(names):
        run.record(n, x[n], img_H[n], name, '_' + model_name + '_PNP_ADMM_L1_D.png', False, '{:.2f}')
    run.averages()
    return run.out


def PNP_ADMM_CNC_D(model_name, mask, noises, **PNP_ADMM_CNC_D_opts):
    alpha = PNP_ADMM_CNC_D_opts.get('alpha', 0.4)                        # S6:85-89
    iter_num = PNP_ADMM_CNC_D_opts.get('iter_num', 46)
    lambda1 = PNP_ADMM_CNC_D_opts.get('lambda1', 2.75)
    reo = PNP_ADMM_CNC_D_opts.get('reo', 1)
    b = PNP_ADMM_CNC_D_opts.get('b', 1)
    extras = _split_opts(PNP_ADMM_CNC_D_opts)
    run = _Run(model_name, extras)
    D = _denoiser(model_name, iter_num, False, noises, extras, run.logger)   # x8 = False (S6:93)
    img_H, names, L_path = _load_images(extras)
    run.check_slots(len(names))
    img_L = _stack(img_H)
    xd = pnp_admm_cnc(torch.as_tensor(img_L), mask, noises, D, None, alpha=alpha, iter_num=iter_num, lambda1=lambda1, reo=reo, b=b,
                      device=extras.get('device'))
    run.score(xd, img_H, True)
    x = xd.cpu().numpy()
    for n, name in enumerate(names):
        run.record(n, x[n], img_H[n], name, 'PNP_ADMM_CNC_D.png', True, '{:.4f}')
    run.averages(alpha)
    return run.out, run.psnr1


def PNP_ADMM_CNC_DnCNN(model_name1, model_name2, mask, noises, **PNP_ADMM_CNC_DnCNN_opts):
    alpha = PNP_ADMM_CNC_DnCNN_opts.get('alpha', 0.4)                    # S6:378-382
    iter_num = PNP_ADMM_CNC_DnCNN_opts.get('iter_num', 46)
    lambda1 = PNP_ADMM_CNC_DnCNN_opts.get('lambda1', 2.75)
    reo = PNP_ADMM_CNC_DnCNN_opts.get('reo', 1)
    b = PNP_ADMM_CNC_DnCNN_opts.get('b', 1)
    extras = _split_opts(PNP_ADMM_CNC_DnCNN_opts)
    run = _Run(model_name1 + '_' + model_name2, extras, slots=21)        # S6:392 has 21 slots
    D1 = _denoiser(model_name1, iter_num, False, noises, extras, run.logger)
    # The reference loads model_path1 into model2 as well (S6:435), so both denoisers share weights;
    # fix_model2=True loads model_zoo/<model_name2>.pth instead.
    w2 = model_name2 if extras.get('fix_model2', False) else model_name1
    D2 = _denoiser(model_name2, iter_num, False, noises, extras, run.logger, weights_name=w2)
    img_H, names, L_path = _load_images(extras)
    run.logger.info(L_path)
    run.check_slots(len(names))
    img_L = _stack(img_H)
    _zero_fill_print(img_L, mask, noises)                                # S6:479-481
    xd = pnp_admm_cnc(torch.as_tensor(img_L), mask, noises, D1, D2, alpha=alpha, iter_num=iter_num, lambda1=lambda1, reo=reo, b=b,
                      device=extras.get('device'))
    run.score(xd, img_H, True)
    x = xd.cpu().numpy()
    for n, name in enumerate(names):
        run.record(n, x[n], img_H[n], name, 'PNP_ADMM_CNC_DnCNN.png', True, '{:.4f}')
    run.averages(alpha)
    return run.out, run.psnr1


def load_cs_mri(root: str = 'CS_MRI'):
    """The reference drivers' input block (S1:177-186): three masks 'Q1' as float64 + noises x 3.0."""
    import scipy.io as sio
    mask = np.array([sio.loadmat(os.path.join(root, f + '.mat')).get('Q1').astype(np.float64)
                     for f in ('Q_Random30', 'Q_Radial30', 'Q_Cartesian30')])
    noises = sio.loadmat(os.path.join(root, 'noises.mat')).get('noises').astype(np.complex128) * 3.0
    return mask, noises
