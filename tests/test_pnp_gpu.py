"""GPU: PnP variants (rows a7-a10 of the scope table) and the drop-in entry points."""
import os

import numpy as np
import pytest
import torch

from oracle import kat_table as kat
from oracle import reference_numpy as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLD, 'pnp_golden.npz'))


@pytest.fixture(scope='module')
def env(cs_inputs):
    import pnp_admm_cnc_mri_b200 as pk
    pk.load_library()
    img = orc.preprocess_uint8(cs_inputs['images'][4])
    return pk, img, cs_inputs['masks'][0], cs_inputs['noises']


def _D(name, it, x8, nz, dtype=torch.float32):
    from pnp_admm_cnc_mri_b200.denoisers import Denoiser
    return Denoiser(name, iter_num=it, x8=x8, noises=nz, dtype=dtype, seed=0)


def test_pnp_cnc_dncnn_vs_unmodified_reference(env, gold):
    """PNP_ADMM_CNC_DnCNN (S6:372) with the same fp32 weights as the unmodified script run."""
    pk, img, m, nz = env
    from pnp_admm_cnc_mri_b200.pnp import pnp_admm_cnc
    it = int(gold['iters'])
    D = _D('dncnn_25', it, False, nz)
    x = pnp_admm_cnc(img, m, nz, D, D, alpha=1.2, iter_num=it, lambda1=4, reo=0.45, b=0.3)
    assert rel(x, gold['cnc_dncnn']) < 1e-4
    assert np.abs(x - gold['cnc_dncnn']).max() < 1e-3


def test_pnp_cnc_drunet_vs_unmodified_reference(env, gold):
    pk, img, m, nz = env
    from pnp_admm_cnc_mri_b200.pnp import pnp_admm_cnc
    it = int(gold['iters'])
    x = pnp_admm_cnc(img, m, nz, _D('drunet_gray', it, False, nz), None, alpha=1, iter_num=it, lambda1=0.8, reo=0.8, b=0.45)
    assert rel(x, gold['cnc_drunet']) < 1e-4


def test_pnp_l1_drunet_x8_vs_unmodified_reference(env, gold):
    pk, img, m, nz = env
    from pnp_admm_cnc_mri_b200.pnp import pnp_admm_l1
    it = int(gold['iters'])
    x = pnp_admm_l1(img, m, nz, _D('drunet_gray', it, True, nz), iter_num=it, reo=0.26)
    assert rel(x, gold['l1_drunet']) < 1e-4


@pytest.mark.parametrize('name,x8', [('ffdnet_gray', False), ('fdncnn_gray', False), ('ircnn_gray', False), ('dncnn_15', False)])
def test_pnp_l1_other_denoisers_vs_oracle(env, name, x8):
    """Every denoiser branch of denoising_step1 (S3:19-68), 4 iterations, batch of 3 vs the oracle per image."""
    pk, img, m, nz = env
    from pnp_admm_cnc_mri_b200.pnp import pnp_admm_l1
    from pnp_admm_cnc_mri_b200.denoisers import Denoiser
    it = 4
    imgs = np.stack([img, img[::-1].copy(), img.T.copy()])
    x = pnp_admm_l1(imgs, m, nz, _D(name, it, x8, nz), iter_num=it, reo=0.25)
    Dc = Denoiser(name, iter_num=it, x8=x8, noises=nz, dtype=torch.float32, device='cpu', seed=0)
    f = lambda a, i: Dc(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))[None, None], i)[0, 0].numpy()
    for k in range(3):
        xr = orc.pnp_admm_l1(imgs[k], m.astype(np.float64), nz, f, iter_num=it, reo=0.25)
        assert rel(x[k], xr) < 1e-4, (name, k)


def test_pnp_512_quadrant_tiling(env):
    """Config 4 shape: DRUNet on 512x512 goes through the 4 x 288^2 quadrant path (utils_model.py:91-108)."""
    pk, *_ = env
    from pnp_admm_cnc_mri_b200 import data
    from pnp_admm_cnc_mri_b200.pnp import pnp_admm_l1
    from pnp_admm_cnc_mri_b200.denoisers import Denoiser
    N, it = 512, 2
    imgs = data.phantoms(2, N, seed0=5)
    m = data.make_mask('random', N, seed=3)
    nz = data.make_noise(N, seed=4)
    x = pnp_admm_l1(imgs, m, nz, _D('drunet_gray', it, True, nz), iter_num=it, reo=0.26)
    Dc = Denoiser('drunet_gray', iter_num=it, x8=True, dtype=torch.float32, device='cpu', seed=0)
    f = lambda a, i: Dc(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))[None, None], i)[0, 0].numpy()
    xr = orc.pnp_admm_l1(imgs[0], m.astype(np.float64), nz, f, iter_num=it, reo=0.26)
    assert rel(x[0], xr) < 1e-4


def test_bf16_denoiser_single_call_error(env):
    """bf16 tensor-core denoisers vs the same fp32 network: reported as single-call relative error."""
    pk, img, m, nz = env
    x = torch.as_tensor(img, device='cuda')[None, None].repeat(4, 1, 1, 1)
    for name in ('dncnn_25', 'drunet_gray', 'ffdnet_gray'):
        a = _D(name, 50, False, nz, torch.float32)(x, 3)
        b = _D(name, 50, False, nz, torch.bfloat16)(x, 3)
        err = float((a - b).norm() / a.norm())
        print(f'{name}: bf16 vs fp32 single-call rel error {err:.2e}')
        assert err < 2e-2


def test_drop_in_entry_points(env, cs_inputs, ref_out, tmp_path, monkeypatch, capsys):
    """Same names / kwargs / return contract as the reference functions (S1:29, S4:31, S6:372)."""
    pk, img, m, nz = env
    from pnp_admm_cnc_mri_b200 import reference_api as api
    monkeypatch.chdir(tmp_path)
    os.makedirs('testsets/Set1')
    import cv2
    cv2.imwrite('testsets/Set1/05.png', cs_inputs['images'][4])
    mask64 = m.astype(np.float64)
    out = api.ADMM_L1(mask64, nz, iter_num=50, lambda1=0.1, reo=0.015)
    assert 'zero-filling psnr = 20.6451' in capsys.readouterr().out           # S1:101 prints it per image
    assert isinstance(out, list) and len(out) == 22 and out[1].dtype == np.uint8 and out[1].shape == (256, 256)
    assert out[0].dtype == np.float64 and rel(out[0], ref_out['l1']) < 1e-4
    assert os.path.exists('results/Set1_dn_ADMM_L1/05_PDG L1.png')
    log = open('results/Set1_dn_ADMM_L1/Set1_dn_ADMM_L1.log').read()
    assert '05.png - PSNR: 23.87 dB; SSIM: 0.5877 ; RE: 0.2028.' in log           # the reference's own log line
    out = api.ADMM_CNC(mask64, nz, **api.PRESETS['ADMM_CNC'])
    assert rel(out[0], ref_out['cnc']) < 1e-4
    assert 'PSNR: 24.5765 dB; SSIM: 0.5600 ; RE: 0.1870.' in open('results/Set1_dn_ADMM_CNC/Set1_dn_ADMM_CNC.log').read()
    # function-level fallbacks when kwargs are omitted (S4:37-41): alpha .4, 4 iterations, lambda .04, reo 2.75, b 1
    out = api.ADMM_CNC(mask64, nz, save_E=False)
    assert rel(out[0], orc.admm_cnc(img, mask64, nz, 0.4, 4, 0.04, 2.75, 1)) < 1e-4
    out, psnr1 = api.PNP_ADMM_CNC_DnCNN('dncnn_25', 'dncnn_15', mask64, nz, iter_num=2, alpha=1.2, lambda1=4, reo=0.45, b=0.3,
                                        save_E=False)
    assert len(out) == 21 and len(psnr1) == 22 and out[0].dtype == np.float32 and psnr1[0] > 15
    out = api.PNP_ADMM_L1_D('drunet_gray', mask64, nz, iter_num=2, reo=0.26, save_E=False)
    assert len(out) == 22 and out[0].shape == (256, 256)
    # keyword-only extras: the fp64 validation build behind the same entry point, and the slot check before the batch is launched
    out = api.ADMM_L1(mask64, nz, iter_num=50, lambda1=0.1, reo=0.015, dtype='float64', save_E=False)
    assert rel(out[0], ref_out['l1']) < 1e-10
    with pytest.raises(IndexError):
        api.ADMM_CNC(mask64, nz, images=[cs_inputs['images'][0]] * 23, save_E=False)
    s = api.soft(np.array([-2.0, -0.5, 0.0, 0.5, 2.0]), 1.0)
    assert np.array_equal(s, orc.soft(np.array([-2.0, -0.5, 0.0, 0.5, 2.0]), 1.0))
    ns = api.analyze_parse_ADMM_CNC(0.45, 50, 0.5, 0.05, 64, argv=['--iter_num', '7'])
    assert (ns.alpha, ns.iter_num, ns.lambda1, ns.reo, ns.b) == (0.45, 7, 0.5, 0.05, 64)


@pytest.mark.gpu
def test_pnp_cnc_with_tensor_core_denoiser_tracks_fp32_denoiser():
    """PnP-ADMM-CNC (S6:491-525) with the DnCNN on the tcgen05 kernels (bf16 operands) against the same loop with the
    float32 PyTorch module and the same weights: the reconstruction stays within bf16-denoiser accuracy (the gate the
    design states for the bf16 denoiser: reported error, here bounded at 2e-2 relative after 5 iterations)."""
    import torch
    from pnp_admm_cnc_mri_b200 import data, denoisers, pnp
    N, B = 256, 3
    imgs = data.phantoms(B, N, seed0=5)
    mask = data.make_mask('radial', N, seed=1)
    noise = data.make_noise(N, seed=9)
    d16 = denoisers.build_denoiser('dncnn_25', seed=2)
    assert d16.fused is not None
    d32 = denoisers.build_denoiser('dncnn_25', seed=2, dtype=torch.float32)
    assert d32.fused is None
    P = dict(alpha=1.2, iter_num=5, lambda1=4.0, reo=0.45, b=0.3)
    x16 = pnp.pnp_admm_cnc(imgs, mask, noise, d16, d16, **P)
    x32 = pnp.pnp_admm_cnc(imgs, mask, noise, d32, d32, **P)
    for k in range(B):
        err = np.linalg.norm(x16[k] - x32[k]) / np.linalg.norm(x32[k])
        assert err < 2e-2, (k, err)


# ---------------------------------------------------------------------------------------------
# Full preset depth (50 iterations) against outputs of the UNMODIFIED scripts S3 / S6 with the same seeded weights
# (tests/golden/pnp_golden_50it.npz, oracle/make_golden_pnp50.py).  The denoiser runs in float32 on both sides; what
# differs is cuDNN on the GPU against oneDNN on the CPU (summation order, ~1e-6 per forward) and how far 50 iterations of
# a random-weight network amplify that.  The generator records the same sensitivity on the CPU alone (the restatement,
# driven by this package's Denoiser, against the script: `diff_vals`): residual DnCNN / FDnCNN / IRCNN / FFDNet-L1 loops
# contract (<= 2e-6), the random-weight DRUNet loops do not (max|diff| 0.1-0.8, rel-L2 ~ 8e-3), FFDNet-CNC sits in
# between (1.4e-3).  Gates follow that: 1e-4 rel-L2 where the loop is stable, a reported bound where it is not.
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def gold50():
    return np.load(os.path.join(GOLD, 'pnp_golden_50it.npz'))


def _img(cs_inputs, name):
    names = cs_inputs['image_names']
    return orc.preprocess_uint8(cs_inputs['images'][names.index(name)])


def _ircnn_sets(seed0):
    from pnp_admm_cnc_mri_b200 import denoisers as dn
    return {str(k): dn.build_model('ircnn_gray', seed=seed0 + k).state_dict() for k in range(25)}


STABLE_TOL, CHAOTIC_TOL = 1e-4, 5e-2


@pytest.mark.parametrize('key,name,tol', [('l1_dncnn', 'dncnn_15', STABLE_TOL), ('l1_fdncnn', 'fdncnn_gray', STABLE_TOL),
                                          ('l1_ffdnet', 'ffdnet_gray', STABLE_TOL), ('l1_ircnn', 'ircnn_gray', STABLE_TOL)])
def test_pnp_l1_presets_50_iterations_vs_unmodified_s3(env, cs_inputs, gold50, key, name, tol):
    """S3 presets (S3:339-346) of the denoisers its driver does not reach, through the script's own PNP_ADMM_L1_D."""
    pk, _, m, nz = env
    from pnp_admm_cnc_mri_b200.pnp import pnp_admm_l1
    from pnp_admm_cnc_mri_b200.denoisers import Denoiser
    it, reo = int(gold50[key + '_params'][0]), float(gold50[key + '_params'][1])
    assert it == 50
    kw = dict(ircnn_weights=_ircnn_sets(int(gold50['ircnn_seed0']))) if 'ircnn' in name else {}
    D = Denoiser(name, iter_num=it, x8=False, noises=nz, dtype=torch.float32, seed=0, **kw)
    x = pnp_admm_l1(_img(cs_inputs, str(gold50['single_image'])), m, nz, D, iter_num=it, reo=reo)
    e = rel(x, gold50[key])
    print(f'{key}: 50 iterations, rel-L2 vs unmodified S3 {e:.2e}, max|diff| {np.abs(x - gold50[key]).max():.2e}')
    assert e < tol


@pytest.mark.parametrize('key,name,tol', [('cnc_fdncnn', 'fdncnn_gray', STABLE_TOL), ('cnc_ircnn', 'ircnn_gray', STABLE_TOL),
                                          ('cnc_ffdnet', 'ffdnet_gray', CHAOTIC_TOL)])
def test_pnp_cnc_presets_50_iterations_vs_unmodified_s6(env, cs_inputs, gold50, key, name, tol):
    """S6 presets (S6:569-575) through the script's own PNP_ADMM_CNC_D."""
    pk, _, m, nz = env
    from pnp_admm_cnc_mri_b200.pnp import pnp_admm_cnc
    from pnp_admm_cnc_mri_b200.denoisers import Denoiser
    a, it, lam, reo, b = (float(v) for v in gold50[key + '_params'])
    assert it == 50
    kw = dict(ircnn_weights=_ircnn_sets(int(gold50['ircnn_seed0']))) if 'ircnn' in name else {}
    D = Denoiser(name, iter_num=int(it), x8=False, noises=nz, dtype=torch.float32, seed=0, **kw)
    x = pnp_admm_cnc(_img(cs_inputs, str(gold50['single_image'])), m, nz, D, None, alpha=a, iter_num=int(it), lambda1=lam, reo=reo, b=b)
    e = rel(x, gold50[key])
    print(f'{key}: 50 iterations, rel-L2 vs unmodified S6 {e:.2e}, max|diff| {np.abs(x - gold50[key]).max():.2e}')
    assert e < tol


def test_pnp_cnc_dncnn_pair_50_iterations_three_images(env, cs_inputs, gold50):
    """BASELINE config 3's loop at its preset depth: PNP_ADMM_CNC_DnCNN (S6:571: 1.2, 50, 4, 0.45, 0.3) on 05 / 01 / 10.png as
    ONE batch, float32 denoiser, against the unmodified script; then the same loop with the DnCNN on the tcgen05 kernels
    (bf16 operands): reported, and bounded at bf16-denoiser accuracy."""
    pk, _, m, nz = env
    from pnp_admm_cnc_mri_b200.pnp import pnp_admm_cnc
    names = [str(s) for s in gold50['images3']]
    imgs = np.stack([_img(cs_inputs, n) for n in names])
    P = dict(alpha=1.2, iter_num=50, lambda1=4, reo=0.45, b=0.3)
    D = _D('dncnn_25', 50, False, nz)
    x = pnp_admm_cnc(imgs, m, nz, D, D, **P)
    for k, n in enumerate(names):
        e = rel(x[k], gold50['cnc_dncnn'][k])
        print(f'cnc_dncnn {n}: 50 iterations, fp32 denoiser, rel-L2 vs unmodified S6 {e:.2e}')
        assert e < STABLE_TOL, (n, e)
    D16 = _D('dncnn_25', 50, False, nz, torch.bfloat16)
    assert D16.fused is not None
    x16 = pnp_admm_cnc(imgs, m, nz, D16, D16, **P)
    for k, n in enumerate(names):
        e = rel(x16[k], gold50['cnc_dncnn'][k])
        print(f'cnc_dncnn {n}: 50 iterations, K5 (bf16 tcgen05) denoiser, rel-L2 vs unmodified S6 {e:.2e}')
        assert e < CHAOTIC_TOL, (n, e)


def test_pnp_drunet_50_iterations_three_images(env, cs_inputs, gold50):
    """S3's own driver (PnP-ADMM-L1, DRUNet, x8 schedule; BASELINE config 4's loop at 256x256) and S6's (PnP-ADMM-CNC, DRUNet) at 50
    iterations on three images.  A random-weight DRUNet inside the loop amplifies rounding differences (see the header of this
    block), so the agreement is reported and bounded, not gated at 1e-4; the 3-iteration goldens above gate the same code at 1e-4."""
    pk, _, m, nz = env
    from pnp_admm_cnc_mri_b200.pnp import pnp_admm_cnc, pnp_admm_l1
    names = [str(s) for s in gold50['images3']]
    imgs = np.stack([_img(cs_inputs, n) for n in names])
    x = pnp_admm_l1(imgs, m, nz, _D('drunet_gray', 50, True, nz), iter_num=50, reo=0.26)
    for k, n in enumerate(names):
        e = rel(x[k], gold50['l1_drunet'][k])
        print(f'l1_drunet {n}: 50 iterations, rel-L2 vs unmodified S3 {e:.2e}')
        assert e < CHAOTIC_TOL, (n, e)
    x = pnp_admm_cnc(imgs, m, nz, _D('drunet_gray', 50, False, nz), None, alpha=1, iter_num=50, lambda1=0.8, reo=0.8, b=0.45)
    for k, n in enumerate(names):
        e = rel(x[k], gold50['cnc_drunet'][k])
        print(f'cnc_drunet {n}: 50 iterations, rel-L2 vs unmodified S6 {e:.2e}')
        assert e < CHAOTIC_TOL, (n, e)


def test_pnp_l1_ffdnet_50_iterations_tensor_core_denoiser(env, cs_inputs, gold50):
    """S3's FFDNet preset at its 50 iterations with the FFDNet on the tcgen05 kernels (bf16 operands, pnpadmm_ffdnet_forward_bf16)
    against the unmodified script's float32 result: reported, and bounded at bf16-denoiser accuracy (the float32 denoiser is gated
    at 1e-4 above); the stock PyTorch bf16 module in the same loop is the yardstick."""
    pk, _, m, nz = env
    from pnp_admm_cnc_mri_b200.pnp import pnp_admm_l1
    it, reo = int(gold50['l1_ffdnet_params'][0]), float(gold50['l1_ffdnet_params'][1])
    img = _img(cs_inputs, str(gold50['single_image']))
    D16 = _D('ffdnet_gray', it, False, nz, torch.bfloat16)
    assert D16.fused is not None
    e = rel(pnp_admm_l1(img, m, nz, D16, iter_num=it, reo=reo), gold50['l1_ffdnet'])
    from pnp_admm_cnc_mri_b200.denoisers import Denoiser
    Dt = Denoiser('ffdnet_gray', iter_num=it, noises=nz, dtype=torch.bfloat16, seed=0, fused=False)
    et = rel(pnp_admm_l1(img, m, nz, Dt, iter_num=it, reo=reo), gold50['l1_ffdnet'])
    print(f'l1_ffdnet: 50 iterations vs unmodified S3: K5 (bf16 tcgen05) {e:.2e}, PyTorch bf16 module {et:.2e}')
    assert e < max(CHAOTIC_TOL, 2 * et), (e, et)


def test_pnp_l1_ircnn_50_iterations_tensor_core_denoiser(env, cs_inputs, gold50):
    """S3's IRCNN preset (25 weight sets switched by sigma, S3:280-288) at 50 iterations with the dilated network on the tcgen05
    kernels (pnpadmm_dncnn_forward_dilated_bf16) against the unmodified script's float32 result: reported, bounded at
    bf16-denoiser accuracy, with the stock PyTorch bf16 module in the same loop as the yardstick."""
    pk, _, m, nz = env
    from pnp_admm_cnc_mri_b200.pnp import pnp_admm_l1
    from pnp_admm_cnc_mri_b200.denoisers import Denoiser
    it, reo = int(gold50['l1_ircnn_params'][0]), float(gold50['l1_ircnn_params'][1])
    img = _img(cs_inputs, str(gold50['single_image']))
    sets = _ircnn_sets(int(gold50['ircnn_seed0']))
    D16 = Denoiser('ircnn_gray', iter_num=it, noises=nz, dtype=torch.bfloat16, seed=0, ircnn_weights=sets)
    assert D16.fused is not None
    e = rel(pnp_admm_l1(img, m, nz, D16, iter_num=it, reo=reo), gold50['l1_ircnn'])
    Dt = Denoiser('ircnn_gray', iter_num=it, noises=nz, dtype=torch.bfloat16, seed=0, ircnn_weights=sets, fused=False)
    et = rel(pnp_admm_l1(img, m, nz, Dt, iter_num=it, reo=reo), gold50['l1_ircnn'])
    print(f'l1_ircnn: 50 iterations vs unmodified S3: K5 (bf16 tcgen05) {e:.2e}, PyTorch bf16 module {et:.2e}')
    assert e < max(CHAOTIC_TOL, 2 * et), (e, et)
