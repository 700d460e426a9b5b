"""CPU: denoiser architectures against the reference's (parameter counts, state-dict key layout, forward
outputs recorded from the reference classes with the same weights), and the PnP restatement of the oracle
against outputs of the unmodified scripts (tests/golden/pnp_golden.npz, oracle/make_golden_pnp.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import kat_table as kat
from oracle import reference_numpy as orc
from pnp_admm_cnc_mri_b200 import denoisers as dn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
NAMES = dict(dncnn='dncnn_25', fdncnn='fdncnn_gray', ircnn='ircnn_gray', ffdnet='ffdnet_gray', drunet='drunet_gray')


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLD, 'pnp_golden.npz'))


@pytest.mark.parametrize('arch', list(NAMES))
def test_architecture_matches_reference(arch, gold):
    keys = json.load(open(os.path.join(GOLD, 'model_keys.json')))
    m = dn.build_model(NAMES[arch], seed=11)
    sd = m.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == keys['keys'][arch]          # KAIR checkpoints load strict
    assert dn.count_params(m) == keys['params'][arch]
    x = torch.from_numpy(gold[arch + '_x'])
    with torch.no_grad():
        y = m(x, torch.full((1, 1, 1, 1), 15 / 255.)) if arch == 'ffdnet' else m(x)
    assert torch.allclose(y, torch.from_numpy(gold[arch + '_y']), atol=2e-6, rtol=1e-5)


def test_dncnn17_param_count_from_reference_logs():
    assert dn.count_params(dn.build_model('dncnn_15')) == kat.DNCNN17_PARAMS         # "Params number: 555137"


def test_sigma_schedule_equals_oracle():
    a = dn.get_rho_sigma(sigma=15 / 255., iter_num=50, modelSigma1=49, modelSigma2=15, w=1.0)[1]
    b = orc.get_rho_sigma(sigma=15 / 255., iter_num=50, modelSigma1=49, modelSigma2=15, w=1.0)[1]
    assert np.array_equal(a, b)


def test_augment_roundtrip():
    x = torch.arange(2 * 1 * 6 * 6, dtype=torch.float32).reshape(2, 1, 6, 6)
    for mode in range(8):
        assert torch.equal(dn.augment(dn.augment(x, mode), dn.augment_inverse_mode(mode)), x)


def test_split_forward_quadrants():
    """512^2 goes through 4 overlapping 288^2 quadrants stitched at the centre (utils_model.py:91-108)."""
    calls = []

    def model(t):
        calls.append(tuple(t.shape))
        return t[:, :1] * 2

    x = torch.rand(2, 2, 512, 512)
    y = dn.split_forward(model, x, 32, 256, 16)
    assert calls == [(8, 2, 288, 288)]                     # one batched pass over the 4 quadrants
    assert torch.equal(y, x[:, :1] * 2)
    calls.clear()
    x = torch.rand(1, 2, 250, 250)                        # <= 256^2: replicate-pad to a multiple of 16
    y = dn.split_forward(model, x, 32, 256, 16)
    assert calls == [(1, 2, 256, 256)] and y.shape == (1, 1, 250, 250)


def _cpu_denoiser(name, iters, x8, noises):
    D = dn.Denoiser(name, iter_num=iters, x8=x8, noises=noises, dtype=torch.float32, device='cpu', seed=0)
    return lambda a, i: D(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))[None, None], i)[0, 0].numpy()


def test_oracle_pnp_matches_unmodified_scripts(cs_inputs, gold):
    it = int(gold['iters'])
    img = orc.preprocess_uint8(cs_inputs['images'][4])
    m = cs_inputs['masks'][0].astype(np.float64)
    nz = cs_inputs['noises']
    D = _cpu_denoiser('dncnn_25', it, False, nz)
    x = orc.pnp_admm_cnc(img, m, nz, D, D, alpha=1.2, iter_num=it, lambda1=4, reo=0.45, b=0.3)      # S6:571 preset
    assert np.abs(x - gold['cnc_dncnn']).max() < 2e-5
    x = orc.pnp_admm_l1(img, m, nz, _cpu_denoiser('drunet_gray', it, True, nz), iter_num=it, reo=0.26)  # S3:347
    assert np.abs(x - gold['l1_drunet']).max() < 2e-5
