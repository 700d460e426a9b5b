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


def test_ircnn_sigma_indexed_weight_switch():
    """S3:280-288 / S6:289-298: before every IRCNN call the reference picks one of 25 weight sets by
    `current_idx = np.int(np.ceil(sigmas[i] * 255. / 2.) - 1)` and reloads only when the index changes
    (`former_idx` starts at 0, so the very first call already swaps the random-init net for set 24).  The
    reference branch itself is dead on NumPy >= 1.24 (`np.int`), so the rule is restated here and the live set is
    identified per iteration from the output: 25 synthetic state-dicts, 50 iterations of the DPIR schedule."""
    it = 50
    sets = {str(k): dn.build_model('ircnn_gray', seed=100 + k).state_dict() for k in range(25)}   # KAIR layout: str keys
    D = dn.Denoiser('ircnn_gray', iter_num=it, x8=False, dtype=torch.float32, device='cpu', seed=0, ircnn_weights=sets)
    sig = orc.get_rho_sigma(sigma=max(0.255 / 255., 15 / 255.), iter_num=it, modelSigma1=49, modelSigma2=15.0, w=1.0)[1]
    want_idx = [int(np.ceil(float(s) * 255. / 2.) - 1) for s in sig]
    assert want_idx[0] == 24 and want_idx[-1] == 7 and sorted(set(want_idx), reverse=True) == list(range(24, 6, -1))
    x = torch.rand(1, 1, 40, 56, generator=torch.Generator().manual_seed(3))
    probe = dn.build_model('ircnn_gray', seed=0)
    loads = []
    orig = D.net.load_state_dict
    D.net.load_state_dict = lambda sd, strict=True: (loads.append(1), orig(sd, strict=strict))[1]
    for i in range(it):
        y = D(x, i)
        probe.load_state_dict(sets[str(want_idx[i])], strict=True)
        with torch.no_grad():                              # (channels_last vs contiguous kernels: equal to rounding)
            assert (y - probe(x)).abs().max() < 1e-6, (i, want_idx[i])
        others = [k for k in (want_idx[i] - 1, want_idx[i] + 1) if 0 <= k < 25]
        for k in others:                                   # and it is not a neighbouring set
            probe.load_state_dict(sets[str(k)], strict=True)
            with torch.no_grad():
                assert (y - probe(x)).abs().max() > 1e-3
    assert len(loads) == len(set(want_idx))                # one reload per index change, none in between
    # a list indexed by int works too, and `weights=` given the 25-set dict is recognised as such
    D2 = dn.Denoiser('ircnn_gray', iter_num=it, dtype=torch.float32, device='cpu', seed=0, weights=sets)
    D3 = dn.Denoiser('ircnn_gray', iter_num=it, dtype=torch.float32, device='cpu', seed=0,
                     ircnn_weights=[sets[str(k)] for k in range(25)])
    for i in (0, 17, 49):
        assert torch.equal(D2(x, i), D(x, i)) and torch.equal(D3(x, i), D(x, i))


@pytest.mark.parametrize('key,name', [('l1_ircnn', 'ircnn_gray'), ('l1_dncnn', 'dncnn_15')])
def test_oracle_pnp_50_iterations_matches_unmodified_s3(cs_inputs, key, name):
    """The PnP restatement at the presets' full depth against outputs of the UNMODIFIED S3 function
    (tests/golden/pnp_golden_50it.npz, oracle/make_golden_pnp50.py): the IRCNN preset exercises the 25-set weight switch
    inside the reference's own loop (S3:280-288, run with the `np.int` alias restored), DnCNN-15 the plain residual branch."""
    g = np.load(os.path.join(GOLD, 'pnp_golden_50it.npz'))
    it, reo = int(g[key + '_params'][0]), float(g[key + '_params'][1])
    assert it == 50
    names = cs_inputs['image_names']
    img = orc.preprocess_uint8(cs_inputs['images'][names.index(str(g['single_image']))])
    kw = {}
    if 'ircnn' in name:
        kw['ircnn_weights'] = {str(k): dn.build_model('ircnn_gray', seed=int(g['ircnn_seed0']) + k).state_dict() for k in range(25)}
    D = dn.Denoiser(name, iter_num=it, x8=False, noises=cs_inputs['noises'], dtype=torch.float32, device='cpu', seed=0, **kw)
    f = lambda a, i: D(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))[None, None], i)[0, 0].numpy()
    x = orc.pnp_admm_l1(img, cs_inputs['masks'][0].astype(np.float64), cs_inputs['noises'], f, iter_num=it, reo=reo)
    assert np.abs(x - g[key]).max() < 2e-5
