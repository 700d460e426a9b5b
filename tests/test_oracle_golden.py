"""CPU: the oracle restatement against the reference's own golden vectors
(SURVEY.md section 4 / 8c): outputs of the unmodified scripts and the PSNR table
mined from results/*.log."""
import numpy as np
import pytest

from oracle import kat_table as kat
from oracle import reference_numpy as orc


def _img(cs, i):
    return orc.preprocess_uint8(cs['images'][i])


def test_restatement_bit_identical_to_unmodified_scripts(cs_inputs, ref_out):
    img = _img(cs_inputs, 4)                       # 05.png = testsets/set1
    m = cs_inputs['masks'][0].astype(np.float64)
    assert np.array_equal(orc.admm_l1(img, m, cs_inputs['noises'], **kat.L1_DEFAULTS), ref_out['l1'])
    assert np.array_equal(orc.admm_cnc(img, m, cs_inputs['noises'], **kat.CNC_DEFAULTS), ref_out['cnc'])


def test_set1_log_values(cs_inputs, ref_out):
    H = cs_inputs['images'][4]
    for algo in ('l1', 'cnc'):
        E = ref_out[algo] * 255                     # S1:133
        p, s, r = kat.SET1[algo]
        assert round(orc.calculate_psnr(E, H), 4) == p
        assert abs(orc.calculate_ssim(E, H) - s) < 5e-5
        assert round(orc.calculate_re(E, H), 4) == r


def test_zero_fill_psnr(cs_inputs):
    img = _img(cs_inputs, 4)
    for mi, name in enumerate(cs_inputs['mask_names']):
        y = orc.acquire(img, cs_inputs['masks'][mi].astype(np.float64), cs_inputs['noises'])
        x0 = np.fft.ifft2(y)
        assert round(float(orc.psnr_zero_fill(x0 * 255, img * 255)), 4) == kat.ZERO_FILL_05[name]


@pytest.mark.parametrize('mask_i', [0, 1, 2])
@pytest.mark.parametrize('algo', ['l1', 'cnc'])
def test_kat_table_all_90_rows(cs_inputs, mask_i, algo):
    name = cs_inputs['mask_names'][mask_i]
    m = cs_inputs['masks'][mask_i].astype(np.float64)
    ps, ss, rs = [], [], []
    for i in range(15):
        img = _img(cs_inputs, i)
        if algo == 'l1':
            x = orc.admm_l1(img, m, cs_inputs['noises'], **kat.L1_DEFAULTS)
        else:
            x = orc.admm_cnc(img, m, cs_inputs['noises'], **kat.CNC_DEFAULTS)
        H = cs_inputs['images'][i]
        p = orc.calculate_psnr(x * 255, H)
        assert abs(p - kat.PSNR[(name, algo)][i]) < 6e-5, (name, algo, i, p)
        ps.append(p)
        ss.append(orc.calculate_ssim(x * 255, H))
        rs.append(orc.calculate_re(x * 255, H))
    ap, as_, ar = kat.AVERAGES[(name, algo)]
    assert abs(np.mean(ps) - ap) < 6e-4
    assert abs(np.mean(ss) - as_) < 6e-4
    assert abs(np.mean(rs) - ar) < 6e-4


def test_soft_semantics():
    x = np.array([-2.0, -0.5, 0.0, 0.5, 2.0])
    assert np.array_equal(orc.soft(x, 1.0), np.array([-1.0, -0.0, 0.0, 0.0, 1.0]))
    assert orc.soft(np.array([0.0]), -1.0)[0] == 0.0          # sign(0) == 0 even for c < 0


def test_sigma_schedule():
    _, s = orc.get_rho_sigma(sigma=15 / 255., iter_num=50, modelSigma1=49, modelSigma2=15, w=1.0)
    assert s.dtype == np.float32 and s.shape == (50,)
    assert abs(s[0] * 255 - 49) < 1e-4 and abs(s[-1] * 255 - 15) < 1e-4
    assert np.all(np.diff(s) < 0)
