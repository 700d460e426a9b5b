"""Generate the committed fixtures under tests/golden/ (build container only).

TEST INFRASTRUCTURE.  Run as ``python oracle/make_golden.py``.  It
  1. converts the reference's DATA files (CS_MRI/*.mat masks + noise, the 15
     test PNGs read exactly as ``utils_image.imread_uint(path, 1)`` does, i.e.
     ``cv2.imread(path, 0)``) into ``tests/golden/cs_mri_inputs.npz`` so the
     GPU box (which has no /root/reference) can run the reference's own configs;
  2. executes the UNMODIFIED scripts S1 / S4 through ``oracle/run_reference.py``
     and asserts ``oracle/reference_numpy.py`` is bit-identical to their
     ``out[0]`` (this is what pins the oracle);
  3. stores those outputs as ``tests/golden/ref_out_05_random.npz``.
No reference SOURCE is copied; only data and outputs.
"""
from __future__ import annotations

import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import reference_numpy as orc      # noqa: E402
from oracle import run_reference as rr         # noqa: E402
from oracle import kat_table as kat            # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def convert_inputs():
    import cv2
    import scipy.io as sio
    ref = rr.REF
    masks = np.stack([sio.loadmat(os.path.join(ref, 'CS_MRI', m + '.mat'))['Q1'].astype(np.uint8)
                      for m in kat.MASKS])
    noises = sio.loadmat(os.path.join(ref, 'CS_MRI', 'noises.mat'))['noises'].astype(np.complex128)
    paths = sorted(glob.glob(os.path.join(ref, 'testsets', 'set', '*.png')))
    imgs = np.stack([cv2.imread(p, 0) for p in paths]).astype(np.uint8)
    assert imgs.shape == (15, 256, 256) and masks.shape == (3, 256, 256)
    np.savez_compressed(os.path.join(GOLD, 'cs_mri_inputs.npz'),
                        masks=masks, mask_names=np.array(kat.MASKS),
                        noises=noises,                    # NOT yet multiplied by 3.0 (S1:186 does that)
                        images=imgs, image_names=np.array([os.path.basename(p) for p in paths]))
    return masks, noises, imgs


def pin_against_unmodified_scripts(masks, noises, imgs):
    n3 = noises * 3.0
    img05 = orc.preprocess_uint8(imgs[4])
    out = {}
    for mi, mfile in enumerate(['Q_Random30.mat', 'Q_Radial30.mat', 'Q_Cartesian30.mat']):
        m = masks[mi].astype(np.float64)
        g1 = rr.run_script('【1】', mask_file=mfile)
        ref_l1 = np.asarray(g1['out'][0])
        mine = orc.admm_l1(img05, m, n3, **kat.L1_DEFAULTS)
        assert np.array_equal(ref_l1, mine), f'L1 restatement differs from unmodified S1 ({mfile})'
        g4 = rr.run_script('【4】', mask_file=mfile)
        ref_cnc = np.asarray(g4['out'][0])
        mine = orc.admm_cnc(img05, m, n3, **kat.CNC_DEFAULTS)
        assert np.array_equal(ref_cnc, mine), f'CNC restatement differs from unmodified S4 ({mfile})'
        print(f'[pin] {mfile}: restatement == unmodified S1 and S4 (bit-identical)')
        if mi == 0:
            out['l1'], out['cnc'] = ref_l1, ref_cnc
    # non-default parameters through the scripts' CLI (S1:21-27, S4:21-29)
    g = rr.run_script('【4】', argv=['--iter_num', 7, '--alpha', 0.3, '--lambda1', 0.2, '--reo', 0.1, '--b', 16])
    mine = orc.admm_cnc(img05, masks[0].astype(np.float64), n3, alpha=0.3, iter_num=7, lambda1=0.2, reo=0.1, b=16)
    assert np.array_equal(np.asarray(g['out'][0]), mine)
    g = rr.run_script('【1】', argv=['--iter_num', 9, '--lambda1', 0.3, '--reo', 0.05])
    mine = orc.admm_l1(img05, masks[0].astype(np.float64), n3, iter_num=9, lambda1=0.3, reo=0.05)
    assert np.array_equal(np.asarray(g['out'][0]), mine)
    print('[pin] non-default CLI parameters: bit-identical')
    np.savez_compressed(os.path.join(GOLD, 'ref_out_05_random.npz'),
                        l1=out['l1'], cnc=out['cnc'])


def main():
    if not rr.reference_available():
        raise SystemExit('reference tree not present; fixtures can only be regenerated in the build container')
    os.makedirs(GOLD, exist_ok=True)
    masks, noises, imgs = convert_inputs()
    pin_against_unmodified_scripts(masks, noises, imgs)
    if '--pnp' in sys.argv:
        from oracle import make_golden_pnp
        make_golden_pnp.main()
    print('golden fixtures written to', GOLD)


if __name__ == '__main__':
    main()
