"""Golden vectors of the PnP variants at the presets' full depth (50 iterations), from the UNMODIFIED
reference scripts S3 / S6 (build container only; TEST INFRASTRUCTURE).  Writes tests/golden/pnp_golden_50it.npz.

  * the scripts' own module-level drivers run as shipped (S3:375-378 -> PNP_ADMM_L1_D('drunet_gray'),
    S6:610-618 -> PNP_ADMM_CNC_D('drunet_gray') and PNP_ADMM_CNC_DnCNN('dncnn_25', 'dncnn_15')) on three
    images of testsets/set (05, 01, 10) with seeded random weights saved as model_zoo/*.pth;
  * the model index is hard-coded in both drivers, so the other presets (FDnCNN, FFDNet, IRCNN, DnCNN-15) are
    reached by calling the scripts' own, unmodified functions with the scripts' own preset dictionaries from the
    globals runpy returns (one image, 05.png).  The IRCNN branch uses `np.int` (S3:281, S6:290), removed in
    NumPy 1.24: the generator restores that alias for the duration of the run (an environment shim, the
    reference source is untouched);
  * every output is compared with oracle.reference_numpy.pnp_* driven by pnp_admm_cnc_mri_b200.denoisers on the
    CPU with the same weights, and the differences are printed and stored (restatement_maxdiff).

CPU cost: about 20 minutes on 8 cores (DRUNet dominates).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import reference_numpy as orc     # noqa: E402
from oracle import run_reference as rr        # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
IMAGES3 = ['05.png', '01.png', '10.png']
IRCNN_SEED0 = 100


def ircnn_zoo():
    """25 seeded IRCNN state-dicts keyed '0'..'24' like KAIR's ircnn_gray.pth (S3:187-189)."""
    from pnp_admm_cnc_mri_b200 import denoisers as dn
    return {str(k): dn.build_model('ircnn_gray', seed=IRCNN_SEED0 + k).state_dict() for k in range(25)}


def cpu_denoiser(model_name, iter_num, x8, noises, seed, **kw):
    from pnp_admm_cnc_mri_b200.denoisers import Denoiser
    D = Denoiser(model_name, iter_num=iter_num, x8=x8, noises=noises, dtype=torch.float32, device='cpu', seed=seed, **kw)
    return lambda a, i: D(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))[None, None], i)[0, 0].numpy()


def main():
    from pnp_admm_cnc_mri_b200 import denoisers as dn
    d = np.load(os.path.join(GOLD, 'cs_mri_inputs.npz'))
    names = [str(s) for s in d['image_names']]
    idx3 = [names.index(n) for n in IMAGES3]
    mask = d['masks'][0].astype(np.float64)
    noises = d['noises'] * 3.0
    imgs = [orc.preprocess_uint8(d['images'][i]) for i in idx3]
    paths3 = [os.path.join(rr.REF, 'testsets', 'set', n) for n in IMAGES3]
    zoo = {n: dn.build_model(n, seed=0).state_dict() for n in
           ('drunet_gray', 'dncnn_25', 'dncnn_15', 'fdncnn_gray', 'ffdnet_gray')}
    zoo['ircnn_gray'] = ircnn_zoo()
    if not hasattr(np, 'int'):
        np.int = int          # S3:281 / S6:290 (NumPy < 1.24 alias); environment shim only
    out, diffs = {}, {}
    t0 = time.time()

    # ---- S3: PnP-ADMM-L1 ------------------------------------------------------------------------
    def post3(g):
        res = {}
        for m, nm in enumerate(g['name']):
            if nm == 'drunet_gray':
                continue
            o = g['PNP_ADMM_L1_D'](nm, g['mask'][0], g['noises'], **g['PNP_ADMM_L1_D_opts'][m])
            res[nm] = (np.asarray(o[0]).copy(), dict(g['PNP_ADMM_L1_D_opts'][m]))
        return res
    # the driver's own run on three images; the extra presets see the same three but only 05.png (sorted first? no:
    # sorted order is 01, 05, 10) is kept
    g3 = rr.run_script('【3】', images=paths3, model_zoo=zoo, post=post3)
    order = sorted(IMAGES3)
    out['l1_drunet'] = np.stack([np.asarray(g3['out'][order.index(n)]) for n in IMAGES3]).astype(np.float32)
    print(f'[S3] driver done, {time.time() - t0:.0f} s', flush=True)
    for k, im in enumerate(imgs):
        mine = orc.pnp_admm_l1(im, mask, noises, cpu_denoiser('drunet_gray', 50, True, noises, 0), iter_num=50, reo=0.26)
        diffs[f'l1_drunet_{k}'] = float(np.abs(mine - out['l1_drunet'][k]).max())
    i01 = order.index('01.png')     # out[0] of the post runs = first image in sorted order = 01.png
    img01 = imgs[IMAGES3.index('01.png')]
    for nm, (o, P) in g3['_post'].items():
        key = 'l1_' + nm.split('_')[0]
        out[key] = o.astype(np.float32)
        kw = dict(ircnn_weights=zoo['ircnn_gray']) if 'ircnn' in nm else {}
        mine = orc.pnp_admm_l1(img01, mask, noises, cpu_denoiser(nm, P['iter_num'], False, noises, 0, **kw), **P)
        diffs[key] = float(np.abs(mine - o).max())
        out[key + '_params'] = np.array([P['iter_num'], P['reo']], dtype=np.float64)
    print(f'[S3] presets done, {time.time() - t0:.0f} s; restatement max|diff|:', {k: v for k, v in diffs.items() if k.startswith("l1")}, flush=True)

    # ---- S6: PnP-ADMM-CNC -----------------------------------------------------------------------
    def post6(g):
        res = {}
        for m, nm in enumerate(g['name']):
            if nm == 'drunet_gray':
                continue
            o, _ = g['PNP_ADMM_CNC_D'](nm, g['mask'][0], g['noises'], **g['PNP_ADMM_CNC_D_opts'][m])
            res[nm] = (np.asarray(o[0]).copy(), dict(g['PNP_ADMM_CNC_D_opts'][m]))
        # the driver keeps only out2[0] of the DnCNN pair (S6:618): call the unmodified function again for all three images
        o2, _ = g['PNP_ADMM_CNC_DnCNN'](g['name1'][1], g['name1'][0], g['mask'][0], g['noises'], **g['PNP_ADMM_CNC_DnCNN_opts'])
        res['_dncnn_pair'] = ([np.asarray(a).copy() for a in o2[:3]], np.asarray(g['out2']).copy())
        return res
    g6 = rr.run_script('【6】', images=paths3, model_zoo=zoo, post=post6)
    out['cnc_drunet'] = np.stack([np.asarray(g6['out'][order.index(n)]) for n in IMAGES3]).astype(np.float32)
    pair3, pair_first = g6['_post'].pop('_dncnn_pair')
    assert np.array_equal(pair3[0], pair_first)                # same function, same inputs: deterministic
    out['cnc_dncnn'] = np.stack([pair3[order.index(n)] for n in IMAGES3]).astype(np.float32)
    print(f'[S6] driver done, {time.time() - t0:.0f} s', flush=True)
    P4 = dict(alpha=1, iter_num=50, lambda1=0.8, reo=0.8, b=0.45)        # S6:577
    P2 = dict(alpha=1.2, iter_num=50, lambda1=4, reo=0.45, b=0.3)        # S6:571
    for k, im in enumerate(imgs):
        mine = orc.pnp_admm_cnc(im, mask, noises, cpu_denoiser('drunet_gray', 50, False, noises, 0), **P4)
        diffs[f'cnc_drunet_{k}'] = float(np.abs(mine - out['cnc_drunet'][k]).max())
        D1 = cpu_denoiser('dncnn_25', 50, False, noises, 0)
        mine = orc.pnp_admm_cnc(im, mask, noises, D1, D1, **P2)          # model2 loads model_path1 (S6:435)
        diffs[f'cnc_dncnn_{k}'] = float(np.abs(mine - out['cnc_dncnn'][k]).max())
    for nm, (o, P) in g6['_post'].items():
        key = 'cnc_' + nm.split('_')[0]
        out[key] = o.astype(np.float32)
        kw = dict(ircnn_weights=zoo['ircnn_gray']) if 'ircnn' in nm else {}
        mine = orc.pnp_admm_cnc(img01, mask, noises, cpu_denoiser(nm, P['iter_num'], False, noises, 0, **kw), **P)
        diffs[key] = float(np.abs(mine - o).max())
        out[key + '_params'] = np.array([P['alpha'], P['iter_num'], P['lambda1'], P['reo'], P['b']], dtype=np.float64)
    print(f'[S6] presets done, {time.time() - t0:.0f} s', flush=True)
    for k, v in diffs.items():
        print(f'[pin] {k:16s} restatement vs unmodified script: max|diff| = {v:.3e}')
    np.savez_compressed(os.path.join(GOLD, 'pnp_golden_50it.npz'), images3=np.array(IMAGES3), single_image=np.array('01.png'),
                        ircnn_seed0=IRCNN_SEED0, diff_keys=np.array(list(diffs)), diff_vals=np.array(list(diffs.values())), **out)
    print('wrote pnp_golden_50it.npz', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
