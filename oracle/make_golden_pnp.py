"""Pin the PnP restatement and the denoiser architectures against the UNMODIFIED reference
(build container only; TEST INFRASTRUCTURE).  Writes tests/golden/pnp_golden.npz and model_keys.json.

  * every architecture of pnp_admm_cnc_mri_b200.denoisers is instantiated with seeded random weights,
    its state_dict is loaded with strict=True into the reference's own class, and both forwards must agree;
  * the same state_dicts are saved as model_zoo/*.pth for the unmodified scripts S3 / S6, which are run
    for 3 iterations; their outputs must equal oracle.reference_numpy.pnp_* driven by the same networks
    and are stored as golden vectors for the GPU parity tests.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import reference_numpy as orc     # noqa: E402
from oracle import run_reference as rr        # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
ITERS = 3


def _ref_models():
    mpl = types.ModuleType('matplotlib'); plt = types.ModuleType('matplotlib.pyplot'); mpl.pyplot = plt
    saved = {k: sys.modules.get(k) for k in ('matplotlib', 'matplotlib.pyplot')}
    sys.modules['matplotlib'] = mpl; sys.modules['matplotlib.pyplot'] = plt
    sys.path.insert(0, rr.REF)
    try:
        from models.network_dncnn import DnCNN, FDnCNN, IRCNN
        from models.network_ffdnet import FFDNet
        from models.network_unet import UNetRes
    finally:
        sys.path.remove(rr.REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return dict(dncnn=lambda: DnCNN(1, 1, 64, 17, 'R'), fdncnn=lambda: FDnCNN(2, 1, 64, 20, 'R'), ircnn=lambda: IRCNN(1, 1, 64),
                ffdnet=lambda: FFDNet(1, 1, 64, 15, 'R'),
                drunet=lambda: UNetRes(2, 1, [64, 128, 256, 512], 4, 'R', 'strideconv', 'convtranspose'))


def cpu_denoiser(model_name, iter_num, x8, noises, seed):
    from pnp_admm_cnc_mri_b200.denoisers import Denoiser
    D = Denoiser(model_name, iter_num=iter_num, x8=x8, noises=noises, dtype=torch.float32, device='cpu', seed=seed)
    return lambda a, i: D(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))[None, None], i)[0, 0].numpy()


def main():
    from pnp_admm_cnc_mri_b200 import denoisers as dn
    ref = _ref_models()
    names = dict(dncnn='dncnn_25', fdncnn='fdncnn_gray', ircnn='ircnn_gray', ffdnet='ffdnet_gray', drunet='drunet_gray')
    keys, fwd = {}, {}
    g = torch.Generator().manual_seed(7)
    for arch, name in names.items():
        mine = dn.build_model(name, seed=11)
        theirs = ref[arch]()
        theirs.load_state_dict(mine.state_dict(), strict=True)           # key / shape compatibility
        theirs.eval()
        keys[arch] = {k: list(v.shape) for k, v in theirs.state_dict().items()}
        assert dn.count_params(mine) == sum(p.numel() for p in theirs.parameters())
        cin = 2 if arch in ('fdncnn', 'drunet') else 1
        x = torch.rand((2, cin, 48, 40), generator=g)
        with torch.no_grad():
            if arch == 'ffdnet':
                x = x[:1]                                    # the reference's sigma.repeat() only supports batch 1
                s = torch.full((1, 1, 1, 1), 15 / 255.)
                a, b = mine(x, s), theirs(x, s)
            else:
                a, b = mine(x), theirs(x)
        assert torch.allclose(a, b, atol=1e-6, rtol=1e-5), arch
        fwd[arch + '_x'] = x.numpy(); fwd[arch + '_y'] = b.numpy()
        print(f'[pin] {arch}: {dn.count_params(mine)} params, keys + forward identical to the reference class')
    json.dump({'keys': keys, 'params': {a: int(sum(int(np.prod(s)) for s in k.values())) for a, k in keys.items()}},
              open(os.path.join(GOLD, 'model_keys.json'), 'w'), indent=0)

    # ---- unmodified scripts with these weights -------------------------------------------------
    d = np.load(os.path.join(GOLD, 'cs_mri_inputs.npz'))
    mask = d['masks'][0].astype(np.float64)
    noises = d['noises'] * 3.0
    img = orc.preprocess_uint8(d['images'][4])
    zoo = {'drunet_gray': dn.build_model('drunet_gray', seed=0).state_dict(),
           'dncnn_25': dn.build_model('dncnn_25', seed=0).state_dict(),
           'dncnn_15': dn.build_model('dncnn_15', seed=0).state_dict()}
    out = {}
    g6 = rr.run_script('【6】', argv=['--iter_num', ITERS], model_zoo=zoo)
    ref_cnc_drunet = np.asarray(g6['out1'])
    ref_cnc_dncnn = np.asarray(g6['out2'])                      # S6:618 already takes out2[0]
    P = dict(alpha=1, iter_num=ITERS, lambda1=0.8, reo=0.8, b=0.45)      # S6:577 preset, iter_num from the CLI
    mine = orc.pnp_admm_cnc(img, mask, noises, cpu_denoiser('drunet_gray', ITERS, False, noises, 0), **P)
    print('[pin] S6 PNP_ADMM_CNC_D drunet : max|diff| vs restatement', np.abs(mine - ref_cnc_drunet).max())
    assert np.abs(mine - ref_cnc_drunet).max() < 2e-5
    P2 = dict(alpha=1.2, iter_num=ITERS, lambda1=4, reo=0.45, b=0.3)     # S6:571
    D1 = cpu_denoiser('dncnn_25', ITERS, False, noises, 0)
    mine = orc.pnp_admm_cnc(img, mask, noises, D1, D1, **P2)            # model2 loads model_path1 (S6:435)
    print('[pin] S6 PNP_ADMM_CNC_DnCNN   : max|diff| vs restatement', np.abs(mine - ref_cnc_dncnn).max())
    assert np.abs(mine - ref_cnc_dncnn).max() < 2e-5
    g3 = rr.run_script('【3】', argv=['--iter_num', ITERS], model_zoo=zoo)
    ref_l1_drunet = np.asarray(g3['out'][0])
    mine = orc.pnp_admm_l1(img, mask, noises, cpu_denoiser('drunet_gray', ITERS, True, noises, 0), iter_num=ITERS, reo=0.26)
    print('[pin] S3 PNP_ADMM_L1_D drunet  : max|diff| vs restatement', np.abs(mine - ref_l1_drunet).max())
    assert np.abs(mine - ref_l1_drunet).max() < 2e-5
    np.savez_compressed(os.path.join(GOLD, 'pnp_golden.npz'), iters=ITERS, cnc_drunet=ref_cnc_drunet, cnc_dncnn=ref_cnc_dncnn,
                        l1_drunet=ref_l1_drunet, **fwd)
    print('PnP golden vectors written')


if __name__ == '__main__':
    main()
