"""Small solves for compute-sanitizer (tools/sanitize.sh): one K1 solve with the chunked schedule, one hybrid
K1 + K2 solve, one K2 solve at N = 512, each checked against the oracle so a sanitizer-induced timing change
that exposed a protocol bug would also show up as a wrong answer."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pnp_admm_cnc_mri_b200 as pk                      # noqa: E402
from pnp_admm_cnc_mri_b200 import data                  # noqa: E402
from oracle import reference_numpy as orc               # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
P = dict(alpha=0.45, iter_num=6, lambda1=0.5, reo=0.05, b=64)


def check(x, imgs, m, nz, tag, n_check=2):
    for k in range(n_check):
        xr = orc.admm_cnc(imgs[k], m.astype(np.float64), nz, **P)
        e = np.linalg.norm(x[k] - xr) / np.linalg.norm(xr)
        assert e < 1e-4, (tag, k, e)
    print(tag, 'ok', flush=True)


N = 256
m = data.make_mask('random', N, seed=0)
nz = data.make_noise(N, seed=3)
if which in ('all', 'k1'):
    os.environ['PNPADMM_K1_CHUNKS'] = '3'               # plane state handed between clusters through L2
    imgs = data.phantoms(5, N, seed0=1)
    check(pk.admm_solve(imgs, m, nz, prox='cnc', kernel='cluster', **P), imgs, m, nz, 'K1 chunked B=5')
    del os.environ['PNPADMM_K1_CHUNKS']
if which in ('all', 'hybrid'):
    os.environ['PNPADMM_HYBRID_P2'] = '2'
    B = 2 * (int(os.environ.get('SAN_NCL', '14')) + 3)  # more planes than resident clusters, so the split is live
    imgs = np.concatenate([data.phantoms(4, N, seed0=1)] * (B // 4 + 1))[:B]
    check(pk.admm_solve(imgs, m, nz, prox='cnc', kernel='auto', **P), imgs, m, nz, f'hybrid B={B}')
if which in ('all', 'fused'):
    # images -> reconstruction in one call: fused prologue inside the cluster kernel, from uint8 images, hybrid split live
    os.environ['PNPADMM_HYBRID_P2'] = '2'
    B = 2 * (int(os.environ.get('SAN_NCL', '14')) + 3) + 1      # odd: the last plane has an empty slot
    imgs = np.concatenate([data.phantoms(4, N, seed0=1)] * (B // 4 + 1))[:B]
    u8 = np.uint8((imgs * 255).round())
    s = pk.AdmmSolver(B, N)
    x = s.reconstruct(u8, m, nz, 'cnc', P['iter_num'], P['lambda1'], P['reo'], P['alpha'], P['b'])[0].cpu().numpy()
    check(x, np.float32(u8 / 255.), m, nz, f'fused reconstruct B={B}')
    x = s.reconstruct(u8, m, nz, 'cnc', P['iter_num'], P['lambda1'], P['reo'], P['alpha'], P['b'], kernel='cluster')[0].cpu().numpy()
    check(x, np.float32(u8 / 255.), m, nz, f'fused reconstruct, cluster only, B={B}')
    del os.environ['PNPADMM_HYBRID_P2']
if which in ('all', 'k3'):
    # row-separable kernels: N = 256 (both implementations are exercised by PNPADMM_K3_K1CODE) and N = 512, odd batches, uint8 in
    mc = data.make_mask('cartesian', N, seed=7)
    imgs = data.phantoms(5, N, seed0=1)
    check(pk.admm_solve(imgs, mc, nz, prox='cnc', **P), imgs, mc, nz, 'K3 rowsep N=256 B=5')
    N3 = 512
    mc3 = data.make_mask('cartesian', N3, seed=7); nz3 = data.make_noise(N3, seed=4)
    im3 = data.phantoms(3, N3, seed0=2)
    check(pk.admm_solve(im3, mc3, nz3, prox='cnc', **P), im3, mc3, nz3, 'K3 rowsepN N=512 B=3', 1)
if which in ('all', 'k2'):
    N2 = 512
    m2 = data.make_mask('radial', N2, seed=1)
    nz2 = data.make_noise(N2, seed=4)
    imgs = data.phantoms(3, N2, seed0=2)
    check(pk.admm_solve(imgs, m2, nz2, prox='cnc', kernel='streaming', **P), imgs, m2, nz2, 'K2 N=512 B=3', 1)
if which in ('k5',):
    # tensor-core denoisers on ragged sizes (partial x tiles, short strips, odd sizes for FFDNet's padding / crop, every dilation):
    # one forward each, checked against the float32 module
    import torch
    from pnp_admm_cnc_mri_b200 import denoisers, dncnn_fused as df
    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm())
    for name, shape in (('dncnn_25', (2, 70, 200)), ('ircnn_gray', (2, 45, 150)), ('fdncnn_gray', (1, 33, 130))):
        net = denoisers.build_model(name, seed=3).cuda()
        cin = df.conv_layers(net)[0].in_channels
        xin = torch.rand(shape[0], cin, shape[1], shape[2], device='cuda')
        got = df.FusedDnCNN(net, residual=(cin == 1))(xin)
        with torch.no_grad():
            want = net(xin)
        assert rel(got, want) < 2e-2, (name, rel(got, want))
        print('K5', name, shape, 'ok', flush=True)
    net = denoisers.build_model('ffdnet_gray', seed=3).cuda()
    xin = torch.rand(2, 1, 65, 131, device='cuda')
    got = df.FusedFFDNet(net)(xin, 15 / 255.)
    with torch.no_grad():
        want = net(xin, torch.full((1, 1, 1, 1), 15 / 255., device='cuda'))
    assert rel(got, want) < 2e-2, rel(got, want)
    print('K5 ffdnet_gray (2, 65, 131) ok', flush=True)
