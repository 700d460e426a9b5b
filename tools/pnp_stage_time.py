import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pnp_admm_cnc_mri_b200 import data, pnp, denoisers
from pnp_admm_cnc_mri_b200.solver import cnc_combine, dual_update_
dev = torch.device('cuda', 0)
B, N = 256, 256
imgs = torch.as_tensor(np.stack([data.phantom(N, i) for i in range(8)] * (B // 8)).astype(np.float32)).to(dev)
m = data.make_mask('random', N, seed=0); nz = data.make_noise(N, seed=3)
D = denoisers.build_denoiser('dncnn_25', iter_num=50, device=dev)
solver, z, w, single = pnp._setup(imgs, m, nz, 0.45, dev)
alpha, coef = 1.2, 1.2 * 0.45 * 4 * 0.3
names = ['xupdate', 'D1', 'combine', 'D2', 'dual']
acc = {k: 0.0 for k in names}
def ev(): return torch.cuda.Event(enable_timing=True)
for it in range(25):
    es = [ev() for _ in range(6)]
    es[0].record()
    x = solver.xupdate(z, w); es[1].record()
    s = D(z[:, None], it)[:, 0].contiguous(); es[2].record()
    t = cnc_combine(z, x, w, s, alpha, coef); es[3].record()
    z = D(t[:, None], it)[:, 0].contiguous(); es[4].record()
    dual_update_(x, z, w, clamp01=True); es[5].record()
    torch.cuda.synchronize()
    if it >= 5:
        for k in range(5): acc[names[k]] += es[k].elapsed_time(es[k + 1])
print({k: round(v / 20, 3) for k, v in acc.items()}, 'ms per iteration; total', round(sum(acc.values()) / 20, 3))
