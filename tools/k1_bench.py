"""Scratch timing of K1 (cluster kernel) only: python tools/k1_bench.py [B ...]. CUDA events, median of 7."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pnp_admm_cnc_mri_b200 as pk
from pnp_admm_cnc_mri_b200 import data

Bs = [int(a) for a in sys.argv[1:]] or [64, 1024]
N = 256
for B in Bs:
    imgs = data.phantoms(8, N, seed0=0)
    imgs = np.concatenate([imgs] * ((B + 7) // 8))[:B]
    m = data.make_mask('random', N, seed=0)
    nz = data.make_noise(N, seed=1)
    s = pk.AdmmSolver(B, N)
    y = s.acquire(imgs, m, nz)
    z0 = s.zero_filled(y)
    s.prepare(y, m, 0.05)
    x = torch.empty_like(z0)
    for prox in ('cnc', 'l1'):
        ts = []
        for r in range(9):
            z = z0.clone(); w = torch.zeros_like(z0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            s.iterate(x, z, w, prox, 50, 0.5, 0.05, 0.45, 64, kernel=os.environ.get('K1B_KERNEL', 'cluster'))
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        t = float(np.median(ts[2:]))
        its = B * 50 / (t * 1e-3)
        print(f'K1 cl={os.environ.get("PNPADMM_K1_CLUSTER","def")} B={B:5d} {prox}: {t:8.3f} ms  {its/1e6:6.3f} M it/s  '
              f'{its*10485760/1e12:6.2f} TFLOP/s nominal ({its*10485760/74.45e12*100:4.1f}% of 74.45)', flush=True)
