"""Scratch timing of the streaming kernels (K2) at large batch: python tools/k2_bench.py N B iters [N B iters ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pnp_admm_cnc_mri_b200 as pk
from pnp_admm_cnc_mri_b200 import data

args = [int(a) for a in sys.argv[1:]] or [1024, 256, 10]
for N, B, iters in zip(args[0::3], args[1::3], args[2::3]):
    imgs = data.phantoms(4, N, seed0=0)
    imgs = np.concatenate([imgs] * ((B + 3) // 4))[:B]
    m = data.make_mask('random', N, seed=0)
    nz = data.make_noise(N, seed=1)
    s = pk.AdmmSolver(B, N)
    y = s.acquire(imgs, m, nz)
    z0 = s.zero_filled(y)
    s.prepare(y, m, 0.05)
    x = torch.empty_like(z0)
    ts = []
    for r in range(5):
        z = z0.clone(); w = torch.zeros_like(z0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        s.iterate(x, z, w, 'cnc', iters, 0.5, 0.05, 0.45, 64, kernel='streaming')
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts[1:]))
    its = B * iters / (t * 1e-3)
    print(f'K2 N={N:5d} B={B:5d} iters={iters}: {t:9.3f} ms  {its/1e6:7.4f} M it/s  {B/(t*1e-3):9.0f} img/s  '
          f'Q-model(57 B/px) {its*57*N*N/1e9:8.1f} GB/s  moved(36.5 B/px) {its*36.5*N*N/1e9:8.1f} GB/s', flush=True)
    del s, y, z0, x, z, w
    torch.cuda.empty_cache()
