#!/usr/bin/env python
"""K3 timing: AdmmSolver.reconstruct under a Cartesian mask (row-separable kernel), CUDA events, min of 7.
usage: python tools/k3_time.py [N B]..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pnp_admm_cnc_mri_b200 as pk
from pnp_admm_cnc_mri_b200 import data

args = [int(a) for a in sys.argv[1:]] or [256, 64, 256, 512, 512, 64]
for N, B in zip(args[0::2], args[1::2]):
    imgs = data.phantoms(8, N, seed0=0)
    imgs = torch.as_tensor(np.concatenate([imgs] * ((B + 7) // 8))[:B]).cuda()
    m = torch.as_tensor(data.make_mask('cartesian', N, seed=0)).cuda()
    nz = torch.as_tensor(data.make_noise(N, seed=1)).cuda().to(torch.complex64)
    s = pk.AdmmSolver(B, N)
    ts = []
    for r in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        s.reconstruct(imgs, m, nz, 'cnc', 50, 0.5, 0.05, 0.45, 64)
        e1.record(); torch.cuda.synchronize()
        if r >= 3:
            ts.append(e0.elapsed_time(e1))
    print(f'K3 N={N} B={B}: {min(ts):.4f} ms (median {sorted(ts)[len(ts) // 2]:.4f})', flush=True)
    del s
