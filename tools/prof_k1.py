"""Profiling driver: a few launches of the iterate kernels on prepared state (run under ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pnp_admm_cnc_mri_b200 as pk
from pnp_admm_cnc_mri_b200 import data

kernel = sys.argv[1] if len(sys.argv) > 1 else 'cluster'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 30
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
N = int(sys.argv[4]) if len(sys.argv) > 4 else 256
imgs = data.phantoms(min(B, 8), N, seed0=0)
imgs = np.concatenate([imgs] * ((B + 7) // 8))[:B]
m = data.make_mask('random', N, seed=0)
nz = data.make_noise(N, seed=1)
s = pk.AdmmSolver(B, N)
y = s.acquire(imgs, m, nz)
z0 = s.zero_filled(y)
s.prepare(y, m, 0.05)
x = torch.empty_like(z0)
for r in range(3):
    z = z0.clone(); w = torch.zeros_like(z0)
    s.iterate(x, z, w, 'cnc', iters, 0.5, 0.05, 0.45, 64, kernel=kernel)
torch.cuda.synchronize()
print('done', float(x.sum()))
