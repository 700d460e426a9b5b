"""Profiling driver (run under ncu): a few pnpadmm_reconstruct_f32 calls, images -> reconstructions.
usage: python tools/prof_reconstruct.py <mask kind: cartesian|radial|random> [B] [iters] [kernel]
cartesian -> K3 rowsep256_kernel; radial / random -> K1 cluster256_kernel with the fused prologue (+ K2 share under kernel=auto)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pnp_admm_cnc_mri_b200 as pk
from pnp_admm_cnc_mri_b200 import data

kind = sys.argv[1] if len(sys.argv) > 1 else 'random'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 28
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
kernel = sys.argv[4] if len(sys.argv) > 4 else 'auto'
N = 256
imgs = data.phantoms(min(B, 8), N, seed0=0)
imgs = torch.as_tensor(np.concatenate([imgs] * ((B + 7) // 8))[:B]).cuda()
m = data.make_mask(kind, N, seed=0)
nz = data.make_noise(N, seed=1)
s = pk.AdmmSolver(B, N)
for r in range(3):
    x, z, w = s.reconstruct(imgs, m, nz, 'cnc', iters, 0.5, 0.05, 0.45, 64, kernel=kernel)
torch.cuda.synchronize()
print('done', float(x.sum()))
