#!/usr/bin/env python
"""Where the non-iterate part of a device-resident step goes (B = 64, 256x256): CUDA events around each C-ABI call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pnp_admm_cnc_mri_b200 as pk
from pnp_admm_cnc_mri_b200 import data
N, B = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 64
imgs = torch.as_tensor(data.phantoms(B, N)).cuda()
m = torch.as_tensor(data.make_mask('random', N)).cuda()
nz = torch.as_tensor(data.make_noise(N)).cuda().to(torch.complex64)
s = pk.AdmmSolver(B, N)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
acc = {}
def ev(): return torch.cuda.Event(enable_timing=True)
for r in range(12):
    flush.fill_(r & 1)
    e = [ev() for _ in range(6)]
    torch.cuda.synchronize()
    e[0].record(); y = s.acquire(imgs, m, nz)
    e[1].record(); z = s.zero_filled(y)
    e[2].record(); w = torch.zeros_like(z); x = torch.empty_like(z)
    e[3].record(); s.prepare(y, m, 0.05)
    e[4].record(); s.iterate(x, z, w, 'cnc', 50, 0.5, 0.05, 0.45, 64)
    e[5].record(); torch.cuda.synchronize()
    if r >= 2:
        for k, nm in enumerate(['acquire', 'zero_filled', 'torch zeros/empty', 'prepare', 'iterate']):
            acc[nm] = acc.get(nm, 0.0) + e[k].elapsed_time(e[k + 1]) / 10
print({k: round(v * 1e3, 1) for k, v in acc.items()}, 'us; total', round(sum(acc.values()) * 1e3, 1))
