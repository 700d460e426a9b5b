import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import pnp_admm_cnc_mri_b200 as pk
from pnp_admm_cnc_mri_b200 import data
B, N = 64, 256
s = pk.AdmmSolver(B, N)
imgs = torch.as_tensor(data.phantoms(B, N)).cuda(); m = torch.as_tensor(data.make_mask('random', N)).cuda()
nz = torch.as_tensor(data.make_noise(N)).to(torch.complex64).cuda()
y = s.acquire(imgs, m, nz)
for kernel in ('auto', 'cluster'):
    for r in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = s.solve(y, m, 'cnc', 50, 0.5, 0.05, 0.45, 64, kernel=kernel)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(kernel, 'host enqueue %.3f ms, total %.3f ms' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
