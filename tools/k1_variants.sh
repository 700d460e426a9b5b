#!/bin/bash
# timing experiments for K1 (results of dbg != 0 runs are invalid by construction)
for d in 0 1 2 3; do
  echo "== PNPADMM_K1_DEBUG=$d"
  PNPADMM_K1_DEBUG=$d timeout 120 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import pnp_admm_cnc_mri_b200 as pk
from pnp_admm_cnc_mri_b200 import data
N=256
for B in (30, 64, 240):
    imgs = data.phantoms(8, N); imgs = np.concatenate([imgs]*((B+7)//8))[:B]
    m = data.make_mask('random', N); nz = data.make_noise(N)
    s = pk.AdmmSolver(B, N); y = s.acquire(imgs, m, nz); z0 = s.zero_filled(y); s.prepare(y, m, 0.05)
    x = torch.empty_like(z0); ts=[]
    for r in range(6):
        z = z0.clone(); w = torch.zeros_like(z0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        s.iterate(x, z, w, 'cnc', 50, 0.5, 0.05, 0.45, 64, kernel='cluster')
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = np.median(ts[2:]); rounds = -(-((B+1)//2)//15)
    print(f'B={B:4d} {t:8.3f} ms  {B*50/t/1e3:7.3f} M it/s   {t*1e3/rounds/50:6.2f} us per plane-iteration', flush=True)
PY
done
