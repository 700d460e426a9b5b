"""Scratch timing of the iterate kernels (CUDA events). Not the contract bench; see bench.py."""
import ctypes
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pnp_admm_cnc_mri_b200 as pk
from pnp_admm_cnc_mri_b200 import data, _abi

lib = _abi.load()
sm, ncl, ma, mi = (ctypes.c_int() for _ in range(4))
_abi.check(lib.pnpadmm_device_info(sm, ncl, ma, mi))
print('device: SMs', sm.value, 'max co-resident 8-CTA clusters', ncl.value, 'cc', ma.value, mi.value, flush=True)
fl = ctypes.c_double()
_abi.check(lib.pnpadmm_measure_fp32_peak(fl, None))
print('measured FP32 FMA peak: %.2f TFLOP/s' % (fl.value / 1e12), flush=True)


def timeit(B, N, kernel, prox='cnc', iters=50, reps=5, dtype='float32'):
    imgs = data.phantoms(min(B, 8), N, seed0=0)
    imgs = np.concatenate([imgs] * ((B + 7) // 8))[:B]
    m = data.make_mask('random', N, seed=0)
    nz = data.make_noise(N, seed=1)
    s = pk.AdmmSolver(B, N, dtype=dtype)
    y = s.acquire(imgs, m, nz)
    z0 = s.zero_filled(y)
    s.prepare(y, m, 0.05)
    x = torch.empty_like(z0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for r in range(reps + 2):
        z = z0.clone(); w = torch.zeros_like(z0)
        torch.cuda.synchronize()
        e0.record()
        s.iterate(x, z, w, prox, iters, 0.5, 0.05, 0.45, 64, kernel=kernel)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = np.median(ts[2:])
    its = B * iters / (t * 1e-3)
    F = 10 * N * N * np.log2(N * N)
    print(f'{kernel:9s} {dtype} N={N:5d} B={B:5d} {prox}: {t:8.3f} ms  {its/1e6:8.3f} M it/s  {B/(t*1e-3):10.0f} img/s  '
          f'nominal {its*F/1e12:6.2f} TFLOP/s  bytes-model {its*57*N*N/1e9:8.1f} GB/s', flush=True)


for B in (2, 32, 64, 256, 1024):
    timeit(B, 256, 'cluster')
for B in (64, 1024):
    timeit(B, 256, 'streaming')
timeit(64, 256, 'cluster', prox='l1')
timeit(64, 512, 'streaming', iters=20)
timeit(32, 1024, 'streaming', iters=10)
timeit(16, 256, 'streaming', dtype='float64')
