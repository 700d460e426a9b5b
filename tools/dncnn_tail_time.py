#!/usr/bin/env python
"""Head + tail only (a 2-layer DnCNN) through the fused path: isolates the N = 16 tail kernel for PNPADMM_TC_DEBUG runs
(library built with PNPADMM_NVCC_EXTRA=-DPNPADMM_TC_EXPERIMENTS)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pnp_admm_cnc_mri_b200 import denoisers, dncnn_fused as df
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
net = denoisers.DnCNN(1, 1, 64, nb=2).cuda()
f = df.FusedDnCNN(net, residual=True)
x = torch.rand(B, 1, 256, 256, device='cuda')
ts = []
for r in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); f(x); e1.record(); torch.cuda.synchronize()
    if r >= 2: ts.append(e0.elapsed_time(e1))
print(f'dbg={os.environ.get("PNPADMM_TC_DEBUG", "0")} head+tail B={B}: {min(ts) * 1e3:.1f} us', flush=True)
if int(os.environ.get('PNPADMM_TC_DEBUG', '0')) & 256:
    buf = (ctypes.c_ulonglong * 8)()
    f.lib.pnpadmm_debug_tc_prof(buf)
    n = 7 * 148
    names = ['producer wait empty', 'mma wait tempty', 'mma wait full', 'epilogue wait tfull', 'mma total', 'producer total', 'epilogue total']
    print('  per CTA and launch, kcycles: ' + ', '.join(f'{nm} {buf[i] / n / 1e3:.0f}' for i, nm in enumerate(names)))
