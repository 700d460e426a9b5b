#!/usr/bin/env python
"""Time one 64->64 tensor-core layer (pnpadmm_conv64_bf16 / pnpadmm_conv64_dilated_bf16): python tools/conv64_time.py [B] [reps] [dilation]
PNPADMM_TC_DEBUG=1|2|4 (or sums) skips input copies / output stores / MMAs, 256 prints the wait-time attribution (timing
experiments, results invalid; needs a library built with PNPADMM_NVCC_EXTRA=-DPNPADMM_TC_EXPERIMENTS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pnp_admm_cnc_mri_b200 import _abi, dncnn_fused as df

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dil = int(sys.argv[3]) if len(sys.argv) > 3 else 1
H = W = 256
lib = _abi.load()
a = torch.randn(B, H, W, 64, device='cuda').to(torch.bfloat16)
wp = df.pack_conv64(torch.randn(64, 64, 3, 3, device='cuda') / 24)
b = torch.zeros(64, device='cuda')
out = torch.empty_like(a)
st = torch.cuda.current_stream().cuda_stream
ts = []
for r in range(reps + 2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    if dil == 1:
        _abi.check(lib.pnpadmm_conv64_bf16(a.data_ptr(), out.data_ptr(), wp.data_ptr(), b.data_ptr(), B, H, W, 1, st))
    else:
        _abi.check(lib.pnpadmm_conv64_dilated_bf16(a.data_ptr(), out.data_ptr(), wp.data_ptr(), b.data_ptr(), B, H, W, 1, dil, st))
    e1.record(); torch.cuda.synchronize()
    if r >= 2: ts.append(e0.elapsed_time(e1))
t = min(ts)
fl = 2.0 * 64 * 64 * 9 * H * W * B
print(f'dbg={os.environ.get("PNPADMM_TC_DEBUG", "0")} B={B} dilation={dil}: {t * 1e3:.1f} us  {fl / t / 1e9:.0f} TFLOP/s  {2 * a.numel() * 2 / t / 1e6:.0f} GB/s', flush=True)

if int(os.environ.get('PNPADMM_TC_DEBUG', '0')) & 256:
    import ctypes
    buf = (ctypes.c_ulonglong * 8)()
    lib.pnpadmm_debug_tc_prof(buf)
    n = (reps + 2) * 148
    names = ['producer wait empty', 'mma wait tempty', 'mma wait full', 'epilogue wait tfull', 'mma total', 'producer total', 'epilogue total']
    print('  per CTA and launch, kcycles: ' + ', '.join(f'{nm} {buf[i] / n / 1e3:.0f}' for i, nm in enumerate(names)))
