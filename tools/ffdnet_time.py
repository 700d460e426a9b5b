#!/usr/bin/env python
"""Two FFDNet forwards through pnpadmm_ffdnet_forward_bf16 (for ncu captures of the pack kernel, the thin first layer and the
pixel-shuffled tail): python tools/ffdnet_time.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pnp_admm_cnc_mri_b200 import denoisers, dncnn_fused as df

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
net = denoisers.build_model('ffdnet_gray', seed=0).cuda()
f = df.FusedFFDNet(net)
x = torch.rand(B, 1, 256, 256, device='cuda')
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = f(x, 15 / 255.); e1.record(); torch.cuda.synchronize()
    print(f'FFDNet forward B={B}: {e0.elapsed_time(e1):.3f} ms', flush=True)
