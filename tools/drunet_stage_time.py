#!/usr/bin/env python
"""Where a DRUNet forward spends its time in PyTorch bf16 (cuDNN, channels_last): python tools/drunet_stage_time.py [B] [H]
Default = the quadrant batch of BASELINE config 4 (16 images of 512^2 -> 64 tiles of 288^2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pnp_admm_cnc_mri_b200 import denoisers

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 288
net = denoisers.build_model('drunet_gray', seed=0).cuda().to(torch.bfloat16).to(memory_format=torch.channels_last)
x0 = torch.rand(B, 2, H, H, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def timed(fn, reps=5):
    for _ in range(2):
        out = fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), out


with torch.no_grad():
    t_all, _ = timed(lambda: net(x0))
    t, x1 = timed(lambda: net.m_head(x0)); print(f'head 2->64            {t:7.3f} ms')
    blocks1 = torch.nn.Sequential(*list(net.m_down1)[:4]); down1 = list(net.m_down1)[4]
    t, y = timed(lambda: blocks1(x1)); print(f'down1: 4 ResBlocks 64 {t:7.3f} ms   ({8 * 2 * 9 * 64 * 64 * B * H * H / t / 1e9:.0f} TFLOP/s)')
    t, x2 = timed(lambda: down1(y)); print(f'down1: strideconv     {t:7.3f} ms')
    t, x3 = timed(lambda: net.m_down2(x2)); print(f'down2 (128)           {t:7.3f} ms')
    t, x4 = timed(lambda: net.m_down3(x3)); print(f'down3 (256)           {t:7.3f} ms')
    t, x = timed(lambda: net.m_body(x4)); print(f'body (512)            {t:7.3f} ms')
    t, x = timed(lambda: net.m_up3(x + x4)); print(f'up3 (256)             {t:7.3f} ms')
    t, x = timed(lambda: net.m_up2(x + x3)); print(f'up2 (128)             {t:7.3f} ms')
    up1 = list(net.m_up1)[0]; blocks_u = torch.nn.Sequential(*list(net.m_up1)[1:])
    t, u = timed(lambda: up1(x + x2)); print(f'up1: convtranspose    {t:7.3f} ms')
    t, x = timed(lambda: blocks_u(u)); print(f'up1: 4 ResBlocks 64   {t:7.3f} ms')
    t, o = timed(lambda: net.m_tail(x + x1)); print(f'tail 64->1            {t:7.3f} ms')
    print(f'whole forward         {t_all:7.3f} ms   B={B} {H}x{H}')
