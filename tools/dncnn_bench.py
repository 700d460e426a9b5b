#!/usr/bin/env python
"""K5 check + timing: tensor-core DnCNN forward (csrc/dncnn_tc.cuh) against PyTorch's bf16 cuDNN path.

    python tools/dncnn_bench.py [B ...]      (default 64 256; 256x256 images, DnCNN-17 random init)

Prints the single-layer and whole-network errors first (cheap, so a broken kernel shows up before the timing),
then CUDA-event times of one forward: ours, torch bf16 channels_last, and the achieved dense-bf16 TFLOP/s.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from pnp_admm_cnc_mri_b200 import denoisers, dncnn_fused as df  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def check_layer(B, H, W):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, H, W, 64, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24.0).to(torch.bfloat16).float().cuda()
    b = torch.randn(64, generator=g).cuda()
    got = df.conv64(x, w, b, relu=True).float()
    torch.cuda.synchronize()
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        want = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w, b, padding=1)).permute(0, 2, 3, 1)
    print(f'layer B={B} H={H} W={W}: rel {rel(got, want):.3e}  max|err| {float((got - want).abs().max()):.3e}', flush=True)
    return rel(got, want)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def main():
    Bs = [int(a) for a in sys.argv[1:]] or [64, 256]      # e.g. "1 15 256": small batches show the strip decomposition
    torch.cuda.set_device(0)
    bad = 0
    for shp in [(1, 8, 128), (2, 64, 256), (1, 70, 200)]:
        bad += check_layer(*shp) > 3e-3
    net = denoisers.build_model('dncnn_25', seed=0).cuda()
    fused = df.FusedDnCNN(net, residual=True)
    net16 = denoisers.build_model('dncnn_25', seed=0).cuda().to(torch.bfloat16).to(memory_format=torch.channels_last)
    x = torch.rand(2, 1, 256, 256, device='cuda')
    got = fused(x)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        want = net(x)
    ref16 = net16(x.to(torch.bfloat16)).float()
    print(f'network: ours vs fp32 {rel(x - got, x - want):.3e} (n(x) part), torch bf16 vs fp32 {rel(x - ref16, x - want):.3e}', flush=True)
    if bad:
        print('LAYER CHECK FAILED - skipping timing')
        return 1
    H = W = 256
    flops_mid = 2.0 * 64 * 64 * 9 * H * W
    for B in Bs:
        x = torch.rand(B, 1, H, W, device='cuda')
        x16 = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        t_ours = timed(lambda: fused(x))
        t_torch = timed(lambda: net16(x16))
        fl = B * (15 * flops_mid + 2.0 * 64 * 9 * H * W * 2)
        print(f'B={B}: ours {t_ours[0]:.3f} ms (avg {t_ours[1]:.3f}) = {fl / t_ours[0] / 1e9:.0f} TFLOP/s | '
              f'torch bf16 cuDNN {t_torch[0]:.3f} ms = {fl / t_torch[0] / 1e9:.0f} TFLOP/s | speed-up {t_torch[0] / t_ours[0]:.2f}x', flush=True)
        # one middle layer alone
        a = torch.randn(B, H, W, 64, device='cuda').to(torch.bfloat16)
        w = torch.randn(64, 64, 3, 3, device='cuda') / 24
        b = torch.zeros(64, device='cuda')
        wp = df.pack_conv64(w)
        out = torch.empty_like(a)
        lib = fused.lib
        st = torch.cuda.current_stream().cuda_stream
        t_l = timed(lambda: lib.pnpadmm_conv64_bf16(a.data_ptr(), out.data_ptr(), wp.data_ptr(), b.data_ptr(), B, H, W, 1, st))
        print(f'      one 64->64 layer: {t_l[0] * 1e3:.1f} us = {B * flops_mid / t_l[0] / 1e9:.0f} TFLOP/s, '
              f'{2 * a.numel() * 2 / t_l[0] / 1e6:.0f} GB/s of activations', flush=True)
    # IRCNN (7 layers, dilations 1, 2, 3, 4, 3, 2, 1)
    inet = denoisers.build_model('ircnn_gray', seed=0).cuda()
    ifused = df.FusedDnCNN(inet, residual=True)
    inet16 = denoisers.build_model('ircnn_gray', seed=0).cuda().to(torch.bfloat16).to(memory_format=torch.channels_last)
    x = torch.rand(2, 1, 256, 256, device='cuda')
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        want = inet(x)
    print(f'IRCNN: ours vs fp32 {rel(x - ifused(x), x - want):.3e} (n(x) part), torch bf16 vs fp32 '
          f'{rel(x - inet16(x.to(torch.bfloat16)).float(), x - want):.3e}', flush=True)
    for B in Bs:
        x = torch.rand(B, 1, H, W, device='cuda')
        x16 = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        t_ours = timed(lambda: ifused(x))
        t_torch = timed(lambda: inet16(x16))
        fl = B * (5 * flops_mid + 2.0 * 64 * 9 * H * W * 2)
        print(f'IRCNN B={B}: ours {t_ours[0]:.3f} ms (avg {t_ours[1]:.3f}) = {fl / t_ours[0] / 1e9:.0f} TFLOP/s | '
              f'torch bf16 cuDNN {t_torch[0]:.3f} ms | speed-up {t_torch[0] / t_ours[0]:.2f}x', flush=True)
        a = torch.randn(B, H, W, 64, device='cuda').to(torch.bfloat16)
        wp = df.pack_conv64(torch.randn(64, 64, 3, 3, device='cuda') / 24)
        b = torch.zeros(64, device='cuda')
        out = torch.empty_like(a)
        st = torch.cuda.current_stream().cuda_stream
        for dil in (2, 4):
            t_l = timed(lambda: ifused.lib.pnpadmm_conv64_dilated_bf16(a.data_ptr(), out.data_ptr(), wp.data_ptr(), b.data_ptr(), B, H, W, 1, dil, st))
            print(f'      one 64->64 layer, dilation {dil}: {t_l[0] * 1e3:.1f} us = {B * flops_mid / t_l[0] / 1e9:.0f} TFLOP/s', flush=True)
    # FFDNet-15 (half resolution, 13 x conv64 between a thin first layer and a four-channel pixel-shuffled tail)
    fnet = denoisers.build_model('ffdnet_gray', seed=0).cuda()
    ffused = df.FusedFFDNet(fnet)
    fnet16 = denoisers.build_model('ffdnet_gray', seed=0).cuda().to(torch.bfloat16).to(memory_format=torch.channels_last)
    sig16 = torch.full((1, 1, 1, 1), 15 / 255., device='cuda', dtype=torch.bfloat16)
    x = torch.rand(2, 1, 256, 256, device='cuda')
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        want = fnet(x, sig16.float())
    print(f'FFDNet: ours vs fp32 {rel(ffused(x, 15 / 255.), want):.3e}, torch bf16 vs fp32 {rel(fnet16(x.to(torch.bfloat16), sig16).float(), want):.3e}',
          flush=True)
    for B in Bs:
        x = torch.rand(B, 1, H, W, device='cuda')
        x16 = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        t_ours = timed(lambda: ffused(x, 15 / 255.))
        t_torch = timed(lambda: fnet16(x16, sig16))
        fl = B * (H // 2) * (W // 2) * 2.0 * 9 * (13 * 64 * 64 + 5 * 64 + 64 * 4)
        print(f'FFDNet B={B}: ours {t_ours[0]:.3f} ms (avg {t_ours[1]:.3f}) = {fl / t_ours[0] / 1e9:.0f} TFLOP/s | '
              f'torch bf16 cuDNN {t_torch[0]:.3f} ms | speed-up {t_torch[0] / t_ours[0]:.2f}x', flush=True)
    return 0


if __name__ == '__main__':
    t0 = time.time()
    rc = main()
    print(f'done in {time.time() - t0:.1f} s')
    sys.exit(rc)
