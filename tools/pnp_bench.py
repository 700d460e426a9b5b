"""BASELINE configs 3 and 4 (PnP variants, random-init bf16 denoisers): time per ADMM iteration and the split
between the denoiser (PyTorch bf16 tensor cores) and our glue kernels.  python tools/pnp_bench.py [c3|c4] [B] [iters]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pnp_admm_cnc_mri_b200 as pk
from pnp_admm_cnc_mri_b200 import data, pnp, denoisers

which = sys.argv[1] if len(sys.argv) > 1 else 'c3'
FUSED = os.environ.get('PNP_FUSED', '1') != '0'      # PNP_FUSED=0: stock PyTorch bf16 (cuDNN) denoiser, the A/B baseline
dev = torch.device('cuda', 0)


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


if which == 'c3':
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    N = 256
    imgs = torch.as_tensor(np.stack([data.phantom(N, i) for i in range(8)] * (B // 8)).astype(np.float32)).to(dev)
    m = data.make_mask('random', N, seed=0)
    nz = data.make_noise(N, seed=3)
    D = denoisers.build_denoiser('dncnn_25', iter_num=iters, device=dev, fused=None if FUSED else False)   # DnCNN-17, nc=64, random init, bf16
    t_all = timed(lambda: pnp.pnp_admm_cnc(imgs, m, nz, D, D, alpha=1.2, iter_num=iters, lambda1=4, reo=0.45, b=0.3))
    x = torch.rand(B, 1, N, N, device=dev)
    t_den = timed(lambda: D(x, 0))
    flops = 2 * 555137 * N * N * B          # per forward (2 x params x pixels)
    print(json.dumps({'config': 3, 'workload': f'PnP-ADMM-CNC DnCNN-17 bf16, B={B}, 256x256', 'iters_timed': iters,
                      'denoiser': 'tcgen05 kernels (csrc/dncnn_tc.cuh)' if D.fused is not None else 'PyTorch bf16 channels_last (cuDNN)',
                      'ms_per_iteration': t_all / iters, 'image_iterations_per_s': B * iters / (t_all * 1e-3),
                      'denoiser_forward_ms': t_den, 'denoiser_share': 2 * t_den * iters / t_all,
                      'denoiser_tflops': flops / (t_den * 1e-3) / 1e12}))
else:
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    N = 512
    imgs = torch.as_tensor(np.stack([data.phantom(N, i) for i in range(4)] * (B // 4)).astype(np.float32)).to(dev)
    m = data.make_mask('random', N, seed=0)
    nz = data.make_noise(N, seed=3)
    D = denoisers.build_denoiser('drunet_gray', iter_num=iters, x8=True, device=dev)
    t_all = timed(lambda: pnp.pnp_admm_l1(imgs, m, nz, D, iter_num=iters, reo=0.26), reps=2)
    x = torch.rand(B, 1, N, N, device=dev)
    t_den = timed(lambda: D(x, 1), reps=2)
    print(json.dumps({'config': 4, 'workload': f'PnP-ADMM-L1 DRUNet bf16 (4 x 288^2 quadrants, x8 schedule), B={B}, 512x512',
                      'iters_timed': iters, 'ms_per_iteration': t_all / iters,
                      'image_iterations_per_s': B * iters / (t_all * 1e-3), 'denoiser_call_ms': t_den,
                      'denoiser_share': t_den * iters / t_all}))
