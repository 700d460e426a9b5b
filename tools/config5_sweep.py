"""BASELINE config 5: ADMM-CNC sweep 256^2 - 1024^2, synthetic phantoms, 512 images per GPU (4096 over 8 GPUs),
cluster-resident (hybrid) vs streaming kernel.  Run alone (1 GPU) or under torchrun (one rank per GPU, no collective
on the data path; max-over-ranks device time).  python tools/config5_sweep.py [B_per_gpu] [iters]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import pnp_admm_cnc_mri_b200 as pk
from pnp_admm_cnc_mri_b200 import data

world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
P = dict(alpha=0.45, lambda1=0.5, reo=0.05, b=64)
rows = []
for N, kernels in ((256, ('auto', 'cluster', 'streaming')), (512, ('streaming',)), (1024, ('streaming',))):
    imgs = np.stack([data.phantom(N, 1000 * rank + i) for i in range(4)] * (B // 4)).astype(np.float32)
    m = data.make_mask('radial', N, seed=1)
    s = pk.AdmmSolver(B, N)
    y = s.acquire(imgs, m, data.make_noise(N, seed=5))
    z0 = s.zero_filled(y)
    s.prepare(y, m, P['reo'])
    x = torch.empty_like(z0)
    for kernel in kernels:
        ts = []
        for r in range(3):
            z, w = z0.clone(), torch.zeros_like(z0)
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            s.iterate(x, z, w, 'cnc', iters, P['lambda1'], P['reo'], P['alpha'], P['b'], kernel=kernel)
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device='cuda', dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        t = float(np.median(ts[1:]))
        its = world * B * iters / (t * 1e-3)
        rows.append(dict(N=N, kernel=kernel, n_gpus=world, images=world * B, iters=iters, ms=t, iterations_per_s=its,
                         images_per_s=world * B / (t * 1e-3), moved_GBps_per_gpu=its / world * 36.5 * N * N / 1e9,
                         nominal_fft_TFLOPs_per_gpu=its / world * 10 * N * N * np.log2(N * N) / 1e12))
    del s, y, z0, x, z, w
    torch.cuda.empty_cache()
if rank == 0:
    for r in rows:
        print(json.dumps(r))
if world > 1:
    dist.destroy_process_group()
