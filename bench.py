#!/usr/bin/env python
"""bench.py — headline benchmark of the ADMM-CNC reconstruction path (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one pass of the hot path over one batch: ADMM-CNC (reference defaults S4:176:
alpha 0.45, 50 iterations, lambda 0.5, reo 0.05, b 64) on 64 synthetic 256x256 phantoms per GPU
with one 30 % sampling mask (cartesian / radial / random cycling step by step) and complex
Gaussian noise.  metric = ADMM iterations/s = images x iter_num / time (whole job, all ranks).

  value    : device-resident images -> acquisition + zero-fill + prepare + 50 iterations, CUDA events
  e2e      : same through the host-buffer C-ABI call (pnpadmm_reconstruct_host_f32): pinned host
             uint8 images in, float32 reconstructions out, copies inside the timed region
  roofline : the cluster-resident kernel alone (K1) against the non-tensor FP32 peak, nominal FFT
             flops 10 N^2 log2(N^2) per image-iteration
  cpu_baseline / --impl reference : the reference's NumPy loop (oracle restatement, bit-identical
             to the unmodified scripts) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N = 256
B_PER_GPU = 64
CNC = dict(alpha=0.45, iter_num=50, lambda1=0.5, reo=0.05, b=64)      # S4:176
MASK_KINDS = ('cartesian', 'radial', 'random')
WORKLOAD = ('BASELINE config 2: ADMM-CNC 256x256, batch 64 per GPU, 30% cartesian/radial/random masks cycling per step, '
            'reference defaults (alpha .45, 50 it, lambda .5, reo .05, b 64)')
FLOP_PER_IMAGE_ITER = 10 * N * N * math.log2(N * N)                   # 10 485 760
FP32_LANES_PER_SM = 128


def synth_inputs(B, seed0=0):
    from pnp_admm_cnc_mri_b200 import data
    base = [np.uint8((data.phantom(N, seed0 + i) * 255.0).round()) for i in range(min(B, 16))]
    imgs = np.stack([base[i % len(base)] for i in range(B)])             # uint8 gray levels
    masks = np.stack([data.make_mask(k, N, seed=s) for s, k in enumerate(MASK_KINDS)])
    noise = data.make_noise(N, seed=1234)                                # sigma 15 per component (= noises.mat x 3)
    return imgs, masks, noise


# --------------------------------------------------------------------------------------------
# CPU reference arm (oracle restatement of S4:101-132, NumPy fp64)
# --------------------------------------------------------------------------------------------
def _cpu_one(args):
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from oracle import reference_numpy as orc
    img, mask, noise = args
    x = orc.admm_cnc(np.float32(img / 255.), mask.astype(np.float64), noise, **CNC)
    return float(x.sum())


def cpu_throughput(n_images, cores):
    """iterations/s of the reference loop over `n_images` images on 1 process (cpu_baseline leg)."""
    imgs, masks, noise = synth_inputs(n_images)
    jobs = [(imgs[i], masks[i % 3], noise) for i in range(n_images)]
    _cpu_one(jobs[0])                      # warm-up (imports, FFT plans)
    t0 = time.perf_counter()
    for j in jobs:
        _cpu_one(j)
    dt = time.perf_counter() - t0
    return n_images * CNC['iter_num'] / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation on all host cores (one image per task)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_img = max(cores * 2, 16)
    imgs, masks, noise = synth_inputs(n_img)
    jobs = [(imgs[i], masks[i % 3], noise) for i in range(n_img)]
    t_tot = 0.0
    with mp.get_context('fork').Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_one, jobs[:cores], chunksize=1)
        for _ in range(args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_one, jobs, chunksize=1)
            t_tot += time.perf_counter() - t0
    value = n_img * CNC['iter_num'] * args.steps / t_tot
    line = {
        'impl': 'reference', 'metric': 'admm_cnc_iterations_per_s', 'value': value, 'unit': 'iterations/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_tot / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'batch_per_gpu': B_PER_GPU, 'iter_num': CNC['iter_num'],
                   'sample': f'{n_img} images of that workload per step (bounded CPU sample), one process per core'},
        'cpu_baseline': {'value': value, 'unit': 'iterations/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{n_img} images x 50 iterations per step, one process per core (NumPy fp64, '
                                   f'oracle restatement bit-identical to the unmodified reference script)'},
        'e2e': {'value': value, 'unit': 'iterations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [s.strip() for s in r.split(',')]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    import pnp_admm_cnc_mri_b200 as pk
    from pnp_admm_cnc_mri_b200 import _abi

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('for --gpus N > 1 launch with torch.distributed.run --nproc-per-node N')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    lib = _abi.load()
    dev = torch.device('cuda', local)

    B = B_PER_GPU
    imgs_u8, masks, noise = synth_inputs(B, seed0=1000 * rank)
    solver = pk.AdmmSolver(B, N)
    d_imgs = torch.as_tensor(np.float32(imgs_u8 / 255.)).to(dev)
    d_masks = torch.as_tensor(masks).to(dev)
    d_noise = torch.as_tensor(noise).to(dev, torch.complex64)
    # host-side buffers of the e2e path (pinned)
    h_img = torch.as_tensor(imgs_u8).pin_memory()
    h_masks = [torch.as_tensor(masks[i]).pin_memory() for i in range(3)]
    h_noise = torch.view_as_real(torch.as_tensor(noise).to(torch.complex64)).contiguous().pin_memory()
    h_x = torch.empty((B, N, N), dtype=torch.float32).pin_memory()
    scratch = torch.empty(lib.pnpadmm_host_scratch_bytes(B, N), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(i):
        y = solver.acquire(d_imgs, d_masks[i % 3], d_noise)
        return solver.solve(y, d_masks[i % 3], 'cnc', CNC['iter_num'], CNC['lambda1'], CNC['reo'], CNC['alpha'], CNC['b'])

    def step_host(i):
        st = torch.cuda.current_stream().cuda_stream
        _abi.check(lib.pnpadmm_reconstruct_host_f32(
            h_img.data_ptr(), h_masks[i % 3].data_ptr(), h_noise.data_ptr(), h_x.data_ptr(), B, N, _abi.PROX_CNC,
            CNC['iter_num'], CNC['lambda1'], CNC['reo'], CNC['alpha'], CNC['b'], _abi.KERNEL_AUTO,
            scratch.data_ptr(), scratch.numel(), solver.ws.data_ptr(), solver.ws_bytes, st))
        torch.cuda.current_stream().synchronize()        # the user reads h_x
        return float(h_x[0, 0, 0])

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for i in range(steps):
            flush.fill_(i & 1)                            # L2 flush between timed iterations (untimed)
            evs[i][0].record()
            fn(i)
            evs[i][1].record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # pipelined host-buffer arm: copies on their own streams, two slots (pnpadmm_reconstruct_host_pipelined_f32)
    pscratch = torch.empty(lib.pnpadmm_host_pipeline_scratch_bytes(B, N), dtype=torch.uint8, device=dev)
    h_x2 = [torch.empty((B, N, N), dtype=torch.float32).pin_memory() for _ in range(2)]
    s_c, s_i, s_o = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed_pipelined(steps, warmup):
        def enqueue(i):
            _abi.check(lib.pnpadmm_reconstruct_host_pipelined_f32(
                h_img.data_ptr(), h_masks[i % 3].data_ptr(), h_noise.data_ptr(), h_x2[i & 1].data_ptr(), B, N, _abi.PROX_CNC,
                CNC['iter_num'], CNC['lambda1'], CNC['reo'], CNC['alpha'], CNC['b'], _abi.KERNEL_AUTO,
                pscratch.data_ptr(), pscratch.numel(), solver.ws.data_ptr(), solver.ws_bytes, i & 1,
                s_c.cuda_stream, s_i.cuda_stream, s_o.cuda_stream))

        def collect(slot):
            _abi.check(lib.pnpadmm_reconstruct_host_wait(slot))
            return float(h_x2[slot][0, 0, 0])             # the user reads the step's result

        for i in range(warmup):
            enqueue(i)
            collect(i & 1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s_i)
        for i in range(steps):
            if i >= 2:
                collect(i & 1)                             # step i-2 has landed before its slot is reused
            enqueue(i)
        for i in range(max(steps - 2, 0), steps):
            collect(i & 1)
        e1.record(s_o)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_device, args.steps, args.warmup)
    ms_e2e_sync = timed(step_host, args.steps, args.warmup)
    ms_e2e = timed_pipelined(args.steps, args.warmup)

    # K1 alone for the roofline: iterate() on prepared state, one launch per timed region
    y = solver.acquire(d_imgs, d_masks[0], d_noise)
    z0 = solver.zero_filled(y)
    solver.prepare(y, d_masks[0], CNC['reo'])
    x = torch.empty_like(z0)
    def time_iterate(kernel):
        ts = []
        for r in range(args.warmup + args.steps):
            z, w = z0.clone(), torch.zeros_like(z0)
            flush.fill_(r & 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            solver.iterate(x, z, w, 'cnc', CNC['iter_num'], CNC['lambda1'], CNC['reo'], CNC['alpha'], CNC['b'], kernel=kernel)
            e1.record()
            torch.cuda.synchronize()
            if r >= args.warmup:
                ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts))

    k1_ms = time_iterate('cluster')        # all 32 planes on the cluster kernel: the K1 roofline leg
    hyb_ms = time_iterate('auto')          # what the step runs: K1 on 28 planes + K2 on 4, concurrently

    # K2 (streaming kernels) against the HBM roofline: ADMM-CNC at N = 1024, 64 images per GPU, 10 iterations
    # per launch sequence (state + workspace = 2.3 GB >> L2, so every pass streams from HBM)
    k2 = None
    if rank == 0:
        from pnp_admm_cnc_mri_b200 import data as pdata
        N2, B2, IT2 = 1024, 64, 10
        s2 = pk.AdmmSolver(B2, N2)
        im2 = np.stack([pdata.phantom(N2, i) for i in range(4)] * (B2 // 4)).astype(np.float32)
        m2 = pdata.make_mask('random', N2, seed=0)
        y2 = s2.acquire(im2, m2, pdata.make_noise(N2, seed=7))
        z20 = s2.zero_filled(y2)
        s2.prepare(y2, m2, CNC['reo'])
        x2 = torch.empty_like(z20)
        t2 = []
        for r in range(5):
            z2, w2 = z20.clone(), torch.zeros_like(z20)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            s2.iterate(x2, z2, w2, 'cnc', IT2, CNC['lambda1'], CNC['reo'], CNC['alpha'], CNC['b'], kernel='streaming')
            e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                t2.append(e0.elapsed_time(e1))
        k2 = dict(N=N2, B=B2, iters=IT2, ms=float(np.mean(t2)))
        del s2, y2, z20, x2, z2, w2
        torch.cuda.empty_cache()
    # K5 (DnCNN on tcgen05, BASELINE config 3's denoiser) against the bf16 tensor peak: one 64->64 layer and one
    # DnCNN-17 forward at B = 256, 256x256 (activations 2 x 2.1 GB >> L2)
    clocks = sampler.stop() if rank == 0 else None     # the headline legs end here; the tensor-core leg below samples its own clocks
    k5 = None
    if rank == 0:
        from pnp_admm_cnc_mri_b200 import denoisers as pden, dncnn_fused as pdf
        sampler5 = ClockSampler(local)
        sampler5.start()
        B5 = 256
        net = pden.build_model('dncnn_25', seed=0)
        fused = pdf.FusedDnCNN(net, residual=True, device=dev)
        x5 = torch.rand(B5, 1, N, N, device=dev)
        a5 = torch.randn(B5, N, 8, N, 8, device=dev).to(torch.bfloat16)
        o5 = torch.empty_like(a5)
        w5 = pdf.pack_conv64(torch.randn(64, 64, 3, 3, device=dev) / 24)
        b5 = torch.zeros(64, device=dev)
        st5 = torch.cuda.current_stream().cuda_stream

        def t_of(fn, reps=5, warm=2):
            ts = []
            for r in range(warm + reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                if r >= warm:
                    ts.append(e0.elapsed_time(e1))
            return float(np.mean(ts))
        k5 = dict(B=B5, layer_ms=t_of(lambda: _abi.check(lib.pnpadmm_conv64_bf16(a5.data_ptr(), o5.data_ptr(), w5.data_ptr(),
                                                                               b5.data_ptr(), B5, N, N, 1, st5))),
                  forward_ms=t_of(lambda: fused(x5)))
        k5['clocks'] = sampler5.stop()
        del fused, x5, a5, o5
        torch.cuda.empty_cache()

    fl = ctypes.c_double()
    _abi.check(lib.pnpadmm_measure_fp32_peak(fl, None))
    sm, ncl = ctypes.c_int(), ctypes.c_int()
    _abi.check(lib.pnpadmm_device_info(sm, ncl, None, None))

    if rank == 0:
        its_step = world * B * CNC['iter_num']
        value = its_step * args.steps / (ms_dev * 1e-3)
        e2e = its_step * args.steps / (ms_e2e * 1e-3)
        achieved = B * CNC['iter_num'] * FLOP_PER_IMAGE_ITER / (k1_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        sm_max = peaks.get('sm_max_mhz', 1965.0)
        nominal_peak = sm.value * FP32_LANES_PER_SM * 2 * sm_max * 1e6 / 1e12
        cores = os.cpu_count() or 1
        cpu_v, cpu_dt = cpu_throughput(8, 1) if world == 1 else (None, None)
        line = {
            'metric': 'admm_cnc_iterations_per_s', 'value': value, 'unit': 'iterations/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'images_per_s': value / CNC['iter_num'],
            'config': {'workload': WORKLOAD, 'batch_per_gpu': B, 'iter_num': CNC['iter_num'], 'kernel': 'hybrid: cluster256 (K1, 16-CTA clusters, 2 CTAs/SM) on 28 of the 32 packed planes + K2 rows2/cols2 for the other 4 on the SMs K1 cannot use; acquisition / zero-fill: K2',
                       'l2': 'flushed (256 MiB fill) between timed steps', 'sharding': f'batch x{world}, no collective'},
            'e2e': {'value': e2e, 'unit': 'iterations/s', 'ms_per_step': ms_e2e / args.steps,
                    'h2d_bytes_per_step': int(h_img.numel() + h_masks[0].numel() + h_noise.numel() * 4),
                    'd2h_bytes_per_step': int(h_x.numel() * 4),
                    'api': 'pnpadmm_reconstruct_host_pipelined_f32 (pinned host buffers in and out every step; H2D / kernels / '
                           'D2H on three streams, two device slots, so the copies of neighbouring steps overlap the kernels; '
                           'every result is waited for and read on the host inside the timed region)',
                    'sync_call': {'value': its_step * args.steps / (ms_e2e_sync * 1e-3), 'ms_per_step': ms_e2e_sync / args.steps,
                                  'api': 'pnpadmm_reconstruct_host_f32: one stream, host synchronises after every step '
                                         '(copies not overlapped; single-call latency)'}},
            'gpu_launches': 110 * args.steps,
            'launches_per_step': {'device': 110, 'e2e': 111,
                                  'kernels': 'rows2<FWD_IMG>, cols2<FWD_ACQ>, cols2<INV>, rows2<INV_ABS>, copy_zero, write_cf, '
                                             'prepare, pack_mcode, cluster256 x1, and for the K2 share rows2<FWD_ZW> x1 + (cols2<BLEND> + rows2<PROX>) x50 '
                                             '(+ u8_to_unit in e2e)'},
            'roofline': {'bound': 'fp32', 'kernel': 'cluster256_kernel', 'achieved': achieved, 'peak': nominal_peak,
                         'unit': 'TFLOP/s', 'frac': achieved / nominal_peak, 'traffic': None,
                         'peak_source': f'nominal non-tensor FP32: {sm.value} SMs x 128 lanes x 2 x {sm_max:.0f} MHz '
                                        f'(MEASURED_PEAKS.json has no fp32 entry); measured FFMA probe in this run: '
                                        f'{fl.value / 1e12:.1f} TFLOP/s',
                         'measured_fma_peak': fl.value / 1e12, 'frac_of_measured_fma': achieved / (fl.value / 1e12),
                         'flop_model': '10 N^2 log2(N^2) = 10485760 per image-iteration (nominal radix-2 count of 2 '
                                       'complex 2-D FFTs; executed flops are lower: pair-packing + radix-16)',
                         'launch_ms': k1_ms, 'resident_clusters': ncl.value,
                         'hybrid_iterate': {
                             'call_ms': hyb_ms,
                             'achieved': B * CNC['iter_num'] * FLOP_PER_IMAGE_ITER / (hyb_ms * 1e-3) / 1e12,
                             'frac': B * CNC['iter_num'] * FLOP_PER_IMAGE_ITER / (hyb_ms * 1e-3) / 1e12 / nominal_peak,
                             'what': 'the iterate call of the timed step (kernel=auto): cluster256 on the planes that fill '
                                     'whole cluster rounds, K2 rows2/cols2 for the rest on a side stream, on the SMs outside '
                                     'the clusters; same FLOP model, all 148 SMs'}},
            'clocks': clocks,
        }
        hbm_peak = peaks.get('hbm_gbs', 6450.0)
        its2 = k2['B'] * k2['iters'] / (k2['ms'] * 1e-3)
        moved = its2 * 36.5 * k2['N'] ** 2 / 1e9
        line['roofline_streaming'] = {
            'bound': 'hbm', 'kernel': 'rows2_kernel<1024> + cols2_tma_kernel<1024> (K2, one pair per iteration; column tiles loaded by 2-D TMA)',
            'achieved': moved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': moved / hbm_peak,
            'traffic': 2319e6 * k2['B'] / 64,
            'traffic_source': 'profiles/r1_k2_tma_ncu_summary.txt: dram read+write of one rows2 + one cols2 launch at N=1024, '
                              'B=64 (1551 MB + 763 MB; r1_k2v2_ncu_summary.txt: 1554 + 765), scaled by B/64',
            'bytes_model': '73 N^2 per packed plane-iteration = 36.5 N^2 per image-iteration (rows pass 48 B/px: K 8+8, '
                           'z,w of two images 16+16; cols pass 25 B/px: K 8+8, G 8, codes 1/4): DESIGN.md 5',
            'achieved_survey_q57': its2 * 57 * k2['N'] ** 2 / 1e9,
            'survey_model': 'SURVEY 8d counts 57 N^2 per image-iteration for an unpaired implementation; pairing two '
                            'real images per complex plane moves 36.5 N^2, so the survey-model figure can exceed the peak',
            'workload': f"ADMM-CNC N={k2['N']}, B={k2['B']}, {k2['iters']} iterations per timed call", 'call_ms': k2['ms'],
            'iterations_per_s': its2,
            'peak_source': 'MEASURED_PEAKS.json hbm_gbs (copy bandwidth)' if 'hbm_gbs' in peaks else 'fallback 6450 GB/s',
        }
        layer_flop = 2.0 * 64 * 64 * 9 * N * N * k5['B']
        fwd_flop = 2.0 * 555137 * N * N * k5['B']
        tpeak = peaks.get('bf16_tflops', 1670.0)
        line['roofline_tensor'] = {
            'bound': 'tensor', 'kernel': 'conv64_tc_kernel<64> (K5: one DnCNN conv3x3 64->64 + bias + ReLU layer, tcgen05 implicit GEMM)',
            'achieved': layer_flop / (k5['layer_ms'] * 1e-3) / 1e12, 'peak': tpeak, 'unit': 'TFLOP/s',
            'frac': layer_flop / (k5['layer_ms'] * 1e-3) / 1e12 / tpeak, 'traffic': 4293e6,
            'traffic_source': 'profiles/r1_k5_conv64_tc_ncu_summary.txt: dram read 2198 MB + write 2095 MB per launch at B=256 '
                              '(algorithmic: 2147 MB in x 1.05 halo + 2147 MB out)',
            'flop_model': '2 x 64 x 64 x 9 = 73728 FLOP per pixel and layer', 'launch_ms': k5['layer_ms'],
            'hbm_gbs_moved': 2 * 128.0 * N * N * k5['B'] / (k5['layer_ms'] * 1e-3) / 1e9,
            'peak_source': 'MEASURED_PEAKS.json bf16_tflops (cuBLAS burst)' if 'bf16_tflops' in peaks else 'fallback 1670 TFLOP/s',
            'workload': f"B={k5['B']}, 256x256, bf16 operands, fp32 accumulation",
            'clocks': k5['clocks'],
            'dncnn17_forward': {'ms': k5['forward_ms'], 'achieved': fwd_flop / (k5['forward_ms'] * 1e-3) / 1e12,
                                'what': 'head (CUDA cores) + 15 x conv64_tc_kernel<64> + tail conv64_tc_kernel<16>; BASELINE config 3 '
                                        'runs two of these per PnP-ADMM-CNC iteration (tools/pnp_bench.py c3)'}}
        if cpu_v is not None:
            line['cpu_baseline'] = {'value': cpu_v, 'unit': 'iterations/s', 'cores': 1, 'kind': 'port',
                                    'sample': '8 images x 50 iterations, 1 thread, NumPy fp64 oracle restatement '
                                              '(bit-identical to the unmodified reference script); host has '
                                              f'{cores} cores, all-cores number: bench.py --impl reference'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
