#!/usr/bin/env python
"""bench.py — headline benchmark of the ADMM-CNC reconstruction path (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--legs all|headline]

A "step" = one pass of the hot path over one batch: ADMM-CNC (reference defaults S4:176:
alpha 0.45, 50 iterations, lambda 0.5, reo 0.05, b 64) on 64 synthetic 256x256 phantoms per GPU
with one 30 % sampling mask (cartesian / radial / random cycling step by step) and complex
Gaussian noise.  metric = ADMM iterations/s = images x iter_num / time (whole job, all ranks).

  value     : device-resident images -> pnpadmm_reconstruct_f32 (acquisition + zero-fill + data term + 50 iterations), CUDA events
  e2e       : same through the host-buffer pipeline (pnp_admm_cnc_mri_b200.HostPipeline over
              pnpadmm_reconstruct_host_pipelined_f32): pinned host uint8 images in, float32
              reconstructions out, copies inside the timed region
  sustained : the same two steps looped for >= 2 s each, with SM clock / power / throttle reasons sampled
  roofline  : the cluster-resident kernel alone (K1) against the non-tensor FP32 peak, nominal FFT
              flops 10 N^2 log2(N^2) per image-iteration
  legs      : the other BASELINE configs, run by every rank at every --gpus N so that the scaling run sees the
              streaming (K2) and tensor-core (K5) kernels too: config 5 (512 images per GPU, N = 256 / 512 /
              1024), config 3 (PnP-ADMM-CNC, DnCNN on the tcgen05 kernels, 256 images per GPU), config 4
              (PnP-ADMM-L1, DRUNet, 512x512, 16 images per GPU) and, for N > 1, the final all_gather
  cpu_baseline / --impl reference : the reference's NumPy loop (oracle restatement, bit-identical
              to the unmodified scripts) on the host cores.
"""
from __future__ import annotations

import argparse
import hashlib
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
CSRC = os.path.join(ROOT, 'pnp_admm_cnc_mri_b200', 'csrc')


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

    def __init__(self, index, period_ms=100):
        self.index, self.rows, self.proc, self.period = index, [], None, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', str(self.period)],
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
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [s.strip() for s in r.split(',')]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            try:
                pw.append(float(f[6]))
            except Exception:
                pass
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm),
                'power_w_median': float(np.median(pw)) if pw else None, 'power_w_max': max(pw) if pw else None}


def source_sha(files):
    """sha256 over kernel source files: ties a number copied from an ncu capture to the code it was captured on."""
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(CSRC, f), 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def recorded_traffic(key):
    """DRAM bytes per launch from profiles/traffic.json (written from `ncu --set full` captures), or (None, why) when the
    kernel sources have changed since the capture."""
    try:
        rec = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))[key]
    except Exception as e:
        return None, f'no record ({type(e).__name__})'
    sha = source_sha(rec['sources'])
    if sha != rec['sources_sha16']:
        return None, f"stale: kernel sources changed since the capture ({rec['capture']}); re-run the ncu capture"
    return rec, rec['capture']


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    import pnp_admm_cnc_mri_b200 as pk
    from pnp_admm_cnc_mri_b200 import _abi
    from pnp_admm_cnc_mri_b200 import data as pdata

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('for --gpus N > 1 launch with torch.distributed.run --nproc-per-node N')
    # one core set per rank: the ranks' launch threads and pinned-copy bookkeeping do not migrate onto each other
    if world > 1 and hasattr(os, 'sched_setaffinity') and os.environ.get('BENCH_NO_PIN') is None:
        try:
            cpus = sorted(os.sched_getaffinity(0))
            per = max(1, len(cpus) // world)
            mine = cpus[local * per:(local + 1) * per] or cpus
            os.sched_setaffinity(0, mine)
        except Exception:
            pass
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    lib = _abi.load()
    dev = torch.device('cuda', local)
    full = args.legs == 'all'

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

    def max_over_ranks(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def step_device(i):
        return solver.reconstruct(d_imgs, d_masks[i % 3], d_noise, 'cnc', CNC['iter_num'], CNC['lambda1'], CNC['reo'], CNC['alpha'],
                                  CNC['b'])

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
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))

    # pipelined host-buffer arm: the package's HostPipeline (copies on their own streams, device slots, captured compute graph)
    SLOTS = 3
    pipe = pk.HostPipeline(B, N, n_slots=SLOTS)
    h_xs = [torch.empty((B, N, N), dtype=torch.float32).pin_memory() for _ in range(SLOTS)]
    enq_s = []                                            # host time spent enqueueing one step (diagnostic)

    def timed_pipelined(steps, warmup, min_seconds=0.0, pipe=pipe, h_xs=h_xs):
        def enqueue(i):
            t0 = time.perf_counter()
            pipe.submit(i % SLOTS, h_img, h_masks[i % 3], h_noise, h_xs[i % SLOTS], prox='cnc', **CNC)
            enq_s.append(time.perf_counter() - t0)

        def collect(i):
            pipe.wait(i % SLOTS)
            return float(h_xs[i % SLOTS][0, 0, 0])        # the user reads the step's result

        for i in range(max(warmup, 2 * SLOTS) if warmup else 0):   # every slot twice: the second use captures its compute graph
            enqueue(i)
            collect(i)
        barrier()
        del enq_s[:]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(pipe.s_h2d)
        t_start = time.perf_counter()
        i = 0
        while i < steps or (min_seconds and time.perf_counter() - t_start < min_seconds):
            if i >= SLOTS:
                collect(i)                                 # step i - SLOTS has landed before its slot is reused
            enqueue(i)
            i += 1
        done = i
        for j in range(max(done - SLOTS, 0), done):
            collect(j)
        e1.record(pipe.s_d2h)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), done

    def sustained_device(min_seconds):
        """The device step back to back for >= min_seconds (no L2 flush: the 48 MB state + 80 MB workspace do not fit in L2
        together with the next step's inputs anyway); CUDA events around the whole run."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < min_seconds:
            for _ in range(20):
                step_device(n)
                n += 1
            torch.cuda.current_stream().synchronize()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), n

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_device, args.steps, args.warmup)
    ms_e2e_sync = timed(step_host, args.steps, args.warmup)
    ms_e2e, _ = timed_pipelined(args.steps, args.warmup)
    enq_us = 1e6 * float(np.median(enq_s)) if enq_s else None
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e_u8 = None
    if full:   # same pipeline copying back img_E (uint8, what the reference saves, S1:133-138) instead of float32 x: 4x fewer D2H bytes
        pipe8 = pk.HostPipeline(B, N, n_slots=SLOTS, output='uint8')
        h_x8 = [torch.empty((B, N, N), dtype=torch.uint8).pin_memory() for _ in range(SLOTS)]
        ms_e2e_u8, _ = timed_pipelined(args.steps, args.warmup, pipe=pipe8, h_xs=h_x8)
        pipe8.close()
    # the step per mask kind (the headline cycles through the three): cartesian masks run on the row-separable kernel K3
    per_mask = {}
    if full:
        for mi, kind in enumerate(MASK_KINDS):
            n_k = max(6, args.steps // 2)
            per_mask[kind] = timed(lambda i, mi=mi: solver.reconstruct(d_imgs, d_masks[mi], d_noise, 'cnc', CNC['iter_num'], CNC['lambda1'],
                                                                      CNC['reo'], CNC['alpha'], CNC['b']), n_k, 3) / n_k

    sustained = None
    if full:
        s2 = ClockSampler(local)
        if rank == 0:
            s2.start()
        ms_sd, n_sd = sustained_device(2.0)
        ms_se, n_se = timed_pipelined(0, 0, min_seconds=2.0)
        # ranks loop for the same wall time, not the same step count: the job's steps are the sum over ranks
        sustained = dict(device_ms=ms_sd, device_steps=sum_over_ranks(n_sd) / world, e2e_ms=ms_se, e2e_steps=sum_over_ranks(n_se) / world,
                         clocks=s2.stop() if rank == 0 else None)

    # K1 alone for the roofline: iterate() on prepared state, one launch per timed region
    y = solver.acquire(d_imgs, d_masks[0], d_noise)
    z0 = solver.zero_filled(y)
    x = torch.empty_like(z0)

    def time_iterate(kernel, mask_i):
        solver.prepare(y, d_masks[mask_i], CNC['reo'])
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

    k1_ms = max_over_ranks(time_iterate('cluster', 2))    # all 32 planes on the cluster kernel: the K1 roofline leg (random mask)
    # FFMA probe right beside the K1 leg (same clock / power state; after the tensor-core legs below the GPU sits power-capped at
    # ~1.45 GHz for a while and the probe would read 53 instead of 72 TFLOP/s)
    fl = ctypes.c_double()
    _abi.check(lib.pnpadmm_measure_fp32_peak(fl, None))
    hyb_ms = max_over_ranks(time_iterate('auto', 2))      # what the step runs: K1 on most planes + K2 on the rest, concurrently
    del y, z0, x

    def ev_time(fn, reps, warm):
        ts = []
        for r in range(warm + reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            if r >= warm:
                ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts))

    # ---- config 5: ADMM-CNC, 512 images per GPU, N = 256 / 512 / 1024 (K1 + K2 hybrid at 256, K2 above; state >> L2) ----
    legs = {}
    if full:
        # K5 layer timed ALONE first, before the long legs heat the GPU into its power cap: the measurement that belongs beside the
        # burst bf16 peak (the same layer is timed again inside the config-3 section below, beside the sustained peak)
        from pnp_admm_cnc_mri_b200 import dncnn_fused as pdf0
        a0 = torch.randn(256, N, 8, N, 8, device=dev).to(torch.bfloat16)
        o0 = torch.empty_like(a0)
        w0 = pdf0.pack_conv64(torch.randn(64, 64, 3, 3, device=dev) / 24)
        b0 = torch.zeros(64, device=dev)
        st0 = torch.cuda.current_stream().cuda_stream
        k5_alone_ms = max_over_ranks(ev_time(lambda: _abi.check(lib.pnpadmm_conv64_bf16(
            a0.data_ptr(), o0.data_ptr(), w0.data_ptr(), b0.data_ptr(), 256, N, N, 1, st0)), 5, 2))
        del a0, o0
        torch.cuda.empty_cache()
        c5 = {}
        for N5 in (256, 512, 1024):
            B5 = 512
            base = np.stack([pdata.phantom(N5, 100 * rank + i) for i in range(4)]).astype(np.float32)
            im5 = torch.as_tensor(base).to(dev).repeat(B5 // 4, 1, 1)
            m5 = torch.as_tensor(pdata.make_mask(MASK_KINDS[(N5 // 256) % 3], N5, seed=N5)).to(dev)
            n5 = torch.as_tensor(pdata.make_noise(N5, seed=7)).to(dev, torch.complex64)
            s5 = pk.AdmmSolver(B5, N5)

            def solve5():
                return s5.reconstruct(im5, m5, n5, 'cnc', CNC['iter_num'], CNC['lambda1'], CNC['reo'], CNC['alpha'], CNC['b'])
            barrier()
            c5[N5] = dict(B=B5, ms=max_over_ranks(ev_time(solve5, 2, 1)), mask=MASK_KINDS[(N5 // 256) % 3])
            # the same batch under a Cartesian mask (full k-space lines): row-separable kernel K3, rows resident on chip
            m5c = torch.as_tensor(pdata.make_mask('cartesian', N5, seed=N5)).to(dev)
            solve5c = lambda: s5.reconstruct(im5, m5c, n5, 'cnc', CNC['iter_num'], CNC['lambda1'], CNC['reo'], CNC['alpha'], CNC['b'])
            barrier()
            c5[N5]['ms_cartesian'] = max_over_ranks(ev_time(solve5c, 2, 1))
            del s5, im5, m5, n5, m5c
            torch.cuda.empty_cache()
        legs['config5'] = c5

        # ---- config 3: PnP-ADMM-CNC, DnCNN-17 x 2 per iteration on the tcgen05 kernels (K5), 256 images per GPU ----
        from pnp_admm_cnc_mri_b200 import denoisers as pden, dncnn_fused as pdf, pnp as ppnp
        s3c = ClockSampler(local)
        if rank == 0:
            s3c.start()
        B3, IT3 = 256, 5
        im3 = torch.as_tensor(np.float32(imgs_u8 / 255.)).to(dev).repeat(B3 // B, 1, 1)
        D3 = pden.build_denoiser('dncnn_25', seed=0, device=dev)
        P3 = dict(alpha=1.2, lambda1=4.0, reo=0.45, b=0.3)               # S6:571
        barrier()
        ppnp.pnp_admm_cnc(im3, d_masks[2], d_noise, D3, D3, iter_num=1, device=dev, **P3)      # warm-up
        ms3 = max_over_ranks(ev_time(lambda: ppnp.pnp_admm_cnc(im3, d_masks[2], d_noise, D3, D3, iter_num=IT3, device=dev, **P3), 1, 0))
        legs['config3'] = dict(B=B3, iters=IT3, ms=ms3)
        # K5 against the bf16 tensor peak: one 64->64 layer and one DnCNN-17 forward at B = 256 (activations 2 x 2.1 GB >> L2)
        a5 = torch.randn(B3, N, 8, N, 8, device=dev).to(torch.bfloat16)
        o5 = torch.empty_like(a5)
        w5 = pdf.pack_conv64(torch.randn(64, 64, 3, 3, device=dev) / 24)
        b5 = torch.zeros(64, device=dev)
        x5 = torch.rand(B3, 1, N, N, device=dev)
        st5 = torch.cuda.current_stream().cuda_stream
        k5 = dict(B=B3,
                  layer_ms=max_over_ranks(ev_time(lambda: _abi.check(lib.pnpadmm_conv64_bf16(
                      a5.data_ptr(), o5.data_ptr(), w5.data_ptr(), b5.data_ptr(), B3, N, N, 1, st5)), 5, 2)),
                  forward_ms=max_over_ranks(ev_time(lambda: D3.fused(x5), 5, 2)))
        # the other 64-channel denoisers of the reference on the same kernels (IRCNN: dilated middle layers; FFDNet: half resolution,
        # thin first layer, pixel-shuffled tail), each against its stock PyTorch bf16 module (cuDNN, channels_last)
        k5['nets'] = {}
        for nm in ('ircnn_gray', 'ffdnet_gray'):
            Dk = pden.build_denoiser(nm, iter_num=50, seed=0, device=dev)
            Dt = pden.build_denoiser(nm, iter_num=50, seed=0, device=dev, fused=False)
            k5['nets'][nm] = dict(ms=max_over_ranks(ev_time(lambda: Dk(x5, 0), 5, 2)), torch_ms=max_over_ranks(ev_time(lambda: Dt(x5, 0), 3, 2)))
            del Dk, Dt
        k5['clocks'] = s3c.stop() if rank == 0 else None
        del D3, im3, a5, o5, x5
        torch.cuda.empty_cache()

        # ---- config 4: PnP-ADMM-L1, DRUNet (PyTorch bf16), 512x512 phantoms, 16 images per GPU, quadrant tiling + x8 schedule ----
        B4, N4, IT4 = 16, 512, 3
        im4 = torch.as_tensor(np.stack([pdata.phantom(N4, 200 * rank + i) for i in range(4)]).astype(np.float32)).to(dev).repeat(B4 // 4, 1, 1)
        m4 = torch.as_tensor(pdata.make_mask('random', N4, seed=3)).to(dev)
        n4 = torch.as_tensor(pdata.make_noise(N4, seed=4)).to(dev, torch.complex64)
        D4 = pden.build_denoiser('drunet_gray', iter_num=50, x8=True, seed=0, device=dev)
        barrier()
        ppnp.pnp_admm_l1(im4, m4, n4, D4, iter_num=2, reo=0.26, device=dev)                     # warm-up (cuDNN plans)
        ms4 = max_over_ranks(ev_time(lambda: ppnp.pnp_admm_l1(im4, m4, n4, D4, iter_num=IT4, reo=0.26, device=dev), 1, 0))
        legs['config4'] = dict(B=B4, N=N4, iters=IT4, ms=ms4)
        del D4, im4, m4, n4
        torch.cuda.empty_cache()

        # ---- the one collective north_star allows: the final gather of the reconstructions (N > 1) ----
        if world > 1:
            from pnp_admm_cnc_mri_b200 import sharding
            all_imgs = d_imgs.repeat(world, 1, 1)           # every rank holds the job's images; it solves its own shard

            def job():
                return sharding.reconstruct_sharded(
                    all_imgs, lambda shard, lo, hi: pk.admm_solve(shard, d_masks[2], d_noise, prox='cnc', **CNC), gather=True)
            job()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = job(); e1.record(); torch.cuda.synchronize()
            t_all = max_over_ranks(e0.elapsed_time(e1))
            job_ng = lambda: sharding.reconstruct_sharded(
                all_imgs, lambda shard, lo, hi: pk.admm_solve(shard, d_masks[2], d_noise, prox='cnc', **CNC), gather=False)
            job_ng()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); job_ng(); e1.record(); torch.cuda.synchronize()
            t_ng = max_over_ranks(e0.elapsed_time(e1))
            legs['gather'] = dict(ms_with_gather=t_all, ms_without=t_ng, gathered_shape=list(out.shape),
                                  bytes_per_rank=int(B * N * N * 4))
            del all_imgs, out

    sm, ncl = ctypes.c_int(), ctypes.c_int()
    _abi.check(lib.pnpadmm_device_info(sm, ncl, None, None))
    pc, ps, ch, la, ls, lr = (ctypes.c_int() for _ in range(6))
    _abi.check(lib.pnpadmm_plan_info(B, N, 0, CNC['iter_num'], _abi.KERNEL_AUTO, pc, ps, ch, la, ls, lr))

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
        launches_dev = lr.value                       # radial / random steps (hybrid K1 + K2, fused prologue)
        launches_k3 = 3                               # cartesian steps: prepare_rowsepN, cols2<INV> over the noise-term planes, rowsepN
        n_cart = len([i for i in range(args.steps) if MASK_KINDS[i % 3] == 'cartesian'])
        line = {
            'metric': 'admm_cnc_iterations_per_s', 'value': value, 'unit': 'iterations/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'images_per_s': value / CNC['iter_num'],
            'config': {'workload': WORKLOAD, 'batch_per_gpu': B, 'iter_num': CNC['iter_num'],
                       'kernel': f'radial / random steps: hybrid, cluster256 (K1, 16-CTA clusters, 2 CTAs/SM) on {pc.value} of the {pc.value + ps.value} packed '
                                 f'planes + K2 rows2/cols2 for the other {ps.value} on the SMs K1 cannot use; acquisition / zero-fill / data term fused into '
                                 'the K1 prologue.  cartesian steps (full k-space lines): rowsep256 (K3), every image row solved on its own, '
                                 'picked by AdmmSolver.reconstruct / the host entry points when the mask is row-separable',
                       'l2': 'flushed (256 MiB fill) between timed steps', 'sharding': f'batch x{world}, no collective'},
            'e2e': {'value': e2e, 'unit': 'iterations/s', 'ms_per_step': ms_e2e / args.steps,
                    'h2d_bytes_per_step': int(h_img.numel() + h_masks[0].numel() + h_noise.numel() * 4),
                    'd2h_bytes_per_step': int(h_x.numel() * 4),
                    'host_enqueue_us_per_step': enq_us,
                    'api': f'pnp_admm_cnc_mri_b200.HostPipeline.submit / wait (pnpadmm_reconstruct_host_pipelined_f32): pinned host buffers in '
                           f'and out every step; H2D / kernels / D2H on three streams, {SLOTS} device slots, the compute section of a step '
                           'replayed as one CUDA graph; every result is waited for and read on the host inside the timed region',
                    'uint8_output': None if ms_e2e_u8 is None else {
                        'value': its_step * args.steps / (ms_e2e_u8 * 1e-3), 'ms_per_step': ms_e2e_u8 / args.steps,
                        'd2h_bytes_per_step': int(h_x.numel()),
                        'what': "HostPipeline(output='uint8'): the D2H read is img_E = uint8(round(255 x)) as the reference saves it (S1:133-138) instead "
                                'of the float32 x; shown because the float32 read makes the 8-GPU e2e run D2H-bound on this box (aggregate pinned D2H '
                                'saturates at 71-94 GB/s for 2-8 GPUs, tools/pcie_probe.py, profiles/r2_pcie_probe.txt)'},
                    'sync_call': {'value': its_step * args.steps / (ms_e2e_sync * 1e-3), 'ms_per_step': ms_e2e_sync / args.steps,
                                  'api': 'pnpadmm_reconstruct_host_f32: one stream, host synchronises after every step '
                                         '(copies not overlapped; single-call latency)'}},
            'gpu_launches': launches_dev * (args.steps - n_cart) + launches_k3 * n_cart,
            'launches_per_step': {'device': launches_dev, 'device_cartesian_step': launches_k3, 'e2e': launches_dev + (1 if ps.value else 0),
                                  'source': 'pnpadmm_plan_info (the library reports the launches of pnpadmm_reconstruct_f32 for this plan)',
                                  'kernels': f'prepare_shared x1, cluster256 x1 ({pc.value} planes, {ch.value} chunk(s); acquisition, zero-fill and data term fused '
                                             f'into its prologue); K2 share of {ps.value} planes on the side stream: rows2<FWD_IMG>, cols2<FWD_ACQ>, cols2<INV>, '
                                             f'rows2<INV_ABS>, copy_zero, prepare, rows2<FWD_ZW> + (cols2<BLEND> + rows2<PROX>) x50 (+ u8_to_unit for its images in e2e); '
                                             f'unfused acquire + solve would be {la.value} + {ls.value}'},
            'roofline': {'bound': 'fp32', 'kernel': 'cluster256_kernel<16>', 'achieved': achieved, 'peak': nominal_peak,
                         'unit': 'TFLOP/s', 'frac': achieved / nominal_peak, 'traffic': None,
                         'peak_source': f'nominal non-tensor FP32: {sm.value} SMs x 128 lanes x 2 x {sm_max:.0f} MHz '
                                        f'(MEASURED_PEAKS.json has no fp32 entry); measured FFMA probe in this run: '
                                        f'{fl.value / 1e12:.1f} TFLOP/s',
                         'measured_fma_peak': fl.value / 1e12, 'frac_of_measured_fma': achieved / (fl.value / 1e12),
                         'flop_model': '10 N^2 log2(N^2) = 10485760 per image-iteration (nominal radix-2 count of 2 '
                                       'complex 2-D FFTs; executed flops are lower: pair-packing + radix-16)',
                         'launch_ms': k1_ms, 'resident_clusters': ncl.value,
                         'workload': {
                             'achieved': value / world * FLOP_PER_IMAGE_ITER / 1e12, 'frac': value / world * FLOP_PER_IMAGE_ITER / 1e12 / nominal_peak,
                             'what': 'SURVEY 8d convention applied to the whole timed step (value x F / FP32 peak, per GPU): nominal flops of the '
                                     "reference's algorithm per delivered image-iteration over the three-mask workload, prologue included. It exceeds the "
                                     'K1 kernel figure because cartesian steps run on K3, which does not execute the column transforms at all, and the '
                                     'hybrid schedule uses all 148 SMs; `frac` above stays the K1 kernel alone'},
                         'traffic_note': 'K1 is not HBM-bound: 22 MB of DRAM reads per launch (profiles/r1_k1_cluster256_ncu_summary.txt)',
                         'hybrid_iterate': {
                             'call_ms': hyb_ms,
                             'achieved': B * CNC['iter_num'] * FLOP_PER_IMAGE_ITER / (hyb_ms * 1e-3) / 1e12,
                             'frac': B * CNC['iter_num'] * FLOP_PER_IMAGE_ITER / (hyb_ms * 1e-3) / 1e12 / nominal_peak,
                             'what': 'the iterate call of the timed step (kernel=auto): cluster256 on the planes that fill '
                                     'whole cluster rounds, K2 rows2/cols2 for the rest on a side stream, on the SMs outside '
                                     'the clusters; same FLOP model, all 148 SMs'}},
            'clocks': clocks,
        }
        if per_mask:
            line['step_ms_by_mask'] = dict(per_mask, what='the device step with one mask kind throughout (ms per 64-image x 50-iteration step): cartesian = K3 '
                                                          'rowsep256 (no column transforms, no transposes; executes about half the nominal FFT work), radial / '
                                                          'random = K1 + K2 hybrid with the fused prologue')
        if sustained:
            line['sustained'] = {
                'device': {'value': its_step * sustained['device_steps'] / (sustained['device_ms'] * 1e-3),
                           'ms_per_step': sustained['device_ms'] / sustained['device_steps'], 'steps': sustained['device_steps'],
                           'seconds': sustained['device_ms'] * 1e-3},
                'e2e': {'value': its_step * sustained['e2e_steps'] / (sustained['e2e_ms'] * 1e-3),
                        'ms_per_step': sustained['e2e_ms'] / sustained['e2e_steps'], 'steps': sustained['e2e_steps'],
                        'seconds': sustained['e2e_ms'] * 1e-3},
                'unit': 'iterations/s', 'clocks': sustained['clocks'],
                'what': 'the same device step / host-pipeline step looped back to back for >= 2 s each (whole job, max over ranks)'}
        hbm_peak = peaks.get('hbm_gbs', 6450.0)
        if 'config5' in legs:
            c5 = legs['config5']
            rows = {}
            for N5, r in c5.items():
                its5 = world * r['B'] * CNC['iter_num'] / (r['ms'] * 1e-3)
                rows[str(N5)] = {'images_per_gpu': r['B'], 'mask': r['mask'], 'ms': r['ms'], 'iterations_per_s': its5, 'images_per_s': its5 / CNC['iter_num'],
                                 'kernel': 'hybrid K1 + K2' if N5 == 256 else 'K2 streaming',
                                 'cartesian_mask': {'ms': r['ms_cartesian'], 'iterations_per_s': world * r['B'] * CNC['iter_num'] / (r['ms_cartesian'] * 1e-3),
                                                    'kernel': 'K3 row-separable (rows resident on chip for the whole solve)'},
                                 'moved_gbs_per_gpu': (its5 / world) * 36.5 * N5 * N5 / 1e9 if N5 > 256 else None,
                                 'nominal_fft_tflops_per_gpu': (its5 / world) * 10 * N5 * N5 * math.log2(N5 * N5) / 1e12}
            line['config5'] = {'what': 'ADMM-CNC, 512 synthetic phantoms per GPU, 50 iterations, acquisition + zero-fill + prepare + solve '
                                       '(device-resident inputs, CUDA events, max over ranks; whole-job rates)', 'sizes': rows}
            r = c5[1024]
            its2 = r['B'] * CNC['iter_num'] / (r['ms'] * 1e-3)                 # per GPU
            moved = its2 * 36.5 * 1024 ** 2 / 1e9
            rec, why = recorded_traffic('k2_n1024_b64_rows2_plus_cols2')
            line['roofline_streaming'] = {
                'bound': 'hbm', 'kernel': 'rows2_kernel<1024> + cols2_tma_kernel<1024> (K2, one pair per iteration; column tiles loaded by 2-D TMA)',
                'achieved': moved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': moved / hbm_peak,
                'traffic': rec['dram_bytes'] * r['B'] / 64 if rec else None, 'traffic_source': why,
                'bytes_model': '73 N^2 per packed plane-iteration = 36.5 N^2 per image-iteration (rows pass 48 B/px: K 8+8, '
                               'z,w of two images 16+16; cols pass 25 B/px: K 8+8, G 8, codes 1/4): DESIGN.md 5',
                'achieved_survey_q57': its2 * 57 * 1024 ** 2 / 1e9,
                'survey_model': 'SURVEY 8d counts 57 N^2 per image-iteration for an unpaired implementation; pairing two '
                                'real images per complex plane moves 36.5 N^2, so the survey-model figure can exceed the peak',
                'workload': f"BASELINE config 5 at N=1024: ADMM-CNC, B={r['B']} per GPU, 50 iterations, whole solve incl. acquisition",
                'call_ms': r['ms'], 'iterations_per_s_per_gpu': its2,
                'peak_source': 'MEASURED_PEAKS.json hbm_gbs (copy bandwidth)' if 'hbm_gbs' in peaks else 'fallback 6450 GB/s',
            }
        if 'config3' in legs:
            r3 = legs['config3']
            line['config3'] = {'what': 'PnP-ADMM-CNC (S6:372 preset 1.2 / 4 / 0.45 / 0.3), DnCNN-17 random-init bf16 on the tcgen05 kernels K5, two '
                                       'forwards per iteration, 256 images per GPU, 256x256',
                               'ms_per_iteration': r3['ms'] / r3['iters'], 'iterations_timed': r3['iters'],
                               'image_iterations_per_s': world * r3['B'] * r3['iters'] / (r3['ms'] * 1e-3)}
            layer_flop = 2.0 * 64 * 64 * 9 * N * N * k5['B']
            fwd_flop = 2.0 * 555137 * N * N * k5['B']
            tpeak = peaks.get('bf16_tflops', 1670.0)
            rec5, why5 = recorded_traffic('k5_conv64_b256')
            line['roofline_tensor'] = {
                'bound': 'tensor', 'kernel': 'conv64_tc_kernel<64> (K5: one DnCNN conv3x3 64->64 + bias + ReLU layer, tcgen05 implicit GEMM)',
                'achieved': layer_flop / (k5_alone_ms * 1e-3) / 1e12, 'peak': tpeak, 'unit': 'TFLOP/s',
                'frac': layer_flop / (k5_alone_ms * 1e-3) / 1e12 / tpeak,
                'launch_ms_alone': k5_alone_ms,
                'timing': 'achieved / frac: the layer timed alone (5 launches, 7 ms in all: too short for the 100 ms clock sampler) before the long legs, against the burst peak; '
                          'launch_ms / frac_of_sustained_peak: the same launch timed inside the config-3 section (GPU at its power cap), '
                          'against the sustained peak',
                'traffic': rec5['dram_bytes'] if rec5 else None, 'traffic_source': why5,
                'flop_model': '2 x 64 x 64 x 9 = 73728 FLOP per pixel and layer', 'launch_ms': k5['layer_ms'],
                'hbm_gbs_moved': 2 * 128.0 * N * N * k5['B'] / (k5['layer_ms'] * 1e-3) / 1e9,
                'peak_source': 'MEASURED_PEAKS.json bf16_tflops (cuBLAS burst)' if 'bf16_tflops' in peaks else 'fallback 1670 TFLOP/s',
                'frac_of_sustained_peak': layer_flop / (k5['layer_ms'] * 1e-3) / 1e12 / peaks.get('bf16_tflops_sustained', 1392.5),
                'workload': f"B={k5['B']}, 256x256, bf16 operands, fp32 accumulation",
                'clocks': k5['clocks'],
                'dncnn17_forward': {'ms': k5['forward_ms'], 'achieved': fwd_flop / (k5['forward_ms'] * 1e-3) / 1e12,
                                    'what': 'head (CUDA cores) + 15 x conv64_tc_kernel<64> + tail conv64_tc_kernel<16>; BASELINE config 3 '
                                            'runs two of these per PnP-ADMM-CNC iteration'},
                'other_networks': {nm: {'forward_ms': v['ms'], 'pytorch_bf16_forward_ms': v['torch_ms'], 'speedup': v['torch_ms'] / v['ms']}
                                   for nm, v in k5['nets'].items()},
                'other_networks_what': f"Denoiser.__call__ at B={k5['B']}, 256x256 (IRCNN: 5 dilated 64->64 layers, pnpadmm_dncnn_forward_dilated_bf16; "
                                       'FFDNet: 14 layers at 128x128 + pack + shuffled tail, pnpadmm_ffdnet_forward_bf16) against the stock module'}
        if 'config4' in legs:
            r4 = legs['config4']
            line['config4'] = {'what': 'PnP-ADMM-L1 (S3:347 preset, reo 0.26), DRUNet random-init bf16 in PyTorch (4 x 288^2 quadrants per image, x8 '
                                       'schedule, sigma schedule 49 -> 15), 512x512 phantoms, 16 images per GPU, batch-sharded, no collective',
                               'ms_per_iteration': r4['ms'] / r4['iters'], 'iterations_timed': r4['iters'],
                               'image_iterations_per_s': world * r4['B'] * r4['iters'] / (r4['ms'] * 1e-3)}
        if 'gather' in legs:
            g = legs['gather']
            line['final_gather'] = {'what': 'sharding.reconstruct_sharded: every rank solves its 64-image shard (admm_solve incl. allocation), then ONE '
                                            'NCCL all_gather of the float32 reconstructions (the only collective on the path)',
                                    'ms_with_gather': g['ms_with_gather'], 'ms_without_gather': g['ms_without'],
                                    'gather_ms': g['ms_with_gather'] - g['ms_without'], 'bytes_per_rank': g['bytes_per_rank'],
                                    'gathered_shape': g['gathered_shape']}
        if cpu_v is not None:
            line['cpu_baseline'] = {'value': cpu_v, 'unit': 'iterations/s', 'cores': 1, 'kind': 'port',
                                    'sample': '8 images x 50 iterations, 1 thread, NumPy fp64 oracle restatement '
                                              '(bit-identical to the unmodified reference script); host has '
                                              f'{cores} cores, all-cores number: bench.py --impl reference'}
        print(json.dumps(line), flush=True)
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--legs', default='all', choices=['all', 'headline'],
                    help="'headline': only the config-2 step, e2e and the K1 roofline (quick runs under ncu)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
