"""Aggregate host <-> device copy bandwidth of the box with N ranks copying at once (torchrun), pinned host memory,
the byte counts of one bench step per rank (D2H 16.8 MB float32 reconstructions, H2D 4.8 MB uint8 images + mask + noise).
Names the limiter of the e2e scaling run: if the aggregate rate saturates, e2e at N GPUs is bound by the host links."""
import os, sys, time
import ctypes
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnp_admm_cnc_mri_b200 import _abi

world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
dev = torch.device('cuda', local)
D2H, H2D = 64 * 256 * 256 * 4, 64 * 256 * 256 + 256 * 256 + 256 * 256 * 8
d_out = torch.empty(D2H, dtype=torch.uint8, device=dev); h_out = torch.empty(D2H, dtype=torch.uint8).pin_memory()
d_in = torch.empty(H2D, dtype=torch.uint8, device=dev); h_in = torch.empty(H2D, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
lib = _abi.load()
lib.pnpadmm_debug_copy.restype = ctypes.c_int
lib.pnpadmm_debug_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
KBLOCKS = int(os.environ.get('PROBE_BLOCKS', '32'))      # CTAs of the zero-copy kernel (it should not need many SMs)


def run(mode, reps=200):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if mode in ('d2h', 'both'):
            with torch.cuda.stream(s1):
                h_out.copy_(d_out, non_blocking=True)
        if mode == 'd2h_kernel':      # a kernel stores straight into the pinned host buffer (UVA: same pointer on the device)
            _abi.check(lib.pnpadmm_debug_copy(h_out.data_ptr(), d_out.data_ptr(), D2H, KBLOCKS, s1.cuda_stream))
        if mode == 'h2d_kernel':
            _abi.check(lib.pnpadmm_debug_copy(d_in.data_ptr(), h_in.data_ptr(), (H2D // 16) * 16, KBLOCKS, s2.cuda_stream))
        if mode in ('h2d', 'both'):
            with torch.cuda.stream(s2):
                d_in.copy_(h_in, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nbytes = reps * ((D2H if mode in ('d2h', 'both', 'd2h_kernel') else 0) + (H2D if mode in ('h2d', 'both', 'h2d_kernel') else 0))
    return nbytes / float(t.item()) / 1e9


for mode in ('d2h', 'd2h_kernel', 'h2d', 'h2d_kernel', 'both'):
    run(mode, 20)
    g = run(mode)
    if rank == 0:
        print(f'ranks {world}: {mode:5s} {g:7.1f} GB/s per rank, {g * world:7.1f} GB/s aggregate; one bench step per rank '
              f'({(D2H + H2D) / 1e6:.1f} MB both ways) would take {((D2H if 'h2d' not in mode else 0) + (H2D if 'd2h' not in mode else 0)) / g / 1e6:.3f} ms', flush=True)
if world > 1:
    dist.destroy_process_group()
