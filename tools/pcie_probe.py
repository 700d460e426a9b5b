"""Aggregate host <-> device copy bandwidth of the box with N ranks copying at once (torchrun), pinned host memory,
the byte counts of one bench step per rank (D2H 16.8 MB float32 reconstructions, H2D 4.8 MB uint8 images + mask + noise).
Names the limiter of the e2e scaling run: if the aggregate rate saturates, e2e at N GPUs is bound by the host links."""
import os, sys, time
import torch
import torch.distributed as dist

world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
dev = torch.device('cuda', local)
D2H, H2D = 64 * 256 * 256 * 4, 64 * 256 * 256 + 256 * 256 + 256 * 256 * 8
d_out = torch.empty(D2H, dtype=torch.uint8, device=dev); h_out = torch.empty(D2H, dtype=torch.uint8).pin_memory()
d_in = torch.empty(H2D, dtype=torch.uint8, device=dev); h_in = torch.empty(H2D, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(mode, reps=200):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if mode in ('d2h', 'both'):
            with torch.cuda.stream(s1):
                h_out.copy_(d_out, non_blocking=True)
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
    nbytes = reps * ((D2H if mode in ('d2h', 'both') else 0) + (H2D if mode in ('h2d', 'both') else 0))
    return nbytes / float(t.item()) / 1e9


for mode in ('d2h', 'h2d', 'both'):
    run(mode, 20)
    g = run(mode)
    if rank == 0:
        print(f'ranks {world}: {mode:5s} {g:7.1f} GB/s per rank, {g * world:7.1f} GB/s aggregate; one bench step per rank '
              f'({(D2H + H2D) / 1e6:.1f} MB both ways) would take {((D2H if mode != "h2d" else 0) + (H2D if mode != "d2h" else 0)) / g / 1e6:.3f} ms', flush=True)
if world > 1:
    dist.destroy_process_group()
