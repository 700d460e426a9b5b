"""CPU, world_size 2 over gloo: the batch-sharding path (contiguous shards, no data-path collective,
one final all_gather) reproduces the single-process result.  The oracle stands in for the GPU solver."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pnp_admm_cnc_mri_b200.sharding import reconstruct_sharded, shard_bounds


def test_shard_bounds_cover_batch():
    for B in (1, 2, 7, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import reference_numpy as orc
    from pnp_admm_cnc_mri_b200 import data
    N = 32
    imgs = torch.from_numpy(data.phantoms(B, N, seed0=3))
    mask = data.make_mask('random', N, seed=1).astype(np.float64)
    nz = data.make_noise(N, seed=2)

    def solve(shard, lo, hi):
        return torch.from_numpy(np.stack([orc.admm_cnc(a.numpy(), mask, nz, 0.45, 5, 0.5, 0.05, 64) for a in shard]))

    # the images arrive as float32 while solve() returns float64: a rank with an empty shard (B < world) must still
    # contribute a float64 buffer (it used to take the input's dtype and the all_gather failed)
    full = reconstruct_sharded(imgs.float(), solve, out_dtype=torch.float64, out_device='cpu')
    if rank == 0:
        q.put(full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize('B', [1, 5, 8])
def test_world2_gather_equals_single_process(B):
    from oracle import reference_numpy as orc
    from pnp_admm_cnc_mri_b200 import data
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    N = 32
    imgs = data.phantoms(B, N, seed0=3)
    mask = data.make_mask('random', N, seed=1).astype(np.float64)
    nz = data.make_noise(N, seed=2)
    want = np.stack([orc.admm_cnc(a.astype(np.float32), mask, nz, 0.45, 5, 0.5, 0.05, 64) for a in imgs])
    assert got.shape == want.shape and np.array_equal(got, want)
