"""GPU, world_size 2 over NCCL: the path's one collective (north_star: "no collective beyond the final gather") on hardware.
Each rank reconstructs its contiguous shard with the CUDA solver, then ONE all_gather of the float32 shards; the result on
every rank must equal the single-GPU reconstruction of the whole batch and meet the parity gate against the oracle.
Needs two GPUs (skipped otherwise; run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

P = dict(alpha=0.45, iter_num=50, lambda1=0.5, reo=0.05, b=64)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(B):
    from pnp_admm_cnc_mri_b200 import data
    N = 256
    return data.phantoms(B, N, seed0=11), data.make_mask('radial', N, seed=2), data.make_noise(N, seed=5)


def _worker(rank, world, port, B, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import pnp_admm_cnc_mri_b200 as pk
    from pnp_admm_cnc_mri_b200.sharding import reconstruct_sharded
    imgs, mask, nz = _inputs(B)
    dev = torch.device('cuda', rank)
    d_imgs = torch.as_tensor(imgs).to(dev)

    def solve(shard, lo, hi):
        return pk.admm_solve(shard, mask, nz, prox='cnc', device=dev, **P)

    full = reconstruct_sharded(d_imgs, solve)
    assert full.is_cuda and full.dtype == torch.float32 and tuple(full.shape) == (B, 256, 256)
    q.put((rank, full.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('B', [1, 7, 40])
def test_nccl_world2_final_gather(B):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    import pnp_admm_cnc_mri_b200 as pk
    from oracle import reference_numpy as orc
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    imgs, mask, nz = _inputs(B)
    assert np.array_equal(got[0], got[1])                                    # every rank holds the same gathered batch
    lo0 = (B + 1) // 2                                                       # rank 0's shard is the first ceil(B / 2) images
    single = pk.admm_solve(imgs[:lo0], mask, nz, prox='cnc', **P)           # same shard on this process' GPU: same kernels, same bits
    assert np.array_equal(got[0][:lo0], single)
    for k in sorted({0, B // 2, B - 1}):
        xr = orc.admm_cnc(imgs[k], mask.astype(np.float64), nz, **P)
        assert np.linalg.norm(got[0][k] - xr) / np.linalg.norm(xr) < 1e-4, k
