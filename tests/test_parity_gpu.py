"""GPU parity: the CUDA path (through the C ABI) against the fp64 oracle on the same inputs.
Gates (BASELINE.json north_star): fp32 rel-L2 <= 1e-4 and |dPSNR| <= 0.01 dB; fp64 rel-L2 <= 1e-10."""
import os
import numpy as np
import pytest

from oracle import kat_table as kat
from oracle import reference_numpy as orc

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')

TOL32, TOL64, TOL_PSNR = 1e-4, 1e-10, 0.01


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    a = a.astype(np.complex128 if np.iscomplexobj(a) else np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _oracle_one(job):
    prox, im, m, n, P = job
    fn = orc.admm_l1 if prox == 'l1' else orc.admm_cnc
    return fn(im, m.astype(np.float64), n, return_state=True, **P)


def oracle_batch(imgs, mask, noises, prox, P, workers=1):
    """The fp64 oracle per image; `workers` > 1 spreads the images over host processes (large N, full depth)."""
    jobs = [(prox, im, mask[i] if mask.ndim == 3 else mask, noises[i] if noises.ndim == 3 else noises, P)
            for i, im in enumerate(imgs)]
    if workers > 1 and len(jobs) > 1:
        import multiprocessing as mp
        with mp.get_context('fork').Pool(min(workers, len(jobs))) as pool:
            out = pool.map(_oracle_one, jobs, chunksize=1)
    else:
        out = [_oracle_one(j) for j in jobs]
    return [np.stack([o[k] for o in out]) for k in range(4)]


@pytest.fixture(scope='module')
def pk():
    import pnp_admm_cnc_mri_b200 as pk
    pk.load_library()
    return pk


def _imgs(cs, idx):
    return np.stack([orc.preprocess_uint8(cs['images'][i]) for i in idx])


# ---------------------------------------------------------------------------------------------
# a2: acquisition and zero-filled start
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dtype,tol', [('float64', 1e-11), ('float32', 2e-6)])
def test_acquire_and_zero_filled(pk, cs_inputs, dtype, tol):
    imgs = _imgs(cs_inputs, [4, 0, 9])
    m = cs_inputs['masks'][0]
    s = pk.AdmmSolver(3, 256, dtype=dtype)
    y = s.acquire(imgs, m, cs_inputs['noises'])
    x0 = s.zero_filled(y)
    for k in range(3):
        yr = orc.acquire(imgs[k], m.astype(np.float64), cs_inputs['noises'])
        assert rel(y[k].cpu().numpy(), yr) < tol
        assert rel(x0[k].cpu().numpy(), orc.zero_filled(yr)) < tol * 4


# ---------------------------------------------------------------------------------------------
# configs 1 and 2 on the reference's own data
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('kernel', ['cluster', 'streaming'])
@pytest.mark.parametrize('prox,P', [('l1', kat.L1_DEFAULTS), ('cnc', kat.CNC_DEFAULTS)])
def test_config1_fp32_reference_outputs(pk, cs_inputs, ref_out, prox, P, kernel):
    """05.png / Q_Random30 / defaults against the UNMODIFIED reference's stored out[0]."""
    img = _imgs(cs_inputs, [4])
    x = pk.admm_solve(img, cs_inputs['masks'][0], cs_inputs['noises'], prox=prox, kernel=kernel, **P)
    assert rel(x[0], ref_out[prox]) < TOL32
    H = cs_inputs['images'][4]
    p = orc.calculate_psnr(x[0].astype(np.float64) * 255, H)
    assert abs(p - kat.SET1[prox][0]) < TOL_PSNR


@pytest.mark.parametrize('prox,P', [('l1', kat.L1_DEFAULTS), ('cnc', kat.CNC_DEFAULTS)])
def test_config1_fp64(pk, cs_inputs, ref_out, prox, P):
    img = _imgs(cs_inputs, [4])
    x = pk.admm_solve(img, cs_inputs['masks'][0], cs_inputs['noises'], prox=prox, dtype='float64', **P)
    assert rel(x[0], ref_out[prox]) < TOL64


@pytest.mark.parametrize('kernel', ['cluster', 'streaming'])
@pytest.mark.parametrize('mask_i', [0, 1, 2])
@pytest.mark.parametrize('prox,P', [('l1', kat.L1_DEFAULTS), ('cnc', kat.CNC_DEFAULTS)])
def test_kat_table_fp32(pk, cs_inputs, mask_i, prox, P, kernel):
    """All 15 set images x 3 masks: rel-L2 vs oracle and PSNR vs the reference's log table."""
    imgs = _imgs(cs_inputs, range(15))
    m = cs_inputs['masks'][mask_i]
    x = pk.admm_solve(imgs, m, cs_inputs['noises'], prox=prox, kernel=kernel, **P)
    xr = oracle_batch(imgs, m, cs_inputs['noises'], prox, P)[0]
    name = cs_inputs['mask_names'][mask_i]
    for k in range(15):
        assert rel(x[k], xr[k]) < TOL32, (name, prox, k)
        p = orc.calculate_psnr(x[k].astype(np.float64) * 255, cs_inputs['images'][k])
        assert abs(p - kat.PSNR[(name, prox)][k]) < TOL_PSNR, (name, prox, k, p)


def test_config2_batch64_cnc(pk, cs_inputs):
    """BASELINE config 2: ADMM-CNC, batch 64 (15 set images tiled cyclically), three masks."""
    idx = [i % 15 for i in range(64)]
    imgs = _imgs(cs_inputs, idx)
    for mi in range(3):
        m = cs_inputs['masks'][mi]
        x, z, w, y = pk.admm_solve(imgs, m, cs_inputs['noises'], prox='cnc', return_state=True, **kat.CNC_DEFAULTS)
        xr, zr, wr, yr = oracle_batch(imgs[:15], m, cs_inputs['noises'], 'cnc', kat.CNC_DEFAULTS)
        for k in range(64):
            assert rel(x[k], xr[k % 15]) < TOL32
            assert rel(z[k], zr[k % 15]) < TOL32
        # every copy of the same image must give the same answer whichever pair slot it lands in
        assert rel(x[0], x[30].astype(np.float64)) < TOL32      # slot a (0) vs slot a (30): 30 = 2*15
        assert rel(x[0], x[15].astype(np.float64)) < TOL32      # slot a (0) vs slot b (15)


# ---------------------------------------------------------------------------------------------
# fp64 validation build over sizes, odd batches, per-image masks / noise
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('N', [16, 32, 64, 128, 256, 512, 1024])
@pytest.mark.parametrize('prox', ['l1', 'cnc'])
def test_fp64_sizes(pk, N, prox):
    from pnp_admm_cnc_mri_b200 import data
    B = 3
    imgs = data.phantoms(B, N, seed0=10)
    m = data.make_mask('random', N, seed=1)
    nz = data.make_noise(N, seed=5)
    P = dict(kat.L1_DEFAULTS) if prox == 'l1' else dict(kat.CNC_DEFAULTS)
    P['iter_num'] = 12
    x, z, w, y = pk.admm_solve(imgs, m, nz, prox=prox, dtype='float64', return_state=True, **P)
    xr, zr, wr, yr = oracle_batch(imgs, m, nz, prox, P, workers=3)
    assert rel(y, yr) < 1e-11
    assert rel(x, xr) < TOL64 and rel(z, zr) < TOL64
    assert np.abs(w - wr).max() < 1e-9


@pytest.mark.parametrize('N,kernel', [(64, 'streaming'), (128, 'streaming'), (256, 'cluster'), (256, 'streaming'),
                                      (512, 'streaming'), (1024, 'streaming')])
@pytest.mark.parametrize('prox', ['l1', 'cnc'])
def test_fp32_sizes(pk, N, kernel, prox):
    from pnp_admm_cnc_mri_b200 import data
    B = 5 if N <= 512 else 2                      # odd: last packed plane has an empty slot
    imgs = data.phantoms(B, N, seed0=20)
    m = data.make_mask(('random', 'radial', 'cartesian')[N % 3], N, seed=2)
    nz = data.make_noise(N, seed=6)
    P = dict(kat.L1_DEFAULTS) if prox == 'l1' else dict(kat.CNC_DEFAULTS)
    P['iter_num'] = 50 if N <= 256 else 10
    x = pk.admm_solve(imgs, m, nz, prox=prox, kernel=kernel, **P)
    xr = oracle_batch(imgs, m, nz, prox, P)[0]
    for k in range(B):
        assert rel(x[k], xr[k]) < TOL32, (N, prox, k)
        assert abs(orc.calculate_psnr(x[k].astype(np.float64) * 255, imgs[k] * 255.)
                   - orc.calculate_psnr(xr[k] * 255, imgs[k] * 255.)) < TOL_PSNR


@pytest.mark.parametrize('N', [512, 1024])
@pytest.mark.parametrize('prox,kind', [('cnc', 'cartesian'), ('cnc', 'radial'), ('cnc', 'random'),
                                       ('l1', 'cartesian'), ('l1', 'radial'), ('l1', 'random')])
def test_full_depth_large_sizes_fp32(pk, N, prox, kind):
    """The reference depth (50 iterations, S1:171 / S4:176; BASELINE config 5) on the K2 streaming kernels at
    N = 512 and 1024, every mask kind, odd batch (last packed plane half empty): SURVEY 7 warns that the fp32
    headroom of CNC shrinks with depth, so the gate is checked where it is tightest."""
    from pnp_admm_cnc_mri_b200 import data
    B = 3
    imgs = data.phantoms(B, N, seed0=40 + N // 512)
    m = data.make_mask(kind, N, seed=4)
    nz = data.make_noise(N, seed=8)
    P = dict(kat.L1_DEFAULTS) if prox == 'l1' else dict(kat.CNC_DEFAULTS)
    assert P['iter_num'] == 50
    x, z, w, y = pk.admm_solve(imgs, m, nz, prox=prox, kernel='streaming', return_state=True, **P)
    xr, zr, wr, yr = oracle_batch(imgs, m, nz, prox, P, workers=B)
    for k in range(B):
        e = rel(x[k], xr[k])
        dp = abs(orc.calculate_psnr(x[k].astype(np.float64) * 255, imgs[k] * 255.) - orc.calculate_psnr(xr[k] * 255, imgs[k] * 255.))
        print(f'N={N} {prox} {kind} image {k}: rel-L2 {e:.2e}, |dPSNR| {dp:.1e} dB')
        assert e < TOL32, (N, prox, kind, k, e)
        assert rel(z[k], zr[k]) < TOL32
        assert dp < TOL_PSNR


@pytest.mark.parametrize('prox', ['l1', 'cnc'])
def test_full_depth_fp64_1024(pk, prox):
    """fp64 validation build at the largest size and the reference depth: <= 1e-10 after 50 iterations."""
    from pnp_admm_cnc_mri_b200 import data
    N, B = 1024, 2
    imgs = data.phantoms(B, N, seed0=60)
    m = data.make_mask('radial', N, seed=5)
    nz = data.make_noise(N, seed=9)
    P = dict(kat.L1_DEFAULTS) if prox == 'l1' else dict(kat.CNC_DEFAULTS)
    x, z, w, y = pk.admm_solve(imgs, m, nz, prox=prox, dtype='float64', return_state=True, **P)
    xr, zr, wr, yr = oracle_batch(imgs, m, nz, prox, P, workers=B)
    assert rel(x, xr) < TOL64 and rel(z, zr) < TOL64


@pytest.mark.parametrize('dtype,kernel,tol', [('float64', 'auto', TOL64), ('float32', 'cluster', TOL32),
                                              ('float32', 'streaming', TOL32)])
def test_per_image_masks_and_noise(pk, dtype, kernel, tol):
    from pnp_admm_cnc_mri_b200 import data
    N, B = 256, 4
    imgs = data.phantoms(B, N, seed0=30)
    masks = np.stack([data.make_mask(k, N, seed=s) for s, k in enumerate(('random', 'radial', 'cartesian', 'random'))])
    nz = data.make_noise(N, seed=7, B=B)
    P = dict(kat.CNC_DEFAULTS, iter_num=20)
    x = pk.admm_solve(imgs, masks, nz, prox='cnc', dtype=dtype, kernel=kernel, **P)
    xr = oracle_batch(imgs, masks, nz, 'cnc', P)[0]
    for k in range(B):
        assert rel(x[k], xr[k]) < tol


def test_non_default_parameters_and_single_image(pk, cs_inputs):
    img = orc.preprocess_uint8(cs_inputs['images'][2])
    m = cs_inputs['masks'][1]
    for prox, P in (('l1', dict(iter_num=9, lambda1=0.3, reo=0.05)),
                    ('cnc', dict(alpha=0.3, iter_num=7, lambda1=0.2, reo=0.1, b=16)),
                    ('cnc', dict(alpha=0.4, iter_num=4, lambda1=0.04, reo=2.75, b=1))):   # S4:37-41 fallbacks
        x = pk.admm_solve(img, m, cs_inputs['noises'], prox=prox, dtype='float64', **P)   # (N,N) in -> (N,N) out
        fn = orc.admm_l1 if prox == 'l1' else orc.admm_cnc
        assert x.shape == (256, 256)
        assert rel(x, fn(img, m.astype(np.float64), cs_inputs['noises'], **P)) < TOL64
        x32 = pk.admm_solve(img, m, cs_inputs['noises'], prox=prox, **P)
        assert rel(x32, fn(img, m.astype(np.float64), cs_inputs['noises'], **P)) < TOL32


# ---------------------------------------------------------------------------------------------
# a3 x-update alone and the step-wise API (what the PnP variants call)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dtype,kernel,tol', [('float64', 'auto', 1e-12), ('float32', 'cluster', 2e-6),
                                              ('float32', 'streaming', 2e-6)])
def test_xupdate(pk, cs_inputs, dtype, kernel, tol):
    rng = np.random.default_rng(0)
    imgs = _imgs(cs_inputs, [1, 2, 3])
    m = cs_inputs['masks'][2]
    s = pk.AdmmSolver(3, 256, dtype=dtype)
    y = s.acquire(imgs, m, cs_inputs['noises'])
    s.prepare(y, m, reo=0.26)
    z = rng.uniform(0, 1, (3, 256, 256))
    w = rng.uniform(-0.1, 0.1, (3, 256, 256))
    x, xpw = s.xupdate(torch.as_tensor(z), torch.as_tensor(w), want_xpw=True, kernel=kernel)
    idx = np.nonzero(m)
    for k in range(3):
        yr = orc.acquire(imgs[k], m.astype(np.float64), cs_inputs['noises'])
        xr = orc.x_update(z[k], w[k], yr, idx, 0.26)
        assert rel(x[k].cpu().numpy(), xr) < tol
        assert rel(xpw[k].cpu().numpy(), xr + w[k]) < tol * 2


def test_stepwise_iterate_equals_solve(pk, cs_inputs):
    imgs = _imgs(cs_inputs, [5, 6])
    m = cs_inputs['masks'][0]
    P = kat.CNC_DEFAULTS
    for kernel in ('cluster', 'streaming'):
        s = pk.AdmmSolver(2, 256)
        y = s.acquire(imgs, m, cs_inputs['noises'])
        x1, z1, w1 = s.solve(y, m, 'cnc', P['iter_num'], P['lambda1'], P['reo'], P['alpha'], P['b'], kernel=kernel)
        z = s.zero_filled(y)
        w = torch.zeros_like(z)
        x = torch.empty_like(z)
        s.prepare(y, m, P['reo'])
        # 50 iterations as 20 + 30 through the warm-start entry point
        s.iterate(x, z, w, 'cnc', 20, P['lambda1'], P['reo'], P['alpha'], P['b'], kernel=kernel)
        s.iterate(x, z, w, 'cnc', 30, P['lambda1'], P['reo'], P['alpha'], P['b'], kernel=kernel)
        assert torch.equal(x, x1) and torch.equal(z, z1) and torch.equal(w, w1)


def test_chunked_cluster_schedule_is_bit_identical(pk, cs_inputs, monkeypatch):
    """The chunked static schedule (plane state handed between clusters through L2) must not change a bit."""
    imgs = _imgs(cs_inputs, range(7))
    m = cs_inputs['masks'][0]
    monkeypatch.setenv('PNPADMM_K1_CHUNKS', '1')
    a = pk.admm_solve(imgs, m, cs_inputs['noises'], prox='cnc', kernel='cluster', return_state=True, **kat.CNC_DEFAULTS)
    for n in ('3', '7', '50'):
        monkeypatch.setenv('PNPADMM_K1_CHUNKS', n)
        b = pk.admm_solve(imgs, m, cs_inputs['noises'], prox='cnc', kernel='cluster', return_state=True, **kat.CNC_DEFAULTS)
        for u, v in zip(a[:3], b[:3]):
            assert np.array_equal(u, v), n
    monkeypatch.delenv('PNPADMM_K1_CHUNKS')


def test_cluster_and_streaming_agree(pk, cs_inputs):
    imgs = _imgs(cs_inputs, range(6))
    m = cs_inputs['masks'][1]
    a = pk.admm_solve(imgs, m, cs_inputs['noises'], prox='l1', kernel='cluster', **kat.L1_DEFAULTS)
    b = pk.admm_solve(imgs, m, cs_inputs['noises'], prox='l1', kernel='streaming', **kat.L1_DEFAULTS)
    assert rel(a, b.astype(np.float64)) < 2e-5


# ---------------------------------------------------------------------------------------------
# pointwise pieces
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dt', [torch.float32, torch.float64])
def test_pointwise(pk, dt):
    g = torch.Generator(device='cuda').manual_seed(0)
    sh = (3, 64, 64)
    z, x, w, s = (torch.rand(sh, generator=g, device='cuda', dtype=dt) - 0.3 for _ in range(4))
    x[0, 0, :5] = 0.0
    npdt = np.float64 if dt == torch.float64 else np.float32
    assert np.array_equal(pk.soft(x, 0.2).cpu().numpy(), orc.soft(x.cpu().numpy(), npdt(0.2)).astype(npdt))
    t = pk.cnc_combine(z, x, w, s, 1.2, 0.54).cpu().numpy()
    zz, xx, ww, ss = (a.cpu().numpy().astype(np.float64) for a in (z, x, w, s))
    tr = (1 - 1.2) * zz + 1.2 * (xx + ww) + 0.54 * (zz - ss)
    assert np.abs(t - tr).max() < (1e-14 if dt == torch.float64 else 1e-6)
    x2, z2, w2 = x.clone(), z.clone(), w.clone()
    pk.dual_update_(x2, z2, w2, True)
    assert np.allclose(w2.cpu().numpy(), np.clip(ww + xx - zz, 0, 1), atol=1e-6)
    assert np.array_equal(x2.cpu().numpy(), np.clip(x.cpu().numpy(), 0, 1))
    x3, z3, w3 = x.clone(), z.clone(), w.clone()
    pk.dual_update_(x3, z3, w3, False)
    assert torch.equal(x3, x) and torch.equal(z3, z)
    assert np.allclose(w3.cpu().numpy(), ww + xx - zz, atol=1e-6)


# ---------------------------------------------------------------------------------------------
# error behaviour of the boundary
# ---------------------------------------------------------------------------------------------
def test_error_codes(pk):
    with pytest.raises(ValueError):
        pk.AdmmSolver(2, 100)                                     # not a power of two -> workspace fine, call fails
        pk.admm_solve(np.zeros((2, 100, 100), np.float32), np.ones((100, 100)), np.zeros((100, 100), complex))
    with pytest.raises(ValueError):
        pk.admm_solve(np.zeros((2, 64, 64), np.float32), np.ones((64, 64)), np.zeros((64, 64), complex), prox='bm3d')
    with pytest.raises(pk.PnpAdmmError):
        pk.admm_solve(np.zeros((2, 64, 64), np.float32), np.ones((64, 64)), np.zeros((64, 64), complex), kernel='cluster')
    with pytest.raises(ValueError):
        pk.admm_solve(np.zeros((2, 64, 64), np.float32), np.ones((64, 64)), np.zeros((64, 64), complex), reo=0.0)


# ---------------------------------------------------------------------------------------------
# 8f rank 1: PSNR / SSIM / RE on the device against the reference's definitions (oracle restatement)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dt', ['float32', 'float64'])
def test_device_metrics_match_reference_definitions(pk, cs_inputs, dt):
    idx = [0, 4, 9]
    H = cs_inputs['images'][idx]
    imgs = _imgs(cs_inputs, idx)
    x = np.stack([orc.admm_cnc(im, cs_inputs['masks'][0].astype(np.float64), cs_inputs['noises'], **kat.CNC_DEFAULTS)
                  for im in imgs]).astype(dt)                      # unclipped, may exceed 1
    for quantize in (False, True):
        got = pk.image_metrics(torch.as_tensor(x).cuda(), torch.as_tensor(H).cuda(), quantize=quantize).cpu().numpy()
        for k in range(len(idx)):
            E = orc.single2uint(x[k].astype(np.float32)) if quantize else x[k].astype(np.float64) * 255
            want = (orc.calculate_psnr(E, H[k]), orc.calculate_ssim(E, H[k]), orc.calculate_re(E, H[k]))
            assert abs(got[k, 0] - want[0]) < 1e-9 and abs(got[k, 1] - want[1]) < 1e-11 and abs(got[k, 2] - want[2]) < 1e-12, \
                (quantize, k, got[k], want)
    # the KAT table row of image 01, Q_Random30, CNC (results/Set_dn_ADMM_CNC.log: 24.7868 / 0.4528 / 0.1486)
    x64 = orc.admm_cnc(imgs[0], cs_inputs['masks'][0].astype(np.float64), cs_inputs['noises'], **kat.CNC_DEFAULTS)
    m = pk.image_metrics(torch.as_tensor(x64[None]).cuda(), torch.as_tensor(H[:1]).cuda()).cpu().numpy()[0]
    assert (round(m[0], 4), round(m[1], 4), round(m[2], 4)) == (24.7868, 0.4528, 0.1486)


def test_device_metrics_sizes_and_errors(pk):
    rng = np.random.default_rng(5)
    for N in (16, 64, 512):
        x = rng.random((2, N, N)).astype(np.float32)
        H = rng.integers(0, 256, (2, N, N), dtype=np.uint8)
        got = pk.image_metrics(torch.as_tensor(x).cuda(), torch.as_tensor(H).cuda()).cpu().numpy()
        for k in range(2):
            E = x[k].astype(np.float64) * 255
            want = (orc.calculate_psnr(E, H[k]), orc.calculate_ssim(E, H[k]), orc.calculate_re(E, H[k]))
            np.testing.assert_allclose(got[k], want, rtol=1e-10, atol=1e-12)
    with pytest.raises(ValueError):
        pk.image_metrics(torch.zeros((2, 8, 8), device='cuda'), torch.zeros((2, 8, 8), dtype=torch.uint8, device='cuda'))


def test_pipelined_host_call_matches_sync_call(pk, cs_inputs):
    """pnpadmm_reconstruct_host_pipelined_f32 (copies on their own streams, two slots) returns the same
    reconstructions as the single-stream host call, step after step."""
    from pnp_admm_cnc_mri_b200 import _abi
    lib = _abi.load()
    B, N = 6, 256
    P = kat.CNC_DEFAULTS
    imgs = [torch.as_tensor(cs_inputs['images'][k:k + B].copy()).pin_memory() for k in (0, 3, 6, 9)]
    masks = [torch.as_tensor(cs_inputs['masks'][k].copy()).pin_memory() for k in range(3)]
    noise = torch.view_as_real(torch.as_tensor(cs_inputs['noises']).to(torch.complex64)).contiguous().pin_memory()
    solver = pk.AdmmSolver(B, N)
    st = torch.cuda.current_stream().cuda_stream
    scratch = torch.empty(lib.pnpadmm_host_scratch_bytes(B, N), dtype=torch.uint8, device='cuda')
    want = {}
    for j in range(4):
        for k in range(3):
            hx = torch.empty((B, N, N), dtype=torch.float32).pin_memory()
            _abi.check(lib.pnpadmm_reconstruct_host_f32(imgs[j].data_ptr(), masks[k].data_ptr(), noise.data_ptr(), hx.data_ptr(),
                                                        B, N, _abi.PROX_CNC, P['iter_num'], P['lambda1'], P['reo'], P['alpha'], P['b'],
                                                        _abi.KERNEL_AUTO, scratch.data_ptr(), scratch.numel(), solver.ws.data_ptr(),
                                                        solver.ws_bytes, st))
            torch.cuda.synchronize()
            want[(j, k)] = hx.clone()
    # the pipelined call through the package's HostPipeline (owns the C-ABI pipeline object, streams, scratch):
    # 2 and 3 slots, more steps than slots so that slots are reused and the captured compute graphs are replayed
    for S in (2, 3):
        pipe = pk.HostPipeline(B, N, n_slots=S)
        hx2 = [torch.empty((B, N, N), dtype=torch.float32).pin_memory() for _ in range(S)]
        got = {}
        steps = 9
        for i in range(steps):
            if i >= S:
                pipe.wait(i % S)
                got[i - S] = hx2[i % S].clone()
            pipe.submit(i % S, imgs[i % 4], masks[i % 3], noise, hx2[i % S], prox='cnc', **P)
        for i in range(steps - S, steps):
            pipe.wait(i % S)
            got[i] = hx2[i % S].clone()
        for i in range(steps):
            assert torch.equal(got[i], want[(i % 4, i % 3)]), (S, i)
        pipe.close()
    # uint8 output (img_E as the reference saves it, S1:133-138): the float32 result rounded half-to-even and saturated
    pipe = pk.HostPipeline(B, N, n_slots=2, output='uint8')
    h8 = [torch.empty((B, N, N), dtype=torch.uint8).pin_memory() for _ in range(2)]
    for i in range(4):
        if i >= 2:
            pipe.wait(i & 1)
        pipe.submit(i & 1, imgs[i % 4], masks[i % 3], noise, h8[i & 1], prox='cnc', **P)
    for i in (2, 3):
        pipe.wait(i & 1)
        w32 = want[(i % 4, i % 3)].numpy()
        assert np.array_equal(h8[i & 1].numpy(), np.uint8(np.clip(np.rint(np.float32(255.0) * w32), 0, 255))), i
    pipe.close()
    # and against the oracle for one image of the last step
    xr = orc.admm_cnc(orc.preprocess_uint8(cs_inputs['images'][9]), cs_inputs['masks'][0].astype(np.float64), cs_inputs['noises'], **P)
    assert rel(want[(3, 0)][0].numpy(), xr) < TOL32


def test_two_pipelines_share_a_device(pk, cs_inputs):
    """the ordering events used to be one global set per device, so two pipelines on one device
    (two host threads, or two scratch / stream sets) re-recorded each other's events and a slot's inputs could be
    overwritten before its compute had read them.  The events now live in a caller-owned pipeline object: two
    HostPipelines driven from two host threads, interleaved, must both reproduce the single-stream results."""
    import threading
    from pnp_admm_cnc_mri_b200 import _abi
    lib = _abi.load()
    B, N = 4, 256
    P = dict(kat.CNC_DEFAULTS, iter_num=12)
    masks = [torch.as_tensor(cs_inputs['masks'][k].copy()).pin_memory() for k in range(3)]
    noise = torch.view_as_real(torch.as_tensor(cs_inputs['noises']).to(torch.complex64)).contiguous().pin_memory()
    sets = {t: [torch.as_tensor(cs_inputs['images'][k:k + B].copy()).pin_memory() for k in ks]
            for t, ks in ((0, (0, 2, 4, 6, 8, 10)), (1, (1, 3, 5, 7, 9, 11)))}
    solver = pk.AdmmSolver(B, N)
    scratch = torch.empty(lib.pnpadmm_host_scratch_bytes(B, N), dtype=torch.uint8, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    want = {}
    for t in (0, 1):
        for i, im in enumerate(sets[t]):
            hx = torch.empty((B, N, N), dtype=torch.float32).pin_memory()
            _abi.check(lib.pnpadmm_reconstruct_host_f32(im.data_ptr(), masks[(i + t) % 3].data_ptr(), noise.data_ptr(), hx.data_ptr(),
                                                        B, N, _abi.PROX_CNC, P['iter_num'], P['lambda1'], P['reo'], P['alpha'], P['b'],
                                                        _abi.KERNEL_AUTO, scratch.data_ptr(), scratch.numel(), solver.ws.data_ptr(),
                                                        solver.ws_bytes, st))
            torch.cuda.synchronize()
            want[(t, i)] = hx.clone()
    got, errs = {}, []

    def worker(t):
        try:
            torch.cuda.set_device(0)
            pipe = pk.HostPipeline(B, N, n_slots=2)
            hx = [torch.empty((B, N, N), dtype=torch.float32).pin_memory() for _ in range(2)]
            n = len(sets[t])
            for i in range(n):
                if i >= 2:
                    pipe.wait(i & 1)
                    got[(t, i - 2)] = hx[i & 1].clone()
                pipe.submit(i & 1, sets[t][i], masks[(i + t) % 3], noise, hx[i & 1], prox='cnc', **P)
            for i in (n - 2, n - 1):
                pipe.wait(i & 1)
                got[(t, i)] = hx[i & 1].clone()
            pipe.close()
        except Exception as e:          # surfaced in the main thread
            errs.append(e)

    th = [threading.Thread(target=worker, args=(t,)) for t in (0, 1)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for k, v in want.items():
        assert torch.equal(got[k], v), k


def test_hybrid_schedule_parity(pk, cs_inputs, monkeypatch):
    """kernel='auto' at N = 256 splits a large batch: most packed planes on the cluster kernel, the rest on the
    K2 streaming kernels on a side stream (the SMs the clusters cannot use).  Both shares must meet the gate,
    including the odd tail image, and the forced splits must agree with the pure cluster run."""
    B = 45                                               # 23 packed planes, last one half empty
    idx = [i % 15 for i in range(B)]
    imgs = _imgs(cs_inputs, idx)
    m = cs_inputs['masks'][1]
    P = kat.CNC_DEFAULTS
    ref = {i: orc.admm_cnc(imgs[i], m.astype(np.float64), cs_inputs['noises'], **P) for i in range(15)}
    pure = pk.admm_solve(imgs, m, cs_inputs['noises'], prox='cnc', kernel='cluster', **P)
    for p2 in ('3', '7', None):                          # forced K2 shares, then the planner's own choice
        if p2 is None:
            monkeypatch.delenv('PNPADMM_HYBRID_P2', raising=False)
        else:
            monkeypatch.setenv('PNPADMM_HYBRID_P2', p2)
        x = pk.admm_solve(imgs, m, cs_inputs['noises'], prox='cnc', kernel='auto', **P)
        for k in range(B):
            assert rel(x[k], ref[idx[k]]) < TOL32, (p2, k)
        assert rel(x, pure) < 2e-5
        if p2 is not None:                               # the K1 share is bit-identical to the pure cluster run
            n1 = 2 * (23 - int(p2))
            assert np.array_equal(x[:n1], pure[:n1])


def test_hybrid_schedule_with_per_image_masks(pk, monkeypatch):
    """Hybrid split when every image has its own mask and noise (one image per plane: the K2 share needs
    the per-plane mask codes at the right offset)."""
    from pnp_admm_cnc_mri_b200 import data
    N, B = 256, 18
    imgs = data.phantoms(B, N, seed0=50)
    kinds = ('random', 'radial', 'cartesian')
    masks = np.stack([data.make_mask(kinds[s % 3], N, seed=s) for s in range(B)])
    nz = data.make_noise(N, seed=11, B=B)
    P = dict(kat.L1_DEFAULTS, iter_num=12)
    xr = oracle_batch(imgs, masks, nz, 'l1', P)[0]
    monkeypatch.setenv('PNPADMM_HYBRID_P2', '3')
    x = pk.admm_solve(imgs, masks, nz, prox='l1', kernel='auto', **P)
    for k in range(B):
        assert rel(x[k], xr[k]) < TOL32, k


def test_solve_under_cuda_graph_capture(pk, cs_inputs):
    """The whole stream-ordered solve (hybrid fork/join onto the library's side stream included) can be
    captured into a CUDA graph and replayed: no allocation, no host synchronisation inside the C ABI."""
    B, N = 40, 256
    idx = [i % 15 for i in range(B)]
    imgs = torch.as_tensor(_imgs(cs_inputs, idx)).cuda()
    m = torch.as_tensor(cs_inputs['masks'][0]).cuda()
    nz = torch.as_tensor(cs_inputs['noises']).to(torch.complex64).cuda()
    P = kat.CNC_DEFAULTS
    solver = pk.AdmmSolver(B, N)
    y = solver.acquire(imgs, m, nz)
    want = solver.solve(y, m, 'cnc', P['iter_num'], P['lambda1'], P['reo'], P['alpha'], P['b'])[0].clone()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            out = solver.solve(y, m, 'cnc', P['iter_num'], P['lambda1'], P['reo'], P['alpha'], P['b'])[0]
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)


@pytest.mark.parametrize('N,B,kernel', [(1024, 256, 'streaming'), (256, 1024, 'auto')])
def test_full_size_batches_replicate_small_ones(pk, N, B, kernel):
    """BASELINE config 5 sizes (hundreds of images per GPU, state far beyond L2 and 2^32 bytes): a batch made of
    4 distinct phantoms tiled B/4 times must give, for every copy, bit-identical results (same pairing partner,
    same kernel share), and the distinct ones must meet the gate against the oracle."""
    from pnp_admm_cnc_mri_b200 import data
    base = data.phantoms(4, N, seed0=70)
    imgs = np.concatenate([base] * (B // 4))
    m = data.make_mask('radial', N, seed=3)
    nz = data.make_noise(N, seed=13)
    P = dict(kat.CNC_DEFAULTS, iter_num=6)
    x = pk.admm_solve(torch.as_tensor(imgs).cuda(), m, nz, prox='cnc', kernel=kernel, **P)
    x = x.cpu().numpy() if isinstance(x, torch.Tensor) else x
    xr = oracle_batch(base, m, nz, 'cnc', P)[0]
    for k in range(4):
        assert rel(x[k], xr[k]) < TOL32, k
    reps = x.reshape(B // 4, 4, N, N)
    if kernel == 'streaming':
        assert np.array_equal(reps, np.broadcast_to(reps[:1], reps.shape))        # bit-identical copies
    else:   # hybrid: copies computed by the cluster kernel and by the streaming kernels differ only by rounding
        assert rel(reps, np.broadcast_to(reps[:1], reps.shape)) < 2e-5
        assert np.array_equal(reps[1], reps[0])                                    # both inside the K1 share


@pytest.mark.gpu
def test_k2_column_pass_variants_bit_identical(tmp_path):
    """The iteration's columns pass has three data-movement variants (2-D TMA tile loads = default, cp.async loads, TMA loads +
    TMA stores); the arithmetic is the same code, so the reconstructions must agree bit for bit.  The variants are selected by
    environment variables read once per process, hence the subprocesses."""
    import subprocess
    import sys
    script = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
        "import pnp_admm_cnc_mri_b200 as pk\n"
        "from pnp_admm_cnc_mri_b200 import data\n"
        "out = []\n"
        "for N in (256, 512, 1024):\n"
        "    imgs = data.phantoms(3, N, seed0=7)\n"
        "    x = pk.admm_solve(imgs, data.make_mask('radial', N, seed=2), data.make_noise(N, seed=4), prox='cnc', kernel='streaming',\n"
        "                      alpha=0.45, iter_num=4, lambda1=0.5, reo=0.05, b=64)\n"
        "    out.append(x.ravel())\n"
        "np.save(sys.argv[1], np.concatenate(out))\n")
    results = []
    for k, env_add in enumerate(({}, {'PNPADMM_COLS_LSU': '1'}, {'PNPADMM_COLS_TMA_STORE': '1'})):
        f = str(tmp_path / f'v{k}.npy')
        env = dict(os.environ, **env_add)
        subprocess.run([sys.executable, '-c', script, f], check=True, env=env, timeout=300)
        results.append(np.load(f))
    assert np.array_equal(results[0], results[1])
    assert np.array_equal(results[0], results[2])


# ---------------------------------------------------------------------------------------------
# pnpadmm_reconstruct_*: images -> reconstructions in one call; at N = 256 (fp32, shared mask + noise) acquisition,
# zero-filled start and the data term run inside the cluster kernel (fused prologue)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('prox,P', [('l1', kat.L1_DEFAULTS), ('cnc', kat.CNC_DEFAULTS)])
@pytest.mark.parametrize('mask_i', [0, 1, 2])
def test_fused_prologue_matches_oracle_and_unfused_path(pk, cs_inputs, prox, P, mask_i):
    idx = list(range(15))
    u8 = np.ascontiguousarray(cs_inputs['images'][idx])
    imgs = _imgs(cs_inputs, idx)
    m, nz = cs_inputs['masks'][mask_i], cs_inputs['noises']
    s = pk.AdmmSolver(15, 256)
    args = (prox, P['iter_num'], P['lambda1'], P['reo'], P.get('alpha', 0.0), P.get('b', 1.0))
    xf, zf, wf = (t.cpu().numpy() for t in s.reconstruct(imgs, m, nz, *args, kernel='cluster'))     # float images
    x8, z8, w8 = (t.cpu().numpy() for t in s.reconstruct(u8, m, nz, *args, kernel='cluster'))       # uint8 gray levels
    assert np.array_equal(xf, x8) and np.array_equal(zf, z8) and np.array_equal(wf, w8)
    y = s.acquire(imgs, m, nz)
    xo, zo, wo = (t.cpu().numpy() for t in s.solve(y, m, *args, kernel='cluster'))                  # unfused: acquire + solve
    xr, zr, wr, _ = oracle_batch(imgs, m, nz, prox, P)
    name = cs_inputs['mask_names'][mask_i]
    for k in range(15):
        assert rel(xf[k], xr[k]) < TOL32 and rel(zf[k], zr[k]) < TOL32, (name, prox, k)
        assert rel(xf[k], xo[k].astype(np.float64)) < TOL32          # two fp32 paths, each ~1e-5 from the oracle under CNC
        p = orc.calculate_psnr(xf[k].astype(np.float64) * 255, cs_inputs['images'][k])
        assert abs(p - kat.PSNR[(name, prox)][k]) < TOL_PSNR, (name, prox, k, p)
    # the dual: same accuracy as the unfused path reaches against the oracle (a few pixels sit on a threshold)
    ew_f, ew_o = rel(wf, wr), rel(wo, wr)
    print(f'{name} {prox}: dual w rel-L2 vs oracle: fused {ew_f:.2e}, unfused {ew_o:.2e}')
    assert ew_f < max(3 * ew_o, 1e-4)


def test_fused_prologue_hybrid_split_and_fallbacks(pk, cs_inputs, monkeypatch):
    """The fused call under the hybrid schedule (streaming share does its own acquisition on the side stream, from float and
    from uint8 images), and the cases that fall back to acquisition + solve: N != 256, per-image masks, float64."""
    from pnp_admm_cnc_mri_b200 import data
    B = 45
    idx = [i % 15 for i in range(B)]
    u8 = np.ascontiguousarray(cs_inputs['images'][idx])
    imgs = _imgs(cs_inputs, idx)
    m, nz = cs_inputs['masks'][0], cs_inputs['noises']
    P = kat.CNC_DEFAULTS
    args = ('cnc', P['iter_num'], P['lambda1'], P['reo'], P['alpha'], P['b'])
    ref = {i: orc.admm_cnc(imgs[i], m.astype(np.float64), nz, **P) for i in range(15)}
    s = pk.AdmmSolver(B, 256)
    pure = s.reconstruct(imgs, m, nz, *args, kernel='cluster')[0].cpu().numpy()
    for p2 in ('5', None):
        if p2 is None:
            monkeypatch.delenv('PNPADMM_HYBRID_P2', raising=False)
        else:
            monkeypatch.setenv('PNPADMM_HYBRID_P2', p2)
        for src in (imgs, u8):
            x = s.reconstruct(src, m, nz, *args, kernel='auto')[0].cpu().numpy()
            for k in range(B):
                assert rel(x[k], ref[idx[k]]) < TOL32, (p2, k)
            if p2 is not None:
                n1 = 2 * (23 - int(p2))
                assert np.array_equal(x[:n1], pure[:n1])             # the K1 share is the pure cluster run, bit for bit
    monkeypatch.delenv('PNPADMM_HYBRID_P2', raising=False)
    # fallbacks
    N2 = 512
    im2 = data.phantoms(3, N2, seed0=3)
    m2, n2 = data.make_mask('radial', N2, seed=1), data.make_noise(N2, seed=2)
    Pq = dict(P, iter_num=8)
    aq = ('cnc', 8, P['lambda1'], P['reo'], P['alpha'], P['b'])
    x = pk.AdmmSolver(3, N2).reconstruct(np.uint8((im2 * 255).round()), m2, n2, *aq)[0].cpu().numpy()
    xr = oracle_batch(np.float32(np.uint8((im2 * 255).round()) / 255.), m2, n2, 'cnc', Pq)[0]
    assert rel(x, xr) < TOL32
    masks = np.stack([cs_inputs['masks'][k % 3] for k in range(4)])
    x = pk.AdmmSolver(4, 256, mask_batched=True).reconstruct(imgs[:4], masks, nz, *aq)[0].cpu().numpy()
    xr = oracle_batch(imgs[:4], masks, nz, 'cnc', Pq)[0]
    assert rel(x, xr) < TOL32
    x = pk.AdmmSolver(2, 256, dtype='float64').reconstruct(imgs[:2], m, nz, *aq)[0].cpu().numpy()
    xr = oracle_batch(imgs[:2], m, nz, 'cnc', Pq)[0]
    assert rel(x, xr) < TOL64


# ---------------------------------------------------------------------------------------------
# K3: row-separable masks (full k-space lines, the reference's Q_Cartesian30): every image row is solved on its own
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('prox,P', [('l1', kat.L1_DEFAULTS), ('cnc', kat.CNC_DEFAULTS)])
def test_rowsep_kernel_kat_table_cartesian(pk, cs_inputs, prox, P):
    """All 15 set images under Q_Cartesian30 on the row-separable kernel (explicitly, and picked by kernel='auto'), float and
    uint8 images, against the oracle, the reference's log table and the general cluster kernel."""
    m, nz = cs_inputs['masks'][2], cs_inputs['noises']
    assert cs_inputs['mask_names'][2] == 'Q_Cartesian30' and pk.mask_is_row_separable(m)
    assert not pk.mask_is_row_separable(cs_inputs['masks'][0]) and not pk.mask_is_row_separable(cs_inputs['masks'][1])
    imgs = _imgs(cs_inputs, range(15))
    u8 = np.ascontiguousarray(cs_inputs['images'][:15])
    s = pk.AdmmSolver(15, 256)
    args = (prox, P['iter_num'], P['lambda1'], P['reo'], P.get('alpha', 0.0), P.get('b', 1.0))
    x, z, w = (t.cpu().numpy() for t in s.reconstruct(imgs, m, nz, *args, kernel='rowsep'))
    xa = s.reconstruct(u8, torch.as_tensor(m).cuda(), nz, *args, kernel='auto')[0].cpu().numpy()     # device mask: tested once, cached
    assert np.array_equal(x, xa)
    xc = s.reconstruct(imgs, m, nz, *args, kernel='cluster')[0].cpu().numpy()
    xr, zr, wr, _ = oracle_batch(imgs, m, nz, prox, P)
    for k in range(15):
        assert rel(x[k], xr[k]) < TOL32 and rel(z[k], zr[k]) < TOL32, (prox, k)
        assert rel(x[k], xc[k].astype(np.float64)) < TOL32
        p = orc.calculate_psnr(x[k].astype(np.float64) * 255, cs_inputs['images'][k])
        assert abs(p - kat.PSNR[('Q_Cartesian30', prox)][k]) < TOL_PSNR, (prox, k, p)
    print(f'rowsep {prox}: worst rel-L2 vs oracle {max(rel(x[k], xr[k]) for k in range(15)):.2e}')
    # admm_solve picks it for a separable mask; results identical
    assert np.array_equal(pk.admm_solve(imgs, m, nz, prox=prox, **P), x)


def test_rowsep_kernel_refuses_other_masks_and_covers_batches(pk, cs_inputs):
    from pnp_admm_cnc_mri_b200 import data
    nz = cs_inputs['noises']
    imgs = _imgs(cs_inputs, range(3))
    P = kat.CNC_DEFAULTS
    args = ('cnc', P['iter_num'], P['lambda1'], P['reo'], P['alpha'], P['b'])
    # asserted separability is verified on the device: a radial mask yields NaN, not a wrong reconstruction
    x, z, w = pk.AdmmSolver(3, 256).reconstruct(imgs, cs_inputs['masks'][1], nz, *args, kernel='rowsep')
    assert torch.isnan(x).all() and torch.isnan(z).all() and torch.isnan(w).all()
    with pytest.raises(pk.PnpAdmmError):                  # sizes without a row-separable kernel say so
        pk.AdmmSolver(2, 128).reconstruct(data.phantoms(2, 128, 0), data.make_mask('cartesian', 128, 0), data.make_noise(128, 1), *args,
                                          kernel='rowsep')
    # a synthetic Cartesian mask, a batch larger than one wave of CTAs (2 x 148 SMs x 16 rows), odd count
    B = 75
    m = data.make_mask('cartesian', 256, seed=7)
    assert pk.mask_is_row_separable(m)
    base = data.phantoms(5, 256, seed0=9)
    big = np.concatenate([base] * 15)[:B]
    x = pk.AdmmSolver(B, 256).reconstruct(big, m, nz, *args)[0].cpu().numpy()
    xr = oracle_batch(base, m, nz, 'cnc', P)[0]
    for k in range(B):
        assert rel(x[k], xr[k % 5]) < TOL32, k
    reps = x[:70].reshape(7, 10, 256, 256)                                    # period 10: same pair slot, same partner image
    assert np.array_equal(reps, np.broadcast_to(reps[:1], reps.shape))        # rows are independent: such copies are bit-identical


@pytest.mark.parametrize('N,prox', [(512, 'cnc'), (512, 'l1'), (1024, 'cnc'), (1024, 'l1')])
def test_rowsep_kernel_large_sizes(pk, N, prox):
    """K3 on the K2 line FFT (rowsepN.cuh): Cartesian masks at N = 512 / 1024, the reference depth (50 iterations), odd batch, uint8
    images in; picked by kernel='auto'.  Against the oracle and against the general streaming kernels."""
    from pnp_admm_cnc_mri_b200 import data
    B = 3
    u8 = np.uint8((data.phantoms(B, N, seed0=80 + N // 512) * 255).round())
    imgs = np.float32(u8 / 255.)
    m = data.make_mask('cartesian', N, seed=6)
    nz = data.make_noise(N, seed=10)
    assert pk.mask_is_row_separable(m)
    P = dict(kat.L1_DEFAULTS) if prox == 'l1' else dict(kat.CNC_DEFAULTS)
    args = (prox, P['iter_num'], P['lambda1'], P['reo'], P.get('alpha', 0.0), P.get('b', 1.0))
    s = pk.AdmmSolver(B, N)
    x, z, w = (t.cpu().numpy() for t in s.reconstruct(u8, m, nz, *args))                  # auto -> rowsep
    xs = s.reconstruct(imgs, m, nz, *args, kernel='streaming')[0].cpu().numpy()
    xr, zr, wr, _ = oracle_batch(imgs, m, nz, prox, P, workers=B)
    for k in range(B):
        e = rel(x[k], xr[k])
        print(f'rowsepN N={N} {prox} image {k}: rel-L2 vs oracle {e:.2e} (streaming kernels {rel(xs[k], xr[k]):.2e})')
        assert e < TOL32 and rel(z[k], zr[k]) < TOL32, (N, prox, k)
        assert rel(x[k], xs[k].astype(np.float64)) < TOL32
    # a non-separable mask is refused loudly at these sizes too
    bad = s.reconstruct(u8, data.make_mask('radial', N, seed=1), nz, *args, kernel='rowsep')[0]
    assert torch.isnan(bad).all()


def test_plan_info_reports_the_schedule(pk):
    """pnpadmm_plan_info (what bench.py derives gpu_launches from): planes per kernel add up, the fused reconstruct launches far
    fewer kernels than acquire + solve, a batch that fits the resident clusters is not split."""
    import ctypes
    from pnp_admm_cnc_mri_b200 import _abi
    lib = _abi.load()
    sm, ncl = ctypes.c_int(), ctypes.c_int()
    _abi.check(lib.pnpadmm_device_info(sm, ncl, None, None))
    v = [ctypes.c_int() for _ in range(6)]
    _abi.check(lib.pnpadmm_plan_info(64, 256, 0, 50, _abi.KERNEL_AUTO, *v))
    p1, p2, chunks, la, ls, lr = (x.value for x in v)
    assert p1 + p2 == 32 and p1 >= ncl.value and chunks >= 1 and la == 2
    assert ls == 6 + (1 if p1 else 0) + ((1 + 100) if p2 else 0)
    assert lr == 1 + (1 if p1 else 0) + ((7 + 100) if p2 else 0) and lr < la + ls
    _abi.check(lib.pnpadmm_plan_info(2 * ncl.value, 256, 0, 50, _abi.KERNEL_AUTO, *v))
    assert v[0].value == ncl.value and v[1].value == 0 and v[5].value == 2          # one preparation launch + one cluster kernel
    _abi.check(lib.pnpadmm_plan_info(8, 1024, 0, 10, _abi.KERNEL_AUTO, *v))
    assert v[0].value == 0 and v[1].value == 4 and v[5].value == v[3].value + v[4].value
    assert lib.pnpadmm_plan_info(8, 100, 0, 10, _abi.KERNEL_AUTO, *v) == _abi.ERR_BAD_SIZE


@pytest.mark.gpu
def test_experiment_switches_keep_parity(tmp_path):
    """The A/B switches of the kernels (read once per process, hence the subprocesses) select other code for the same
    arithmetic: the blocked-tile / bulk-copy transposes of K1, the first row-separable kernel, the K2 share on two streams,
    plain launches instead of programmatic dependent launch, acquisition + solve instead of the fused prologue, the literal
    planner constants.  Each must reproduce the default build's reconstructions within the gate (bit for bit where the
    arithmetic is literally the same code)."""
    import subprocess
    import sys
    script = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
        "import pnp_admm_cnc_mri_b200 as pk\n"
        "from pnp_admm_cnc_mri_b200 import data\n"
        "out = []\n"
        "nz = data.make_noise(256, seed=4)\n"
        "for kind, B in (('random', 37), ('cartesian', 5)):\n"
        "    imgs = np.concatenate([data.phantoms(5, 256, seed0=7)] * 8)[:B]\n"
        "    x = pk.admm_solve(imgs, data.make_mask(kind, 256, seed=2), nz, prox='cnc', alpha=0.45, iter_num=12, lambda1=0.5, reo=0.05, b=64)\n"
        "    out.append(x.ravel())\n"
        "np.save(sys.argv[1], np.concatenate(out))\n")
    # Bit-exact (tolerance 0.0) expectations need the SAME hybrid plan in both runs.  The planner's constants are measured per process
    # (calibrate_hybrid) with whatever kernels the switches select, so a switch can move a plane between the K1 and the K2 share and with
    # it the last bits of two images: every run but the last uses the literal constants (PNPADMM_NO_CALIBRATE=1); the last one is the
    # default library (calibration on) against the same base at the gate.
    NC = {'PNPADMM_NO_CALIBRATE': '1'}
    variants = [(NC, None), (dict(NC, PNPADMM_K1_BULK='1', PNPADMM_NO_FUSED_PROLOGUE='1'), 1e-4), (dict(NC, PNPADMM_K1_BULK='1'), 0.0),
                (dict(NC, PNPADMM_K3_K1CODE='1'), 1e-4), (dict(NC, PNPADMM_K2_SPLIT='1'), 0.0), (dict(NC, PNPADMM_NO_PDL='1'), 0.0),
                (dict(NC, PNPADMM_NO_FUSED_PROLOGUE='1', PNPADMM_NO_ROWSEP='1'), 1e-4), ({}, 1e-4)]
    base = None
    for k, (env_add, tol) in enumerate(variants):
        f = str(tmp_path / f'v{k}.npy')
        env = {kk: vv for kk, vv in os.environ.items() if kk != 'PNPADMM_NO_CALIBRATE'}
        subprocess.run([sys.executable, '-c', script, f], check=True, env=dict(env, **env_add), timeout=300)
        got = np.load(f)
        assert np.isfinite(got).all(), env_add
        if base is None:
            base = got
        elif tol == 0.0:
            assert np.array_equal(got, base), env_add
        else:
            assert rel(got, base.astype(np.float64)) < tol, env_add
