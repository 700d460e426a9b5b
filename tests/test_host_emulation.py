"""CPU: run the K1 cluster kernel's per-thread phase code (cluster256_core.cuh is host+device)
for 8 emulated CTAs x 512 threads and compare with the fp64 oracle.  Validates the 16x16 FFT
decomposition, swizzles, thread<->pixel maps and DSMEM offsets without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import kat_table as kat
from oracle import reference_numpy as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, 'tests', 'host_emu')
FP = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope='module')
def emu():
    so = os.path.join(EMU_DIR, 'k1_emu.so')
    src = os.path.join(EMU_DIR, 'k1_emu.cpp')
    core = os.path.join(ROOT, 'pnp_admm_cnc_mri_b200', 'csrc', 'cluster256_core.cuh')
    common = os.path.join(ROOT, 'pnp_admm_cnc_mri_b200', 'csrc', 'common.cuh')
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in (src, core, common)):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-mfma', '-ffp-contract=fast', '-shared', '-fPIC',
                               '-o', so, src])
    return ctypes.CDLL(so)


def mirror(a):
    return np.roll(np.flip(a, (0, 1)), 1, (0, 1))


def prepare_np(ys, m, reo):
    """NumPy statement of prepare_kernel (streaming.cuh): packed Hermitian-symmetrised data term."""
    N = m.shape[0]
    g = 1 / (1 + 1 / (2 * reo))
    mm = m.astype(np.float64)
    Ys = lambda y: 0.5 * (mm * y + mirror(mm) * np.conj(mirror(y)))
    G = []
    for p in range((len(ys) + 1) // 2):
        a = Ys(ys[2 * p])
        b = Ys(ys[2 * p + 1]) if 2 * p + 1 < len(ys) else 0
        G.append(g / N ** 2 * (a + 1j * b))
    return np.stack(G).astype(np.complex64), (m + mirror(m)).astype(np.uint8), [(g * c / 2) / N ** 2 for c in range(3)]


def run_emu(emu, imgs, m, noises, prox, P, cluster=8):
    ys = [orc.acquire(im, m.astype(np.float64), noises) for im in imgs]
    z0 = np.stack([orc.zero_filled(y) for y in ys]).astype(np.float32)
    w0 = np.zeros_like(z0)
    G, mc, cf = prepare_np(ys, m, P['reo'])
    B = len(imgs)
    x, z, w = np.zeros_like(z0), np.zeros_like(z0), np.zeros_like(z0)
    a, l, reo, b = P.get('alpha', 0.), P['lambda1'], P['reo'], P.get('b', 1.)
    f = ctypes.c_float
    emu.k1_emulate(z0.ctypes.data_as(FP), w0.ctypes.data_as(FP), x.ctypes.data_as(FP), z.ctypes.data_as(FP),
                   w.ctypes.data_as(FP), None, G.ctypes.data_as(FP), mc.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), 0,
                   f(cf[0]), f(cf[1]), f(cf[2]), B, (B + 1) // 2, 0, P['iter_num'], 0 if prox == 'l1' else 1,
                   f(reo * l), f(1 / b), f(1 - a), f(a), f(a * reo * l * b), f(a * reo * l), cluster)
    return x, z, w


def test_fft256_line(emu):
    rng = np.random.default_rng(0)
    for inv in (0, 1):
        x = (rng.standard_normal(256) + 1j * rng.standard_normal(256)).astype(np.complex64)
        out = np.zeros(256, np.complex64)
        emu.k1_fft256_line(x.ctypes.data_as(FP), out.ctypes.data_as(FP), inv)
        ref = np.fft.ifft(x.astype(np.complex128)) * 256 if inv else np.fft.fft(x.astype(np.complex128))
        assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 3e-7


@pytest.mark.parametrize('cluster', [8, 16, 116])      # 116 = 16-CTA geometry with blocked tiles + bulk-copy transposes
@pytest.mark.parametrize('prox,P', [('l1', kat.L1_DEFAULTS), ('cnc', kat.CNC_DEFAULTS)])
def test_k1_emulated_solve_matches_oracle(emu, cs_inputs, prox, P, cluster):
    idx = [4, 0, 7]                                      # odd count: last plane has an empty b slot
    imgs = [orc.preprocess_uint8(cs_inputs['images'][i]) for i in idx]
    m = cs_inputs['masks'][1]
    x, z, w = run_emu(emu, imgs, m, cs_inputs['noises'], prox, P, cluster)
    fn = orc.admm_l1 if prox == 'l1' else orc.admm_cnc
    for k, i in enumerate(idx):
        xr, zr, wr, _ = fn(imgs[k], m.astype(np.float64), cs_inputs['noises'], return_state=True, **P)
        assert np.linalg.norm(x[k] - xr) / np.linalg.norm(xr) < 1e-4
        assert np.linalg.norm(z[k] - zr) / np.linalg.norm(zr) < 1e-4
        p = orc.calculate_psnr(x[k].astype(np.float64) * 255, cs_inputs['images'][i])
        assert abs(p - kat.PSNR[('Q_Radial30', prox)][i]) < 0.01


@pytest.mark.parametrize('cluster', [8, 16])
@pytest.mark.parametrize('prox,P', [('l1', kat.L1_DEFAULTS), ('cnc', kat.CNC_DEFAULTS)])
def test_k1_emulated_fused_prologue_matches_oracle(emu, cs_inputs, prox, P, cluster):
    """The fused prologue (acquisition, zero-filled start and data term inside the cluster kernel: packed F = fft2(a + i b),
    G = cf F + noise term, |ifft2(y)| from the symmetric / antisymmetric mask branches) followed by the loop, from the uint8 images."""
    idx = [4, 0, 7]                                      # odd count: the last plane's imaginary slot must stay zero
    m = cs_inputs['masks'][2]
    nz = cs_inputs['noises']
    u8 = np.ascontiguousarray(cs_inputs['images'][idx])
    noise = np.ascontiguousarray(nz.astype(np.complex64)).view(np.float32)
    B = len(idx)
    x, z, w = (np.zeros((B, 256, 256), np.float32) for _ in range(3))
    a, l, reo, b = P.get('alpha', 0.), P['lambda1'], P['reo'], P.get('b', 1.)
    f = ctypes.c_float
    U8 = ctypes.POINTER(ctypes.c_uint8)
    mask = np.ascontiguousarray(m.astype(np.uint8))
    for it in (0, P['iter_num']):
        emu.k1_emulate_fused(None, u8.ctypes.data_as(U8), mask.ctypes.data_as(U8), noise.ctypes.data_as(FP), f(reo),
                             x.ctypes.data_as(FP), z.ctypes.data_as(FP), w.ctypes.data_as(FP), B, max(it, 1), 0 if prox == 'l1' else 1,
                             f(reo * l), f(1 / b), f(1 - a), f(a), f(a * reo * l * b), f(a * reo * l), cluster)
        if it == 0:
            continue
        fn = orc.admm_l1 if prox == 'l1' else orc.admm_cnc
        for k, i in enumerate(idx):
            xr, zr, wr, _ = fn(orc.preprocess_uint8(cs_inputs['images'][i]), m.astype(np.float64), nz, return_state=True, **P)
            assert np.linalg.norm(x[k] - xr) / np.linalg.norm(xr) < 1e-4
            assert np.linalg.norm(z[k] - zr) / np.linalg.norm(zr) < 1e-4
            p = orc.calculate_psnr(x[k].astype(np.float64) * 255, cs_inputs['images'][i])
            assert abs(p - kat.PSNR[('Q_Cartesian30', prox)][i]) < 0.01


@pytest.mark.parametrize('prox,P', [('l1', kat.L1_DEFAULTS), ('cnc', kat.CNC_DEFAULTS)])
def test_k3_emulated_rowsep_solve_matches_oracle(emu, cs_inputs, prox, P):
    """K3 (rowsep256.cuh): under the reference's Q_Cartesian30 mask (76 full columns) the blend commutes with the column
    transforms and every image row is solved on its own.  Same per-thread code as the device kernel, from the uint8 images."""
    idx = [4, 0, 7]
    m = cs_inputs['masks'][2]
    assert cs_inputs['mask_names'][2] == 'Q_Cartesian30' and (m == m[:1]).all()
    nz = cs_inputs['noises']
    u8 = np.ascontiguousarray(cs_inputs['images'][idx])
    noise = np.ascontiguousarray(nz.astype(np.complex64)).view(np.float32)
    mask = np.ascontiguousarray(m.astype(np.uint8))
    B = len(idx)
    a, l, reo, b = P.get('alpha', 0.), P['lambda1'], P['reo'], P.get('b', 1.)
    g = 1 / (1 + 1 / (2 * reo))
    f = ctypes.c_float
    U8 = ctypes.POINTER(ctypes.c_uint8)
    planes = np.zeros((3, 256, 256), np.complex64)
    emu.k3_noise_terms(mask.ctypes.data_as(U8), noise.ctypes.data_as(FP), f(g / 65536.), planes.view(np.float32).ctypes.data_as(FP))
    planes = np.ascontiguousarray((np.fft.ifft(planes.astype(np.complex128), axis=1) * 256).astype(np.complex64))   # column IFFT, unnormalised
    x, z, w = (np.zeros((B, 256, 256), np.float32) for _ in range(3))
    rc = emu.k3_emulate(None, u8.ctypes.data_as(U8), mask.ctypes.data_as(U8), planes.view(np.float32).ctypes.data_as(FP), f(reo),
                        x.ctypes.data_as(FP), z.ctypes.data_as(FP), w.ctypes.data_as(FP), B, P['iter_num'], 0 if prox == 'l1' else 1,
                        f(reo * l), f(1 / b), f(1 - a), f(a), f(a * reo * l * b), f(a * reo * l))
    assert rc == 0
    fn = orc.admm_l1 if prox == 'l1' else orc.admm_cnc
    for k, i in enumerate(idx):
        xr, zr, wr, _ = fn(orc.preprocess_uint8(cs_inputs['images'][i]), m.astype(np.float64), nz, return_state=True, **P)
        assert np.linalg.norm(x[k] - xr) / np.linalg.norm(xr) < 1e-4
        assert np.linalg.norm(z[k] - zr) / np.linalg.norm(zr) < 1e-4
        p = orc.calculate_psnr(x[k].astype(np.float64) * 255, cs_inputs['images'][i])
        assert abs(p - kat.PSNR[('Q_Cartesian30', prox)][i]) < 0.01
    # a mask that is not made of full lines is refused
    assert emu.k3_emulate(None, u8.ctypes.data_as(U8), np.ascontiguousarray(cs_inputs['masks'][0].astype(np.uint8)).ctypes.data_as(U8),
                          planes.view(np.float32).ctypes.data_as(FP), f(reo), x.ctypes.data_as(FP), z.ctypes.data_as(FP),
                          w.ctypes.data_as(FP), B, 1, 0, f(0), f(1), f(1), f(0), f(0), f(0)) == 1


# ---------------------------------------------------------------------------------------------
# K2 streaming kernels: the per-thread FFT stages (stream2_core.cuh) against a naive double DFT
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def s2emu():
    so = os.path.join(EMU_DIR, 's2_emu.so')
    src = os.path.join(EMU_DIR, 's2_emu.cpp')
    deps = [src] + [os.path.join(ROOT, 'pnp_admm_cnc_mri_b200', 'csrc', f)
                    for f in ('stream2_core.cuh', 'cluster256_core.cuh', 'common.cuh', 'rowsep_core.cuh')]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in deps):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', so, src])
    lib = ctypes.CDLL(so)
    lib.s2_fft_check.restype = ctypes.c_double
    lib.s2_fft_check.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint]
    return lib


@pytest.mark.parametrize('N', [256, 512, 1024])
@pytest.mark.parametrize('inv', [0, 1])
@pytest.mark.parametrize('layout', [0, 1], ids=['rows', 'cols'])
def test_k2_fft_stages(s2emu, N, inv, layout):
    # radix 16 x 16 [x 2 | x 4] Stockham stages, row (padded) and column ([N][C]) shared-memory layouts
    for seed in (1, 2):
        assert 0 <= s2emu.s2_fft_check(N, inv, layout, seed) < 3e-7


def test_k2_packed_codes(s2emu):
    rng = np.random.default_rng(0)
    for N in (256, 512, 1024):
        mc = rng.integers(0, 3, (N, N), dtype=np.uint8)
        T = N // 16
        for t, kc in ((0, 0), (T - 1, N - 1), (3, 17)):
            w = s2emu.s2_pack_codes(mc.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), N, t, kc)
            assert [(w >> (2 * m)) & 3 for m in range(16)] == [int(mc[t + T * m, kc]) for m in range(16)]


@pytest.mark.parametrize('N,prox', [(256, 'cnc'), (512, 'cnc'), (512, 'l1'), (1024, 'cnc')])
def test_k3n_emulated_rowsep_solve_matches_oracle(s2emu, N, prox):
    """K3 on the K2 line FFT (rowsepN.cuh, N = 512 / 1024; 256 for A/B): synthetic Cartesian mask (full k-space lines), uint8 phantoms,
    odd batch; the per-thread code of the kernel run line by line against the fp64 oracle."""
    from pnp_admm_cnc_mri_b200 import data
    B = 3 if N < 1024 else 1
    P = dict(kat.L1_DEFAULTS) if prox == 'l1' else dict(kat.CNC_DEFAULTS)
    P['iter_num'] = 50 if N <= 512 else 12
    u8 = np.ascontiguousarray(np.uint8((data.phantoms(B, N, seed0=3) * 255).round()))
    m = data.make_mask('cartesian', N, seed=5)
    nz = data.make_noise(N, seed=6)
    noise = np.ascontiguousarray(nz.astype(np.complex64)).view(np.float32)
    mask = np.ascontiguousarray(m.astype(np.uint8))
    a, l, reo, b = P.get('alpha', 0.), P['lambda1'], P['reo'], P.get('b', 1.)
    g = 1 / (1 + 1 / (2 * reo))
    f = ctypes.c_float
    U8 = ctypes.POINTER(ctypes.c_uint8)
    planes = np.zeros((3, N, N), np.complex64)
    s2emu.k3n_noise_terms(mask.ctypes.data_as(U8), noise.ctypes.data_as(FP), N, f(g / (N * N)), planes.view(np.float32).ctypes.data_as(FP))
    planes = np.ascontiguousarray((np.fft.ifft(planes.astype(np.complex128), axis=1) * N).astype(np.complex64))   # column IFFT, unnormalised
    x, z, w = (np.zeros((B, N, N), np.float32) for _ in range(3))
    rc = s2emu.k3n_emulate(N, u8.ctypes.data_as(U8), mask.ctypes.data_as(U8), planes.view(np.float32).ctypes.data_as(FP), f(reo),
                           x.ctypes.data_as(FP), z.ctypes.data_as(FP), w.ctypes.data_as(FP), B, P['iter_num'], 0 if prox == 'l1' else 1,
                           f(reo * l), f(1 / b), f(1 - a), f(a), f(a * reo * l * b), f(a * reo * l))
    assert rc == 0
    fn = orc.admm_l1 if prox == 'l1' else orc.admm_cnc
    for k in range(B):
        xr, zr, wr, _ = fn(np.float32(u8[k] / 255.), m.astype(np.float64), nz, return_state=True, **P)
        assert np.linalg.norm(x[k] - xr) / np.linalg.norm(xr) < 1e-4, (N, prox, k)
        assert np.linalg.norm(z[k] - zr) / np.linalg.norm(zr) < 1e-4


def test_blend_coefficient_product_is_exact():
    """common.cuh blend_coef: the kernels compute cf[code] as (float)code * cf[1] (and cf[1] + cf[1] in the select form) instead of
    reading cf[2].  pnpadmm.cu rounds cf[1] = 0.5 g / N^2 and cf[2] = g / N^2 from the same double, so the three forms agree bit for bit
    for every reo and N the library accepts (checked here in NumPy float32 / float64 arithmetic, which is IEEE like the device's)."""
    import numpy as np
    rng = np.random.default_rng(0)
    for N in (16, 64, 256, 512, 1024, 2048):
        for reo in np.concatenate([[0.015, 0.05, 0.26, 0.45, 0.8], rng.uniform(1e-4, 10.0, 200)]):
            g = 1.0 / (1.0 + 1.0 / (2.0 * reo))                    # g = 1 / (1 + La2), La2 = 1 / (2 reo)      (S1:117)
            for denom in (float(N) * N, float(N)):                  # cf = g ms / N^2 (K1 / K2) and N cf = g ms / N (K3)
                c1, c2 = np.float32(0.5 * g / denom), np.float32(g / denom)
                assert np.float32(2.0) * c1 == c2 and c1 + c1 == c2 and np.float32(0.0) * c1 == 0.0
                c1d, c2d = 0.5 * g / denom, g / denom
                assert 2.0 * c1d == c2d
