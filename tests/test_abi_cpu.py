"""CPU: the C-ABI library loads and exports every symbol include/pnpadmm.h declares; argument
validation that needs no GPU; the package refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from pnp_admm_cnc_mri_b200 import build
    build.build_library()                # no-op when the in-tree .so is fresh
    from pnp_admm_cnc_mri_b200 import _abi
    return _abi.load()


def test_header_symbols_all_exported(lib):
    from pnp_admm_cnc_mri_b200 import _abi
    hdr = open(os.path.join(ROOT, 'include', 'pnpadmm.h')).read()
    declared = sorted(set(re.findall(r'\b(pnpadmm_[a-z0-9_]+)\s*\(', hdr)))
    assert declared == sorted(_abi.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.pnpadmm_abi_version() == _abi.ABI_VERSION == 3


def test_workspace_sizes(lib):
    nn = 256 * 256
    w = lib.pnpadmm_workspace_bytes(64, 256, 0, 0)
    assert w >= (64 + 32 + 32) * nn * 8 + nn
    assert lib.pnpadmm_workspace_bytes(64, 256, 1, 0) > w
    assert lib.pnpadmm_workspace_bytes(64, 256, 0, 1) > w          # per-image masks: no pairing
    assert lib.pnpadmm_workspace_bytes(0, 256, 0, 0) == 0
    assert lib.pnpadmm_host_scratch_bytes(64, 256) >= 64 * nn * (1 + 4 + 8 + 12)


def test_argument_validation_without_gpu(lib):
    # bad sizes / NULLs are rejected before any CUDA call
    rc = lib.pnpadmm_solve_f32(None, None, None, None, None, 1, 256, 0, 0, 1, 0.1, 0.1, 0.0, 1.0, 0, None, 0, None)
    assert rc == -1
    buf = (ctypes.c_char * 1024)()
    p = ctypes.addressof(buf)
    rc = lib.pnpadmm_solve_f32(p, p, p, p, p, 1, 100, 0, 0, 1, 0.1, 0.1, 0.0, 1.0, 0, p, 1024, None)
    assert rc == -2
    assert b'power of two' in lib.pnpadmm_last_error_string()
    rc = lib.pnpadmm_iterate_f32(p, p, p, 1, 256, 0, 7, 1, 0.1, 0.1, 0.0, 1.0, 0, p, 1024, None)
    assert rc == -1 and b'prox' in lib.pnpadmm_last_error_string()
    rc = lib.pnpadmm_soft_f32(None, None, 0.1, 10, None)
    assert rc == -1
    # pipelined host call: the pipeline object is caller-owned (two pipelines may share a device); without one the
    # call is rejected before anything else, and creating one needs a device
    assert lib.pnpadmm_reconstruct_host_pipelined_f32(None, p, p, p, p, 4, 256, 1, 5, 0.5, 0.05, 0.45, 64.0, 0,
                                                      p, 1024, p, 1024, 0, None, None, None) == -1
    assert b'pipeline' in lib.pnpadmm_last_error_string()
    assert lib.pnpadmm_reconstruct_host_wait(None, 0) == -1
    assert lib.pnpadmm_pipeline_destroy(None) == 0
    h = ctypes.c_void_p()
    assert lib.pnpadmm_pipeline_create(ctypes.byref(h), 7) == -1 and b'n_slots' in lib.pnpadmm_last_error_string()
    assert lib.pnpadmm_pipeline_create(None, 2) == -1
    assert lib.pnpadmm_host_pipeline_scratch_bytes(64, 256, 2) > lib.pnpadmm_host_scratch_bytes(64, 256)
    assert lib.pnpadmm_host_pipeline_scratch_bytes(64, 256, 3) > lib.pnpadmm_host_pipeline_scratch_bytes(64, 256, 2)
    assert lib.pnpadmm_host_pipeline_scratch_bytes(64, 256, 1) == 0 and lib.pnpadmm_host_pipeline_scratch_bytes(64, 256, 5) == 0
    big = (ctypes.c_char * 4096)()
    q = (ctypes.addressof(big) + 255) // 256 * 256
    # device metrics: NULLs, unsupported size, scratch too small
    assert lib.pnpadmm_metrics_f32(None, None, 1, 256, 0, None, None, 0, None) == -1
    assert lib.pnpadmm_metrics_f32(p, p, 1, 8, 0, p, p, 1024, None) == -2
    assert lib.pnpadmm_metrics_f64(p, p, 64, 256, 0, p, p, 16, None) == -3
    assert lib.pnpadmm_metrics_scratch_bytes(64) == 64 * 32 and lib.pnpadmm_metrics_scratch_bytes(0) == 0
    # tensor-core denoiser: NULLs, misaligned pointers, channel counts, degenerate sizes
    assert lib.pnpadmm_conv64_bf16(None, None, None, None, 1, 8, 8, 1, None) == -1
    assert lib.pnpadmm_conv64_bf16(q + 2, q, q, q, 1, 8, 8, 1, None) == -1
    assert b'aligned' in lib.pnpadmm_last_error_string()
    assert lib.pnpadmm_dncnn_forward_bf16(None, None, 1, 1, 8, 8, 15, None, None, None, None, None, None, 1, None, None, None) == -1
    assert lib.pnpadmm_dncnn_forward_bf16(q, q, 1, 3, 8, 8, 15, q, q, q, q, q, q, 1, q, q, None) == -5
    assert b'input channels' in lib.pnpadmm_last_error_string()
    assert lib.pnpadmm_dncnn_forward_bf16(q, q, 1, 1, 8, 8, -1, q, q, q, q, q, q, 1, q, q, None) == -1
    assert lib.pnpadmm_dncnn_activation_bytes(256, 256, 256) >= 256 * 256 * 256 * 128
    assert lib.pnpadmm_dncnn_activation_bytes(0, 256, 256) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import pnp_admm_cnc_mri_b200 as pk
    with pytest.raises(pk.PnpAdmmError):
        pk.admm_solve(np.zeros((2, 64, 64), np.float32), np.ones((64, 64)), np.zeros((64, 64), complex))


def test_product_never_imports_oracle():
    """The product package must not import / reference anything under oracle/."""
    pkg = os.path.join(ROOT, 'pnp_admm_cnc_mri_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f
                assert 'reference_numpy' not in src, f
