"""ctypes binding of include/pnpadmm.h.  There is NO fallback: if the CUDA library is missing or a
call fails, this raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_size_t, c_void_p

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, 'libpnpadmm.so')

OK, ERR_BAD_ARG, ERR_BAD_SIZE, ERR_WORKSPACE, ERR_CUDA, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
PROX_L1, PROX_CNC = 0, 1
KERNEL_AUTO, KERNEL_CLUSTER, KERNEL_STREAMING, KERNEL_ROWSEP = 0, 1, 2, 3
OUT_F32, OUT_U8 = 0, 1
ABI_VERSION = 3

# every symbol include/pnpadmm.h declares (tests check the .so exports all of them)
SYMBOLS = [
    'pnpadmm_abi_version', 'pnpadmm_last_error_string', 'pnpadmm_device_info', 'pnpadmm_plan_info', 'pnpadmm_workspace_bytes',
    'pnpadmm_acquire_f32', 'pnpadmm_acquire_f64', 'pnpadmm_zero_filled_f32', 'pnpadmm_zero_filled_f64',
    'pnpadmm_prepare_f32', 'pnpadmm_prepare_f64', 'pnpadmm_xupdate_f32', 'pnpadmm_xupdate_f64',
    'pnpadmm_iterate_f32', 'pnpadmm_iterate_f64', 'pnpadmm_solve_f32', 'pnpadmm_solve_f64',
    'pnpadmm_reconstruct_f32', 'pnpadmm_reconstruct_f64',
    'pnpadmm_host_scratch_bytes', 'pnpadmm_reconstruct_host_f32',
    'pnpadmm_pipeline_create', 'pnpadmm_pipeline_destroy', 'pnpadmm_pipeline_set_output',
    'pnpadmm_host_pipeline_scratch_bytes', 'pnpadmm_reconstruct_host_pipelined_f32', 'pnpadmm_reconstruct_host_wait',
    'pnpadmm_soft_f32', 'pnpadmm_soft_f64', 'pnpadmm_cnc_combine_f32', 'pnpadmm_cnc_combine_f64',
    'pnpadmm_dual_update_f32', 'pnpadmm_dual_update_f64', 'pnpadmm_measure_fp32_peak',
    'pnpadmm_metrics_scratch_bytes', 'pnpadmm_metrics_f32', 'pnpadmm_metrics_f64',
    'pnpadmm_dncnn_activation_bytes', 'pnpadmm_conv64_bf16', 'pnpadmm_dncnn_forward_bf16', 'pnpadmm_ffdnet_forward_bf16',
    'pnpadmm_conv64_dilated_bf16', 'pnpadmm_dncnn_forward_dilated_bf16',
]

_lib = None


class PnpAdmmError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libpnpadmm.so (built in-tree by ``pnp_admm_cnc_mri_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PnpAdmmError(
            f'{LIB_PATH} is missing: the CUDA extension has not been built. Run '
            f'`python -m pnp_admm_cnc_mri_b200.build` (needs nvcc). There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    p, i, d, z = c_void_p, c_int, c_double, c_size_t
    lib.pnpadmm_abi_version.restype = i
    lib.pnpadmm_abi_version.argtypes = []
    lib.pnpadmm_last_error_string.restype = c_char_p
    lib.pnpadmm_last_error_string.argtypes = []
    lib.pnpadmm_device_info.restype = i
    lib.pnpadmm_device_info.argtypes = [POINTER(c_int)] * 4
    lib.pnpadmm_plan_info.restype = i
    lib.pnpadmm_plan_info.argtypes = [i, i, i, i, i] + [POINTER(c_int)] * 6
    lib.pnpadmm_workspace_bytes.restype = z
    lib.pnpadmm_workspace_bytes.argtypes = [i, i, i, i]
    lib.pnpadmm_host_scratch_bytes.restype = z
    lib.pnpadmm_host_scratch_bytes.argtypes = [i, i]
    for sfx in ('f32', 'f64'):
        f = getattr(lib, 'pnpadmm_acquire_' + sfx); f.restype = i
        f.argtypes = [p, p, p, p, i, i, i, i, p, z, p] if sfx == 'f32' else [p, p, p, p, i, i, i, i, i, p, z, p]
        f = getattr(lib, 'pnpadmm_zero_filled_' + sfx); f.restype = i
        f.argtypes = [p, p, i, i, p, z, p]
        f = getattr(lib, 'pnpadmm_prepare_' + sfx); f.restype = i
        f.argtypes = [p, p, i, i, i, d, p, z, p]
        f = getattr(lib, 'pnpadmm_xupdate_' + sfx); f.restype = i
        f.argtypes = [p, p, p, p, i, i, i, i, p, z, p]
        f = getattr(lib, 'pnpadmm_iterate_' + sfx); f.restype = i
        f.argtypes = [p, p, p, i, i, i, i, i, d, d, d, d, i, p, z, p]
        f = getattr(lib, 'pnpadmm_solve_' + sfx); f.restype = i
        f.argtypes = [p, p, p, p, p, i, i, i, i, i, d, d, d, d, i, p, z, p]
        f = getattr(lib, 'pnpadmm_reconstruct_' + sfx); f.restype = i
        f.argtypes = [p, p, p, p, p, p, p, i, i, i, i, i, i, d, d, d, d, i, p, z, p]
        f = getattr(lib, 'pnpadmm_soft_' + sfx); f.restype = i
        f.argtypes = [p, p, d, z, p]
        f = getattr(lib, 'pnpadmm_cnc_combine_' + sfx); f.restype = i
        f.argtypes = [p, p, p, p, p, d, d, z, p]
        f = getattr(lib, 'pnpadmm_dual_update_' + sfx); f.restype = i
        f.argtypes = [p, p, p, i, z, p]
    lib.pnpadmm_reconstruct_host_f32.restype = i
    lib.pnpadmm_reconstruct_host_f32.argtypes = [p, p, p, p, i, i, i, i, d, d, d, d, i, p, z, p, z, p]
    lib.pnpadmm_metrics_scratch_bytes.restype = z
    lib.pnpadmm_metrics_scratch_bytes.argtypes = [i]
    for sfx in ('f32', 'f64'):
        f = getattr(lib, 'pnpadmm_metrics_' + sfx); f.restype = i
        f.argtypes = [p, p, i, i, i, p, p, z, p]
    lib.pnpadmm_host_pipeline_scratch_bytes.restype = z
    lib.pnpadmm_host_pipeline_scratch_bytes.argtypes = [i, i, i]
    lib.pnpadmm_pipeline_create.restype = i
    lib.pnpadmm_pipeline_create.argtypes = [POINTER(c_void_p), i]
    lib.pnpadmm_pipeline_set_output.restype = i
    lib.pnpadmm_pipeline_set_output.argtypes = [p, i]
    lib.pnpadmm_pipeline_destroy.restype = i
    lib.pnpadmm_pipeline_destroy.argtypes = [p]
    lib.pnpadmm_reconstruct_host_pipelined_f32.restype = i
    lib.pnpadmm_reconstruct_host_pipelined_f32.argtypes = [p, p, p, p, p, i, i, i, i, d, d, d, d, i, p, z, p, z, i, p, p, p]
    lib.pnpadmm_reconstruct_host_wait.restype = i
    lib.pnpadmm_reconstruct_host_wait.argtypes = [p, i]
    lib.pnpadmm_measure_fp32_peak.restype = i
    lib.pnpadmm_measure_fp32_peak.argtypes = [POINTER(c_double), p]
    lib.pnpadmm_dncnn_activation_bytes.restype = z
    lib.pnpadmm_dncnn_activation_bytes.argtypes = [i, i, i]
    lib.pnpadmm_conv64_bf16.restype = i
    lib.pnpadmm_conv64_bf16.argtypes = [p, p, p, p, i, i, i, i, p]
    lib.pnpadmm_dncnn_forward_bf16.restype = i
    lib.pnpadmm_dncnn_forward_bf16.argtypes = [p, p, i, i, i, i, i, p, p, p, p, p, p, i, p, p, p]
    lib.pnpadmm_conv64_dilated_bf16.restype = i
    lib.pnpadmm_conv64_dilated_bf16.argtypes = [p, p, p, p, i, i, i, i, i, p]
    lib.pnpadmm_dncnn_forward_dilated_bf16.restype = i
    lib.pnpadmm_dncnn_forward_dilated_bf16.argtypes = [p, p, i, i, i, i, i, POINTER(c_int), p, p, p, p, p, p, i, p, p, p]
    lib.pnpadmm_ffdnet_forward_bf16.restype = i
    lib.pnpadmm_ffdnet_forward_bf16.argtypes = [p, p, i, i, i, c_float, i, p, p, p, p, p, p, p, p, p]
    if lib.pnpadmm_abi_version() != ABI_VERSION:
        raise PnpAdmmError(f'ABI version mismatch: library {lib.pnpadmm_abi_version()} != binding {ABI_VERSION}; rebuild')
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Map a status code to the Python exception the wrapper contract promises
    (ValueError for bad arguments / sizes, PnpAdmmError otherwise)."""
    if rc == OK:
        return
    msg = load().pnpadmm_last_error_string().decode('utf-8', 'replace')
    if rc in (ERR_BAD_ARG, ERR_BAD_SIZE):
        raise ValueError(f'pnpadmm: {msg} (code {rc})')
    raise PnpAdmmError(f'pnpadmm: {msg} (code {rc})')
