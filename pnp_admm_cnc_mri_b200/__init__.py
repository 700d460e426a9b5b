"""pnp_admm_cnc_mri_b200 — B200-native ADMM / PnP-ADMM CS-MRI reconstruction path.

Host side of the C ABI in include/pnpadmm.h (hand-written sm_100a CUDA in csrc/).  Keeps the call
signatures of the reference's entry points (see reference_api) and adds a batched core API.
"""
from ._abi import PnpAdmmError, load as load_library  # noqa: F401
from .solver import AdmmSolver, HostPipeline, admm_solve, mask_is_row_separable, soft, cnc_combine, dual_update_, image_metrics  # noqa: F401

__all__ = ['AdmmSolver', 'HostPipeline', 'admm_solve', 'mask_is_row_separable', 'soft', 'cnc_combine', 'dual_update_', 'image_metrics', 'PnpAdmmError', 'load_library']
