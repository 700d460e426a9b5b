"""In-tree build of the C-ABI library (libpnpadmm.so) for sm_100a.

``python -m pnp_admm_cnc_mri_b200.build`` or ``build_library()``.  The built ``.so`` is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libpnpadmm.so')
SOURCES = ['pnpadmm.cu', 'common.cuh', 'streaming.cuh', 'stream2.cuh', 'stream2_core.cuh', 'cluster256.cuh', 'cluster256_core.cuh',
           'metrics.cuh', 'dncnn_tc.cuh', 'rowsep256.cuh', 'rowsepN.cuh', 'rowsep_core.cuh']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(os.path.dirname(PKG), 'include', 'pnpadmm.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    # PNPADMM_NVCC_EXTRA: extra flags, e.g. "-DPNPADMM_TC_EXPERIMENTS -DPNPADMM_K1_EXPERIMENTS" to compile the timing switches in
    extra = os.environ.get('PNPADMM_NVCC_EXTRA', '').split()
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB, os.path.join(CSRC, 'pnpadmm.cu')]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
