"""Run the UNMODIFIED reference scripts as the ground-truth oracle (build container only).

TEST INFRASTRUCTURE.  Needs ``/root/reference`` (absent on the GPU box), so it
is only used by ``oracle/make_golden.py`` to pin ``oracle/reference_numpy.py``
and to generate the fixtures committed under ``tests/golden/``.

Recipe: scratch cwd with ``CS_MRI`` and
``testsets/Set1`` symlinked from the read-only reference tree, a stub
``matplotlib`` (``utils/utils_image.py:10`` imports it, never calls it), then
``runpy.run_path`` executes the script's module-level driver (S1:171-194).
"""
from __future__ import annotations

import glob
import os
import runpy
import shutil
import sys
import tempfile
import types

REF = os.environ.get('PNPADMM_REFERENCE', '/root/reference')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF, 'CS_MRI'))


def _script(tag: str) -> str:
    hits = [p for p in glob.glob(os.path.join(REF, '*.py')) if tag in os.path.basename(p)]
    if len(hits) != 1:
        raise FileNotFoundError(f'reference script {tag!r} not found under {REF}')
    return hits[0]


def run_script(tag: str, argv=(), images=None, mask_file=None, model_zoo=None, post=None):
    """Execute reference script ``tag`` ('【1】', '【3】', '【4】', '【6】').

    images    : list of PNG paths to expose as testsets/Set1 (default: reference set1 = 05.png)
    mask_file : basename of the CS_MRI mask the script's k=0 slot should see
                (default Q_Random30.mat).  The scripts hard-code k=0 (S1:191).
    model_zoo : optional dict name -> state_dict saved as model_zoo/<name>.pth (PnP scripts)
    post      : optional callable(g) run after the script, still inside the scratch cwd, e.g. to call the
                script's own (unmodified) functions for the presets its driver does not reach (S3:375-376,
                S6:610-611 hard-code the model index); its return value is stored as g['_post'].
    Returns the script's globals dict (``g['out']`` is the 22-slot list).
    """
    if not reference_available():
        raise RuntimeError(f'{REF} not present')
    scratch = tempfile.mkdtemp(prefix='pnpadmm_ref_')
    cwd0, argv0, path0 = os.getcwd(), list(sys.argv), list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ('matplotlib', 'matplotlib.pyplot')}
    try:
        os.makedirs(os.path.join(scratch, 'CS_MRI'))
        for f in os.listdir(os.path.join(REF, 'CS_MRI')):
            os.symlink(os.path.join(REF, 'CS_MRI', f), os.path.join(scratch, 'CS_MRI', f))
        if mask_file is not None and mask_file != 'Q_Random30.mat':
            tgt = os.path.join(scratch, 'CS_MRI', 'Q_Random30.mat')
            os.remove(tgt)
            os.symlink(os.path.join(REF, 'CS_MRI', mask_file), tgt)
        os.makedirs(os.path.join(scratch, 'testsets', 'Set1'))
        if images is None:
            images = sorted(glob.glob(os.path.join(REF, 'testsets', 'set1', '*.png')))
        for p in images:
            os.symlink(p, os.path.join(scratch, 'testsets', 'Set1', os.path.basename(p)))
        if model_zoo:
            import torch
            os.makedirs(os.path.join(scratch, 'model_zoo'))
            for name, sd in model_zoo.items():
                torch.save(sd, os.path.join(scratch, 'model_zoo', name + '.pth'))
        mpl = types.ModuleType('matplotlib')
        plt = types.ModuleType('matplotlib.pyplot')
        mpl.pyplot = plt
        sys.modules['matplotlib'] = mpl
        sys.modules['matplotlib.pyplot'] = plt
        sys.path.insert(0, REF)
        script = _script(tag)
        sys.argv = [script] + [str(a) for a in argv]
        os.chdir(scratch)
        g = runpy.run_path(script, run_name='__main__')
        if post is not None:
            g['_post'] = post(g)
        return g
    finally:
        os.chdir(cwd0)
        sys.argv = argv0
        sys.path[:] = path0
        for k, v in saved_mods.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        # the reference's utils/models packages must not linger in sys.modules
        for k in [k for k in sys.modules if k == 'utils' or k.startswith('utils.')
                  or k == 'models' or k.startswith('models.')]:
            sys.modules.pop(k, None)
        shutil.rmtree(scratch, ignore_errors=True)
