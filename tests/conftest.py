import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def cs_inputs():
    """Reference DATA converted by oracle/make_golden.py: masks (3,256,256) u8,
    noises (256,256) c128 already multiplied by 3.0 (S1:186), images (15,256,256) u8."""
    d = np.load(os.path.join(GOLDEN, 'cs_mri_inputs.npz'))
    return dict(masks=d['masks'], mask_names=[str(s) for s in d['mask_names']],
                noises=d['noises'] * 3.0, images=d['images'],
                image_names=[str(s) for s in d['image_names']])


@pytest.fixture(scope='session')
def ref_out():
    d = np.load(os.path.join(GOLDEN, 'ref_out_05_random.npz'))
    return dict(l1=d['l1'], cnc=d['cnc'])
