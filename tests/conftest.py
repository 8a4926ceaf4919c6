import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
GOLDEN_CASES = ['gc_small', 'gc_multivar_offtime', 'noloc_small', 'gc_4deg', 'gc_inflate', 'gc_dense300']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA (B200) device; run with -m gpu on the GPU box')


def load_golden(name):
    import json
    import numpy as np
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    p = json.loads(str(g['params']))
    return g, p


@pytest.fixture(scope='session')
def lib():
    from efa_xray_b200 import _lib
    return _lib.load()
