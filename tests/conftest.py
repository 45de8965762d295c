import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))
    return load


@pytest.fixture(params=['cluster', 'three_kernels'])
def bn_path(request, monkeypatch):
    """Both implementations behind cpgb_bn_relu_fwd / _bwd: the stats -> finalize -> apply sequence (default) and the
    single-launch cluster kernels (CPGB_BN_CLUSTER=1; the library reads the variable per call)."""
    monkeypatch.setenv('CPGB_BN_CLUSTER', '1' if request.param == 'cluster' else '0')
    return request.param
