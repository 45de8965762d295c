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
    """Both implementations behind cpgb_bn_relu_fwd / _bwd: the single-launch cluster kernels (default for tensors up
    to 9 MB; forced for every size here) and the stats -> finalize -> apply sequence (CPGB_BN_CLUSTER=0).  The library
    reads the variables per call."""
    monkeypatch.setenv('CPGB_BN_CLUSTER', '1' if request.param == 'cluster' else '0')
    monkeypatch.setenv('CPGB_BN_CLUSTER_MB', '1e9')
    return request.param
