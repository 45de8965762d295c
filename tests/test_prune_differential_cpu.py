"""Three-way differential test of the a7 prune step on seeded adversarial inputs (tests/_prune_cases.py): the numpy
oracle, the plain-C oracle and -- when a reference checkout is present -- the UNMODIFIED reference method
SparsePruner._pruning_mask (utils/prune.py:30-53, Tensor.cuda made the identity) must produce the same task mask bit
for bit, or all take the exit-2 path."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import cpg_oracle as O
from tests._prune_cases import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_root():
    for p in ('/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if os.path.isfile(os.path.join(p, 'utils', 'prune.py')):
            return p
    return None


def _numpy_oracle(w, t, cur, ratio):
    tt = torch.from_numpy(t.copy())
    try:
        O.pruning_mask(torch.from_numpy(w.copy()), tt, cur, ratio)
    except O.NotEnoughWeights:
        return None
    return tt.numpy()


@pytest.fixture(scope='module')
def orc():
    subprocess.run(['make', '-s', '-C', os.path.join(ROOT, 'oracle')], check=True)
    lib = ctypes.CDLL(os.path.join(ROOT, 'oracle', '_build', 'libcpg_oracle.so'))
    lib.orc_pruning_mask.restype = ctypes.c_int
    return lib


def _c_oracle(orc, w, t, cur, ratio):
    tt = t.copy()
    rc = orc.orc_pruning_mask(w.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                              tt.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), ctypes.c_int64(tt.size), cur,
                              ctypes.c_double(ratio), None, None, None)
    return None if rc == 2 else tt


def test_numpy_and_c_oracles_agree(orc):
    n_exit = 0
    for i, (w, t, cur, ratio) in enumerate(cases()):
        a, b = _numpy_oracle(w, t, cur, ratio), _c_oracle(orc, w, t, cur, ratio)
        assert (a is None) == (b is None), (i, ratio)
        if a is None:
            n_exit += 1
        else:
            assert np.array_equal(a, b), i
    assert 10 < n_exit < 200                    # both outcomes are exercised


REF_CODE = r'''
import sys
import numpy as np
import torch
REF, ROOT, OUT = sys.argv[1], sys.argv[2], sys.argv[3]
torch.Tensor.cuda = lambda self, *a, **k: self
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
from utils.prune import SparsePruner
from tests._prune_cases import cases


class Stub:                                    # _pruning_mask only reads current_dataset_idx
    pass


res = {}
for i, (w, t, cur, ratio) in enumerate(cases()):
    s = Stub()
    s.current_dataset_idx = cur
    tt = torch.from_numpy(t.copy())
    try:
        SparsePruner._pruning_mask(s, torch.from_numpy(w.copy()), tt, 'layer', ratio)
        res['t%d' % i] = tt.numpy()
    except SystemExit as e:
        assert e.code == 2
        res['t%d' % i] = np.zeros(0, dtype=np.uint8) - 0
        res['x%d' % i] = np.array(2)
np.savez(OUT, **res)
print('ok')
'''


@pytest.mark.skipif(_ref_root() is None, reason='no reference checkout')
def test_oracles_agree_with_the_live_reference(orc, tmp_path):
    out = os.path.join(str(tmp_path), 'ref.npz')
    r = subprocess.run([sys.executable, '-c', REF_CODE, _ref_root(), ROOT, out], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stderr[-3000:]
    ref = dict(np.load(out))
    for i, (w, t, cur, ratio) in enumerate(cases()):
        want = None if ('x%d' % i) in ref else ref['t%d' % i]
        for name, got in (('numpy', _numpy_oracle(w, t, cur, ratio)), ('c', _c_oracle(orc, w, t, cur, ratio))):
            assert (got is None) == (want is None), (name, i, ratio, int(((t == cur) | (t == 0)).sum()))
            if want is not None:
                assert np.array_equal(got, want), (name, i)
