"""Differential tests of the a6 gradient epilogue, the a1 Binarizer and a9 / a10 on seeded adversarial inputs
(tests/_epilogue_cases.py): the torch oracle and the UNMODIFIED reference code (SparsePruner.
do_weight_decay_and_make_grads_zero on reference layers, utils/prune.py:195-211; Binarizer.apply,
models/layers.py:15-19) must agree bit for bit -- NaN / inf gradients under masked positions come out as exact zeros
because the reference assigns there -- and the plain-C oracle to one rounding of the `dW + wd * W` term (a C compiler may
contract it into an FMA)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import cpg_oracle as O
from tests._epilogue_cases import binarizer_cases, epilogue_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_root():
    for p in ('/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if os.path.isfile(os.path.join(p, 'utils', 'prune.py')):
            return p
    return None


@pytest.fixture(scope='module')
def orc():
    subprocess.run(['make', '-s', '-C', os.path.join(ROOT, 'oracle')], check=True)
    return ctypes.CDLL(os.path.join(ROOT, 'oracle', '_build', 'libcpg_oracle.so'))


def _oracle_epilogue(w, g, gp, t, cur, mode, wd):
    dW = torch.from_numpy(g.copy())
    dP = torch.from_numpy(gp.copy()) if gp is not None else None
    O.weight_decay_and_mask_grads(dW, dP, torch.from_numpy(w), torch.from_numpy(t), cur, wd, mode)
    return dW.numpy(), (dP.numpy() if dP is not None else None)


def test_c_oracle_epilogue_and_binarizer(orc):
    P = lambda a, ty=ctypes.c_float: a.ctypes.data_as(ctypes.POINTER(ty))
    for i, (w, g, gp, t, cur, mode, wd) in enumerate(epilogue_cases()):
        want_w, want_p = _oracle_epilogue(w, g, gp, t, cur, mode, wd)
        dW, dP = g.copy(), (gp.copy() if gp is not None else None)
        orc.orc_grad_epilogue(P(dW), P(dP) if dP is not None else None, P(w), P(t, ctypes.c_uint8), ctypes.c_int64(t.size),
                              cur, ctypes.c_float(wd), 1 if mode == 'finetune' else 2)
        zero = t != cur
        assert np.array_equal(dW[zero], want_w[zero]) and not dW[zero].any(), i          # exact zeros, NaN / inf included
        with np.errstate(invalid='ignore', over='ignore'):
            a, b = dW[~zero].astype(np.float64), want_w[~zero].astype(np.float64)
            # one rounding of the product: |difference| <= ulp of the larger operand (the sum may cancel)
            bound = 2.0 ** -23 * (np.abs(g[~zero].astype(np.float64)) + np.abs(wd * w[~zero].astype(np.float64)))
            finite = np.isfinite(a) & np.isfinite(b)
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~finite & ~np.isnan(a)], b[~finite & ~np.isnan(b)]), i
            assert (np.abs(a[finite] - b[finite]) <= bound[finite]).all(), i
        if gp is not None:
            assert np.array_equal(dP, want_p, equal_nan=True), i
    for p in binarizer_cases():
        out = np.empty_like(p)
        orc.orc_binarize(P(p), P(out), ctypes.c_int64(p.size), ctypes.c_float(5e-3))
        assert np.array_equal(out, O.binarize(torch.from_numpy(p), 5e-3).numpy(), equal_nan=True)


REF_CODE = r'''
import argparse, sys
import numpy as np
import torch
import torch.nn as nn
REF, ROOT, OUT = sys.argv[1], sys.argv[2], sys.argv[3]
torch.Tensor.cuda = lambda self, *a, **k: self
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
import models.layers as nl
from utils.prune import SparsePruner
from tests._epilogue_cases import binarizer_cases, epilogue_cases


class Stub:
    pass


res = {}
for i, (w, g, gp, t, cur, mode, wd) in enumerate(epilogue_cases()):
    if w.ndim == 4:
        layer = nl.SharableConv2d(w.shape[1], w.shape[0], w.shape[2], bias=False)
    else:
        layer = nl.SharableLinear(w.shape[1], w.shape[0], bias=False)
    with torch.no_grad():
        layer.weight.copy_(torch.from_numpy(w))
    layer.weight.grad = torch.from_numpy(g.copy())
    if gp is not None:
        layer.piggymask = nn.Parameter(torch.full(w.shape, 0.01))
        layer.piggymask.grad = torch.from_numpy(gp.copy())
    model = nn.Sequential(layer)
    s = Stub()
    s.model, s.masks, s.current_dataset_idx = model, {'0': torch.from_numpy(t.copy())}, cur
    s.args = argparse.Namespace(weight_decay=wd, mode=mode)
    SparsePruner.do_weight_decay_and_make_grads_zero(s)
    res['w%d' % i] = layer.weight.grad.numpy().copy()
    if gp is not None:
        res['p%d' % i] = layer.piggymask.grad.numpy().copy()
for j, p in enumerate(binarizer_cases()):
    res['b%d' % j] = nl.Binarizer.apply(torch.from_numpy(p.copy()), 5e-3).numpy().copy()
# a9 / a10 (utils/prune.py:213-243) on the same weights and masks (NaN / inf weights must become exact zeros too)
for i, (w, g, gp, t, cur, mode, wd) in enumerate(epilogue_cases()):
    mk = (lambda: nl.SharableConv2d(w.shape[1], w.shape[0], w.shape[2], bias=False)) if w.ndim == 4 else \
         (lambda: nl.SharableLinear(w.shape[1], w.shape[0], bias=False))
    for tag, call in (('a9z', 'make_pruned_zero'), ('a9a', 'apply_mask'), ('a10', 'make_finetuning_mask')):
        layer = mk()
        with torch.no_grad():
            layer.weight.copy_(torch.from_numpy(g))          # the special-valued array as weights
        s = Stub()
        s.model, s.masks = nn.Sequential(layer), {'0': torch.from_numpy(t.copy())}
        s.current_dataset_idx, s.inference_dataset_idx = cur, cur
        getattr(SparsePruner, call)(s)
        res['%s_w%d' % (tag, i)] = layer.weight.detach().numpy().copy()
        res['%s_t%d' % (tag, i)] = s.masks['0'].numpy().copy()
        res['%s_c%d' % (tag, i)] = np.array(s.current_dataset_idx)
np.savez(OUT, **res)
print('ok')
'''


@pytest.mark.skipif(_ref_root() is None, reason='no reference checkout')
def test_oracle_equals_the_live_reference(tmp_path):
    out = os.path.join(str(tmp_path), 'ref.npz')
    r = subprocess.run([sys.executable, '-c', REF_CODE, _ref_root(), ROOT, out], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stderr[-3000:]
    ref = dict(np.load(out))
    bits = lambda a: a.view(np.uint32)
    for i, (w, g, gp, t, cur, mode, wd) in enumerate(epilogue_cases()):
        got_w, got_p = _oracle_epilogue(w, g, gp, t, cur, mode, wd)
        assert np.array_equal(bits(got_w), bits(ref['w%d' % i])), ('dW', i)
        if gp is not None:
            assert np.array_equal(bits(got_p), bits(ref['p%d' % i])), ('dP', i)
    for j, p in enumerate(binarizer_cases()):
        got = O.binarize(torch.from_numpy(p), 5e-3).numpy()
        assert np.array_equal(bits(got), bits(ref['b%d' % j])), ('binarizer', j)
    for i, (w, g, gp, t, cur, mode, wd) in enumerate(epilogue_cases()):
        for tag in ('a9z', 'a9a', 'a10'):
            ww, tt, c = torch.from_numpy(g.copy()), torch.from_numpy(t.copy()), cur
            if tag == 'a9z':
                O.make_pruned_zero(ww, tt)
            elif tag == 'a9a':
                O.apply_mask(ww, tt, cur)
            else:
                c = O.make_finetuning_mask(tt, cur)
            assert np.array_equal(bits(ww.numpy()), bits(ref['%s_w%d' % (tag, i)])), (tag, i)
            assert np.array_equal(tt.numpy(), ref['%s_t%d' % (tag, i)]) and c == int(ref['%s_c%d' % (tag, i)]), (tag, i)
