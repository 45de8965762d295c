"""a3 / a4 / a5 on random layer configurations: the oracle's forward and autograd (oracle.cpg_oracle.conv2d_* /
linear_*) against the UNMODIFIED reference layers (models/layers.py:43-218) run in a second process -- kernel sizes 1 / 3 /
5 / 7, strides, paddings, dilations, groups, with and without bias and piggymask, piggymask values on both sides of
the threshold: outputs and all gradients to 1e-5 of the largest element (the same torch CPU kernels underneath; their
summation order depends on the process's thread state).  Also the constructors'
error behaviour of the drop-in classes (ValueError on non-divisible groups, models/layers.py:65-68)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import cpg_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_root():
    for p in ('/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if os.path.isfile(os.path.join(p, 'models', 'layers.py')):
            return p
    return None


def layer_cases(n=36):
    rng = np.random.RandomState(2718)
    out = []
    for i in range(n):
        if i % 4 == 3:                                      # linear
            I, Oo, B = int(rng.randint(1, 40)), int(rng.randint(1, 40)), int(rng.randint(1, 6))
            cfg = dict(kind='linear', I=I, O=Oo, bias=bool(i % 8 != 7))
            x = rng.standard_normal((B, I)).astype(np.float32)
            w = rng.standard_normal((Oo, I)).astype(np.float32)
            dy = rng.standard_normal((B, Oo)).astype(np.float32)
        else:
            groups = int(rng.choice([1, 1, 2, 3]))
            C, K = groups * int(rng.randint(1, 5)), groups * int(rng.randint(1, 5))
            R = int(rng.choice([1, 3, 5, 7]))
            stride, dil = int(rng.choice([1, 1, 2, 3])), int(rng.choice([1, 1, 2]))
            pad = int(rng.randint(0, R))
            H = W = dil * (R - 1) + 1 + int(rng.randint(0, 9))
            N = int(rng.randint(1, 4))
            cfg = dict(kind='conv', C=C, K=K, R=R, stride=stride, pad=pad, dil=dil, groups=groups, bias=bool(i % 3 == 0))
            x = rng.standard_normal((N, C, H, W)).astype(np.float32)
            w = rng.standard_normal((K, C // groups, R, R)).astype(np.float32)
            P = (H + 2 * pad - dil * (R - 1) - 1) // stride + 1
            dy = rng.standard_normal((N, K, P, P)).astype(np.float32)
        b = rng.standard_normal(w.shape[0]).astype(np.float32) if cfg['bias'] else None
        p = None
        if i % 5 != 0:
            p = rng.uniform(0, 0.01, w.shape).astype(np.float32)
            p.reshape(-1)[::7] = np.float32(5e-3)            # exactly the threshold: binarises to 0
        out.append((cfg, x, w, b, p, dy))
    return out


REF_CODE = r'''
import sys
import numpy as np
import torch
import torch.nn as nn
REF, ROOT, OUT = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
import models.layers as nl
from tests.test_layers_differential_cpu import layer_cases
res = {}
T = lambda a: torch.from_numpy(a.copy())
for i, (cfg, x, w, b, p, dy) in enumerate(layer_cases()):
    if cfg['kind'] == 'conv':
        m = nl.SharableConv2d(cfg['C'], cfg['K'], cfg['R'], stride=cfg['stride'], padding=cfg['pad'], dilation=cfg['dil'],
                              groups=cfg['groups'], bias=cfg['bias'])
    else:
        m = nl.SharableLinear(cfg['I'], cfg['O'], bias=cfg['bias'])
    with torch.no_grad():
        m.weight.copy_(T(w))
        if b is not None:
            m.bias.copy_(T(b))
    if p is not None:
        m.piggymask = nn.Parameter(T(p))
    xs = T(x).requires_grad_(True)
    y = m(xs)
    y.backward(T(dy))
    res['y%d' % i], res['dx%d' % i], res['dW%d' % i] = y.detach().numpy(), xs.grad.numpy(), m.weight.grad.numpy()
    if b is not None:
        res['db%d' % i] = m.bias.grad.numpy()
    if p is not None:
        res['dP%d' % i] = m.piggymask.grad.numpy()
np.savez(OUT, **res)
print('ok')
'''


@pytest.mark.skipif(_ref_root() is None, reason='no reference checkout')
def test_oracle_layers_equal_the_reference(tmp_path):
    out = os.path.join(str(tmp_path), 'ref.npz')
    r = subprocess.run([sys.executable, '-c', REF_CODE, _ref_root(), ROOT, out], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stderr[-3000:]
    ref = dict(np.load(out))
    T = lambda a: torch.from_numpy(a.copy()) if a is not None else None
    for i, (cfg, x, w, b, p, dy) in enumerate(layer_cases()):
        if cfg['kind'] == 'conv':
            geo = (cfg['stride'], cfg['pad'], cfg['dil'], cfg['groups'])
            y = O.conv2d_forward(T(x), T(w), T(p), T(b), *geo)
            dx, dW, dP, db, _ = O.conv2d_backward(T(x), T(w), T(p), T(b), T(dy), *geo)
        else:
            y = O.linear_forward(T(x), T(w), T(p), T(b))
            dx, dW, dP, db, _ = O.linear_backward(T(x), T(w), T(p), T(b), T(dy))
        for name, got in (('y', y), ('dx', dx), ('dW', dW), ('db', db), ('dP', dP)):
            key = '%s%d' % (name, i)
            assert (got is None) == (key not in ref), (key, cfg)
            if got is not None:
                a, b = got.detach().numpy().astype(np.float64), ref[key].astype(np.float64)
                # the same torch CPU kernels underneath, but their reduction order depends on the thread state of the
                # process (oneDNN / OpenMP): equal to fp32 summation order, not bit for bit
                assert a.shape == b.shape and np.abs(a - b).max() <= 1e-5 * max(np.abs(b).max(), 1e-30), (key, cfg)


def test_constructor_errors_of_the_drop_in_layers():
    import cpg_b200.layers as nl
    for args in ((6, 8, 3, 1, 0, 1, 4), (8, 6, 3, 1, 0, 1, 4)):       # in / out channels not divisible by groups
        with pytest.raises(ValueError):
            nl.SharableConv2d(*args)
    m = nl.SharableConv2d(8, 12, (3, 5), stride=2, padding=(1, 2), dilation=1, groups=4, bias=False)
    assert tuple(m.weight.shape) == (12, 2, 3, 5) and m.bias is None and m.piggymask is None
    assert m.info == {'threshold_fn': 'binarizer', 'threshold': nl.DEFAULT_THRESHOLD}
