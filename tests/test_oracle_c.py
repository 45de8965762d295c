"""The plain-C restatement (oracle/cpg_oracle.c) against the golden vectors of the live
reference and against the numpy/torch oracle: two independent restatements must agree."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def orc():
    subprocess.run(['make', '-s', '-C', os.path.join(ROOT, 'oracle')], check=True)
    lib = ctypes.CDLL(os.path.join(ROOT, 'oracle', '_build', 'libcpg_oracle.so'))
    lib.orc_pruning_mask.restype = ctypes.c_int
    return lib


def P(a, ty=ctypes.c_float):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def test_c_binarize(orc, golden):
    g = golden('binarizer')
    out = np.empty_like(g['p'])
    orc.orc_binarize(P(g['p']), P(out), ctypes.c_int64(out.size), ctypes.c_float(5e-3))
    assert np.array_equal(out, g['b'], equal_nan=True)


@pytest.mark.parametrize('mode', ['finetune', 'prune'])
def test_c_grad_epilogue(orc, golden, mode):
    g = golden('pruner')
    for n in g['names']:
        k = str(n).replace('.', '_')
        dW, dP = g['G_' + k].copy(), g['GP_' + k].copy()
        orc.orc_grad_epilogue(P(dW), P(dP), P(g['W_' + k]), P(g['T_' + k], ctypes.c_uint8),
                              ctypes.c_int64(dW.size), 2, ctypes.c_float(4e-5), 1 if mode == 'finetune' else 2)
        assert np.array_equal(dW, g[f'a6_{mode}_dW_{k}']) and np.array_equal(dP, g[f'a6_{mode}_dP_{k}'])


def test_c_pruning_mask_and_masks(orc, golden):
    g = golden('pruner')
    for i, ratio in enumerate(g['a7_ratios']):
        for n in g['names']:
            k = str(n).replace('.', '_')
            t = g['T_' + k].copy()
            cut = ctypes.c_float()
            rc = orc.orc_pruning_mask(P(g['W_' + k]), P(t, ctypes.c_uint8), ctypes.c_int64(t.size), 2,
                                      ctypes.c_double(float(ratio)), ctypes.byref(cut), None, None)
            assert rc == int(g[f'a7_{i}_exit_{k}'])
            if rc == 0:
                assert np.array_equal(t, g[f'a7_{i}_T_{k}'])
    for n in g['names']:
        k = str(n).replace('.', '_')
        w = g['W_' + k].copy()
        orc.orc_apply_mask(P(w), P(g['T_' + k], ctypes.c_uint8), ctypes.c_int64(w.size), 2)
        assert np.array_equal(w, g['a9_apply_' + k])
        t = g['T_' + k].copy()
        orc.orc_make_finetuning_mask(P(t, ctypes.c_uint8), ctypes.c_int64(t.size), 3)
        assert np.array_equal(t, g['a10_T_' + k])


@pytest.mark.parametrize('name', ['conv_s2_g2', 'conv_dil2', 'conv_1x1_s2', 'conv_7x7_s2'])
def test_c_conv_forward(orc, golden, name):
    g = golden(name)
    stride, pad, dil, groups = [int(v) for v in g['conv']]
    x, w = g['x'], g['w']
    y = np.empty_like(g['y'])
    N, C, H, W = x.shape
    K, _, R, S = w.shape
    orc.orc_conv2d_fwd(P(x), P(w), P(g['p']) if 'p' in g else None, P(g['b']) if 'b' in g else None, P(y),
                       N, C, H, W, K, R, S, stride, pad, dil, groups, ctypes.c_float(5e-3))
    assert np.abs(y - g['y']).max() <= 1e-5 * np.abs(g['y']).max()
