"""The CPU oracle (oracle/cpg_oracle.py) against the golden vectors produced by the live,
unmodified reference (tests/golden/make_golden.py).  Bit-exact everywhere: same torch CPU
kernels underneath, single-threaded generation, integer masks."""
import zlib

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import cpg_oracle as O
from cpg_b200.vgg_cifar import VGGCifar, fill_params_deterministic

torch.set_num_threads(1)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a))

CONV_CASES = ['conv_cfg1', 'conv_cfg1_nopiggy', 'conv_c32', 'conv_s2_g2', 'conv_1x1_s2', 'conv_dil2',
              'conv_7x7_s2']


@pytest.mark.parametrize('name', CONV_CASES)
def test_conv_fwd_bwd(golden, name):
    g = golden(name)
    stride, pad, dil, groups = [int(v) for v in g['conv']]
    x, w, dy = T(g['x']), T(g['w']), T(g['dy'])
    p = T(g['p']) if 'p' in g else None
    b = T(g['b']) if 'b' in g else None
    y = O.conv2d_forward(x, w, p, b, stride, pad, dil, groups)
    assert torch.equal(y, T(g['y']))
    dx, dW, dP, db, _ = O.conv2d_backward(x, w, p, b, dy, stride, pad, dil, groups)
    # autograd's conv backward and torch.nn.grad use the same kernels but may differ in
    # summation order for some shapes; demand 1e-6 relative and report exactness.
    for got, key in ((dx, 'dx'), (dW, 'dW'), (dP, 'dP'), (db, 'db')):
        if key in g:
            ref = T(g[key])
            assert got is not None
            err = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
            assert err == 0.0, (key, err)
    if 'bin' in g:
        assert torch.equal(O.binarize(p), T(g['bin']))


@pytest.mark.parametrize('name', ['linear_small', 'linear_nopiggy'])
def test_linear_fwd_bwd(golden, name):
    g = golden(name)
    x, w, b, dy = T(g['x']), T(g['w']), T(g['b']), T(g['dy'])
    p = T(g['p']) if 'p' in g else None
    assert torch.equal(O.linear_forward(x, w, p, b), T(g['y']))
    dx, dW, dP, db, _ = O.linear_backward(x, w, p, b, dy)
    for got, key in ((dx, 'dx'), (dW, 'dW'), (dP, 'dP'), (db, 'db')):
        if key in g:
            ref = T(g[key])
            err = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
            assert err == 0.0, (key, err)


def test_binarizer(golden):
    g = golden('binarizer')
    b = O.binarize(T(g['p'])).numpy()
    assert np.array_equal(b, g['b'], equal_nan=True)
    assert np.isnan(b[np.isnan(g['p'])]).all()
    assert np.array_equal(O.binarize_backward(T(g['g'])).numpy(), g['dp'])


def _names(g):
    return [str(n) for n in g['names']]


@pytest.mark.parametrize('mode', ['finetune', 'prune'])
def test_a6_grad_mask(golden, mode):
    g = golden('pruner')
    for n in _names(g):
        k = n.replace('.', '_')
        dW, dP = T(g['G_' + k].copy()), T(g['GP_' + k].copy())
        O.weight_decay_and_mask_grads(dW, dP, T(g['W_' + k]), T(g['T_' + k]), 2, 4e-5, mode)
        assert np.array_equal(dW.numpy(), g[f'a6_{mode}_dW_{k}'])
        assert np.array_equal(dP.numpy(), g[f'a6_{mode}_dP_{k}'])


def test_a7_pruning_mask(golden):
    g = golden('pruner')
    for i, ratio in enumerate(g['a7_ratios']):
        for n in _names(g):
            k = n.replace('.', '_')
            t = T(g['T_' + k].copy())
            if int(g[f'a7_{i}_exit_{k}']) == 2:
                with pytest.raises(O.NotEnoughWeights):
                    O.pruning_mask(T(g['W_' + k]), t, 2, float(ratio))
            else:
                O.pruning_mask(T(g['W_' + k]), t, 2, float(ratio))
                assert np.array_equal(t.numpy(), g[f'a7_{i}_T_{k}']), (ratio, n)


def test_a8_schedule(golden):
    g = golden('pruner')
    names = _names(g)
    masks = {n: T(g['T_' + n.replace('.', '_')].copy()) for n in names}
    last, ratios, zeros = 0, [], []
    for step in range(12):
        if O.time_to_update_masks(step, 0, 8, last, 2):
            last = step
            r = O.adjust_sparsity(step, 0, 8, 0.0, 0.5)
            for n in names:
                O.pruning_mask(T(g['W_' + n.replace('.', '_')]), masks[n], 2, r)
        else:
            r = O.adjust_sparsity(last, 0, 8, 0.0, 0.5)
        ratios.append(r)
        zeros.append([int(masks[n].eq(0).sum()) for n in names])
    assert np.array_equal(np.array(ratios), g['a8_ratios'])
    assert np.array_equal(np.array(zeros), g['a8_zero_counts'])
    for n in names:
        assert np.array_equal(masks[n].numpy(), g['a8_T_' + n.replace('.', '_')])


def test_a9_a10(golden):
    g = golden('pruner')
    for n in _names(g):
        k = n.replace('.', '_')
        w = T(g['W_' + k].copy())
        assert np.array_equal(O.apply_mask(w, T(g['T_' + k]), 2).numpy(), g['a9_apply_' + k])
        w = T(g['W_' + k].copy())
        assert np.array_equal(O.make_pruned_zero(w, T(g['T_' + k])).numpy(), g['a9_zero_' + k])
        t = T(g['T_' + k].copy())
        assert O.make_finetuning_mask(t, 2) == int(g['a10_cur'])
        assert np.array_equal(t.numpy(), g['a10_T_' + k])


@pytest.mark.parametrize('mode', ['prune', 'finetune'])
def test_trajectory_oracle(golden, mode):
    """Oracle modules + OraclePruner reproduce the reference Manager.train trajectory."""
    from tests.trajectory import run_trajectory
    g = golden('traj_' + mode)
    model, masks = run_trajectory(O.OracleSharableConv2d, O.OracleSharableLinear, mode, device='cpu',
                                  pruner_factory='oracle')
    for n, p in model.named_parameters():
        a = p.detach().numpy().astype(np.float64)
        ref = g['sum_module.' + n]
        assert abs(a.sum() - ref[0]) <= 1e-6 * max(1.0, abs(ref[1])), n
        assert abs(np.abs(a).sum() - ref[1]) <= 1e-6 * max(1.0, abs(ref[1])), n
    for n in masks:
        assert int((masks[n].numpy() == 0).sum()) == int(g['maskzeros_module.' + n])
        assert zlib.crc32(masks[n].numpy().tobytes()) == int(g['maskcrc_module.' + n])


def test_a11_one_shot_prune(golden):
    """oracle.one_shot_prune against the live reference's SparsePruner.one_shot_prune (utils/prune.py:94-109)."""
    g = golden('one_shot')
    names = [str(n) for n in g['names']]
    ws = [T(g['W_' + n.replace('.', '_')].copy()) for n in names]
    ts = [T(g['T_' + n.replace('.', '_')].copy()) for n in names]
    O.one_shot_prune(ws, ts, int(g['cur']), float(g['ratio']))
    for n, w, t in zip(names, ws, ts):
        k = n.replace('.', '_')
        assert np.array_equal(t.numpy(), g['T1_' + k]), n
        assert np.array_equal(w.numpy(), g['W1_' + k]), n
        assert (w.numpy()[g['T1_' + k] == 0] == 0).all()
