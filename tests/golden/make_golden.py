"""Generate tests/golden/*.npz by running the UNMODIFIED reference (ivclab/CPG at
/root/reference) on seeded inputs.  Runs only in the build container (the GPU box
has no /root/reference); the resulting fixtures are committed.

    python tests/golden/make_golden.py

Only shim: ``torch.Tensor.cuda`` is made a no-op so the three ``.cuda()`` calls in
utils/prune.py:39,188,228 work on a CUDA-less host.  No reference file
is edited or copied.
"""
import argparse
import hashlib
import os
import sys
import zlib

import numpy as np
import torch
import torch.nn as nn

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, '..', '..'))

torch.Tensor.cuda = lambda self, *a, **k: self  # CPU shim (see docstring)
torch.set_num_threads(1)  # deterministic reduction order for the fixtures

import models.layers as nl  # noqa: E402  (reference)
import utils.prune as ref_prune  # noqa: E402
from utils.manager import Manager  # noqa: E402
from utils import Optimizers  # noqa: E402
from cpg_b200.vgg_cifar import VGGCifar, fill_params_deterministic  # noqa: E402


def rs(seed):
    return np.random.RandomState(seed)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def piggy_like(rng, shape, frac_low=0.5):
    """Piggymask values straddling the 5e-3 threshold, some exactly at fp32(5e-3)."""
    p = rng.uniform(0.0, 0.01, size=shape).astype(np.float32)
    flat = p.reshape(-1)
    flat[::17] = np.float32(5e-3)                       # == thr -> 0
    flat[1::17] = np.nextafter(np.float32(5e-3), np.float32(1))  # just above -> 1
    return p


def conv_case(name, N, C, H, W, K, R, stride, pad, dil, groups, bias, seed, piggy=True):
    rng = rs(seed)
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((K, C // groups, R, R)) * 0.1).astype(np.float32)
    b = (rng.standard_normal((K,)) * 0.1).astype(np.float32) if bias else None
    p = piggy_like(rng, w.shape) if piggy else None
    m = nl.SharableConv2d(C, K, R, stride=stride, padding=pad, dilation=dil, groups=groups, bias=bias)
    with torch.no_grad():
        m.weight.copy_(t(w))
        if bias:
            m.bias.copy_(t(b))
    if piggy:
        m.piggymask = nn.Parameter(t(p.copy()))
    xt = t(x).requires_grad_(True)
    y = m(xt)
    dy = np.cos(np.arange(y.numel(), dtype=np.float64) * 0.37).astype(np.float32).reshape(tuple(y.shape))
    y.backward(t(dy))
    out = dict(x=x, w=w, dy=dy, y=y.detach().numpy(), dx=xt.grad.numpy(), dW=m.weight.grad.numpy(),
               conv=np.array([stride, pad, dil, groups], dtype=np.int64))
    if bias:
        out.update(b=b, db=m.bias.grad.numpy())
    if piggy:
        out.update(p=p, dP=m.piggymask.grad.numpy(),
                   bin=nl.Binarizer.apply(t(p), nl.DEFAULT_THRESHOLD).numpy())
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'y', tuple(y.shape))


def linear_case(name, N, I, O, seed, piggy=True):
    rng = rs(seed)
    x = rng.standard_normal((N, I)).astype(np.float32)
    w = (rng.standard_normal((O, I)) * 0.05).astype(np.float32)
    b = (rng.standard_normal((O,)) * 0.1).astype(np.float32)
    p = piggy_like(rng, w.shape) if piggy else None
    m = nl.SharableLinear(I, O)
    with torch.no_grad():
        m.weight.copy_(t(w))
        m.bias.copy_(t(b))
    if piggy:
        m.piggymask = nn.Parameter(t(p.copy()))
    xt = t(x).requires_grad_(True)
    y = m(xt)
    dy = np.sin(np.arange(y.numel(), dtype=np.float64) * 0.11).astype(np.float32).reshape(tuple(y.shape))
    y.backward(t(dy))
    out = dict(x=x, w=w, b=b, dy=dy, y=y.detach().numpy(), dx=xt.grad.numpy(),
               dW=m.weight.grad.numpy(), db=m.bias.grad.numpy())
    if piggy:
        out.update(p=p, dP=m.piggymask.grad.numpy())
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name)


def binarizer_case():
    thr32 = np.float32(5e-3)
    vals = np.array([0.0, -0.0, -1.0, 1.0, 5e-3, thr32, np.nextafter(thr32, np.float32(1)),
                     np.nextafter(thr32, np.float32(0)), 0.01, 0.00499, 0.00501, np.inf, -np.inf,
                     np.nan, 1e-38, -1e-38, 1e-45, 3.4e38], dtype=np.float32)
    rng = rs(5)
    vals = np.concatenate([vals, rng.uniform(-0.01, 0.02, 1000).astype(np.float32)])
    out = nl.Binarizer.apply(t(vals), nl.DEFAULT_THRESHOLD).numpy()
    g = rng.standard_normal(vals.shape).astype(np.float32)
    pin = t(vals.copy()).requires_grad_(True)
    nl.Binarizer.apply(pin, nl.DEFAULT_THRESHOLD).backward(t(g))
    np.savez_compressed(os.path.join(HERE, 'binarizer.npz'), p=vals, b=out, g=g, dp=pin.grad.numpy())
    print('binarizer')


class _Args(argparse.Namespace):
    pass


def make_args(mode, dataset='t2', wd=4e-5, freq=2, init_s=0.0, target_s=0.5):
    a = _Args()
    a.mode, a.dataset, a.cuda, a.weight_decay = mode, dataset, False, wd
    a.pruning_frequency, a.initial_sparsity, a.target_sparsity = freq, init_s, target_s
    a.network_width_multiplier, a.log_path, a.finetune_again = 1.0, None, False
    return a


class _Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.c1 = nl.SharableConv2d(3, 8, 3, padding=1, bias=False)
        self.c2 = nl.SharableConv2d(8, 12, 3, padding=1, bias=True)
        self.fc = nl.SharableLinear(12, 10)
        self.datasets = ['t1', 't2', 't3']

    def forward(self, x):
        x = torch.relu(self.c1(x))
        x = torch.relu(self.c2(x)).mean((2, 3))
        return self.fc(x)


def pruner_cases():
    """a6, a7, a8, a9, a10 on a toy model through the reference SparsePruner."""
    rng = rs(11)
    model = nn.DataParallel(_Toy())
    inner = model.module
    names, W, T, P, G, GP = [], {}, {}, {}, {}, {}
    for name, mod in model.named_modules():
        if isinstance(mod, (nl.SharableConv2d, nl.SharableLinear)):
            names.append(name)
            shp = tuple(mod.weight.shape)
            w = rng.standard_normal(shp).astype(np.float32)
            w.reshape(-1)[::7] = w.reshape(-1)[3]          # force magnitude ties
            w.reshape(-1)[5::11] *= -1
            W[name] = w
            T[name] = rng.randint(0, 4, size=shp).astype(np.uint8)  # tasks 0..3
            P[name] = piggy_like(rng, shp)
            G[name] = rng.standard_normal(shp).astype(np.float32)
            GP[name] = rng.standard_normal(shp).astype(np.float32)
            if mod.bias is not None:
                with torch.no_grad():
                    mod.bias.zero_()
    out = {'names': np.array(names)}

    def load(mode):
        masks = {}
        for name, mod in model.named_modules():
            if name in W:
                with torch.no_grad():
                    mod.weight.copy_(t(W[name]))
                mod.piggymask = nn.Parameter(t(P[name].copy()))
                mod.weight.grad = t(G[name].copy())
                mod.piggymask.grad = t(GP[name].copy())
                masks[name] = t(T[name].copy())
        args = make_args(mode, dataset='t2')   # prune: cur = index('t2')+1 = 2; finetune: cur = len-1 = 2
        pr = ref_prune.SparsePruner(model, masks, args, 0, 8, 2)
        return pr, masks

    for name in names:
        key = name.replace('.', '_')
        out['W_' + key], out['T_' + key] = W[name], T[name]
        out['P_' + key], out['G_' + key], out['GP_' + key] = P[name], G[name], GP[name]

    # a6 in both modes
    for mode in ('finetune', 'prune'):
        pr, masks = load(mode)
        assert pr.current_dataset_idx == 2
        pr.do_weight_decay_and_make_grads_zero()
        for name, mod in model.named_modules():
            if name in W:
                key = name.replace('.', '_')
                out[f'a6_{mode}_dW_{key}'] = mod.weight.grad.numpy().copy()
                out[f'a6_{mode}_dP_{key}'] = mod.piggymask.grad.numpy().copy()

    # a7 at several ratios (incl. banker's-rounding half cases) + exit-2 path
    ratios = [0.1, 0.25, 0.4375, 0.5, 0.75, 1.0, 0.0, 1e-9]
    out['a7_ratios'] = np.array(ratios)
    for i, ratio in enumerate(ratios):
        pr, masks = load('prune')
        for name, mod in model.named_modules():
            if name in W:
                key = name.replace('.', '_')
                try:
                    m = pr._pruning_mask(mod.weight.data, masks[name], name, ratio)
                    out[f'a7_{i}_T_{key}'] = m.numpy().copy()
                    out[f'a7_{i}_exit_{key}'] = np.array(0)
                except SystemExit as e:
                    out[f'a7_{i}_exit_{key}'] = np.array(int(e.code))

    # a8 schedule: gradually_prune over steps 0..11, freq 2, window [0, 8]
    pr, masks = load('prune')
    ratios_seen, zero_counts = [], []
    for step in range(12):
        ratios_seen.append(pr.gradually_prune(step))
        zero_counts.append([int(masks[n].eq(0).sum()) for n in names])
    out['a8_ratios'] = np.array(ratios_seen, dtype=np.float64)
    out['a8_zero_counts'] = np.array(zero_counts, dtype=np.int64)
    for name in names:
        out['a8_T_' + name.replace('.', '_')] = masks[name].numpy().copy()

    # a9 apply_mask (inference idx 2) / make_pruned_zero ; a10 make_finetuning_mask
    pr, masks = load('prune')
    pr.apply_mask()
    for name, mod in model.named_modules():
        if name in W:
            out['a9_apply_' + name.replace('.', '_')] = mod.weight.data.numpy().copy()
    pr, masks = load('prune')
    pr.make_pruned_zero()
    for name, mod in model.named_modules():
        if name in W:
            out['a9_zero_' + name.replace('.', '_')] = mod.weight.data.numpy().copy()
    pr, masks = load('prune')
    pr.make_finetuning_mask()
    out['a10_cur'] = np.array(pr.current_dataset_idx)
    for name in names:
        out['a10_T_' + name.replace('.', '_')] = masks[name].numpy().copy()
    # stats
    pr, masks = load('prune')
    out['stats'] = np.array([pr.calculate_sparsity(), pr.calculate_curr_task_ratio(),
                             pr.calculate_zero_ratio(), pr.calculate_shared_part_ratio()], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'pruner.npz'), **out)
    print('pruner')


def one_shot_case():
    """a11 one_shot_prune (utils/prune.py:94-109) on the toy model through the reference SparsePruner."""
    rng = rs(31)
    model = nn.DataParallel(_Toy())
    out, names, masks = {}, [], {}
    for name, mod in model.named_modules():
        if isinstance(mod, (nl.SharableConv2d, nl.SharableLinear)):
            names.append(name)
            shp = tuple(mod.weight.shape)
            w = rng.standard_normal(shp).astype(np.float32)
            w.reshape(-1)[::9] = w.reshape(-1)[2]           # magnitude ties
            tm = rng.randint(0, 4, size=shp).astype(np.uint8)
            with torch.no_grad():
                mod.weight.copy_(t(w))
            masks[name] = t(tm.copy())
            key = name.replace('.', '_')
            out['W_' + key], out['T_' + key] = w, tm
    pr = ref_prune.SparsePruner(model, masks, make_args('prune', dataset='t2'), 0, 8, 2)
    pr.one_shot_prune(0.35)
    for name, mod in model.named_modules():
        if name in masks:
            key = name.replace('.', '_')
            out['W1_' + key] = mod.weight.data.numpy().copy()
            out['T1_' + key] = pr.masks[name].numpy().copy()
    out['names'], out['ratio'], out['cur'] = np.array(names), np.array(0.35), np.array(pr.current_dataset_idx)
    np.savez_compressed(os.path.join(HERE, 'one_shot.npz'), **out)
    print('one_shot')


def trajectory_case(mode, name, steps=3, width=0.125, batch=8):
    """N training steps of a narrow VGG16-BN-cifar (task 2: piggymasks on every sharable
    layer) through the reference's unmodified Manager.train + SparsePruner."""
    torch.manual_seed(1)
    model = VGGCifar(nl.SharableConv2d, nl.SharableLinear, width=width)
    model.add_dataset('t1', 5)
    model.add_dataset('t2', 5)
    model.set_dataset('t2')
    fill_params_deterministic(model, seed=3)
    model = nn.DataParallel(model)
    rng = rs(21)
    masks = {}
    for n, m in model.named_modules():
        if isinstance(m, (nl.SharableConv2d, nl.SharableLinear)):
            tm = rng.randint(1, 3, size=tuple(m.weight.shape)).astype(np.uint8)  # tasks 1..2
            masks[n] = t(tm)
            pm = np.full(tuple(m.weight.shape), 0.01, dtype=np.float32)
            old = tm < 2
            pm[old] = rng.uniform(0, 0.01, size=int(old.sum())).astype(np.float32)
            m.piggymask = nn.Parameter(t(pm))
    args = make_args(mode, dataset='t2', freq=2, init_s=0.0, target_s=0.3)
    if mode == 'finetune':
        args.finetune_again = True      # cur = index('t2')+1 = 2
    loader = []
    for i in range(steps):
        data = rng.standard_normal((batch, 3, 32, 32)).astype(np.float32)
        target = rng.randint(0, 5, size=(batch,)).astype(np.int64)
        loader.append((t(data), t(target)))
    mgr = Manager(args, model, {}, masks, loader, loader, 0, 4)
    sgd_params = [p for n, p in model.named_parameters()
                  if 'piggymask' not in n and ('classifiers' not in n or '.1.' in n)]
    adam_params = [p for n, p in model.named_parameters() if 'piggymask' in n]
    opts = Optimizers()
    opts.add(torch.optim.SGD(sgd_params, lr=1e-2, weight_decay=0.0, momentum=0.9, nesterov=True), 1e-2)
    opts.add(torch.optim.Adam(adam_params, lr=5e-4), 5e-4)
    acc, step = mgr.train(opts, 0, [1e-2], 0)
    out = {'steps': np.array(steps), 'final_step': np.array(step)}
    for n, p in model.named_parameters():
        a = p.detach().numpy()
        out['sum_' + n] = np.array([a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum()])
    for n in masks:
        b = masks[n].numpy()
        out['maskzeros_' + n] = np.array(int((b == 0).sum()))
        out['maskcrc_' + n] = np.array(zlib.crc32(b.tobytes()))
    # the first sharable conv's final weight in full, as a direct tensor check
    first = [m for _, m in model.named_modules() if isinstance(m, nl.SharableConv2d)][0]
    out['w_first'] = first.weight.detach().numpy().copy()
    out['p_first'] = first.piggymask.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'final step', step)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'one_shot':     # added in round 2; leaves the other fixtures untouched
        one_shot_case()
        sys.exit(0)
    # BASELINE.json configs[0]: single SharableConv2d 3x3, batch 4x3x32x32
    conv_case('conv_cfg1', 4, 3, 32, 32, 64, 3, 1, 1, 1, 1, True, seed=1)
    conv_case('conv_cfg1_nopiggy', 4, 3, 32, 32, 64, 3, 1, 1, 1, 1, False, seed=2, piggy=False)
    conv_case('conv_c32', 2, 32, 8, 8, 64, 3, 1, 1, 1, 1, False, seed=3)
    conv_case('conv_s2_g2', 2, 8, 9, 9, 12, 3, 2, 1, 1, 2, True, seed=4)
    conv_case('conv_1x1_s2', 2, 16, 7, 7, 24, 1, 2, 0, 1, 1, False, seed=5)
    conv_case('conv_dil2', 1, 4, 10, 10, 6, 3, 1, 2, 2, 1, True, seed=6)
    conv_case('conv_7x7_s2', 1, 3, 20, 20, 8, 7, 2, 3, 1, 1, False, seed=7)
    linear_case('linear_small', 8, 48, 40, seed=8)
    linear_case('linear_nopiggy', 4, 32, 16, seed=9, piggy=False)
    binarizer_case()
    pruner_cases()
    trajectory_case('prune', 'traj_prune')
    trajectory_case('finetune', 'traj_finetune')
    one_shot_case()
    h = hashlib.sha256()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            h.update(open(os.path.join(HERE, f), 'rb').read())
    print('fixtures sha256', h.hexdigest()[:16])
