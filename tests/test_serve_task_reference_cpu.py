"""SparsePruner.serve_task against the UNMODIFIED reference's per-task evaluation flow, on the CPU.

Process A runs reference code only (models/vgg.py on models/layers.py, utils.manager.Manager, utils.prune.SparsePruner;
one shim: Tensor.cuda is the identity on this CUDA-less host): it writes a two-task checkpoint the way the training
runs do (Manager.save_checkpoint, utils/manager.py:198-231), then evaluates each task the way the reference does --
a fresh model per task, Manager.load_checkpoint_only_for_evaluate (:266-325), the piggymask binding of
CPG_cifar100_main_normal.py:282-289 and the destructive pruner.apply_mask() (utils/prune.py:223-231) -- and dumps
every per-task tensor plus the masked weights.

Process B builds ONE resident model from the same unmodified model file on cpg_b200.layers (cpg_b200.install()), loads
the checkpoint once, and switches between the tasks with SparsePruner.serve_task (the one kernel it launches,
cpgb_apply_mask, replaced by its host restatement): every tensor the layers / batch-norms / classifier would use must
equal the reference's dump bit for bit, and weight.data must stay whole.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_root():
    for p in ('/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if os.path.isfile(os.path.join(p, 'utils', 'manager.py')):
            return p
    return None


COMMON = r'''
import argparse, os, sys
import numpy as np
import torch
import torch.nn as nn
REF, OUT, ROOT = sys.argv[1], sys.argv[2], sys.argv[3]
CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M']
WIDTH = 0.125
FMT = '{save_folder}/checkpoint-{epoch}.pth.tar'
_load = torch.load
torch.load = lambda f, *a, **k: _load(f, *a, **dict(k, weights_only=False))     # the reference calls bare torch.load(path)


def args_for(dataset, mode):
    a = argparse.Namespace()
    a.mode, a.dataset, a.cuda, a.weight_decay = mode, dataset, False, 4e-5
    a.pruning_frequency, a.initial_sparsity, a.target_sparsity = 2, 0.0, 0.5
    a.network_width_multiplier, a.log_path, a.finetune_again = WIDTH, None, False
    a.checkpoint_format = FMT
    return a


def build(models, dataset, history, d2n, sli):
    torch.manual_seed(1)
    m = models.custom_vgg_cifar100(CFG, dataset_history=history, dataset2num_classes=d2n,
                                   network_width_multiplier=WIDTH, shared_layer_info=sli)
    m.add_dataset(dataset, 5)
    m.set_dataset(dataset)
    return nn.DataParallel(m)


def empty_info():
    return {'bias': {}, 'bn_layer_running_mean': {}, 'bn_layer_running_var': {}, 'bn_layer_weight': {},
            'bn_layer_bias': {}, 'piggymask': {},
            'network_width_multiplier': WIDTH}           # CPG_cifar100_main_normal.py:290
'''

PROC_A = COMMON + r'''
torch.Tensor.cuda = lambda self, *a, **k: self
sys.path.insert(0, REF)
import models
import models.layers as nl
from utils.manager import Manager
from torch.nn.parameter import Parameter

rng = np.random.RandomState(17)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a))


def sharable(model):
    return [(n, m) for n, m in model.named_modules() if isinstance(m, (nl.SharableConv2d, nl.SharableLinear))]


def perturb_task_tensors(model, scale):
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.weight.copy_(T((1.0 + scale * rng.standard_normal(m.weight.shape)).astype(np.float32)))
                m.bias.copy_(T((scale * rng.standard_normal(m.bias.shape)).astype(np.float32)))
                m.running_mean.copy_(T((scale * rng.standard_normal(m.bias.shape)).astype(np.float32)))
                m.running_var.copy_(T((1.0 + scale * rng.uniform(0, 1, m.bias.shape)).astype(np.float32)))
            elif isinstance(m, nl.SharableLinear):
                m.bias.copy_(T((scale * rng.standard_normal(m.bias.shape)).astype(np.float32)))


# ---- task 1: train + prune left T in {0, 1}; save_checkpoint
model = build(models, 't1', [], {}, {})
masks = {n: T(rng.randint(0, 2, tuple(m.weight.shape)).astype(np.uint8)) for n, m in sharable(model)}
perturb_task_tensors(model, 0.1)
mgr = Manager(args_for('t1', 'prune'), model, {'t1': empty_info()}, masks, [], [], 0, 4)
mgr.save_checkpoint(None, 0, OUT)

# ---- task 2 in a "new process": resume from the checkpoint, finetune + prune (T == 0 -> 2 or stays 0), piggymasks
ck = torch.load(FMT.format(save_folder=OUT, epoch=1))
sli = ck['shared_layer_info']
sli['t2'] = empty_info()
model = build(models, 't2', ck['dataset_history'], ck['dataset2num_classes'], sli)
masks = ck['masks']
mgr = Manager(args_for('t2', 'prune'), model, sli, masks, [], [], 0, 4)
mgr.load_checkpoint(None, 1, OUT)
for n, m in sharable(model):
    t = masks[n]
    free = (t == 0) & T(rng.rand(*t.shape) < 0.6)
    t[free] = 2
    with torch.no_grad():
        m.weight[t == 2] += T((0.05 * rng.standard_normal(tuple(t.shape))).astype(np.float32))[t == 2]
    m.piggymask = Parameter(T(rng.uniform(0, 0.01, tuple(t.shape)).astype(np.float32)))
perturb_task_tensors(model, 0.3)
mgr.save_checkpoint(None, 1, OUT)

# ---- the reference's evaluation of each task: one fresh model + one checkpoint load per task
x = T(rng.standard_normal((4, 3, 32, 32)).astype(np.float32))
for d in ('t1', 't2'):
    ck = torch.load(FMT.format(save_folder=OUT, epoch=2))
    sli, masks = ck['shared_layer_info'], ck['masks']
    model = build(models, d, ck['dataset_history'], ck['dataset2num_classes'], sli)
    mgr = Manager(args_for(d, 'inference'), model, sli, masks, [], [], 0, 4)
    mgr.load_checkpoint_only_for_evaluate(2, OUT)
    task_id = model.module.datasets.index(d) + 1
    if task_id > 1:                                   # CPG_cifar100_main_normal.py:282-289
        for n, m in model.module.named_modules():
            if isinstance(m, (nl.SharableConv2d, nl.SharableLinear)):
                m.piggymask = sli[d]['piggymask'][n]
    mgr.pruner.apply_mask()
    model.eval()
    out = {'task_id': np.array(task_id)}
    with torch.no_grad():
        out['logits'] = model(x).numpy()
    for n, m in model.module.named_modules():
        if isinstance(m, (nl.SharableConv2d, nl.SharableLinear)):
            out['w:' + n] = m.weight.detach().numpy().copy()
            if m.bias is not None:
                out['b:' + n] = m.bias.detach().numpy().copy()
            if m.piggymask is not None:
                out['p:' + n] = m.piggymask.detach().numpy().copy()
        elif isinstance(m, nn.BatchNorm2d):
            for k in ('weight', 'bias', 'running_mean', 'running_var'):
                out[k + ':' + n] = getattr(m, k).detach().numpy().copy()
    out['cw'] = model.module.classifier.weight.detach().numpy().copy()
    np.savez(os.path.join(OUT, 'expected_%s.npz' % d), **out)
np.save(os.path.join(OUT, 'x.npy'), x.numpy())
print('A ok')
'''

PROC_B = COMMON + r'''
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
import cpg_b200
layers, prune = cpg_b200.install()
import models
import models.layers as nl
from utils.manager import Manager
from cpg_b200 import _lib
assert nl is layers


class FakeLib:                                           # cpgb_apply_mask restated on host tensors, utils/prune.py:229-230
    def cpgb_apply_mask(self, w, t, n, idx, stream):
        w[(t == 0) | (t > idx)] = 0.0
        return 0


class NullCtx:
    def __init__(self, *a): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False


_lib.load = lambda: FakeLib()
_lib.ptr = lambda t: t
_lib.stream_ptr = lambda: 0
torch.cuda.device = NullCtx

ck = torch.load(FMT.format(save_folder=OUT, epoch=2))                   # ONE load
sli, masks = ck['shared_layer_info'], ck['masks']
model = build(models, 't2', ck['dataset_history'], ck['dataset2num_classes'], sli)
mgr = Manager(args_for('t2', 'inference'), model, sli, masks, [], [], 0, 4)
assert isinstance(mgr.pruner, prune.SparsePruner)
mgr.load_checkpoint_only_for_evaluate(2, OUT)
whole = {n: m.weight.detach().clone() for n, m in model.module.named_modules()
         if isinstance(m, (nl.SharableConv2d, nl.SharableLinear))}
model.eval()
checked = 0
for d in ('t1', 't2', 't1'):
    want = dict(np.load(os.path.join(OUT, 'expected_%s.npz' % d)))
    assert mgr.pruner.serve_task(d, sli) == int(want['task_id'])
    eq = lambda a, key: np.array_equal(a.detach().numpy(), want[key])
    for n, m in model.module.named_modules():
        if isinstance(m, (nl.SharableConv2d, nl.SharableLinear)):
            w, p = m._effective()
            assert eq(w, 'w:' + n), ('weight', d, n)
            assert torch.equal(m.weight.detach(), whole[n]), ('weight.data destroyed', d, n)
            if m.bias is not None:
                assert eq(m.bias, 'b:' + n), ('bias', d, n)
            assert (p is None) == (('p:' + n) not in want), ('piggymask presence', d, n)
            if p is not None:
                assert eq(p, 'p:' + n), ('piggymask', d, n)
            checked += 1
        elif isinstance(m, nn.BatchNorm2d):
            for k in ('weight', 'bias', 'running_mean', 'running_var'):
                assert eq(getattr(m, k), k + ':' + n), (k, d, n)
            checked += 1
    assert eq(model.module.classifier.weight, 'cw'), ('classifier', d)
assert checked == 3 * (15 + 13), checked
print('B ok')
'''


@pytest.mark.skipif(_ref_root() is None, reason='no reference checkout')
def test_serve_task_equals_the_reference_evaluation_flow(tmp_path):
    ref = _ref_root()
    for name, code in (('A', PROC_A), ('B', PROC_B)):
        r = subprocess.run([sys.executable, '-c', code, ref, str(tmp_path), ROOT], capture_output=True, text=True,
                           timeout=600, cwd=str(tmp_path))
        assert r.returncode == 0 and (name + ' ok') in r.stdout, (name, r.stdout[-1500:], r.stderr[-3000:])
