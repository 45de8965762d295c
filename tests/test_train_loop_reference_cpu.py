"""cpg_b200.train_loop.train_sync_free against the UNMODIFIED utils.manager.Manager.train (utils/manager.py:39-100),
both on the reference's own layers and pruner on the CPU (one shim: Tensor.cuda is the identity): same parameters,
piggymasks, batch-norm buffers and task masks after an epoch with prune events, and bit-identical return values
(average training accuracy, prune step counter) -- the loop only moves the host reads, not the arithmetic."""
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


CODE = r'''
import argparse, sys
import numpy as np
import torch
import torch.nn as nn
REF, ROOT, MODE, DATASET = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
torch.Tensor.cuda = lambda self, *a, **k: self
torch.set_num_threads(1)                                 # one summation order for both arms
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
import models
import models.layers as nl
from utils import Optimizers
from utils.manager import Manager
from cpg_b200.train_loop import install_sync_free_train
assert 'cpg_b200.layers' not in sys.modules              # reference layers and pruner on both sides

CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M']
T = lambda a: torch.from_numpy(np.ascontiguousarray(a))


def setup():
    torch.manual_seed(1)
    rng = np.random.RandomState(21)
    sli = {'t1': {'network_width_multiplier': 0.125}, DATASET: {'network_width_multiplier': 0.125}}
    model = models.custom_vgg_cifar100(CFG, dataset_history=[], dataset2num_classes={}, network_width_multiplier=0.125,
                                       shared_layer_info=sli)
    model.add_dataset('t1', 5)
    model.add_dataset(DATASET, 5)
    model.set_dataset(DATASET)
    model = nn.DataParallel(model)
    masks = {}
    for n, m in model.named_modules():
        if isinstance(m, (nl.SharableConv2d, nl.SharableLinear)):
            tm = rng.randint(1, 3, size=tuple(m.weight.shape)).astype(np.uint8)
            masks[n] = T(tm)
            pm = np.full(tuple(m.weight.shape), 0.01, dtype=np.float32)
            old = tm < 2
            pm[old] = rng.uniform(0, 0.01, size=int(old.sum())).astype(np.float32)
            m.piggymask = nn.Parameter(T(pm))
    a = argparse.Namespace()
    a.mode, a.dataset, a.cuda, a.weight_decay = MODE, DATASET, False, 4e-5
    a.pruning_frequency, a.initial_sparsity, a.target_sparsity = 2, 0.0, 0.3
    a.network_width_multiplier, a.log_path, a.finetune_again = 0.125, None, MODE == 'finetune'
    loader = [(T(rng.standard_normal((8, 3, 32, 32)).astype(np.float32)), T(rng.randint(0, 5, size=(8,)).astype(np.int64)))
              for _ in range(5)]
    mgr = Manager(a, model, sli, masks, loader, loader, 0, 4)
    if DATASET == 'face_verification':                  # only the metric branch of utils/manager.py:59 is of interest here;
        mgr.criterion = nn.CrossEntropyLoss()           # AngleLoss needs the SphereNet head's (cos, phi) pair
    sgd = [p for n, p in model.named_parameters() if 'piggymask' not in n and ('classifiers' not in n or '.1.' in n)]
    adam = [p for n, p in model.named_parameters() if 'piggymask' in n]
    opts = Optimizers()
    opts.add(torch.optim.SGD(sgd, lr=1e-2, weight_decay=0.0, momentum=0.9, nesterov=True), 1e-2)
    opts.add(torch.optim.Adam(adam, lr=5e-4), 5e-4)
    return mgr, model, masks, opts


ma, model_a, masks_a, opts_a = setup()
ra = ma.train(opts_a, 0, [1e-2], 0)
mb, model_b, masks_b, opts_b = setup()
install_sync_free_train(mb, postfix_every=2)
rb = mb.train(opts_b, 0, [1e-2], 0)
same = lambda u, v: (u == v) or (u != u and v != v)
assert same(ra[0], rb[0]) and ra[1] == rb[1], (ra, rb)
# ... and the validation pass that follows every epoch (utils/manager.py:103-152: apply_mask, eval, accuracy)
va, vb = ma.validate(0), mb.validate(0)
assert same(va, vb), (va, vb)
assert mb.validate.__func__.__name__ == 'validate_sync_free' and ma.validate.__func__.__name__ == 'validate'
for (na, pa), (nb, pb) in zip(model_a.named_parameters(), model_b.named_parameters()):
    assert na == nb and torch.equal(pa, pb), na
for (na, ba), (nb, bb) in zip(model_a.named_buffers(), model_b.named_buffers()):
    assert na == nb and torch.equal(ba, bb), na
for n in masks_a:
    assert torch.equal(masks_a[n], masks_b[n]), n
print('ok', ra)
'''


@pytest.mark.skipif(_ref_root() is None, reason='no reference checkout')
@pytest.mark.parametrize('mode,dataset', [('prune', 't2'), ('finetune', 't2'), ('finetune', 'face_verification')])
def test_sync_free_loop_equals_manager_train(mode, dataset):
    r = subprocess.run([sys.executable, '-c', CODE, _ref_root(), ROOT, mode, dataset], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and 'ok' in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
