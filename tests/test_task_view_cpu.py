"""Host logic of the non-destructive evaluation predicate (cpg_b200.prune.SparsePruner.select_task) with a
stand-in for the one kernel it launches: evaluation-mode forward passes read W * [1 <= T <= task] from a resident
copy, training-mode passes and `weight.data` are untouched, the destructive `apply_mask()` of utils/prune.py:223-231
stays the default."""
import numpy as np
import torch

import cpg_b200._lib as _lib
import cpg_b200.layers as nl
import cpg_b200.prune as cpg_prune
from tests.toy import Toy, Wrap, make_args


class _NullCtx:
    def __init__(self, *a):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _FakeLib:
    """cpgb_apply_mask on host tensors: utils/prune.py:229-230."""
    calls = 0

    def cpgb_apply_mask(self, w, t, n, idx, stream):
        assert w.numel() == n == t.numel()
        w[(t == 0) | (t > idx)] = 0.0
        _FakeLib.calls += 1
        return 0


def _setup(monkeypatch):
    monkeypatch.setattr(_lib, 'load', lambda: _FakeLib())
    monkeypatch.setattr(_lib, 'ptr', lambda t: t)
    monkeypatch.setattr(_lib, 'stream_ptr', lambda: 0)
    monkeypatch.setattr(torch.cuda, 'device', _NullCtx)
    _FakeLib.calls = 0
    rng = np.random.RandomState(5)
    model = Wrap(Toy(nl))
    masks, w0 = {}, {}
    for name, mod in model.named_modules():
        if isinstance(mod, (nl.SharableConv2d, nl.SharableLinear)):
            with torch.no_grad():
                mod.weight.copy_(torch.from_numpy(rng.standard_normal(tuple(mod.weight.shape)).astype(np.float32) + 3.0))
            masks[name] = torch.from_numpy(rng.randint(0, 4, tuple(mod.weight.shape)).astype(np.uint8))
            w0[name] = mod.weight.detach().clone()
    pr = cpg_prune.SparsePruner(model, masks, make_args('inference', dataset='t3'), 0, 8, 2)
    return model, pr, masks, w0


def _layers(model):
    return [(n, m) for n, m in model.named_modules() if isinstance(m, (nl.SharableConv2d, nl.SharableLinear))]


def test_select_task_leaves_the_weights_and_serves_every_task(monkeypatch):
    model, pr, masks, w0 = _setup(monkeypatch)
    for task in (1, 2, 3, 1):
        assert pr.select_task(task) == task
        model.eval()
        for name, m in _layers(model):
            w, _ = m._effective()
            keep = (masks[name] >= 1) & (masks[name] <= task)
            assert torch.equal(w, w0[name] * keep)
            assert w is not m.weight and not w.requires_grad
            assert torch.equal(m.weight.detach(), w0[name])          # nothing destroyed
        model.train()
        for name, m in _layers(model):
            assert m._effective()[0] is m.weight                     # training ignores the view
    assert _FakeLib.calls == 4 * len(_layers(model))
    views = [m._cpg_task_view for _, m in _layers(model)]
    pr.select_task(2)
    assert all(v is m._cpg_task_view for v, (_, m) in zip(views, _layers(model)))      # buffers are reused
    pr.clear_task_view()
    model.eval()
    assert all(m._effective()[0] is m.weight for _, m in _layers(model))


def test_default_task_is_the_pruners_inference_index_and_the_context_manager_clears(monkeypatch):
    model, pr, masks, w0 = _setup(monkeypatch)
    model.eval()
    with pr.task_view() as p:
        assert p is pr and pr._task_view_idx == pr.inference_dataset_idx == 2
        name, m = _layers(model)[0]
        assert torch.equal(m._effective()[0], w0[name] * ((masks[name] >= 1) & (masks[name] <= 2)))
    assert pr._task_view_idx is None and all(m._cpg_task_view is None for _, m in _layers(model))


def test_apply_mask_is_destructive_unless_switched(monkeypatch):
    model, pr, masks, w0 = _setup(monkeypatch)
    model.eval()
    monkeypatch.setattr(cpg_prune, 'NONDESTRUCTIVE_APPLY_MASK', True)
    pr.apply_mask()
    for name, m in _layers(model):
        assert torch.equal(m.weight.detach(), w0[name])
        assert torch.equal(m._effective()[0], w0[name] * ((masks[name] >= 1) & (masks[name] <= 2)))
    monkeypatch.setattr(cpg_prune, 'NONDESTRUCTIVE_APPLY_MASK', False)
    pr.apply_mask()                                                  # the reference's semantics: weight.data is rewritten
    for name, m in _layers(model):
        assert m._cpg_task_view is None
        assert torch.equal(m.weight.detach(), w0[name] * ((masks[name] >= 1) & (masks[name] <= 2)))


def test_a_view_from_another_device_or_shape_is_refused(monkeypatch):
    import pytest
    model, pr, masks, w0 = _setup(monkeypatch)
    pr.select_task(1)
    model.eval()
    name, m = _layers(model)[0]
    m._cpg_task_view = torch.zeros(3)
    with pytest.raises(_lib.CpgbError):
        m._effective()


def test_serve_task_rebinds_the_per_task_tensors(monkeypatch):
    """One resident model, every task: classifier, biases, piggymasks, batch-norm tensors and PReLU slopes come out of
    the checkpoint's shared_layer_info (utils/manager.py:198-225 writes it, :305-325 and
    CPG_cifar100_main_normal.py:282-289 read it back, one process per task), the weights through select_task."""
    import pytest
    import torch.nn as nn
    _setup(monkeypatch)                                              # stand-ins for the CUDA pieces

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.c1 = nl.SharableConv2d(3, 8, 3, padding=1, bias=True)
            self.bn = nn.BatchNorm2d(8)
            self.act = nn.PReLU(8)
            self.fc = nl.SharableLinear(8, 6)
            self.datasets = ['t1', 't2', 't3']
            self.classifiers = nn.ModuleList([nn.Linear(6, 2) for _ in self.datasets])
            self.classifier = None

        def set_dataset(self, d):
            self.classifier = self.classifiers[self.datasets.index(d)]

    torch.manual_seed(2)
    net = Net()
    with torch.no_grad():
        net.c1.weight.normal_(); net.fc.weight.normal_()
    wrap = Wrap(net)
    masks = {'module.c1': torch.randint(0, 4, net.c1.weight.shape, dtype=torch.uint8),
             'module.fc': torch.randint(0, 4, net.fc.weight.shape, dtype=torch.uint8)}
    pr = cpg_prune.SparsePruner(wrap, masks, make_args('inference', dataset='t3'), 0, 8, 3)
    # what Manager.save_checkpoint leaves behind after each task (names relative to model.module)
    info = {}
    for t, d in enumerate(net.datasets, 1):
        info[d] = {'bias': {'c1': nn.Parameter(torch.full((8,), float(t))), 'fc': nn.Parameter(torch.full((6,), -float(t)))},
                   'piggymask': {} if t == 1 else {'c1': nn.Parameter(torch.full_like(net.c1.weight, 0.01 * t)),
                                                   'fc': nn.Parameter(torch.full_like(net.fc.weight, 0.01 * t))},
                   'bn_layer_running_mean': {'bn': torch.full((8,), 0.1 * t)},
                   'bn_layer_running_var': {'bn': torch.full((8,), 1.0 + t)},
                   'bn_layer_weight': {'bn': nn.Parameter(torch.full((8,), 2.0 * t))},
                   'bn_layer_bias': {'bn': nn.Parameter(torch.full((8,), 3.0 * t))},
                   'prelu_layer_weight': {'act': nn.Parameter(torch.full((8,), 0.05 * t))}}
    w_c1 = net.c1.weight.detach().clone()
    keys = list(net.state_dict().keys())
    wrap.eval()
    for d in ('t2', 't1', 't3', 't1'):
        t = net.datasets.index(d) + 1
        assert pr.serve_task(d, info) == t and pr.inference_dataset_idx == t
        assert net.classifier is net.classifiers[t - 1]
        assert net.c1.bias is info[d]['bias']['c1'] and net.fc.bias is info[d]['bias']['fc']
        assert (net.c1.piggymask is None) if t == 1 else (net.c1.piggymask is info[d]['piggymask']['c1'])
        assert net.bn.running_mean is info[d]['bn_layer_running_mean']['bn'] and net.bn.weight is info[d]['bn_layer_weight']['bn']
        assert net.bn.running_var is info[d]['bn_layer_running_var']['bn'] and net.bn.bias is info[d]['bn_layer_bias']['bn']
        assert net.act.weight is info[d]['prelu_layer_weight']['act']
        keep = (masks['module.c1'] >= 1) & (masks['module.c1'] <= t)
        assert torch.equal(net.c1._effective()[0], w_c1 * keep) and torch.equal(net.c1.weight.detach(), w_c1)
        fixed = lambda ks: [k for k in ks if 'piggymask' not in k and not k.startswith('classifier.')]
        assert fixed(net.state_dict().keys()) == fixed(keys)          # (set_dataset registers `classifier`, as in models/vgg.py:90-93)
    # a task trained at another width is refused, a missing entry is named
    bad = dict(info)
    bad['t2'] = dict(info['t2'], bn_layer_weight={'bn': nn.Parameter(torch.ones(12))})
    with pytest.raises(_lib.CpgbError, match='another network width'):
        pr.serve_task('t2', bad)
    bad['t2'] = dict(info['t2'], bias={'c1': info['t2']['bias']['c1']})
    with pytest.raises(_lib.CpgbError, match=r"no bias\[fc\]"):
        pr.serve_task('t2', bad)
