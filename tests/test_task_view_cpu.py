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
