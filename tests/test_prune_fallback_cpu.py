"""Host logic of a prune event (cpg_b200.prune.SparsePruner.gradually_prune) with a stand-in for the CUDA launch: the
two-pass select is tried first, the layers it reports as status 3 are finished by the four-pass radix select, status 2
is the reference's sys.exit(2) path (utils/prune.py:38-42, 78-92)."""
import pytest
import torch

import cpg_b200.layers as nl
import cpg_b200.prune as cpg_prune
from tests.toy import Toy, Wrap, make_args


def _pruner():
    model = Wrap(Toy(nl))
    masks = {}
    for name, mod in model.named_modules():
        if isinstance(mod, (nl.SharableConv2d, nl.SharableLinear)):
            with torch.no_grad():
                mod.weight.zero_()
            masks[name] = torch.full(tuple(mod.weight.shape), 2, dtype=torch.uint8)
    args = make_args('prune', freq=1)
    return cpg_prune.SparsePruner(model, masks, args, 0, 8, 2), masks


def test_status_3_layers_are_finished_by_the_radix_select(monkeypatch):
    pr, masks = _pruner()
    calls = []

    def fake_launch(layers, ratio, infos, sampled=None):
        calls.append(([n for n, _ in layers], sampled))
        infos.zero_()
        if sampled is None:                 # first attempt: the two-pass select cannot bracket the middle layer
            infos[1, 0] = 3
    monkeypatch.setattr(pr, '_launch_prune_batched', fake_launch)
    ratio = pr.gradually_prune(4)
    names = [n for n, _ in pr._sharable()]
    assert 0.0 < ratio <= 0.5
    assert calls == [(names, None), ([names[1]], False)]


def test_status_2_exits_like_the_reference(monkeypatch, capsys):
    pr, masks = _pruner()

    def fake_launch(layers, ratio, infos, sampled=None):
        infos.zero_()
        infos[0, 0] = 2
    monkeypatch.setattr(pr, '_launch_prune_batched', fake_launch)
    with pytest.raises(SystemExit) as e:
        pr.gradually_prune(4)
    assert e.value.code == 2
    assert 'Not enough weights for pruning' in capsys.readouterr().out


def test_status_2_after_the_fallback_still_exits(monkeypatch):
    pr, masks = _pruner()

    def fake_launch(layers, ratio, infos, sampled=None):
        infos.zero_()
        infos[0, 0] = 3 if sampled is None else 2
    monkeypatch.setattr(pr, '_launch_prune_batched', fake_launch)
    with pytest.raises(SystemExit) as e:
        pr.gradually_prune(4)
    assert e.value.code == 2


def test_no_event_outside_the_schedule(monkeypatch):
    pr, masks = _pruner()
    monkeypatch.setattr(pr, '_launch_prune_batched', lambda *a, **k: (_ for _ in ()).throw(AssertionError('launched')))
    pr.last_prune_step = 4
    pr.gradually_prune(4)                   # pruning_frequency steps have not passed: ratio only
