"""The a8 prune schedule of cpg_b200.prune.SparsePruner (python doubles) against the UNMODIFIED reference methods
(utils/prune.py:55-92) on seeded random schedules: `_adjust_sparsity`, `_time_to_update_masks`, and whole
`gradually_prune` runs (which steps fire a prune event, with which ratio, what is returned in between) must be
identical -- doubles compared bit for bit."""
import json
import os
import struct
import subprocess
import sys

import numpy as np
import pytest
import torch

import cpg_b200.layers as nl
import cpg_b200.prune as cpg_prune
from tests.toy import Toy, Wrap, make_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_root():
    for p in ('/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if os.path.isfile(os.path.join(p, 'utils', 'prune.py')):
            return p
    return None


CTOR_CASES = [(mode, ds, again, has) for mode in ('prune', 'inference', 'finetune', 'train')
              for ds in ('t1', 't3') for again, has in ((False, True), (True, True), (False, False))]


def schedules(n=150):
    rng = np.random.RandomState(88)
    out = []
    for i in range(n):
        begin = int(rng.randint(0, 50))
        end = begin + int(rng.choice([1, 2, 7, 79, 316, 1000]))
        freq = int(rng.choice([1, 2, 10, 25]))
        init = float(rng.choice([0.0, 0.1, 0.5, rng.uniform(0, 0.9)]))
        target = float(min(0.99, init + rng.choice([0.0, 0.1, 0.3, rng.uniform(0, 0.5)])))
        steps = sorted(set(int(s) for s in rng.randint(max(0, begin - 5), end + 20, size=40)))
        out.append(dict(begin=begin, end=end, freq=freq, init=init, target=target, steps=steps))
    return out


REF_CODE = r'''
import argparse, json, struct, sys
import torch
REF, ROOT, OUT = sys.argv[1], sys.argv[2], sys.argv[3]
torch.Tensor.cuda = lambda self, *a, **k: self
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
from utils.prune import SparsePruner
from tests.test_schedule_differential_cpu import CTOR_CASES, schedules
bits = lambda x: struct.unpack('<Q', struct.pack('<d', float(x)))[0]


class Stub:
    sparsity_func_exponent = 3
    _adjust_sparsity = SparsePruner._adjust_sparsity
    _time_to_update_masks = SparsePruner._time_to_update_masks
    gradually_prune = SparsePruner.gradually_prune

    def __init__(self):
        self.model = torch.nn.Sequential()          # no sharable layers: the event loop body is a7, tested elsewhere
        self.masks = {}


res = []
for sc in schedules():
    s = Stub()
    s.begin_prune_step, s.end_prune_step, s.last_prune_step = sc['begin'], sc['end'], sc['begin']
    s.args = argparse.Namespace(pruning_frequency=sc['freq'], initial_sparsity=sc['init'], target_sparsity=sc['target'])
    row = []
    for step in sc['steps']:
        fire = bool(s._time_to_update_masks(step))
        row.append([step, fire, bits(s._adjust_sparsity(step)), bits(s.gradually_prune(step)), s.last_prune_step])
    res.append(row)
ctor = []
for mode, dataset, again, has_attr in CTOR_CASES:
    m = torch.nn.Module()
    m.module = torch.nn.Module()
    m.module.datasets = ['t1', 't2', 't3']
    a = argparse.Namespace(mode=mode, dataset=dataset)
    if has_attr:
        a.finetune_again = again
    try:
        p = SparsePruner(m, {}, a, 3, 9, 2)
        ctor.append([p.current_dataset_idx, p.inference_dataset_idx, p.last_prune_step, p.sparsity_func_exponent])
    except SystemExit as e:
        ctor.append(['exit', e.code])
json.dump({'schedules': res, 'ctor': ctor}, open(OUT, 'w'))
print('ok')
'''


@pytest.mark.skipif(_ref_root() is None, reason='no reference checkout')
def test_schedule_equals_the_reference(tmp_path, monkeypatch):
    out = os.path.join(str(tmp_path), 'ref.json')
    r = subprocess.run([sys.executable, '-c', REF_CODE, _ref_root(), ROOT, out], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stderr[-3000:]
    both = json.load(open(out))
    want = both['schedules']
    # the constructor's task index per mode (utils/prune.py:18-25), incl. the unsupported-mode exit
    for (mode, dataset, again, has_attr), exp in zip(CTOR_CASES, both['ctor']):
        model = Wrap(Toy(nl))
        import argparse
        a = argparse.Namespace(mode=mode, dataset=dataset)
        if has_attr:
            a.finetune_again = again
        try:
            p = cpg_prune.SparsePruner(model, {}, a, 3, 9, 2)
            got = [p.current_dataset_idx, p.inference_dataset_idx, p.last_prune_step, p.sparsity_func_exponent]
        except SystemExit as e:
            got = ['exit', e.code]
        assert got == exp, (mode, dataset, again, has_attr, got, exp)
    bits = lambda x: struct.unpack('<Q', struct.pack('<d', float(x)))[0]
    fired = 0
    for sc, rows in zip(schedules(), want):
        model = Wrap(Toy(nl))
        masks = {n: torch.ones(tuple(m.weight.shape), dtype=torch.uint8) for n, m in model.named_modules()
                 if isinstance(m, (nl.SharableConv2d, nl.SharableLinear))}
        args = make_args('prune', freq=sc['freq'], init_s=sc['init'], target_s=sc['target'])
        pr = cpg_prune.SparsePruner(model, masks, args, sc['begin'], sc['end'], 2)
        events = []

        def fake_launch(layers, ratio, infos, sampled=None, _ev=events):
            _ev.append(ratio)
            infos.zero_()
        monkeypatch.setattr(pr, '_launch_prune_batched', fake_launch)
        for step, fire, adj, ret, last in rows:
            assert bool(pr._time_to_update_masks(step)) == fire, (sc, step)
            assert bits(pr._adjust_sparsity(step)) == adj, (sc, step)
            n_before = len(events)
            got = pr.gradually_prune(step)
            assert bits(got) == ret and pr.last_prune_step == last, (sc, step)
            assert (len(events) > n_before) == fire                    # the kernels run exactly at the reference's events
            if fire:
                assert bits(events[-1]) == ret
                fired += 1
    assert fired > 100
