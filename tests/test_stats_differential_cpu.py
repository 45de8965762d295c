"""K12 mask statistics (utils/prune.py:111-193): the oracle's counters (oracle.cpg_oracle.mask_stats) fed to the four
calculate_* methods of cpg_b200.prune.SparsePruner, against the UNMODIFIED reference methods on random task masks and
piggymasks whose values sit at, one ulp above and one ulp below the 0.005 cut of calculate_shared_part_ratio."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import argparse

import cpg_b200.prune as cpg_prune
from oracle import cpg_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(4, 3, 3, 3), (6, 4, 1, 1), (5, 8)]


def _ref_root():
    for p in ('/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if os.path.isfile(os.path.join(p, 'utils', 'prune.py')):
            return p
    return None


def stat_cases(n=60):
    rng = np.random.RandomState(314)
    thr = np.float32(0.005)
    near = np.array([thr, np.nextafter(thr, np.float32(1)), np.nextafter(thr, np.float32(0)), 0.0, 0.01, np.nan],
                    dtype=np.float32)
    out = []
    for i in range(n):
        idx = int(rng.randint(1, 5))
        width = float(rng.choice([1.0, 1.5 ** 0.5, 0.5]))
        hi = int(rng.choice([1, idx, idx + 2])) + 1
        masks = [rng.randint(0, hi, s).astype(np.uint8) for s in SHAPES]
        if i % 7 == 0:
            masks = [np.full(s, idx + 1, dtype=np.uint8) for s in SHAPES]       # nothing free, nothing shared
        piggies = [np.where(rng.rand(*s) < 0.4, rng.choice(near, s), rng.uniform(0, 0.01, s)).astype(np.float32)
                   for s in SHAPES]
        out.append((idx, width, masks, piggies))
    return out


REF_CODE = r'''
import argparse, json, sys
import torch
import torch.nn as nn
REF, ROOT, OUT = sys.argv[1], sys.argv[2], sys.argv[3]
torch.Tensor.cuda = lambda self, *a, **k: self
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
import models.layers as nl
from utils.prune import SparsePruner
from tests.test_stats_differential_cpu import SHAPES, stat_cases


class Stub:
    calculate_sparsity = SparsePruner.calculate_sparsity
    calculate_curr_task_ratio = SparsePruner.calculate_curr_task_ratio
    calculate_zero_ratio = SparsePruner.calculate_zero_ratio
    calculate_shared_part_ratio = SparsePruner.calculate_shared_part_ratio


res = []
for idx, width, masks, piggies in stat_cases():
    layers = [nl.SharableConv2d(s[1], s[0], s[2], bias=False) if len(s) == 4 else nl.SharableLinear(s[1], s[0], bias=False)
              for s in SHAPES]
    for l, p in zip(layers, piggies):
        l.piggymask = nn.Parameter(torch.from_numpy(p.copy()))
    s = Stub()
    s.model = nn.Sequential(*layers)
    s.masks = {str(i): torch.from_numpy(m.copy()) for i, m in enumerate(masks)}
    s.inference_dataset_idx = idx
    s.args = argparse.Namespace(network_width_multiplier=width)
    res.append([s.calculate_sparsity(), s.calculate_curr_task_ratio(), s.calculate_zero_ratio(),
                s.calculate_shared_part_ratio()])
json.dump(res, open(OUT, 'w'))
print('ok')
'''


@pytest.mark.skipif(_ref_root() is None, reason='no reference checkout')
def test_statistics_equal_the_reference(tmp_path):
    out = os.path.join(str(tmp_path), 'ref.json')
    r = subprocess.run([sys.executable, '-c', REF_CODE, _ref_root(), ROOT, out], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and 'ok' in r.stdout, r.stderr[-3000:]
    want = json.load(open(out))
    seen_shared = 0
    for (idx, width, masks, piggies), exp in zip(stat_cases(), want):
        tot = dict(zero=0, cur=0, shared=0, shared_picked=0, numel=0)
        for m, p in zip(masks, piggies):
            for k, v in O.mask_stats(torch.from_numpy(m), idx, torch.from_numpy(p)).items():
                tot[k] += v
        # the product's own methods over these counters (their kernel, cpgb_mask_stats_batched, is checked against the
        # same golden values on the GPU)
        class Stub:
            calculate_sparsity = cpg_prune.SparsePruner.calculate_sparsity
            calculate_curr_task_ratio = cpg_prune.SparsePruner.calculate_curr_task_ratio
            calculate_zero_ratio = cpg_prune.SparsePruner.calculate_zero_ratio
            calculate_shared_part_ratio = cpg_prune.SparsePruner.calculate_shared_part_ratio
            args = argparse.Namespace(network_width_multiplier=width)

            def _stats(self, with_piggy=False, _t=tot):
                return [_t['zero'], _t['cur'], _t['shared'], _t['shared_picked'] if with_piggy else 0, _t['numel']]
        st = Stub()
        got = [st.calculate_sparsity(), st.calculate_curr_task_ratio(), st.calculate_zero_ratio(),
               st.calculate_shared_part_ratio()]
        assert got == exp, (idx, width, got, exp)
        seen_shared += tot['shared'] > 0
    assert seen_shared > 20
