"""Checkpoint layout of the drop-in layers against the UNMODIFIED reference (SURVEY 8b "State"): the reference's own model
files (models/vgg.py, models/resnet.py, models/spherenet.py) built once on models/layers.py and once, after
cpg_b200.install(), on cpg_b200.layers must expose the same state_dict keys, shapes and dtypes in the same order -- with
and without piggymasks, and after the optional fused-norm rewrites -- and the same module names for the mask
dictionary (`module.features.0`, ...)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_root():
    for p in ('/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if os.path.isfile(os.path.join(p, 'models', 'vgg.py')):
            return p
    return None


CODE = r'''
import json, sys
import torch
import torch.nn as nn
REF, ROOT, WHICH = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
if WHICH == 'product':
    import cpg_b200
    cpg_b200.install()
import models
import models.layers as nl
CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M']
out = {}
for arch in ('custom_vgg_cifar100', 'custom_vgg', 'resnet50', 'spherenet20'):
    kw = dict(dataset_history=[], dataset2num_classes={}, network_width_multiplier=0.25, shared_layer_info={})
    m = models.__dict__[arch](CFG, **kw) if 'vgg' in arch else models.__dict__[arch](**kw)
    m.add_dataset('t1', 5)
    m.add_dataset('t2', 7)
    m.set_dataset('t2')
    def describe(model):
        names = [n for n, mod in model.named_modules() if isinstance(mod, (nl.SharableConv2d, nl.SharableLinear))]
        return {'keys': [[k, list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()], 'sharable': names,
                'repr': [repr(mod) for mod in model.modules() if isinstance(mod, (nl.SharableConv2d, nl.SharableLinear))]}
    out[arch] = describe(m)
    for n, mod in m.named_modules():
        if isinstance(mod, (nl.SharableConv2d, nl.SharableLinear)):
            mod.piggymask = nn.Parameter(torch.full(tuple(mod.weight.shape), 0.01))
    out[arch + '+piggymask'] = describe(m)
    if WHICH == 'product':
        from cpg_b200.fused_norm import fuse_bn_relu, fuse_prelu, fuse_resnet_blocks
        fuse_bn_relu(m); fuse_prelu(m); fuse_resnet_blocks(m)
        out[arch + '+piggymask+fused'] = describe(m)
print('JSON' + json.dumps(out))
'''


@pytest.mark.skipif(_ref_root() is None, reason='no reference checkout')
def test_state_dict_layout_equals_the_reference():
    res = {}
    for which in ('reference', 'product'):
        r = subprocess.run([sys.executable, '-c', CODE, _ref_root(), ROOT, which], capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        res[which] = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith('JSON')][0][4:])
    ref, ours = res['reference'], res['product']
    assert len(ref) == 8
    for name, want in ref.items():
        assert ours[name] == want, name
        if '+piggymask' in name:
            assert ours[name + '+fused']['keys'] == want['keys'], name          # the rewrites move no key
            assert ours[name + '+fused']['sharable'] == want['sharable'], name
    assert any('piggymask' in k[0] for k in ref['resnet50+piggymask']['keys'])
    assert len(ref['custom_vgg_cifar100']['sharable']) == 15 and len(ref['resnet50']['sharable']) == 53
