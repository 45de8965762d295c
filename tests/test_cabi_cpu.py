"""CPU-side checks: the C-ABI library loads, exports every symbol include/cpgb200.h declares,
argument validation works without a GPU, and the product refuses to run on CPU tensors
(no silent fallback)."""
import ctypes
import os
import re

import pytest
import torch

from cpg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()


def test_library_exports_every_declared_symbol():
    _ensure_built()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(ROOT, 'include', 'cpgb200.h')).read()
    declared = sorted(set(re.findall(r'\b(cpgb_[a-z0-9_]+)\s*\(', header)))
    assert declared, 'no declarations parsed'
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/cpgb200.h but not exported'
    assert set(_lib.EXPORTS) <= set(declared)


def test_version_and_validation_without_gpu():
    _ensure_built()
    lib = _lib.load()
    assert lib.cpgb_version() == 100
    d = _lib.ConvDesc()
    lib.cpgb_linear_desc(d, 4, 8, 16)
    assert (d.N, d.C, d.K, d.H, d.R, d.P) == (4, 8, 16, 1, 1, 1)
    assert lib.cpgb_workspace_bytes(d) >= 8 * 16 * 4
    # groups not dividing channels -> EINVAL (the module raises ValueError before that)
    bad = _lib.conv_desc((1, 6, 4, 4), (96, 16, 4, 1), (8, 2, 3, 3), (1, 8, 4, 4), (128, 16, 4, 1),
                         (1, 1), (1, 1), (1, 1), 4)
    rc = lib.cpgb_conv2d_fprop(bad, None, None, None, None, None, 5e-3, None, None, 0, None)
    assert rc == -1 and b'groups' in lib.cpgb_last_error()
    # wrong output extent
    bad = _lib.conv_desc((1, 4, 4, 4), (64, 16, 4, 1), (8, 4, 3, 3), (1, 8, 9, 9), (648, 81, 9, 1),
                         (1, 1), (1, 1), (1, 1), 1)
    assert lib.cpgb_conv2d_fprop(bad, None, None, None, None, None, 5e-3, None, None, 0, None) == -1
    assert lib.cpgb_set_path(_lib.PATH_AUTO) in (0, 1, 2)


def test_module_surface_matches_reference():
    import cpg_b200.layers as nl
    conv = nl.SharableConv2d(4, 8, 3, padding=1, bias=False)
    lin = nl.SharableLinear(8, 4)
    assert [n for n, _ in conv.named_parameters()] == ['weight']
    assert conv.piggymask is None and conv.info == {'threshold_fn': 'binarizer', 'threshold': 5e-3}
    conv.piggymask = torch.nn.Parameter(torch.full_like(conv.weight, 0.01))
    assert list(conv.state_dict().keys()) == ['weight', 'piggymask']
    assert list(lin.state_dict().keys()) == ['weight', 'bias']
    assert repr(conv) == 'SharableConv2d (4, 8, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), bias=False)'
    assert repr(lin) == 'SharableLinear(in_features=8, out_features=4)'
    with pytest.raises(ValueError):
        nl.SharableConv2d(6, 8, 3, groups=4)
    assert nl.DEFAULT_THRESHOLD == 5e-3


def test_no_cpu_fallback():
    import cpg_b200.layers as nl
    conv = nl.SharableConv2d(4, 8, 3, padding=1)
    torch.nn.init.normal_(conv.weight)
    torch.nn.init.zeros_(conv.bias)
    with pytest.raises(_lib.CpgbError):
        conv(torch.zeros(1, 4, 5, 5))
    with pytest.raises(_lib.CpgbError):
        nl.Binarizer.apply(torch.zeros(4), 5e-3)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'cpg_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f


@pytest.mark.skipif(not os.path.isdir('/root/reference/models'), reason='reference checkout not present (GPU box)')
def test_install_makes_reference_models_build_on_our_layers():
    """after cpg_b200.install() the UNMODIFIED reference model files build their
    networks out of this package's layers (isinstance checks of utils/prune.py / utils/manager.py keep
    firing) and the reference's SparsePruner name resolves to ours.  Runs in a subprocess so the
    sys.modules aliasing does not leak into the other tests; no kernel is launched (CPU box)."""
    import subprocess
    import sys
    code = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, '/root/reference')
import cpg_b200
layers, prune = cpg_b200.install()
import models, models.vgg, models.resnet, models.spherenet
import models.layers as nl
assert nl is layers and nl.SharableConv2d is layers.SharableConv2d
net = models.vgg.custom_vgg_cifar100(custom_cfg=[64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M'],
                                     dataset_history=[], dataset2num_classes={}, network_width_multiplier=1.0,
                                     shared_layer_info={})
convs = [m for m in net.modules() if isinstance(m, nl.SharableConv2d)]
lins = [m for m in net.modules() if isinstance(m, nl.SharableLinear)]
assert len(convs) == 13 and len(lins) == 2, (len(convs), len(lins))
assert sorted(k for k in convs[0].state_dict()) == ['weight'] and convs[0].piggymask is None
r50 = models.resnet.resnet50(dataset_history=[], dataset2num_classes={}, network_width_multiplier=1.0, shared_layer_info={})
assert sum(isinstance(m, nl.SharableConv2d) for m in r50.modules()) == 53
from cpg_b200.fused_norm import FusedBatchNormReLU2d, fuse_bn_relu, fuse_resnet_blocks
keys = list(r50.state_dict().keys())
fuse_bn_relu(r50)
assert fuse_resnet_blocks(r50) == 16 and fuse_resnet_blocks(r50) == 0      # every Bottleneck of models/resnet.py:60-100
assert list(r50.state_dict().keys()) == keys
assert all(isinstance(b, models.resnet.Bottleneck) and isinstance(b.bn3, FusedBatchNormReLU2d)
           for b in r50.modules() if type(b).__name__.startswith('FusedBottleneck'))
s20 = models.spherenet.spherenet20(dataset_history=[], dataset2num_classes={}, network_width_multiplier=1.0,
                                   shared_layer_info={})
assert sum(isinstance(m, nl.SharableConv2d) for m in s20.modules()) == 20
import utils.prune, utils.manager
assert utils.prune.SparsePruner is prune.SparsePruner and utils.manager.SparsePruner is prune.SparsePruner
print('ok')
''' % ROOT
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and 'ok' in out.stdout, out.stderr[-2000:]
