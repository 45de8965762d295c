"""The reference's OWN training controller driving the cpg_b200 layers on the GPU.

``baseline/_ref/`` holds the unmodified ivclab/CPG checkout (staged by tools/stage_reference.py; git-ignored).
After ``cpg_b200.install()`` the reference's ``utils.manager.Manager`` is imported from there, untouched, and

  * ``Manager.train`` (utils/manager.py:39-100) runs the 3-step golden trajectories of tests/golden/traj_*.npz --
    which the same ``Manager.train`` produced on the CPU with the reference's own layers -- on cuda:0 with the
    product layers and the product ``SparsePruner``;
  * ``save_checkpoint`` -> ``load_checkpoint`` (utils/manager.py:198-264) round-trips weights, task masks and
    piggymasks bit-exactly, with the reference's checkpoint keys (``module.``-prefixed mask names).
"""
import os
import sys
import zlib

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def ref_env():
    if not os.path.isdir(os.path.join(REF, 'utils')):
        pytest.skip('baseline/_ref not staged (python tools/stage_reference.py)')
    import cpg_b200
    sys.path.insert(0, REF)
    try:
        layers, prune = cpg_b200.install()
        import models                      # the reference package, now built on cpg_b200.layers
        import utils.manager as ref_manager
        from utils import Optimizers
        assert models.layers is layers and ref_manager.SparsePruner is prune.SparsePruner
        assert ref_manager.__file__.startswith(REF)
        _torch_version_shims()
        yield {'models': models, 'manager': ref_manager, 'Optimizers': Optimizers, 'nl': layers}
    finally:
        sys.path.remove(REF)


def _torch_version_shims():
    """The GPU twin of the `.cuda` no-op shim make_golden.py needs on a CUDA-less host (cpg_b200.torch_compat_shims);
    no reference file is edited."""
    import cpg_b200
    applied = cpg_b200.torch_compat_shims(DEV)
    print('torch compat shims applied:', applied)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _setup(env, mode, device=DEV):
    """tests/golden/make_golden.py::trajectory_case, input for input, with the model on the GPU."""
    from cpg_b200.vgg_cifar import VGGCifar, fill_params_deterministic
    from tests.trajectory import make_args
    nl = env['nl']
    torch.manual_seed(1)
    model = VGGCifar(nl.SharableConv2d, nl.SharableLinear, width=0.125)
    model.add_dataset('t1', 5)
    model.add_dataset('t2', 5)
    model.set_dataset('t2')
    fill_params_deterministic(model, seed=3)
    model = nn.DataParallel(model.to(device), device_ids=[0])
    rng = np.random.RandomState(21)
    masks = {}
    for n, m in model.named_modules():
        if isinstance(m, (nl.SharableConv2d, nl.SharableLinear)):
            tm = rng.randint(1, 3, size=tuple(m.weight.shape)).astype(np.uint8)
            masks[n] = torch.from_numpy(tm).to(device)
            pm = np.full(tuple(m.weight.shape), 0.01, dtype=np.float32)
            old = tm < 2
            pm[old] = rng.uniform(0, 0.01, size=int(old.sum())).astype(np.float32)
            m.piggymask = nn.Parameter(torch.from_numpy(pm).to(device))
    args = make_args(mode, dataset='t2', freq=2, init_s=0.0, target_s=0.3)   # finetune_again = (mode == 'finetune')
    args.cuda = True
    loader = []
    for _ in range(3):
        data = rng.standard_normal((8, 3, 32, 32)).astype(np.float32)
        target = rng.randint(0, 5, size=(8,)).astype(np.int64)
        loader.append((torch.from_numpy(data), torch.from_numpy(target)))
    shared = {'t2': {'bias': {}, 'bn_layer_running_mean': {}, 'bn_layer_running_var': {}, 'bn_layer_weight': {},
                     'bn_layer_bias': {}, 'piggymask': {}}}
    mgr = env['manager'].Manager(args, model, shared, masks, loader, loader, 0, 4)
    sgd_params = [p for n, p in model.named_parameters()
                  if 'piggymask' not in n and ('classifiers' not in n or '.1.' in n)]
    adam_params = [p for n, p in model.named_parameters() if 'piggymask' in n]
    opts = env['Optimizers']()
    opts.add(torch.optim.SGD(sgd_params, lr=1e-2, weight_decay=0.0, momentum=0.9, nesterov=True), 1e-2)
    opts.add(torch.optim.Adam(adam_params, lr=5e-4), 5e-4)
    return model, masks, mgr, opts, args


@pytest.mark.parametrize('mode', ['prune', 'finetune'])
def test_unmodified_manager_train_matches_golden_trajectory(ref_env, golden, mode):
    from cpg_b200 import _lib
    import cpg_b200.prune as cpg_prune
    g = golden('traj_' + mode)
    _lib.set_path(_lib.PATH_SIMT)          # the fixture is an fp32 CPU run: compare on the fp32 kernels
    try:
        model, masks, mgr, opts, args = _setup(ref_env, mode)
        assert type(mgr.pruner) is cpg_prune.SparsePruner and mgr.pruner.current_dataset_idx == 2
        acc, step = mgr.train(opts, 0, [1e-2], 0)
        torch.cuda.synchronize()
    finally:
        _lib.set_path(_lib.PATH_AUTO)
    assert step == int(g['final_step'])
    first = [m for _, m in model.named_modules() if isinstance(m, ref_env['nl'].SharableConv2d)][0]
    assert rel(first.weight, torch.from_numpy(g['w_first'])) <= 1e-3
    assert rel(first.piggymask, torch.from_numpy(g['p_first'])) <= 1e-3
    worst = 0.0
    for n, p in model.named_parameters():
        a = p.detach().double().cpu().numpy()
        ref = g['sum_' + n]
        worst = max(worst, abs(np.abs(a).sum() - ref[1]) / max(ref[1], 1e-12))
    assert worst <= 1e-3, worst
    for n in masks:
        b = mgr.pruner.masks[n].cpu().numpy()
        assert int((b == 0).sum()) == int(g['maskzeros_' + n]), n
        assert zlib.crc32(b.tobytes()) == int(g['maskcrc_' + n]), n        # bit-exact task masks after the prune event


def test_unmodified_manager_train_on_tensor_core_path(ref_env, golden):
    """Same driver on the benched path (PATH_AUTO): the narrow net amplifies TF32 rounding chaotically over three
    steps (see test_trajectory_vs_reference_golden), so this is a wiring check with a loose bar."""
    g = golden('traj_prune')
    model, masks, mgr, opts, args = _setup(ref_env, 'prune')
    acc, step = mgr.train(opts, 0, [1e-2], 0)
    torch.cuda.synchronize()
    assert step == int(g['final_step'])
    worst = 0.0
    for n, p in model.named_parameters():
        a = p.detach().double().cpu().numpy()
        assert np.isfinite(a).all(), n
        ref = g['sum_' + n]
        worst = max(worst, abs(np.abs(a).sum() - ref[1]) / max(ref[1], 1e-12))
    assert worst <= 0.3, worst
    for n in masks:
        z, zr = int((mgr.pruner.masks[n] == 0).sum()), int(g['maskzeros_' + n])
        assert abs(z - zr) <= max(2, 0.02 * zr), (n, z, zr)       # k = round(ratio * pool) is data-independent


def test_checkpoint_round_trip_through_reference_manager(ref_env, tmp_path):
    model, masks, mgr, opts, args = _setup(ref_env, 'prune')
    args.checkpoint_format = '{save_folder}/checkpoint-{epoch}.pth.tar'
    mgr.train(opts, 0, [1e-2], 0)
    mgr.validate(0)                          # apply_mask (destructive zeroing) + eval forward + the four statistics
    folder = str(tmp_path)
    mgr.save_checkpoint(opts, 0, folder)
    path = args.checkpoint_format.format(save_folder=folder, epoch=1)
    ck = torch.load(path, weights_only=False)
    assert sorted(ck.keys()) == ['dataset2num_classes', 'dataset_history', 'masks', 'model_state_dict', 'shared_layer_info']
    assert ck['dataset_history'] == ['t1', 't2']
    # reference checkpoint layout: uint8 task masks under module.-prefixed names, weight shapes
    for n, m in model.named_modules():
        if isinstance(m, (ref_env['nl'].SharableConv2d, ref_env['nl'].SharableLinear)):
            assert n.startswith('module.') and ck['masks'][n].dtype == torch.uint8
            assert ck['masks'][n].shape == m.weight.shape
            assert torch.equal(ck['masks'][n].cpu(), mgr.pruner.masks[n].cpu())
            inner = n[len('module.'):]
            assert torch.equal(ck['shared_layer_info']['t2']['piggymask'][inner].detach().cpu(), m.piggymask.detach().cpu())
            assert torch.equal(ck['model_state_dict'][inner + '.weight'].cpu(), m.weight.detach().cpu())
            assert inner + '.piggymask' in ck['model_state_dict']
    # a fresh model + Manager resumes from it through the reference's own load_checkpoint
    model2, masks2, mgr2, opts2, args2 = _setup(ref_env, 'prune')
    args2.checkpoint_format = args.checkpoint_format
    with torch.no_grad():
        for p in model2.parameters():
            p.add_(1.0)                      # make sure the load really overwrites
    mgr2.load_checkpoint(opts2, 1, folder)
    sd1, sd2 = model.module.state_dict(), model2.module.state_dict()
    for k in sd1:
        if 'piggymask' in k or k.startswith('classifier.'):
            continue                         # utils/manager.py:243-246 skips these on purpose
        assert torch.equal(sd1[k].cpu(), sd2[k].cpu()), k
    # and the resumed model computes the same thing through the product kernels
    model.eval(); model2.eval()
    for n, m in model2.named_modules():
        if isinstance(m, (ref_env['nl'].SharableConv2d, ref_env['nl'].SharableLinear)):
            m.piggymask = ck['shared_layer_info']['t2']['piggymask'][n[len('module.'):]]
    x = torch.randn(4, 3, 32, 32, device=DEV)
    with torch.no_grad():
        assert torch.equal(model(x), model2(x))


# --------------------------------------------------------------------------------------------------------
# the reference's other model files, unmodified, built on the product layers (BASELINE.json configs[3], [4])
# --------------------------------------------------------------------------------------------------------
class _torch_ops_layers:
    """Context manager: evaluate SharableConv2d / SharableLinear with the reference's expressions in stock torch ops
    (models/layers.py:98-109, 184-194) -- same module objects, same weights -- as the fp32 yardstick."""

    def __init__(self, nl):
        self.nl = nl

    def __enter__(self):
        import torch.nn.functional as F
        nl = self.nl
        self.saved = (nl.SharableConv2d.forward, nl.SharableLinear.forward)

        def eff(m):
            if m.piggymask is None:
                return m.weight
            b = (m.piggymask > 5e-3).float()
            return (b - m.piggymask).detach() * m.weight + m.piggymask * m.weight

        nl.SharableConv2d.forward = lambda m, x, *a, **k: F.conv2d(x, eff(m), m.bias, m.stride, m.padding, m.dilation, m.groups)
        nl.SharableLinear.forward = lambda m, x: F.linear(x, eff(m), m.bias)
        self.flags = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        return self

    def __exit__(self, *exc):
        self.nl.SharableConv2d.forward, self.nl.SharableLinear.forward = self.saved
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.flags


REF_ARCHS = {
    # name: (constructor, constructor args, input shape, dataset name)
    'resnet50': ('resnet50', (), (8, 3, 224, 224), 'cubs'),
    'spherenet20': ('spherenet20', (), (8, 3, 112, 112), 'age'),
    'spherenet20_face': ('spherenet20', (), (8, 3, 112, 112), 'face_verification'),
    'custom_vgg_224': ('custom_vgg', ([64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M'],),
                       (4, 3, 224, 224), 'imagenet'),
    'custom_vgg_cifar100_x1.5': ('custom_vgg_cifar100',
                                 ([64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M'],),
                                 (16, 3, 32, 32), 'task'),
}


@pytest.mark.parametrize('arch', sorted(REF_ARCHS))
def test_unmodified_reference_models_forward_backward(ref_env, arch):
    """models/{resnet,spherenet,vgg}.py of the staged checkout, untouched, after cpg_b200.install(): forward + backward
    on the GPU through the product kernels (piggymasks on every sharable layer) against the same modules evaluated
    with the reference's torch expressions in fp32."""
    models, nl = ref_env['models'], ref_env['nl']
    ctor, cargs, shape, dataset = REF_ARCHS[arch]
    width = 1.5 ** 0.5 if arch.endswith('x1.5') else 1.0
    torch.manual_seed(3)
    model = getattr(models, ctor)(*cargs, dataset_history=[], dataset2num_classes={}, network_width_multiplier=width,
                                  shared_layer_info={})
    model.add_dataset(dataset, 10)
    model.set_dataset(dataset)
    model = model.to(DEV)
    if arch == 'resnet50':                # the reference initialises its convolutions to N(0, 0.001): signal dies in fp32
        for m in model.modules():
            if isinstance(m, nl.SharableConv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
    sharable = [m for m in model.modules() if isinstance(m, (nl.SharableConv2d, nl.SharableLinear))]
    assert len(sharable) == {'resnet50': 53, 'spherenet20': 20, 'spherenet20_face': 20, 'custom_vgg_224': 15,
                             'custom_vgg_cifar100_x1.5': 15}[arch]
    g = torch.Generator().manual_seed(1)
    for m in sharable:
        m.piggymask = nn.Parameter((torch.rand(m.weight.shape, generator=g) * 0.01).to(DEV))
    model.eval()                           # dropout off, BN on running statistics: a deterministic function of x
    x = torch.randn(*shape, generator=g).to(DEV)

    def run():
        for p in model.parameters():
            p.grad = None
        out = model(x)
        if isinstance(out, tuple):         # AngleLinear returns (cos_theta, phi_theta)
            out = out[0]
        out.square().mean().backward()
        return out.detach().clone(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    from cpg_b200 import _lib
    lib = _lib.load()
    before = lib.cpgb_launch_count()
    out, grads = run()
    assert lib.cpgb_launch_count() - before >= 3 * len(sharable)
    with _torch_ops_layers(nl):
        out_ref, grads_ref = run()
    assert out.shape[0] == shape[0] and torch.isfinite(out).all()
    e_out = rel(out, out_ref)
    assert e_out <= 5e-3, (arch, e_out)
    assert set(grads) == set(grads_ref)
    for n in grads:
        a, b = grads[n].double(), grads_ref[n].double()
        l2 = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
        assert l2 <= 0.15, (arch, n, l2)      # first-layer gradients sit behind every ReLU gate flip of the network
