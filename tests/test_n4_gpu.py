"""GPU tests of the SURVEY 8(f) N4 remainder:

  * batch-norm + residual add + ReLU (cpgb_bn_add_relu_fwd / _bwd, the tail of models/resnet.py's blocks) against the
    stock modules on the same GPU, and the block rewrite `fuse_resnet_blocks` against the blocks' own forward code;
  * the non-destructive evaluation predicate (SparsePruner.select_task) against the destructive `apply_mask()` of
    utils/prune.py:223-231: bit-identical logits for every task from ONE resident model, weights untouched.

relative error = max|a - b| / max|b| as in test_gpu_parity.py.  The ReLU gate of an element whose pre-activation is
zero to fp32 accuracy is undecidable (the two implementations round the batch-norm differently, and its gradient
flips between 0 and dy with the last bit of the statistics), so the stock modules' gradients are taken THROUGH THE
GATE OUR KERNELS CHOSE, after checking that the two gates only disagree where |pre-activation| < 1e-4.
"""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

from cpg_b200 import _lib  # noqa: E402
import cpg_b200.layers as nl  # noqa: E402
import cpg_b200.prune as cpg_prune  # noqa: E402
from cpg_b200.fused_norm import FusedBatchNormReLU2d, fuse_bn_relu, fuse_resnet_blocks  # noqa: E402
from cpg_b200.vgg_cifar import VGGCifar  # noqa: E402
from tests.toy import Wrap, make_args  # noqa: E402

DEV = 'cuda:0'


def rel(a, b, keep=None):
    a, b = a.detach().double(), b.detach().double()
    d = (a - b).abs()
    if keep is not None:
        d = d * keep
    return d.max().item() / max(b.abs().max().item(), 1e-30)


@pytest.fixture(autouse=True)
def _reset_path():
    _lib.set_path(_lib.PATH_AUTO)
    yield
    _lib.set_path(_lib.PATH_AUTO)


@pytest.mark.parametrize('shape', [(8, 64, 14, 14), (4, 78, 7, 7), (16, 256, 28, 28)])
@pytest.mark.parametrize('train', [True, False])
def test_bn_add_relu_vs_torch(shape, train):
    """y = relu(bn(x) + r) and its gradients (x, r, gamma, beta), running statistics over two steps."""
    N, C, H, W = shape
    torch.manual_seed(C)
    g = torch.Generator().manual_seed(C + H)
    CL = torch.channels_last
    x0 = (torch.randn(shape, generator=g) * 1.3 + 0.4).to(DEV).contiguous(memory_format=CL)
    r0 = torch.randn(shape, generator=g).to(DEV).contiguous(memory_format=CL)
    dy = torch.randn(shape, generator=g).to(DEV).contiguous(memory_format=CL)
    ref = nn.BatchNorm2d(C).to(DEV)
    with torch.no_grad():
        ref.weight.copy_(torch.linspace(0.5, 1.5, C))
        ref.bias.copy_(torch.linspace(-0.3, 0.3, C))
        ref.running_mean.copy_(torch.randn(C, generator=g) * 0.1)
        ref.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    ours = FusedBatchNormReLU2d(C, tf32_out=False).to(DEV)
    ours.load_state_dict(ref.state_dict())
    ref.train(train); ours.train(train)
    lib = _lib.load()
    for step in range(2):
        xa, ra = x0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
        xb, rb = x0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
        for p in list(ref.parameters()) + list(ours.parameters()):
            p.grad = None
        before = lib.cpgb_launch_count()
        yb = ours(xb, residual=rb)
        assert lib.cpgb_launch_count() - before == (3 if train else 2)
        before = lib.cpgb_launch_count()
        yb.backward(dy)
        assert lib.cpgb_launch_count() - before == 3
        pre = ref(xa) + ra
        gate = yb.detach() > 0
        assert int(((gate != (pre.detach() > 0)) & (pre.detach().abs() > 1e-4)).sum()) == 0
        (pre * gate).backward(dy)                            # relu(pre) differentiated through our gate
        torch.cuda.synchronize()
        assert yb.shape == pre.shape and bool(torch.isfinite(yb).all())
        assert rel(yb, torch.relu(pre)) <= 2e-5
        assert rel(rb.grad, ra.grad) == 0.0                  # dres = dy * [y > 0]
        assert rel(xb.grad, xa.grad) <= 1e-4
        assert rel(ours.weight.grad, ref.weight.grad) <= 1e-4
        assert rel(ours.bias.grad, ref.bias.grad) <= 1e-4
    assert rel(ours.running_mean, ref.running_mean) <= 1e-5
    assert rel(ours.running_var, ref.running_var) <= 1e-5
    assert int(ours.num_batches_tracked) == int(ref.num_batches_tracked) == (2 if train else 0)


def test_bn_add_relu_tf32_out_and_foreign_layouts():
    """tf32_out rounds the stored output (the next block's convolution input); a residual / output gradient in plain
    NCHW is re-packed."""
    shape = (4, 32, 10, 10)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(shape, generator=g).to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    r = torch.randn(shape, generator=g).to(DEV).requires_grad_(True)            # NCHW
    dy = torch.randn(shape, generator=g).to(DEV)                                # NCHW
    a = FusedBatchNormReLU2d(32, tf32_out=False).to(DEV)
    b = FusedBatchNormReLU2d(32, tf32_out=True).to(DEV)
    ya = a(x, residual=r)
    yb = b(x, residual=r)
    from cpg_b200.functional import is_tf32
    assert is_tf32(yb) and not is_tf32(ya)
    assert bool(((yb.detach().contiguous().view(torch.int32) & 0x1FFF) == 0).all())
    assert rel(yb, ya) <= 2.0 ** -11
    ga = torch.autograd.grad(ya, (x, r), dy)
    gb = torch.autograd.grad(yb, (x, r), dy)
    gate_same = ((ya > 0) == (yb > 0)).double()
    assert rel(gb[1], ga[1], gate_same) == 0.0
    assert rel(gb[0], ga[0], gate_same) <= 2e-3


def _resnet_blocks(conv):
    """Blocks with the attribute surface and forward code of models/resnet.py:20-100 on the given convolution class."""
    class BasicBlock(nn.Module):
        def __init__(self, c, down):
            super().__init__()
            self.conv1 = conv(c, c, 3, stride=2 if down else 1, padding=1, bias=False)
            self.bn1 = nn.BatchNorm2d(c)
            self.relu = nn.ReLU(inplace=True)
            self.conv2 = conv(c, c, 3, padding=1, bias=False)
            self.bn2 = nn.BatchNorm2d(c)
            self.downsample = nn.Sequential(conv(c, c, 1, stride=2, bias=False), nn.BatchNorm2d(c)) if down else None

        def forward(self, x):
            identity = x
            out = self.relu(self.bn1(self.conv1(x)))
            out = self.bn2(self.conv2(out))
            if self.downsample is not None:
                identity = self.downsample(x)
            out += identity
            return self.relu(out)

    class Bottleneck(nn.Module):
        def __init__(self, c, down):
            super().__init__()
            self.conv1 = conv(c, c // 2, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(c // 2)
            self.conv2 = conv(c // 2, c // 2, 3, stride=2 if down else 1, padding=1, bias=False)
            self.bn2 = nn.BatchNorm2d(c // 2)
            self.conv3 = conv(c // 2, c, 1, bias=False)
            self.bn3 = nn.BatchNorm2d(c)
            self.relu = nn.ReLU(inplace=True)
            self.downsample = nn.Sequential(conv(c, c, 1, stride=2, bias=False), nn.BatchNorm2d(c)) if down else None

        def forward(self, x):
            identity = x
            out = self.relu(self.bn1(self.conv1(x)))
            out = self.relu(self.bn2(self.conv2(out)))
            out = self.bn3(self.conv3(out))
            if self.downsample is not None:
                identity = self.downsample(x)
            out += identity
            return self.relu(out)

    return BasicBlock, Bottleneck


def test_fuse_resnet_blocks_matches_the_blocks_own_forward(monkeypatch):
    """Three copies of a small residual stack on the masked convolutions: (a) batch-norms swapped only (fuse_bn_relu:
    what bench_workloads ran so far), (b) block forward rewritten as well (fuse_resnet_blocks), (c) stock modules.
    (a) and (b) run the same convolution kernels on the same operands and, on the three-kernel batch-norm path, the
    same statistics kernels: the rewritten tail computes fmaf(x, a, b) + r exactly as bn -> `out += identity` does, so
    the two agree to rounding of the sums; (c) bounds both against torch's batch-norm."""
    monkeypatch.setenv('CPGB_BN_CLUSTER', '0')
    BasicBlock, Bottleneck = _resnet_blocks(nl.SharableConv2d)

    def net():
        return nn.Sequential(BasicBlock(64, False), BasicBlock(64, True), Bottleneck(64, False), Bottleneck(64, True))
    torch.manual_seed(5)
    a, b, c = net().to(DEV), net().to(DEV), net().to(DEV)
    with torch.no_grad():
        for m in a.modules():
            if isinstance(m, nl.SharableConv2d):                     # the reference layer leaves its weight uninitialised
                m.weight.normal_(0, (2.0 / (m.weight[0].numel())) ** 0.5)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.2, 0.2)
    b.load_state_dict(a.state_dict())
    c.load_state_dict(a.state_dict())
    keys = list(b.state_dict().keys())
    fuse_bn_relu(a, tf32_out=False)
    fuse_bn_relu(b, tf32_out=False)
    assert fuse_resnet_blocks(b, tf32_out=False) == 4
    assert list(b.state_dict().keys()) == keys
    g = torch.Generator().manual_seed(9)
    x = torch.randn(8, 64, 16, 16, generator=g).to(DEV)
    dy = torch.randn(8, 64, 4, 4, generator=g).to(DEV)
    l2 = lambda u, v: ((u.double() - v.double()).norm() / v.double().norm().clamp_min(1e-30)).item()
    for train in (True, False):
        res = []
        for net_ in (a, b, c):
            net_.train(train)
            for p in net_.parameters():
                p.grad = None
            xv = x.clone().requires_grad_(True)
            y = net_(xv)
            y.backward(dy)
            torch.cuda.synchronize()
            res.append((y.detach().clone(), xv.grad.clone(), [(n, p.grad.clone()) for n, p in net_.named_parameters()]))
        (ya, ga, pa), (yb, gb, pb), (yc, gc, pc) = res
        assert bool(torch.isfinite(yb).all()) and bool(torch.isfinite(gb).all())
        assert rel(yb, ya) <= 1e-5, rel(yb, ya)
        assert l2(gb, ga) <= 1e-4, l2(gb, ga)
        for (n1, u), (n2, v) in zip(pb, pa):
            assert n1 == n2 and l2(u, v) <= 1e-4, (n1, l2(u, v))
        # against the stock modules: values tightly; gradients in the L2 norm (an odd ReLU gate may sit on the other
        # side of zero there, which moves that element's gradient by dy)
        assert rel(yb, yc) <= 1e-3, rel(yb, yc)
        assert l2(gb, gc) <= 5e-2, l2(gb, gc)
        for (n1, u), (n2, v) in zip(pb, pc):
            assert n1 == n2 and l2(u, v) <= 5e-2, (n1, l2(u, v))
    for (ka, va), (kb, vb), (kc, vc) in zip(a.state_dict().items(), b.state_dict().items(), c.state_dict().items()):
        assert ka == kb == kc and rel(vb.float(), va.float()) <= 1e-5 and rel(vb.float(), vc.float()) <= 1e-3, ka


def test_select_task_equals_destructive_apply_mask():
    """One resident VGG16 (width 0.25) holding three tasks: logits of every task through select_task(j) -- weights
    untouched -- equal, bit for bit, what the reference's destructive apply_mask() gives (utils/prune.py:223-231)."""
    torch.manual_seed(13)
    inner = VGGCifar(nl.SharableConv2d, nl.SharableLinear, width=0.25)
    for t in ('t1', 't2', 't3'):
        inner.add_dataset(t, 10)
    inner.set_dataset('t3')
    model = Wrap(inner).to(DEV)
    rng = np.random.RandomState(21)
    masks, w0 = {}, {}
    layers = [(n, m) for n, m in model.named_modules() if isinstance(m, (nl.SharableConv2d, nl.SharableLinear))]
    for i, (name, m) in enumerate(layers):
        masks[name] = torch.from_numpy(rng.randint(0, 4, tuple(m.weight.shape)).astype(np.uint8)).to(DEV)
        if i % 2 == 1:                                               # piggymasks on every other layer
            m.piggymask = nn.Parameter(torch.from_numpy(rng.uniform(0, 0.01, tuple(m.weight.shape)).astype(np.float32)).to(DEV))
        w0[name] = m.weight.detach().clone()
    pr = cpg_prune.SparsePruner(model, masks, make_args('inference', dataset='t3'), 0, 8, 3)
    x = torch.randn(16, 3, 32, 32, device=DEV)
    model.eval()
    outs = {}
    with torch.no_grad():
        full = model(x).clone()
        for task in (1, 3, 2):
            pr.select_task(task)
            outs[task] = model(x).clone()
            assert all(torch.equal(m.weight.detach(), w0[n]) for n, m in layers)
        assert not torch.equal(outs[1], outs[2]) and not torch.equal(outs[2], outs[3]) and not torch.equal(outs[3], full)
        model.train()
        pr.select_task(1)
        model.eval()
        with pr.task_view(2):
            assert torch.equal(model(x), outs[2])
        assert torch.equal(model(x), full)                           # view dropped: weight.data again
        # the reference's way, most permissive task first (each call destroys more)
        for task in (3, 2, 1):
            pr.inference_dataset_idx = task
            pr.apply_mask()
            assert torch.equal(model(x), outs[task]), task
    keep = (masks[layers[0][0]] == 1)
    assert torch.equal(layers[0][1].weight.detach(), w0[layers[0][0]] * keep)
    pr.detach()
