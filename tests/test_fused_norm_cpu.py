"""Host-side checks of cpg_b200.fused_norm (no GPU): the model rewrite keeps names / tensors, and inputs
the kernels do not take go through stock torch with identical results."""
import torch
import torch.nn as nn

from cpg_b200.fused_norm import FusedBatchNormReLU2d, fuse_bn_relu


def _net():
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8), nn.ReLU(inplace=True),
                         nn.MaxPool2d(2), nn.Conv2d(8, 12, 3, padding=1), nn.BatchNorm2d(12), nn.Sigmoid())


def test_fuse_keeps_names_tensors_and_results():
    torch.manual_seed(0)
    ref, net = _net(), _net()
    net.load_state_dict(ref.state_dict())
    keys = list(net.state_dict().keys())
    bn1 = net[1]
    assert fuse_bn_relu(net) == (1, 1)
    assert list(net.state_dict().keys()) == keys                      # checkpoint layout unchanged
    assert isinstance(net[1], FusedBatchNormReLU2d) and isinstance(net[1], nn.BatchNorm2d) and net[1].relu
    assert isinstance(net[2], nn.Identity) and isinstance(net[5], FusedBatchNormReLU2d) and not net[5].relu
    assert net[1].weight is bn1.weight and net[1].running_mean is bn1.running_mean
    assert fuse_bn_relu(net) == (0, 0)                                # idempotent
    x = torch.randn(4, 3, 8, 8)
    for train in (True, False):
        ref.train(train); net.train(train)
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ya, yb = ref(xa), net(xb)
        assert torch.equal(ya, yb)                                    # CPU: stock path, bit for bit
        ya.sum().backward(); yb.sum().backward()
        assert torch.equal(xa.grad, xb.grad)
    for (ka, va), (kb, vb) in zip(ref.state_dict().items(), net.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka                   # running statistics / step counters too


def test_fused_module_state_dict_roundtrip():
    m = FusedBatchNormReLU2d(8, relu=True)
    plain = nn.BatchNorm2d(8)
    plain.load_state_dict(m.state_dict())
    m.load_state_dict(plain.state_dict())
    assert 'relu=True' in repr(m)


def test_fuse_handles_attribute_style_blocks():
    """ResNet-style blocks keep their ReLU as a shared attribute and call it from forward(): only the
    batch-norm is swapped there (relu=False), nothing else moves."""
    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(4, 4, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(4)
            self.relu = nn.ReLU(inplace=True)

        def forward(self, x):
            return self.relu(self.bn1(self.conv1(x)) + x)

    torch.manual_seed(1)
    a, b = Block(), Block()
    b.load_state_dict(a.state_dict())
    assert fuse_bn_relu(b) == (0, 1)
    assert isinstance(b.bn1, FusedBatchNormReLU2d) and not b.bn1.relu and not b.bn1.pool
    assert isinstance(b.relu, nn.ReLU)
    x = torch.randn(2, 4, 5, 5)
    assert torch.equal(a(x), b(x))
    assert list(a.state_dict().keys()) == list(b.state_dict().keys())


def _blocks():
    """Blocks with the attribute surface and forward code of models/resnet.py:20-100 (stock convolutions)."""
    class BasicBlock(nn.Module):
        def __init__(self, c, down):
            super().__init__()
            self.conv1 = nn.Conv2d(c, c, 3, stride=2 if down else 1, padding=1, bias=False)
            self.bn1 = nn.BatchNorm2d(c)
            self.relu = nn.ReLU(inplace=True)
            self.conv2 = nn.Conv2d(c, c, 3, padding=1, bias=False)
            self.bn2 = nn.BatchNorm2d(c)
            self.downsample = nn.Sequential(nn.Conv2d(c, c, 1, stride=2, bias=False), nn.BatchNorm2d(c)) if down else None

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
            self.conv1 = nn.Conv2d(c, c // 2, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(c // 2)
            self.conv2 = nn.Conv2d(c // 2, c // 2, 3, stride=2 if down else 1, padding=1, bias=False)
            self.bn2 = nn.BatchNorm2d(c // 2)
            self.conv3 = nn.Conv2d(c // 2, c, 1, bias=False)
            self.bn3 = nn.BatchNorm2d(c)
            self.relu = nn.ReLU(inplace=True)
            self.downsample = nn.Sequential(nn.Conv2d(c, c, 1, stride=2, bias=False), nn.BatchNorm2d(c)) if down else None

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


def test_fuse_resnet_blocks_keeps_names_and_results():
    """The rewritten block forward (ReLU and residual passed into the batch-norm modules) equals the reference's
    forward code; on the CPU the modules run stock torch, so bit for bit."""
    from cpg_b200.fused_norm import fuse_resnet_blocks
    BasicBlock, Bottleneck = _blocks()
    torch.manual_seed(3)

    def net():
        return nn.Sequential(BasicBlock(8, False), BasicBlock(8, True), Bottleneck(8, False), Bottleneck(8, True))
    ref, new = net(), net()
    new.load_state_dict(ref.state_dict())
    keys = list(new.state_dict().keys())
    fuse_bn_relu(new[0])                                   # blocks already converted by fuse_bn_relu are taken as well
    assert fuse_resnet_blocks(new) == 4
    assert fuse_resnet_blocks(new) == 0                    # idempotent
    assert list(new.state_dict().keys()) == keys
    assert all(isinstance(b, (BasicBlock, Bottleneck)) for b in new)     # still the reference's classes (subclassed)
    import copy
    assert type(copy.deepcopy(new)[2]) is type(new[2])
    assert all(isinstance(m, FusedBatchNormReLU2d) for b in new for m in (b.bn1, b.bn2))
    assert all(isinstance(b.downsample[1], nn.BatchNorm2d) for b in new if b.downsample is not None)
    x = torch.randn(4, 8, 12, 12)
    for train in (True, False):
        ref.train(train); new.train(train)
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ya, yb = ref(xa), new(xb)
        assert torch.equal(ya, yb)
        (ya * ya).sum().backward(); (yb * yb).sum().backward()
        assert torch.equal(xa.grad, xb.grad)
        for pa, pb in zip(ref.parameters(), new.parameters()):
            assert torch.equal(pa.grad, pb.grad)
    for (ka, va), (kb, vb) in zip(ref.state_dict().items(), new.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka


def test_fuse_resnet_blocks_skips_what_it_does_not_know():
    from cpg_b200.fused_norm import fuse_resnet_blocks

    class Bottleneck(nn.Module):                            # same name, different surface (no bn3 / relu attribute)
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(4, 4, 1)
            self.bn1 = nn.GroupNorm(2, 4)

        def forward(self, x):
            return self.bn1(self.conv1(x))

    m = nn.Sequential(Bottleneck())
    assert fuse_resnet_blocks(m) == 0 and type(m[0]) is Bottleneck
