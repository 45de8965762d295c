"""Layer parity at the shapes of a half-width VGG16 (batch 16 / 32) against F.conv2d: which plan is wrong?"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpg_b200.layers as nl  # noqa: E402

dev = 'cuda:0'
torch.backends.cudnn.allow_tf32 = False
for N in (16, 32, 128):
    for (C, K, HW) in ((32, 32, 32), (32, 64, 16), (64, 64, 16), (64, 128, 8), (128, 128, 8), (128, 256, 4), (256, 256, 4),
                       (256, 256, 2), (512, 512, 2), (512, 512, 4)):
        torch.manual_seed(C + K + HW)
        m = nl.SharableConv2d(C, K, 3, padding=1, bias=False).to(dev)
        x = torch.randn(N, C, HW, HW, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = m(x)
        dy = torch.randn_like(y)
        y.backward(dy)
        xr = x.detach().clone().requires_grad_(True)
        yr = F.conv2d(xr, m.weight.detach(), None, 1, 1)
        yr.backward(dy)
        e = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
        print('N %3d conv %3d->%3d @%2d  y %.2e  dx %.2e  dW %.2e' % (N, C, K, HW, e(y, yr), e(x.grad, xr.grad),
                                                                     e(m.weight.grad, torch.autograd.grad(F.conv2d(x.detach(), m.weight, None, 1, 1), m.weight, dy)[0])), flush=True)
