"""Small single-kernel cases for compute-sanitizer (memcheck) runs on the GPU box:
    compute-sanitizer --tool memcheck python tools/sanitize_cases.py <case>
cases: cluster_dgrad, cluster_fprop, cluster_fc, ragged_linear, ragged_conv"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpg_b200 import _lib  # noqa: E402
import cpg_b200.layers as nl  # noqa: E402

DEV = 'cuda:0'


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def conv_case(N, C, K, HW):
    torch.manual_seed(0)
    m = nl.SharableConv2d(C, K, 3, padding=1, bias=False).to(DEV)
    with torch.no_grad():
        m.weight.normal_(0, 0.05)
    x = torch.randn(N, C, HW, HW, device=DEV).requires_grad_(True)
    y = m(x)
    dy = torch.randn(N, K, HW, HW, device=DEV)
    y.backward(dy)
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    yr = torch.nn.functional.conv2d(x.detach(), m.weight.detach(), None, 1, 1)
    dxr = torch.nn.grad.conv2d_input(x.shape, m.weight.detach(), dy, 1, 1)
    gr = torch.nn.grad.conv2d_weight(x.detach(), m.weight.shape, dy, 1, 1)
    print('conv', (N, C, K, HW), 'y %.2e dx %.2e dW %.2e' % (rel(y, yr), rel(x.grad, dxr), rel(m.weight.grad, gr)), flush=True)


def linear_case(M, I, O):
    torch.manual_seed(0)
    m = nl.SharableLinear(I, O).to(DEV)
    with torch.no_grad():
        m.weight.normal_(0, 0.02)
        m.bias.zero_()
    x = torch.randn(M, I, device=DEV, requires_grad=True)
    y = m(x)
    torch.cuda.synchronize()
    print('linear fprop ok', flush=True)
    dy = torch.randn(M, O, device=DEV)
    y.backward(dy)
    torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = False
    print('linear', (M, I, O), 'y %.2e dx %.2e dW %.2e' % (rel(y, x.detach() @ m.weight.detach().t()),
                                                          rel(x.grad, dy @ m.weight.detach()),
                                                          rel(m.weight.grad, dy.t() @ x.detach())), flush=True)


case = sys.argv[1]
_lib.set_path(_lib.PATH_TCGEN05)
if case == 'cluster_dgrad':
    conv_case(128, 128, 256, 8)
elif case == 'cluster_fprop':
    conv_case(128, 256, 512, 4)
elif case == 'cluster_2x2':
    conv_case(128, 512, 512, 2)
elif case == 'cluster_fc':
    linear_case(128, 4096, 4096)
elif case == 'ragged_linear':
    linear_case(128, 627, 5016)
elif case == 'ragged_conv':
    conv_case(32, 78, 156, 16)
print('done', case)
