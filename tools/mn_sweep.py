"""Bring-up sweep of the MN-major operand descriptor fields (diagnostic; GPU box only)."""
import ctypes
import torch
import torch.nn as nn
import torch.nn.functional as F
from cpg_b200 import _lib
import cpg_b200.layers as nl

DEV = 'cuda:0'
VARIANTS = [  # layout, lbo, sbo, kadv, tma swizzle enum (3 = 128B, 4 = 128B_ATOM_32B)
    (1, 4096, 512, 1024, 4), (1, 512, 4096, 1024, 4), (1, 4096, 1024, 1024, 4), (1, 4096, 256, 1024, 4),
    (1, 4096, 512, 1024, 3), (2, 4096, 1024, 1024, 3), (2, 1024, 4096, 1024, 3), (2, 4096, 512, 1024, 4),
    (1, 256, 512, 1024, 4), (1, 512, 512, 1024, 4),
]


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def main():
    torch.backends.cudnn.allow_tf32 = False
    lib = _lib.load()
    lib.cpgb_debug_set_mn.argtypes = [ctypes.c_int] * 5
    lib.cpgb_debug_set_mn.restype = None
    N, C, H, W, K = 4, 64, 8, 8, 128
    torch.manual_seed(0)
    m = nl.SharableConv2d(C, K, 3, padding=1, bias=False).to(DEV)
    with torch.no_grad():
        m.weight.normal_(0, 0.05)
    x = torch.randn(N, C, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
    yr = F.conv2d(x, m.weight, None, 1, 1)
    dy = torch.randn_like(yr).contiguous(memory_format=torch.channels_last)
    dxr = torch.nn.grad.conv2d_input(x.shape, m.weight, dy, 1, 1)
    gr = torch.nn.grad.conv2d_weight(x, m.weight.shape, dy, 1, 1)
    d = _lib.conv_desc(x.shape, x.stride(), m.weight.shape, dy.shape, dy.stride(), (1, 1), (1, 1), (1, 1), 1)
    ws = torch.empty(max(lib.cpgb_workspace_bytes(d), 256), dtype=torch.uint8, device=DEV)
    P, st = _lib.ptr, _lib.stream_ptr()
    _lib.set_path(_lib.PATH_TCGEN05)
    for v in VARIANTS:
        lib.cpgb_debug_set_mn(*v)
        dx = torch.zeros_like(x)
        _lib.check(lib.cpgb_conv2d_dgrad(d, P(dy), P(m.weight), None, P(dx), 5e-3, None, P(ws), ws.numel(), st), 'dgrad')
        dW = torch.zeros_like(m.weight)
        _lib.check(lib.cpgb_conv2d_wgrad_fused(d, P(x), P(dy), P(m.weight), None, None, 0, 0.0, _lib.GRAD_RAW, P(dW),
                                               None, None, 5e-3, P(ws), ws.numel(), st), 'wgrad')
        torch.cuda.synchronize()
        print(f'variant {v}: dgrad rel {rel(dx, dxr):.3e}   wgrad rel {rel(dW, gr):.3e}', flush=True)


if __name__ == '__main__':
    main()
