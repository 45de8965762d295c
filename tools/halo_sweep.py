"""Bring-up of the halo-mode wgrad (row-shifted MN-major descriptors): which base_offset rule works."""
import ctypes
import torch
import torch.nn.functional as F
from cpg_b200 import _lib

DEV = 'cuda:0'


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def main():
    torch.backends.cudnn.allow_tf32 = False
    lib = _lib.load()
    lib.cpgb_debug_set_mn.argtypes = [ctypes.c_int] * 5
    lib.cpgb_debug_set_mn.restype = None
    P, st = _lib.ptr, _lib.stream_ptr()
    _lib.set_path(_lib.PATH_TCGEN05)
    for (N, C, H, W, K, pad, dil) in [(4, 64, 32, 32, 64, 1, 1), (8, 128, 16, 16, 128, 1, 1), (16, 32, 8, 8, 64, 1, 1),
                                      (2, 32, 12, 12, 32, 2, 2), (3, 96, 7, 7, 160, 1, 1)]:
        torch.manual_seed(0)
        x = torch.randn(N, C, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
        w = torch.randn(K, C, 3, 3, device=DEV) * 0.05
        yr = F.conv2d(x, w, None, 1, pad, dil)
        dy = torch.randn_like(yr).contiguous(memory_format=torch.channels_last)
        gr = torch.nn.grad.conv2d_weight(x, w.shape, dy, 1, pad, dil)
        d = _lib.conv_desc(x.shape, x.stride(), w.shape, dy.shape, dy.stride(), (1, 1), (pad, pad), (dil, dil), 1)
        for mode, enable in [(0, 0), (0, 1), (1, 1), (2, 1)]:
            lib.cpgb_debug_set_mn(-1, mode, enable, 0, 0)
            ws = torch.empty(max(lib.cpgb_workspace_bytes(d), 256), dtype=torch.uint8, device=DEV)
            dW = torch.zeros_like(w)
            _lib.check(lib.cpgb_conv2d_wgrad_fused(d, P(x), P(dy), P(w), None, None, 0, 0.0, _lib.GRAD_RAW, P(dW), None,
                                                   None, 5e-3, P(ws), ws.numel(), st), 'wgrad')
            torch.cuda.synchronize()
            print(f'N{N} C{C} {H}x{W} K{K} p{pad} d{dil}  halo={enable} base_mode={mode}: rel {rel(dW, gr):.3e}', flush=True)


if __name__ == '__main__':
    main()
