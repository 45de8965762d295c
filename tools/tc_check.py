"""On-GPU sweep of the tcgen05 path against torch fp32 (cuDNN, TF32 off) -- a diagnostic, not a
pytest file.  usage: python -m tests.tc_check [fprop|dgrad|wgrad|all] [quick]
Each op group should run in its own process: a trapped kernel poisons the CUDA context."""
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

from cpg_b200 import _lib
import cpg_b200.layers as nl

DEV = 'cuda:0'

# (N, C, H, W, K, R, pad, dil[, stride])
CASES = [
    (2, 32, 8, 8, 64, 3, 1, 1),
    (4, 64, 32, 32, 64, 3, 1, 1),
    (8, 64, 16, 16, 128, 3, 1, 1),
    (8, 128, 8, 8, 256, 3, 1, 1),
    (16, 256, 4, 4, 512, 3, 1, 1),
    (32, 512, 2, 2, 512, 3, 1, 1),
    (3, 96, 7, 7, 160, 3, 1, 1),
    (2, 64, 14, 14, 64, 1, 0, 1),
    (2, 32, 12, 12, 32, 3, 2, 2),
    (2, 36, 9, 9, 40, 3, 1, 1),
    (128, 512, 1, 1, 4096, 1, 0, 1),
    (128, 4096, 1, 1, 4096, 1, 0, 1),
    (128, 64, 32, 32, 64, 3, 1, 1),
    (128, 512, 4, 4, 512, 3, 1, 1),
    (128, 512, 2, 2, 512, 3, 1, 1),
    (128, 3, 32, 32, 64, 3, 1, 1),
    (5, 3, 20, 20, 32, 3, 1, 1),
    # explicit-im2col tier: strided layers and stems
    (4, 64, 16, 16, 128, 3, 1, 1, 2),
    (4, 64, 16, 16, 128, 1, 0, 1, 2),
    (3, 3, 33, 33, 64, 7, 3, 1, 2),
    (6, 128, 14, 14, 256, 3, 1, 1, 2),
    (2, 8, 9, 9, 32, 3, 1, 2, 1),
]


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def proj(a, b):
    """<a,b>/<b,b> - 1: the systematic scale error of a against b"""
    a, b = a.detach().double(), b.detach().double()
    return ((a * b).sum() / (b * b).sum()).item() - 1.0


def bad_frac(a, b):
    a, b = a.detach().double(), b.detach().double()
    tol = 1e-2 * b.abs().max().item()
    return ((a - b).abs() > tol).double().mean().item()


def run(op, cases):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = _lib.load()
    worst = 0.0
    # linear / 1x1 cases are run twice: with a piggymask (staged operand) and without (raw weights as operand)
    cases = [c + ((1,) if len(c) == 8 else ()) + (True,) for c in cases] + \
            [c + ((1,) if len(c) == 8 else ()) + (False,) for c in cases if c[5] == 1]
    for case in cases:
        (N, C, H, W, K, R, pad, dil), stride, with_piggy = case[:8], case[8], case[9]
        torch.manual_seed(N * 7 + C + K)
        m = nl.SharableConv2d(C, K, R, stride=stride, padding=pad, dilation=dil, bias=True).to(DEV)
        with torch.no_grad():
            m.weight.normal_(0, (2.0 / (C * R * R)) ** 0.5)
            m.bias.normal_()
        m.piggymask = nn.Parameter(torch.rand_like(m.weight) * 0.01) if with_piggy else None
        x = torch.randn(N, C, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
        if C % 4:
            xp = torch.empty((N, (C + 3) // 4 * 4, H, W), device=DEV).contiguous(memory_format=torch.channels_last)
            xp.fill_(float('nan'))      # pad lanes must never be read
            xv = xp[:, :C]
            xv.copy_(x)
            x = xv
        weff = ((m.piggymask > 5e-3).float() * m.weight).detach() if with_piggy else m.weight.detach()
        yr = F.conv2d(x, weff, m.bias, stride, pad, dil)
        dy = torch.randn_like(yr).contiguous(memory_format=torch.channels_last)
        d = _lib.conv_desc(x.shape, x.stride(), m.weight.shape, dy.shape, dy.stride(), (stride, stride), (pad, pad),
                           (dil, dil), 1)
        ws = torch.empty(max(lib.cpgb_workspace_bytes(d), 256), dtype=torch.uint8, device=DEV)
        P, st = _lib.ptr, _lib.stream_ptr()
        tag = f'N{N} C{C} {H}x{W} K{K} R{R} p{pad} d{dil} s{stride}' + ('' if with_piggy else ' nopiggy')
        _lib.set_path(_lib.PATH_TCGEN05)
        try:
            t0 = time.time()
            if op == 'fprop':
                y = torch.full_like(dy, float('nan'))
                _lib.check(lib.cpgb_conv2d_fprop(d, P(x), P(m.weight), P(m.piggymask), P(m.bias), P(y), 5e-3, None,
                                                 P(ws), ws.numel(), st), 'fprop')
                torch.cuda.synchronize()
                e, bf = rel(y - m.bias.view(1, -1, 1, 1), yr - m.bias.view(1, -1, 1, 1)), proj(y - m.bias.view(1, -1, 1, 1), yr - m.bias.view(1, -1, 1, 1))
            elif op == 'dgrad':
                dxr = torch.nn.grad.conv2d_input(x.shape, weff, dy, stride, pad, dil)
                dx = torch.full_like(x, float('nan')) if C % 4 == 0 else torch.empty_strided(x.shape, x.stride(), device=DEV)
                _lib.check(lib.cpgb_conv2d_dgrad(d, P(dy), P(m.weight), P(m.piggymask), P(dx), 5e-3, None, P(ws),
                                                 ws.numel(), st), 'dgrad')
                torch.cuda.synchronize()
                e, bf = rel(dx, dxr), proj(dx, dxr)
            else:
                gr = torch.nn.grad.conv2d_weight(x, m.weight.shape, dy, stride, pad, dil)
                dW = torch.full_like(m.weight, float('nan'))
                dP = torch.full_like(m.weight, float('nan')) if with_piggy else None
                tm = torch.randint(0, 4, m.weight.shape, device=DEV, dtype=torch.uint8)
                cur, wd = 3, 0.05
                _lib.check(lib.cpgb_conv2d_wgrad_fused(d, P(x), P(dy), P(m.weight), P(m.piggymask), P(tm), cur, wd,
                                                       _lib.GRAD_FINETUNE, P(dW), P(dP), None, 5e-3, P(ws), ws.numel(),
                                                       st), 'wgrad')
                torch.cuda.synchronize()
                b = (m.piggymask > 5e-3).float() if with_piggy else 1.0
                rW = (gr * b + wd * m.weight) * (tm == cur)
                rP = gr * m.weight * ((tm >= 1) & (tm < cur))
                e = max(rel(dW, rW), rel(dP, rP)) if with_piggy else rel(dW, rW)
                bf = proj(dP, rP) if with_piggy else proj(dW, rW)
            print(f'{op:6s} {tag:46s} rel {e:.3e}  proj {bf:+.3e}  {"OK" if e <= 1e-3 else "FAIL"} '
                  f'({(time.time() - t0) * 1e3:.1f} ms)', flush=True)
            worst = max(worst, e)
        except Exception as ex:  # noqa: BLE001
            print(f'{op:6s} {tag:46s} EXCEPTION {type(ex).__name__}: {ex}', flush=True)
            if 'CUDA' in str(ex) or 'cuda' in str(ex):
                print('context is gone; stopping', flush=True)
                return
        finally:
            _lib.set_path(_lib.PATH_AUTO)
    print(f'{op}: worst rel {worst:.3e}', flush=True)


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    cases = CASES[:3] if len(sys.argv) > 2 else CASES
    for op in (['fprop', 'dgrad', 'wgrad'] if which == 'all' else [which]):
        run(op, cases)
