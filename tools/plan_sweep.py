"""Tile-width / split-K sweep of the fprop and dgrad GEMMs on the VGG16 layer shapes (diagnostic)."""
import os
import statistics
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpg_b200 import _lib

DEV = 'cuda:0'
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
P = _lib.ptr


def timeit(fn, iters=5):
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


shapes = [(128, 128, 16), (128, 256, 8), (256, 256, 8), (256, 512, 4), (512, 512, 4), (512, 512, 2)]
configs = [(None, None), (256, 1), (256, 2), (256, 3), (256, 4), (128, 1), (128, 2), (128, 3), (128, 6)]
for C, K, HW in shapes:
    x = torch.randn(128, C, HW, HW, device=DEV).contiguous(memory_format=torch.channels_last)
    w = torch.randn(K, C, 3, 3, device=DEV) * 0.05
    y = torch.empty(128, K, HW, HW, device=DEV).contiguous(memory_format=torch.channels_last)
    dy = torch.randn_like(y)
    dx = torch.empty_like(x)
    st = _lib.stream_ptr()
    out = []
    for bn, sp in configs:
        for k, v in (('CPGB_GEMM_BN', bn), ('CPGB_GEMM_SPLITS', sp)):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        d = _lib.conv_desc(x.shape, x.stride(), w.shape, y.shape, y.stride(), (1, 1), (1, 1), (1, 1), 1)
        ws = torch.empty(lib.cpgb_workspace_bytes(d), dtype=torch.uint8, device=DEV)
        nst = lib.cpgb_staged_weight_bytes(d)
        staged = torch.empty(nst, dtype=torch.uint8, device=DEV)
        _lib.check(lib.cpgb_stage_weights(d, P(w), None, 5e-3, P(staged), nst, st), 's')
        try:
            tf = timeit(lambda: _lib.check(lib.cpgb_conv2d_fprop(d, P(x), P(w), None, None, P(y), 5e-3, P(staged), P(ws), ws.numel(), st), 'f'))
            td = timeit(lambda: _lib.check(lib.cpgb_conv2d_dgrad(d, P(dy), P(w), None, P(dx), 5e-3, P(staged), P(ws), ws.numel(), st), 'd'))
            out.append(f'bn{bn}/s{sp}: {tf:5.1f} {td:5.1f}')
        except Exception as ex:
            out.append(f'bn{bn}/s{sp}: ERR {str(ex)[:40]}')
    print(f'conv{C}x{K}@{HW}  ' + ' | '.join(out), flush=True)
