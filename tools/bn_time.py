"""Device time of BatchNorm2d + ReLU, forward and backward: stock torch modules vs cpg_b200.fused_norm."""
import os
import statistics
import sys
import torch
import torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpg_b200.fused_norm import FusedBatchNormReLU2d

DEV = 'cuda:0'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=5):
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        torch.cuda._sleep(400000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


for C, HW in ((64, 32), (128, 16), (256, 8), (512, 4), (512, 2)):
    x = torch.randn(128, C, HW, HW, device=DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    dy = torch.randn(128, C, HW, HW, device=DEV).contiguous(memory_format=torch.channels_last)
    mb = x.numel() * 4 / 1e6
    out = []
    for name, mod in (('torch', nn.Sequential(nn.BatchNorm2d(C), nn.ReLU(inplace=True)).to(DEV)),
                      ('fused', FusedBatchNormReLU2d(C, relu=True).to(DEV))):
        y = mod(x)
        tf = timeit(lambda: mod(x))
        y = mod(x)
        tb = timeit(lambda: y.backward(dy, retain_graph=True))
        out.append(f'{name}: fwd {tf:6.1f} bwd {tb:6.1f}')
    print(f'C{C}@{HW} ({mb:.1f} MB; floors fwd {3 * mb / 6.5e3 * 1e3 / 1e3:.1f} us bwd {5 * mb / 6.5e3:.1f} us)  ' + ' | '.join(out), flush=True)
