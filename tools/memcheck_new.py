"""Small-shape pass over the round-1 additions (stem kernels, BatchNorm+ReLU+pool kernels) for
compute-sanitizer:  compute-sanitizer --tool memcheck python tools/memcheck_new.py"""
import os
import sys
import torch
import torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpg_b200.layers as nl
from cpg_b200.fused_norm import FusedBatchNormReLU2d

DEV = 'cuda:0'
torch.manual_seed(0)
for (N, H, W, K, stride, pad, dil, bias) in ((2, 8, 8, 64, 1, 1, 1, False), (3, 7, 5, 8, 1, 1, 1, True),
                                             (2, 12, 12, 64, 2, 1, 1, True), (1, 9, 9, 128, 1, 2, 2, False)):
    m = nl.SharableConv2d(3, K, 3, stride=stride, padding=pad, dilation=dil, bias=bias).to(DEV)
    with torch.no_grad():
        for p in m.parameters():
            p.normal_(0, 0.1)
    m.piggymask = nn.Parameter(torch.rand_like(m.weight) * 0.01)
    y = m(torch.randn(N, 3, H, W, device=DEV))
    y.backward(torch.randn_like(y))
for (N, C, H, W, pool) in ((2, 8, 6, 6, True), (2, 64, 6, 10, True), (3, 12, 5, 7, False), (2, 1028, 2, 2, True)):
    for train in (True, False):
        bn = FusedBatchNormReLU2d(C, relu=True, pool=pool).to(DEV).train(train)
        x = torch.randn(N, C, H, W, device=DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = bn(x)
        y.backward(torch.randn_like(y))
torch.cuda.synchronize()
print('memcheck workload done')
