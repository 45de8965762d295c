"""Small-shape pass over the round-2 kernels for compute-sanitizer (memcheck, racecheck on the shared-memory ones):
  compute-sanitizer --tool memcheck python tools/memcheck_r2.py
two-pass prune select (ties, tiny pool, exit-2, ragged sizes), fused optimizers (ragged tails, packed words), cluster
batch-norm kernels (train / eval, pooled, padded channels), PReLU, two-phase bias gradient, wgrad with the epilogue on a
second stream, a biased convolution + linear layer end to end."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpg_b200 import _lib  # noqa: E402
import cpg_b200.layers as nl  # noqa: E402
from cpg_b200.fused_norm import FusedBatchNormReLU2d, FusedPReLU  # noqa: E402
from cpg_b200.optim import SGD, Adam  # noqa: E402

DEV = 'cuda:0'
lib = _lib.load()
torch.manual_seed(0)
rng = np.random.RandomState(0)

# a7, two-pass select
sizes = [1, 7, 1000, 4099, 36867, 262145]
ws_ = [torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(DEV) for n in sizes]
ws_[3][::3] = 0.0
ts_ = [torch.from_numpy(rng.randint(0, 4, size=n).astype(np.uint8)).to(DEV) for n in sizes]
ts_[1][:] = 1
n_ = len(sizes)
W = (ctypes.c_void_p * n_)(*[t.data_ptr() for t in ws_])
T = (ctypes.c_void_p * n_)(*[t.data_ptr() for t in ts_])
N = (ctypes.c_int64 * n_)(*sizes)
info = torch.zeros(n_, 4, dtype=torch.int64, device=DEV)
wsb = torch.empty(lib.cpgb_prune_sampled_workspace_bytes(n_), dtype=torch.uint8, device=DEV)
for ratio in (0.0015, 0.37, 1.0):
    _lib.check(lib.cpgb_prune_select_sampled(n_, W, T, N, 2, ratio, info.data_ptr(), wsb.data_ptr(), wsb.numel(),
                                             _lib.stream_ptr()), 'sampled')
torch.cuda.synchronize()

# N2 optimizers: ragged tails, packed words
ps = [nn.Parameter(torch.randn(*s, device=DEV) * 0.01) for s in ((4096, 8), (33, 5, 3, 3), (7,), (4099,))]
for p in ps:
    p.grad = torch.randn_like(p) * 0.01
sgd, adam = SGD(ps[:2], lr=1e-2), Adam(ps[2:] + [ps[1]], lr=5e-4,
                                         pack={ps[3]: (torch.zeros((4099 + 31) // 32, dtype=torch.int64, device=DEV), None)})
for _ in range(2):
    sgd.step(); adam.step()
torch.cuda.synchronize()

# BN (cluster kernels for these sizes), PReLU
for (Nn, C, H, Wd, pool) in ((2, 8, 6, 6, True), (2, 64, 6, 10, True), (3, 78, 5, 7, False), (2, 1028, 2, 2, True)):
    for train in (True, False):
        bn = FusedBatchNormReLU2d(C, relu=True, pool=pool).to(DEV).train(train)
        x = torch.randn(Nn, C, H, Wd, device=DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = bn(x)
        y.backward(torch.randn_like(y))
    pr = FusedPReLU(C, tf32_out=True).to(DEV)
    x = torch.randn(Nn, C, H, Wd, device=DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = pr(x)
    y.backward(torch.randn_like(y))
torch.cuda.synchronize()

# biased convolution + linear through the module API (bias gradient scratch, epilogue stream inside backward())
for (Nn, C, K, HW) in ((4, 32, 64, 8), (2, 78, 156, 6), (8, 64, 64, 16)):
    m = nl.SharableConv2d(C, K, 3, padding=1, bias=True).to(DEV)
    m.piggymask = nn.Parameter(torch.rand_like(m.weight) * 0.01)
    x = torch.randn(Nn, C, HW, HW, device=DEV).requires_grad_(True)
    m(x).square().mean().backward()
lin = nl.SharableLinear(512, 1024).to(DEV)
lin.piggymask = nn.Parameter(torch.rand_like(lin.weight) * 0.01)
x = torch.randn(128, 512, device=DEV, requires_grad=True)
lin(x).square().mean().backward()
torch.cuda.synchronize()
print('memcheck workload done')
