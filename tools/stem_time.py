"""Device time of the stem (3->64, 3x3 @32x32, batch 128) through the C ABI: direct kernels (AUTO)
against the tensor-core explicit-im2col tier (PATH_TCGEN05) and the CUDA-core path."""
import statistics
import sys
import os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpg_b200 import _lib

lib = _lib.load()
dev = 'cuda:0'
N, HW, K = 128, 32, 64
x = torch.randn(N, 3, HW, HW, device=dev).contiguous(memory_format=torch.channels_last)
xp = torch.empty(N, 4, HW, HW, device=dev).contiguous(memory_format=torch.channels_last)
xp[:, :3].copy_(x)
xv = xp[:, :3]
w = torch.randn(K, 3, 3, 3, device=dev) * 0.1
y = torch.empty(N, K, HW, HW, device=dev).contiguous(memory_format=torch.channels_last)
dy = torch.randn_like(y)
t = torch.ones(w.shape, dtype=torch.uint8, device=dev)
dW = torch.empty_like(w)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
P = _lib.ptr


def timeit(fn, iters=7):
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        torch.cuda._sleep(200000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


for name, path, xin in (('direct-vec4', _lib.PATH_AUTO, xv), ('direct-scalar', _lib.PATH_AUTO, x), ('tcgen05-xcol', _lib.PATH_TCGEN05, xv), ('simt', _lib.PATH_SIMT, x)):
    _lib.set_path(path)
    d = _lib.conv_desc(xin.shape, xin.stride(), w.shape, y.shape, y.stride(), (1, 1), (1, 1), (1, 1), 1)
    ws = torch.empty(max(lib.cpgb_workspace_bytes(d), 256), dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr()
    try:
        f = timeit(lambda: _lib.check(lib.cpgb_conv2d_fprop(d, P(xin), P(w), None, None, P(y), 5e-3, None, P(ws), ws.numel(), st), 'f'))
        g = timeit(lambda: _lib.check(lib.cpgb_conv2d_wgrad_fused(d, P(xin), P(dy), P(w), None, P(t), 1, 4e-5, 1, P(dW), None, None,
                                                                 5e-3, P(ws), ws.numel(), st), 'w'))
        print(f'{name:14s} fprop {f:7.1f} us   wgrad {g:7.1f} us   (Y / dY = {y.numel() * 4 / 1e6:.1f} MB -> HBM floor {y.numel() * 4 / 6.5e6:.1f} us)')
    except Exception as ex:
        print(name, 'failed:', ex)
_lib.set_path(_lib.PATH_AUTO)
