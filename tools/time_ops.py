"""Device-time of single hot-path calls (diagnostic; GPU box only).
usage: python -m tests.time_ops"""
import statistics
import torch
from cpg_b200 import _lib

DEV = 'cuda:0'


def timeit(fn, iters=7):
    flush = timeit.flush
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


def main():
    lib = _lib.load()
    timeit.flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    P, st = _lib.ptr, _lib.stream_ptr()
    print('empty event pair: %.1f us' % timeit(lambda: None))
    for name, N, C, H, W, K, R, pad in [('fc2', 128, 4096, 1, 1, 4096, 1, 0), ('fc2_n32', 32, 4096, 1, 1, 4096, 1, 0),
                                         ('fc1', 128, 512, 1, 1, 4096, 1, 0),
                                         ('conv256@8', 128, 256, 8, 8, 256, 3, 1), ('conv512@2', 128, 512, 2, 2, 512, 3, 1),
                                         ('conv64@32', 128, 64, 32, 32, 64, 3, 1)]:
        x = torch.randn(N, C, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
        w = torch.randn(K, C, R, R, device=DEV) * 0.05
        y = torch.empty(N, K, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
        dy = torch.randn_like(y)
        dx = torch.empty_like(x)
        t = torch.ones(w.shape, dtype=torch.uint8, device=DEV)
        dW = torch.empty_like(w)
        d = _lib.conv_desc(x.shape, x.stride(), w.shape, y.shape, y.stride(), (1, 1), (pad, pad), (1, 1), 1)
        ws = torch.empty(lib.cpgb_workspace_bytes(d), dtype=torch.uint8, device=DEV)
        nst = lib.cpgb_staged_weight_bytes(d)
        staged = torch.empty(nst, dtype=torch.uint8, device=DEV)
        _lib.check(lib.cpgb_stage_weights(d, P(w), None, 5e-3, P(staged), nst, st), 's')
        tf = timeit(lambda: _lib.check(lib.cpgb_conv2d_fprop(d, P(x), P(w), None, None, P(y), 5e-3, P(staged), P(ws), ws.numel(), st), 'f'))
        td = timeit(lambda: _lib.check(lib.cpgb_conv2d_dgrad(d, P(dy), P(w), None, P(dx), 5e-3, P(staged), P(ws), ws.numel(), st), 'd'))
        tw = timeit(lambda: _lib.check(lib.cpgb_conv2d_wgrad_fused(d, P(x), P(dy), P(w), None, P(t), 1, 4e-5, _lib.GRAD_FINETUNE, P(dW), None, None, 5e-3, P(ws), ws.numel(), st), 'w'))
        print(f'{name:10s} fprop {tf:7.1f} us  dgrad {td:7.1f} us  wgrad {tw:7.1f} us', flush=True)


if __name__ == '__main__':
    main()
