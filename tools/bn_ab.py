"""A/B of the two batch-norm implementations behind cpgb_bn_relu_fwd / _bwd (single-launch cluster kernels vs
stats -> finalize -> apply) per VGG16 layer shape at batch 128, inside CUDA graphs (launch latencies as in the training
step).  Each graph holds REP x [producer copy (leaves x / dy in L2, like the convolution before), BN call]; the
copy-only graph is subtracted.  usage: python tools/bn_ab.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpg_b200.fused_norm import FusedBatchNormReLU2d  # noqa: E402

DEV = 'cuda:0'
REP = 10


def graph_ms(body):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            body()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 / REP * 1e3      # us per body


def main():
    shapes = [(64, 32, False), (64, 32, True), (128, 16, False), (128, 16, True), (256, 8, False), (256, 8, True),
              (512, 4, False), (512, 4, True), (512, 2, False), (512, 2, True)]
    if len(sys.argv) > 1:
        shapes = shapes[int(sys.argv[1]):]
    tot = {'1': [0.0, 0.0], '0': [0.0, 0.0]}
    for C, HW, pool in shapes:
        src = torch.randn(128, C, HW, HW, device=DEV).contiguous(memory_format=torch.channels_last)
        x = torch.empty_like(src).requires_grad_(True)
        ho = HW // 2 if pool else HW
        dsrc = torch.randn(128, C, ho, ho, device=DEV).contiguous(memory_format=torch.channels_last)
        dy = torch.empty_like(dsrc)
        mod = FusedBatchNormReLU2d(C, relu=True, pool=pool, tf32_out=True).to(DEV)
        line = 'C%-4d @%-2d pool=%d (%5.1f MB)' % (C, HW, pool, src.numel() * 4 / 1e6)
        with torch.no_grad():
            base_f = graph_ms(lambda: x.copy_(src))
        base_b = graph_ms(lambda: dy.copy_(dsrc))
        for mode in ('1', '0'):
            os.environ['CPGB_BN_CLUSTER'] = mode

            def fwd():
                with torch.no_grad():
                    x.copy_(src)
                return mod(x)
            tf = graph_ms(fwd) - base_f

            def fwd_bwd():
                x.grad = None
                y = fwd()
                dy.copy_(dsrc)
                y.backward(dy)
            tb = graph_ms(fwd_bwd) - base_f - base_b - tf
            tot[mode][0] += tf
            tot[mode][1] += tb
            line += '  | %s fwd %6.1f bwd %6.1f' % ('cluster' if mode == '1' else '3-kern ', tf, tb)
        print(line, flush=True)
    print('sum over these 10 shapes: cluster fwd %.1f bwd %.1f | 3-kernel fwd %.1f bwd %.1f us' %
          (tot['1'][0], tot['1'][1], tot['0'][0], tot['0'][1]))


if __name__ == '__main__':
    main()
