"""Launch a fixed set of hot-path ops a few times each (for ncu captures; GPU box only).
usage: python -m tests.prof_ops [reps]"""
import sys
import torch
from cpg_b200 import _lib

DEV = 'cuda:0'
# name, N, C, H, W, K, R, pad
OPS = [
    ('fc2', 128, 4096, 1, 1, 4096, 1, 0),
    ('conv128@16', 128, 128, 16, 16, 128, 3, 1),
    ('conv64@32', 128, 64, 32, 32, 64, 3, 1),
    ('conv512@4', 128, 512, 4, 4, 512, 3, 1),
    ('conv256@8', 128, 256, 8, 8, 256, 3, 1),
]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    lib = _lib.load()
    P, st = _lib.ptr, _lib.stream_ptr()
    for name, N, C, H, W, K, R, pad in OPS:
        x = torch.randn(N, C, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
        w = torch.randn(K, C, R, R, device=DEV) * 0.05
        y = torch.empty(N, K, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
        dy = torch.randn_like(y)
        dx = torch.empty_like(x)
        t = torch.ones(w.shape, dtype=torch.uint8, device=DEV)
        dW = torch.empty_like(w)
        d = _lib.conv_desc(x.shape, x.stride(), w.shape, y.shape, y.stride(), (1, 1), (pad, pad), (1, 1), 1)
        ws = torch.empty(lib.cpgb_workspace_bytes(d), dtype=torch.uint8, device=DEV)
        for _ in range(reps):
            torch.cuda.nvtx.range_push(name)
            _lib.check(lib.cpgb_conv2d_fprop(d, P(x), P(w), None, None, P(y), 5e-3, None, P(ws), ws.numel(), st), 'f')
            _lib.check(lib.cpgb_conv2d_dgrad(d, P(dy), P(w), None, P(dx), 5e-3, None, P(ws), ws.numel(), st), 'd')
            _lib.check(lib.cpgb_conv2d_wgrad_fused(d, P(x), P(dy), P(w), None, P(t), 1, 4e-5, _lib.GRAD_FINETUNE, P(dW),
                                                   None, None, 5e-3, P(ws), ws.numel(), st), 'w')
            torch.cuda.nvtx.range_pop()
        torch.cuda.synchronize()
    print('done')


if __name__ == '__main__':
    main()
