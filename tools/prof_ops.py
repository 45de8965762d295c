"""Launch a fixed set of hot-path ops a few times each (for ncu captures; GPU box only).
usage: python -m tools.prof_ops [reps] [vgg]"""
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
# the 15 sharable layers of VGG16-BN-cifar at batch 128, in network order
VGG = [('conv%dx%d@%d' % (c, k, hw), 128, c, hw, hw, k, 3, 1) for (c, k, hw) in
       [(3, 64, 32), (64, 64, 32), (64, 128, 16), (128, 128, 16), (128, 256, 8), (256, 256, 8), (256, 256, 8),
        (256, 512, 4), (512, 512, 4), (512, 512, 4), (512, 512, 2), (512, 512, 2), (512, 512, 2)]] + \
      [('fc512x4096', 128, 512, 1, 1, 4096, 1, 0), ('fc4096x4096', 128, 4096, 1, 1, 4096, 1, 0)]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    ops = VGG if len(sys.argv) > 2 and sys.argv[2] == 'vgg' else OPS
    lib = _lib.load()
    P, st = _lib.ptr, _lib.stream_ptr()
    for name, N, C, H, W, K, R, pad in ops:
        x = torch.randn(N, C, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
        if C % 4:
            xp = torch.empty(N, 4, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
            xp[:, :C].copy_(x)
            x = xp[:, :C]
        w = torch.randn(K, C, R, R, device=DEV) * 0.05
        y = torch.empty(N, K, H, W, device=DEV).contiguous(memory_format=torch.channels_last)
        dy = torch.randn_like(y)
        dx = torch.empty_strided(x.shape, x.stride(), device=DEV)
        t = torch.ones(w.shape, dtype=torch.uint8, device=DEV)
        dW = torch.empty_like(w)
        d = _lib.conv_desc(x.shape, x.stride(), w.shape, y.shape, y.stride(), (1, 1), (pad, pad), (1, 1), 1)
        ws = torch.empty(lib.cpgb_workspace_bytes(d), dtype=torch.uint8, device=DEV)
        for _ in range(reps):
            torch.cuda.nvtx.range_push(name)
            _lib.check(lib.cpgb_conv2d_fprop(d, P(x), P(w), None, None, P(y), 5e-3, None, P(ws), ws.numel(), st), 'f')
            if C % 4 == 0:
                _lib.check(lib.cpgb_conv2d_dgrad(d, P(dy), P(w), None, P(dx), 5e-3, None, P(ws), ws.numel(), st), 'd')
            _lib.check(lib.cpgb_conv2d_wgrad_fused(d, P(x), P(dy), P(w), None, P(t), 1, 4e-5, _lib.GRAD_FINETUNE, P(dW),
                                                   None, None, 5e-3, P(ws), ws.numel(), st), 'w')
            torch.cuda.nvtx.range_pop()
        torch.cuda.synchronize()
    print('done')


if __name__ == '__main__':
    main()
