"""One GPU: gradients of a batch-32 step vs the mean of two batch-16 steps on the same model (what data parallelism
computes), for kernel-variant switches given in the environment.  usage: python tools/shard_check.py"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from ddp_nccl_worker import build, grads_of, rel  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(5)
    data = torch.randn(32, 3, 32, 32, generator=g).to(dev)
    target = torch.randint(0, 5, (32,), generator=g).to(dev)
    crit = nn.CrossEntropyLoss()
    for regime in ('task2', 'task1'):
        net, masks, pruner = build(regime, dev)

        def step(x, t):
            for p in net.parameters():
                p.grad = None
            crit(net(x), t).backward()
            pruner.do_weight_decay_and_make_grads_zero()
            torch.cuda.synchronize()
            return grads_of(net)
        full = step(data, target)
        a, b = step(data[:16], target[:16]), step(data[16:], target[16:])
        bad = []
        for n in full:
            e = rel((a[n] + b[n]) / 2, full[n])
            if e > 2e-4:
                bad.append((n, '%.3g' % e))
        print(regime, 'bad:', bad[:6], '...' if len(bad) > 6 else '', len(bad), 'of', len(full), flush=True)


if __name__ == '__main__':
    main()
