"""Eager training steps of a bench_workloads model (resnet50 / spherenet20) for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/prof_workload.py resnet50 task1 2"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_workloads as bw  # noqa: E402


def main():
    workload, regime, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
    dev = torch.device('cuda:0')
    net, masks, pruner, opts, shape, batch, classes, gflop = bw._build(workload, dev, regime)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, *shape, generator=g).to(dev)
    t = torch.randint(0, classes, (batch,), generator=g).to(dev)
    crit = nn.CrossEntropyLoss()
    for i in range(steps + 1):
        if i == steps:
            torch.cuda.synchronize()
            torch.cuda.profiler.start() if hasattr(torch.cuda, 'profiler') else None
        for o in opts:
            o.zero_grad(set_to_none=True)
        loss = crit(net(x), t)
        loss.backward()
        pruner.do_weight_decay_and_make_grads_zero()
        for o in opts:
            o.step()
    torch.cuda.synchronize()
    print('loss', float(loss))


if __name__ == '__main__':
    main()
