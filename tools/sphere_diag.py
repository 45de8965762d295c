"""SphereNet-20 task-2 regime: per-step loss of the product path next to the reference's torch expressions on the same
weights / inputs (does a NaN come from the kernels or from the optimisation itself?).  Run on a GPU box:
python tools/sphere_diag.py [steps] [lr]"""
import copy
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import bench_workloads as bw  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    lr = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
    bw.LR_OF['spherenet20'] = lr
    dev = torch.device('cuda:0')
    from test_reference_manager_gpu import _torch_ops_layers
    import cpg_b200.layers as nl
    out = {}
    for arm in ('ours', 'torch'):
        net, masks, pruner, opts, shape, batch, classes, gflop = bw._build('spherenet20', dev, 'task2')
        g = torch.Generator().manual_seed(100)
        data = [(torch.randn(batch, *shape, generator=g).to(dev), torch.randint(0, classes, (batch,), generator=g).to(dev))
                for _ in range(4)]
        crit = nn.CrossEntropyLoss()
        losses = []

        def loop():
            for i in range(steps):
                for o in opts:
                    o.zero_grad(set_to_none=True)
                x, t = data[i % 4]
                y = net(x)
                loss = crit(y, t)
                loss.backward()
                pruner.do_weight_decay_and_make_grads_zero()
                gn = max(float(p.grad.abs().max()) for p in net.parameters() if p.grad is not None)
                for o in opts:
                    o.step()
                losses.append((float(loss), float(y.abs().max()), gn))
        if arm == 'torch':
            with _torch_ops_layers(nl):
                loop()
        else:
            loop()
        out[arm] = losses
    for i in range(steps):
        a, b = out['ours'][i], out['torch'][i]
        print('step %2d  ours loss %.5f |y| %.3e |g| %.3e   torch loss %.5f |y| %.3e |g| %.3e' % ((i,) + a + b))


if __name__ == '__main__':
    main()
