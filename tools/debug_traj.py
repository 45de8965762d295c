"""Lock-step trajectory debugger (run on the GPU box): the product layers/pruner on cuda:0
against the oracle modules on the CPU, same inputs, diff after every phase of every step.
Not a pytest file; usage: python -m tests.debug_traj [prune|finetune] [simt|auto]"""
import sys

import numpy as np
import torch
import torch.nn as nn

from tests.trajectory import build, make_args, Wrap


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def main(mode='prune', path='simt'):
    import cpg_b200.layers as nl
    from cpg_b200 import _lib
    from cpg_b200.prune import SparsePruner
    from oracle import cpg_oracle as O
    if path == 'simt':
        _lib.set_path(_lib.PATH_SIMT)
    torch.backends.cudnn.allow_tf32 = False
    dev = 'cuda:0'
    mp, masks_p, loader = build(nl.SharableConv2d, nl.SharableLinear, dev)
    mo, masks_o, _ = build(O.OracleSharableConv2d, O.OracleSharableLinear, 'cpu')
    args = make_args(mode)
    netp = Wrap(mp)
    masks_pp = {'module.' + n: v for n, v in masks_p.items()}
    pr_p = SparsePruner(netp, masks_pp, args, 0, 4, 2)
    pr_o = O.OraclePruner(mo, masks_o, mode=mode, weight_decay=args.weight_decay, cur=2, inference_idx=2,
                          begin_prune_step=0, end_prune_step=4, initial_sparsity=0.0, target_sparsity=0.3,
                          pruning_frequency=2)

    def opts(model):
        sgd = [p for n, p in model.named_parameters() if 'piggymask' not in n and ('classifiers' not in n or '.1.' in n)]
        adam = [p for n, p in model.named_parameters() if 'piggymask' in n]
        return (torch.optim.SGD(sgd, lr=1e-2, weight_decay=0.0, momentum=0.9, nesterov=True),
                torch.optim.Adam(adam, lr=5e-4))
    op, oo = opts(mp), opts(mo)
    crit = nn.CrossEntropyLoss()
    mp.train(); mo.train()

    def cmp(tag, what):
        worst, wname = 0.0, ''
        po, pp = dict(mo.named_parameters()), dict(mp.named_parameters())
        for n in po:
            a, b = (pp[n], po[n]) if what == 'param' else (pp[n].grad, po[n].grad)
            if a is None or b is None:
                if (a is None) != (b is None):
                    print(f'   {tag}: {n} None mismatch {a is None} {b is None}')
                continue
            r = rel(a, b)
            if r > worst:
                worst, wname = r, n
        print(f'  {tag:28s} worst rel {worst:.3e}  ({wname})')

    step = 0
    for i, (data, target) in enumerate(loader):
        print(f'step {i}')
        for o in (*op, *oo):
            o.zero_grad()
        yp = netp(data.to(dev)); yo = mo(data)
        print(f'  output rel {rel(yp, yo):.3e}')
        lp = crit(yp, target.to(dev)); lo = crit(yo, target)
        lp.backward(); lo.backward()
        pr_p.do_weight_decay_and_make_grads_zero(); pr_o.do_weight_decay_and_make_grads_zero()
        cmp('grads after a6', 'grad')
        for o in (*op, *oo):
            o.step()
        cmp('params after step', 'param')
        if mode == 'prune':
            pr_p.gradually_prune(step); pr_o.gradually_prune(step)
            step += 1
            nd = sum(int((masks_p[n].cpu() != masks_o[n]).sum()) for n in masks_o)
            print(f'  mask elements differing: {nd}')


if __name__ == '__main__':
    main(*(sys.argv[1:3]))
