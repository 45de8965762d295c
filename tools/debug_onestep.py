"""diagnostic: per-parameter gradient difference TC (AUTO) vs SIMT for one fwd+bwd"""
import torch, torch.nn as nn
from cpg_b200 import _lib
import cpg_b200.layers as nl
from tests.trajectory import build

DEV = 'cuda:0'
def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)

outs = {}
for path in (_lib.PATH_SIMT, _lib.PATH_AUTO):
    _lib.set_path(path)
    model, masks, loader = build(nl.SharableConv2d, nl.SharableLinear, DEV, width=0.5, batch=16)
    model.train()
    acts = {}
    hooks = []
    for n, m in model.named_modules():
        if isinstance(m, (nl.SharableConv2d, nl.SharableLinear, nn.BatchNorm2d)):
            hooks.append(m.register_forward_hook(lambda mod, i, o, n=n: acts.__setitem__(n, o.detach().clone())))
    data, target = loader[0]
    loss = nn.CrossEntropyLoss()(model(data.to(DEV)), target.to(DEV))
    loss.backward()
    outs[path] = (acts, {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}, loss.item())
a0, g0, l0 = outs[_lib.PATH_SIMT]; a1, g1, l1 = outs[_lib.PATH_AUTO]
print('loss', l0, l1)
for n in a0:
    print(f'act  {n:28s} rel {rel(a1[n], a0[n]):.3e}  shape {tuple(a0[n].shape)}')
for n in g0:
    print(f'grad {n:28s} rel {rel(g1[n], g0[n]):.3e}  proj {((g1[n].double()*g0[n].double()).sum()/(g0[n].double()**2).sum()).item()-1:+.3e}')
