"""Toy 3-layer model mirroring tests/golden/make_golden.py::_Toy with injectable layer classes."""
import argparse

import numpy as np
import torch
import torch.nn as nn


class Toy(nn.Module):
    def __init__(self, nl):
        super().__init__()
        self.c1 = nl.SharableConv2d(3, 8, 3, padding=1, bias=False)
        self.c2 = nl.SharableConv2d(8, 12, 3, padding=1, bias=True)
        self.fc = nl.SharableLinear(12, 10)
        self.datasets = ['t1', 't2', 't3']

    def forward(self, x):
        x = torch.relu(self.c1(x))
        x = torch.relu(self.c2(x)).mean((2, 3))
        return self.fc(x)


class Wrap(nn.Module):
    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


def make_args(mode, dataset='t2', wd=4e-5, freq=2, init_s=0.0, target_s=0.5):
    a = argparse.Namespace()
    a.mode, a.dataset, a.cuda, a.weight_decay = mode, dataset, True, wd
    a.pruning_frequency, a.initial_sparsity, a.target_sparsity = freq, init_s, target_s
    a.network_width_multiplier, a.log_path, a.finetune_again = 1.0, None, False
    return a


def load_toy(g, nl, prune_mod, mode, device):
    """Rebuild the state of make_golden.pruner_cases.load(mode) with product classes."""
    model = Wrap(Toy(nl)).to(device)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    masks = {}
    for name, mod in model.named_modules():
        if isinstance(mod, (nl.SharableConv2d, nl.SharableLinear)):
            k = name.replace('.', '_')
            with torch.no_grad():
                mod.weight.copy_(T(g['W_' + k]))
                if mod.bias is not None:
                    mod.bias.zero_()
            mod.piggymask = nn.Parameter(T(g['P_' + k].copy()))
            mod.weight.grad = T(g['G_' + k].copy())
            mod.piggymask.grad = T(g['GP_' + k].copy())
            masks[name] = T(g['T_' + k].copy())
    pr = prune_mod.SparsePruner(model, masks, make_args(mode), 0, 8, 2)
    return model, pr, masks
