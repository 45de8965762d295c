"""Shared N-step training trajectory (the step sequence of the reference's
utils/manager.py:54-75) with injectable layer classes and pruner.  Mirrors
tests/golden/make_golden.py::trajectory_case input-for-input."""
import argparse

import numpy as np
import torch
import torch.nn as nn

from cpg_b200.vgg_cifar import VGGCifar, fill_params_deterministic


class Wrap(nn.Module):
    """Stand-in for nn.DataParallel on a single device: exposes ``.module`` and the
    ``module.``-prefixed names the reference's mask dict uses."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


def make_args(mode, dataset='t2', wd=4e-5, freq=2, init_s=0.0, target_s=0.3):
    a = argparse.Namespace()
    a.mode, a.dataset, a.cuda, a.weight_decay = mode, dataset, False, wd
    a.pruning_frequency, a.initial_sparsity, a.target_sparsity = freq, init_s, target_s
    a.network_width_multiplier, a.log_path = 1.0, None
    a.finetune_again = (mode == 'finetune')
    return a


def build(conv_cls, linear_cls, device="cpu", width=0.125, steps=3, batch=8):
    torch.manual_seed(1)
    model = VGGCifar(conv_cls, linear_cls, width=width)
    model.add_dataset('t1', 5)
    model.add_dataset('t2', 5)
    model.set_dataset('t2')
    fill_params_deterministic(model, seed=3)
    rng = np.random.RandomState(21)
    masks, piggies = {}, {}
    for n, m in model.named_modules():
        if isinstance(m, (conv_cls, linear_cls)):
            tm = rng.randint(1, 3, size=tuple(m.weight.shape)).astype(np.uint8)
            masks[n] = torch.from_numpy(tm)
            pm = np.full(tuple(m.weight.shape), 0.01, dtype=np.float32)
            old = tm < 2
            pm[old] = rng.uniform(0, 0.01, size=int(old.sum())).astype(np.float32)
            piggies[n] = pm
    loader = []
    for _ in range(steps):
        data = rng.standard_normal((batch, 3, 32, 32)).astype(np.float32)
        target = rng.randint(0, 5, size=(batch,)).astype(np.int64)
        loader.append((torch.from_numpy(data), torch.from_numpy(target)))
    model = model.to(device)
    masks = {n: v.to(device) for n, v in masks.items()}
    for n, m in model.named_modules():
        if n in piggies:
            m.piggymask = nn.Parameter(torch.from_numpy(piggies[n]).to(device))
    return model, masks, loader


def run_trajectory(conv_cls, linear_cls, mode, device='cpu', pruner_factory="oracle", steps=3):
    model, masks, loader = build(conv_cls, linear_cls, device, steps=steps)
    args = make_args(mode)
    if pruner_factory == 'oracle':
        from oracle.cpg_oracle import OraclePruner
        pruner = OraclePruner(model, masks, mode=mode, weight_decay=args.weight_decay, cur=2,
                              inference_idx=2, begin_prune_step=0, end_prune_step=4,
                              initial_sparsity=0.0, target_sparsity=0.3, pruning_frequency=2)
        net = model
    else:
        from cpg_b200.prune import SparsePruner
        net = Wrap(model)
        masks = {'module.' + n: v for n, v in masks.items()}
        pruner = SparsePruner(net, masks, args, 0, 4, 2)
        assert pruner.current_dataset_idx == 2
    sgd_params = [p for n, p in model.named_parameters()
                  if 'piggymask' not in n and ('classifiers' not in n or '.1.' in n)]
    adam_params = [p for n, p in model.named_parameters() if 'piggymask' in n]
    opt_w = torch.optim.SGD(sgd_params, lr=1e-2, weight_decay=0.0, momentum=0.9, nesterov=True)
    opt_m = torch.optim.Adam(adam_params, lr=5e-4)
    crit = nn.CrossEntropyLoss()
    model.train()
    step = 0
    for data, target in loader:
        data, target = data.to(device), target.to(device)
        opt_w.zero_grad()
        opt_m.zero_grad()
        loss = crit(net(data), target)
        loss.backward()
        pruner.do_weight_decay_and_make_grads_zero()
        opt_w.step()
        opt_m.step()
        if mode == 'prune':
            pruner.gradually_prune(step)
            step += 1
    if pruner_factory != 'oracle':
        masks = {n[len('module.'):]: v for n, v in masks.items()}
    return model, {n: v.cpu() for n, v in masks.items()}
