"""The other workloads of BASELINE.json (configs[2..4]) behind ``bench.py --workload ...``: same contract (W warm-up steps,
K timed steps between device events, one JSON line), same product path (the UNMODIFIED reference model files of
baseline/_ref built on cpg_b200.layers after cpg_b200.install(), product SparsePruner, fused BN kernels), synthetic data.

  resnet50            configs[3]  ResNet-50 masked, fine-grained synthetic 224x224, batch 256 over 8 GPUs -> 32 per GPU
                      (experiment2/CPG_imagenet.sh:46); models/resnet.py:241 through cpg_b200.install()
  spherenet20         configs[4]  SphereNet-20 masked, face synthetic 112x112 (the 112x96 crop is rejected by the reference's
                      own flatten), batch 512 over 8 GPUs -> 64 per GPU (experiment3/FvGeEmAg0_CPG_face.sh:59)
  vgg16_prune_cycle   configs[2]  one task of the CPG cycle on VGG16-BN: retrain steps with piggymasks (task >= 2), then
                      gradual pruning -- a prune event every `pruning_frequency` steps on the cubic schedule
                      (utils/prune.py:55-92) -- then make_finetuning_mask for the next task (experiment1/...sh:101-132)

FLOPs per image are SURVEY 8d's probed figures (forward hooks on the reference models).
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')
WD, LR, LR_MASK = 4e-5, 1e-2, 5e-4

SPECS = {
    # name: (constructor, input shape per image, batch per GPU, classes, dataset, train GFLOP per image)
    'resnet50': ('resnet50', (3, 224, 224), 32, 200, 'cubs_cropped', 24.29),
    'spherenet20': ('spherenet20', (3, 112, 112), 64, 8, 'age', 12.16),
}
# SphereNet-20 has no normalisation layers: at the VGG learning rate of 1e-2 its loss diverges on random data within a
# few steps (tools/sphere_diag.py: the reference's torch expressions diverge the same way).
# experiment3/FvGeEmAg0_CPG_face.sh:22-28 trains it with 1e-3 (task 1) / 5e-4 (later tasks); on four memorised random
# batches even 5e-4 spikes after ~50 steps in both arms, so the timing runs use 1e-4 (the work per step is the same).
LR_OF = {'resnet50': LR, 'spherenet20': 1e-4}


def _install():
    if not os.path.isfile(os.path.join(REF_DIR, 'models', 'resnet.py')):
        raise SystemExit('baseline/_ref is not staged: python tools/stage_reference.py')
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import cpg_b200
    nl, _ = cpg_b200.install()
    import models
    return nl, models


class _Wrap(nn.Module):
    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


def _args(dataset, mode='finetune'):
    import argparse
    a = argparse.Namespace()
    a.mode, a.dataset, a.weight_decay, a.finetune_again = mode, dataset, WD, True
    a.pruning_frequency, a.initial_sparsity, a.target_sparsity = 10, 0.0, 0.1
    a.network_width_multiplier, a.log_path, a.cuda = 1.0, None, True
    return a


def _build(workload, device, regime):
    """(net, masks, pruner, optimizers, input shape, batch, classes, gflop per image)."""
    nl, models = _install()
    from cpg_b200.fused_norm import fuse_bn_relu, fuse_prelu
    from cpg_b200.prune import SparsePruner
    ctor, shape, batch, classes, dataset, gflop = SPECS[workload]
    torch.manual_seed(1)
    model = getattr(models, ctor)(dataset_history=[], dataset2num_classes={}, network_width_multiplier=1.0,
                                  shared_layer_info={})
    datasets = ['task1'] if regime == 'task1' else ['task1', dataset]
    for d in datasets:
        model.add_dataset(d, classes)
    model.set_dataset(datasets[-1])
    if workload == 'resnet50':        # models/resnet.py:147-149 draws N(0, 0.001): activations underflow after 50 layers;
        for m in model.modules():     # timing does not care, but keep the numbers finite
            if isinstance(m, nl.SharableConv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
    model = model.to(device)
    fuse_bn_relu(model)
    if os.environ.get('CPGB_FUSE_RESNET_BLOCKS', '1') != '0':
        from cpg_b200.fused_norm import fuse_resnet_blocks
        fuse_resnet_blocks(model)             # ResNet: bn -> relu and bn -> (+ identity) -> relu inside the batch-norm kernels
    if os.environ.get('CPGB_FUSE_PRELU', '1') != '0':
        fuse_prelu(model)                     # SphereNet-20: nn.PReLU after every masked convolution
    cur = len(datasets)
    rng = np.random.RandomState(7)
    masks = {}
    for name, m in model.named_modules():
        if isinstance(m, (nl.SharableConv2d, nl.SharableLinear)):
            shp = tuple(m.weight.shape)
            if regime == 'task1':
                t = np.ones(shp, dtype=np.uint8)
            else:
                t = np.where(rng.rand(*shp) < 0.5, 1, cur).astype(np.uint8)
                p = np.full(shp, 0.01, dtype=np.float32)
                old = t < cur
                p[old] = rng.uniform(0, 0.01, size=int(old.sum())).astype(np.float32)
                m.piggymask = nn.Parameter(torch.from_numpy(p).to(device))
            masks['module.' + name] = torch.from_numpy(t).to(device)
    net = _Wrap(model)
    pruner = SparsePruner(net, masks, _args(datasets[-1]), 0, 1, cur)
    sgd, adam = [], []
    head = '.{}.'.format(len(datasets) - 1)
    for name, p in net.named_parameters():
        if 'classifiers' in name:
            if head in name:
                sgd.append(p)
        elif 'piggymask' in name:
            adam.append(p)
        else:
            sgd.append(p)
    if os.environ.get('CPGB_OPTIM', 'cpgb') == 'cpgb':          # the product's fused optimizer kernels
        from cpg_b200.optim import SGD, Adam
        opts = [SGD(sgd, lr=LR_OF.get(workload, LR), weight_decay=0.0, momentum=0.9, nesterov=True)]
        if adam:
            opts.append(Adam(adam, lr=LR_MASK))
            opts[-1].emit_packed_masks(net)
    else:
        opts = [torch.optim.SGD(sgd, lr=LR_OF.get(workload, LR), weight_decay=0.0, momentum=0.9, nesterov=True, fused=True)]
        if adam:
            opts.append(torch.optim.Adam(adam, lr=LR_MASK, capturable=True, fused=True))
    net.train()
    return net, masks, pruner, opts, shape, batch, classes, gflop


def _timed(step, feed, steps, warmup, world, finish=None):
    import torch.distributed as dist
    for i in range(warmup):
        feed(i); step()
    if finish is not None:
        finish()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        feed(warmup + i); step()
    if finish is not None:
        finish()                               # e.g. the host read of the last step's loss
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def run_model_workload(workload, args, device, world, rank):
    """ResNet-50 / SphereNet-20: the step of utils/manager.py:54-75 in one CUDA graph, both mask regimes."""
    from cpg_b200 import _lib
    from cpg_b200.ddp import GradAllReducer
    lib = _lib.load()
    out = {}
    for regime in ('task1', 'task2'):
        net, masks, pruner, opts, shape, batch, classes, gflop = _build(workload, device, regime)
        reducer = GradAllReducer(net, world) if world > 1 else None
        x = torch.zeros(batch, *shape, device=device)
        t = torch.zeros(batch, dtype=torch.int64, device=device)
        loss_buf = torch.zeros((), device=device)
        crit = nn.CrossEntropyLoss()

        def body():
            for o in opts:
                o.zero_grad(set_to_none=True)
            loss = crit(net(x), t)
            loss.backward()
            if reducer is not None:
                reducer.reduce()
            pruner.do_weight_decay_and_make_grads_zero()
            for o in opts:
                o.step()
            loss_buf.copy_(loss.detach())

        s = torch.cuda.Stream(device=device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        before = lib.cpgb_launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body()
        launches = lib.cpgb_launch_count() - before
        g = torch.Generator().manual_seed(100 + rank)
        host = [(torch.randn(batch, *shape, generator=g).pin_memory(), torch.randint(0, classes, (batch,), generator=g).pin_memory())
                for _ in range(4)]
        dev = [(a.to(device), b.to(device)) for a, b in host]

        def feed_dev(i):
            x.copy_(dev[i % 4][0]); t.copy_(dev[i % 4][1])

        def feed_host(i):
            x.copy_(host[i % 4][0], non_blocking=True); t.copy_(host[i % 4][1], non_blocking=True)

        ms = _timed(graph.replay, feed_dev, args.steps, args.warmup, world)

        # end to end: bench.HostFeed (H2D of every batch on a copy stream under the previous step, D2H of every loss)
        import types

        import bench
        shim = types.SimpleNamespace(x=x, t=t, loss=loss_buf, step=graph.replay, device=device)
        hf = bench.HostFeed(shim, host)
        state = {'i': 0}

        def step_e2e():
            hf.step(state['i'])
            state['i'] += 1
        ms_e2e = _timed(step_e2e, lambda i: None, args.steps, args.warmup, world,
                        finish=lambda: hf.drain(state['i'] - 1))
        imgs = batch * world * args.steps
        out[regime] = {'value': imgs / (ms * 1e-3), 'ms_per_step': ms / args.steps, 'e2e_value': imgs / (ms_e2e * 1e-3),
                       'algorithmic_tflops': gflop * 1e9 * batch * world / (ms / args.steps * 1e-3) / 1e12,
                       'gpu_launches': int(launches * args.steps), 'loss': float(loss_buf.item()),
                       'h2d': host[0][0].numel() * 4 + host[0][1].numel() * 8}
        del graph, net, pruner, opts, reducer
        torch.cuda.empty_cache()
    if rank != 0:
        return None
    ctor, shape, batch, classes, dataset, gflop = SPECS[workload]
    r1 = out['task1']
    return {
        'metric': 'masked-%s train images/sec' % workload, 'value': r1['value'], 'unit': 'images/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r1['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32 (fp32 in/out, fp32 accumulate)', 'data': 'synthetic',
        'config': {'workload': '%s masked (models/%s.py of the reference, unmodified, on cpg_b200.layers), synthetic %dx%d, '
                               'batch %d per GPU' % (workload, 'resnet' if 'resnet' in workload else 'spherenet', shape[1],
                                                     shape[2], batch),
                   'regime': 'task1 (R1: no piggymask, T==1, cur=1)', 'global_batch': batch * world,
                   'parallelism': 'dp%d' % world,
                   'step': 'utils/manager.py:54-75 sequence, CUDA graph; fused BN kernels (no ReLU fold outside nn.Sequential)'},
        'e2e': {'value': r1['e2e_value'], 'unit': 'images/s', 'h2d_bytes_per_step': r1['h2d'], 'd2h_bytes_per_step': 4},
        'gpu_launches': r1['gpu_launches'], 'loss': r1['loss'], 'algorithmic_tflops': r1['algorithmic_tflops'],
        'train_gflop_per_image': gflop, 'regime_task2': out['task2'],
    }


def run_prune_cycle(args, device, world, rank):
    """configs[2]: one task of the CPG cycle on VGG16-BN at batch 128 per GPU: `retrain` steps of the task-2 regime
    (piggymasks, Adam on them), then `prune` steps in mode 'prune' with a prune event every 10 steps (cubic schedule
    0 -> 0.1 over the window, utils/prune.py:55-92; all 15 layers in one batched radix-select), then
    make_finetuning_mask for the next task.  Steps are CUDA-graph replays; prune events run between them."""
    import bench
    from cpg_b200 import _lib
    from cpg_b200.ddp import GradAllReducer, assert_masks_identical
    from cpg_b200.prune import SparsePruner
    lib = _lib.load()
    tr = bench.Trainer('task2', device, world, use_graph=True)
    tr.prepare()
    g = torch.Generator().manual_seed(100 + rank)
    dev = [(torch.randn(bench.BATCH, 3, 32, 32, generator=g).to(device), torch.randint(0, 5, (bench.BATCH,), generator=g).to(device))
           for _ in range(8)]

    def feed(i):
        tr.x.copy_(dev[i % 8][0]); tr.t.copy_(dev[i % 8][1])

    n_retrain = n_prune = args.steps
    ms_retrain = _timed(tr.step, feed, n_retrain, args.warmup, world)
    # prune phase: same model and masks, pruner in 'prune' mode (piggymask grads are zeroed, utils/prune.py:209-210)
    pargs = bench.make_args(['task1', 'task2'], mode='prune')
    pargs.pruning_frequency, pargs.initial_sparsity, pargs.target_sparsity = 10, 0.0, 0.1
    pruner = SparsePruner(tr.net, tr.masks, pargs, 0, n_prune, tr.cur)
    tr.pruner = pruner
    tr.graph = None
    tr.prepare(eager_warmup=1)          # re-capture with the prune-mode epilogue
    state = {'step': 0, 'events': 0}

    def prune_step():
        tr.step()
        before = pruner.last_prune_step
        pruner.gradually_prune(state['step'])
        state['events'] += int(pruner.last_prune_step != before)
        state['step'] += 1

    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_prune):
        feed(i); prune_step()
    pruner.make_finetuning_mask()
    e1.record()
    torch.cuda.synchronize()
    ms_prune = e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_prune], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_prune = float(t.item())
        assert_masks_identical(tr.masks)
    zeros = sum(int((m == 0).sum()) for m in tr.masks.values())
    if rank != 0:
        return None
    imgs = bench.BATCH * world
    total_ms = ms_retrain + ms_prune
    return {
        'metric': 'masked-VGG16 CPG task cycle (retrain + gradual prune) images/sec',
        'value': imgs * (n_retrain + n_prune) / (total_ms * 1e-3), 'unit': 'images/s', 'n_gpus': world,
        'steps': n_retrain + n_prune, 'warmup': args.warmup, 'ms_per_step': total_ms / (n_retrain + n_prune),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32 (fp32 in/out, fp32 accumulate)',
        'data': 'synthetic',
        'config': {'workload': 'VGG16-BN CPG prune->retrain cycle, synthetic CIFAR-100 32x32, batch 128 per GPU',
                   'regime': 'task 2: %d retrain steps (piggymasks) + %d prune-mode steps, prune event every 10 steps '
                             '(cubic schedule 0 -> 0.1), make_finetuning_mask at the end' % (n_retrain, n_prune),
                   'global_batch': imgs, 'parallelism': 'dp%d' % world},
        'retrain': {'images_per_s': imgs * n_retrain / (ms_retrain * 1e-3), 'ms_per_step': ms_retrain / n_retrain},
        'prune': {'images_per_s': imgs * n_prune / (ms_prune * 1e-3), 'ms_per_step': ms_prune / n_prune,
                  'prune_events': state['events'], 'mask_zeros_after_make_finetuning_mask': zeros},
        'gpu_launches': int(tr.launches_per_step * (n_retrain + n_prune)),
    }


def main(args, device, world, rank):
    if args.workload == 'vgg16_prune_cycle':
        line = run_prune_cycle(args, device, world, rank)
    else:
        line = run_model_workload(args.workload, args, device, world, rank)
    if rank == 0:
        print(json.dumps(line))
