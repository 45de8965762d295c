#!/usr/bin/env python
"""bench.py -- masked-VGG16-BN training images/s (BASELINE.json metric) on N B200s.

  python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA path through the C ABI)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)
  torchrun --nproc-per-node N bench.py --gpus N ...        # weak scaling, one rank per GPU

A "step" is one pass of the hot path over one synthetic batch, exactly the step sequence of the
reference's utils/manager.py:54-75:
  zero_grad -> model(data) -> criterion -> backward -> do_weight_decay_and_make_grads_zero -> step.
Workload = BASELINE.json configs[1]: VGG16-BN masked, CIFAR-100 task-1 synthetic 32x32, batch
128 per GPU (regime R1 of SURVEY 8d: no piggymask, T==1).  The same run also measures regime
R2 (task 2: piggymask on every sharable layer) and reports it under "regime_task2".

One JSON line is printed by rank 0 (see the contract in the round prompt).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cpg_b200.vgg_cifar import VGGCifar  # noqa: E402

METRIC = 'masked-VGG16 train images/sec'
WORKLOAD = 'VGG16-BN masked, CIFAR-100 task-1 synthetic 32x32, batch 128 per GPU'
BATCH = 128
WD, LR, LR_MASK = 4e-5, 1e-2, 5e-4   # experiment1/CPG_cifar100_scratch_mul_1.5.sh:38-39,59
VGG_FWD_FLOP_PER_IMG = 0.6641e9      # SURVEY 8d, probed with forward hooks
VGG_TRAIN_FLOP_PER_IMG = 1.989e9


# --------------------------------------------------------------------------------------------
# model / mask construction shared by both arms (SURVEY 8d "Synthetic inputs")
# --------------------------------------------------------------------------------------------
class _Wrap(nn.Module):
    """nn.DataParallel stand-in: the reference pruner expects ``model.module.datasets``."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


REF_DIR = os.path.join(ROOT, 'baseline', '_ref')     # unmodified reference checkout (tools/stage_reference.py)
VGG_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M']   # CPG_cifar100_main_normal.py:188


def have_reference():
    return os.path.isfile(os.path.join(REF_DIR, 'models', 'vgg.py'))


def reference_vgg(width):
    """The reference's own models.custom_vgg_cifar100 (models/vgg.py:276-278).  In the product arm
    cpg_b200.install() has aliased models.layers first, so the unmodified model file builds itself out of the
    product layers; in the reference arm nothing is aliased."""
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import models
    return models.custom_vgg_cifar100(list(VGG_CFG), dataset_history=[], dataset2num_classes={},
                                      network_width_multiplier=width, shared_layer_info={})


def build_model(conv_cls, linear_cls, regime, device, width=1.0, from_reference=False):
    torch.manual_seed(1)                      # --seed default, CPG_cifar100_main_normal.py:79
    model = reference_vgg(width) if from_reference else VGGCifar(conv_cls, linear_cls, width=width)
    datasets = ['task1'] if regime == 'task1' else ['task1', 'task2']
    for d in datasets:
        model.add_dataset(d, 5)
    model.set_dataset(datasets[-1])
    model = model.to(device)
    cur = len(datasets)
    rng = np.random.RandomState(7)
    masks = {}
    for name, m in model.named_modules():
        if isinstance(m, (conv_cls, linear_cls)):
            shape = tuple(m.weight.shape)
            if regime == 'task1':
                t = np.ones(shape, dtype=np.uint8)                 # after make_finetuning_mask
            else:
                t = np.where(rng.rand(*shape) < 0.5, 1, cur).astype(np.uint8)
                p = np.full(shape, 0.01, dtype=np.float32)
                old = t < cur
                p[old] = rng.uniform(0, 0.01, size=int(old.sum())).astype(np.float32)
                m.piggymask = nn.Parameter(torch.from_numpy(p).to(device))
            masks['module.' + name] = torch.from_numpy(t).to(device)
    return _Wrap(model), masks, datasets, cur


def make_args(datasets, mode='finetune'):
    a = argparse.Namespace()
    a.mode, a.dataset, a.weight_decay = mode, datasets[-1], WD
    a.finetune_again = True           # cur = index(dataset)+1
    a.pruning_frequency, a.initial_sparsity, a.target_sparsity = 10, 0.0, 0.1
    a.network_width_multiplier, a.log_path, a.cuda = 1.0, None, True
    return a


FUSE_BN_RELU = os.environ.get('CPGB_FUSE_BN', '1') != '0'   # cpg_b200.fused_norm on the product arm
# optimizer of the product arm: 'cpgb' = cpg_b200.optim (one launch per optimizer, bit-identical to
# torch's multi-tensor update), 'torch' = torch.optim.{SGD,Adam}(fused=True)
OPTIM = os.environ.get('CPGB_OPTIM', 'cpgb')


def split_params(net):
    sgd_params, adam_params = [], []
    head = '.{}.'.format(len(net.module.datasets) - 1)
    for name, p in net.named_parameters():        # routing of CPG_cifar100_main_normal.py:326-346
        if 'classifiers' in name:
            if head in name:
                sgd_params.append(p)
        elif 'piggymask' in name:
            adam_params.append(p)
        else:
            sgd_params.append(p)
    return sgd_params, adam_params


def make_optimizers(net, capturable, product=False):
    """The two optimizers of CPG_cifar100_main_normal.py:339-346.  product: the arm under test (cpg_b200.optim unless
    CPGB_OPTIM=torch); otherwise stock torch.optim (fused=True on the GPU: the same update rule in one pass)."""
    sgd_params, adam_params = split_params(net)
    on_gpu = bool(sgd_params) and sgd_params[0].is_cuda
    if product and on_gpu and OPTIM == 'cpgb':
        from cpg_b200.optim import SGD, Adam
        opts = [SGD(sgd_params, lr=LR, weight_decay=0.0, momentum=0.9, nesterov=True)]
        if adam_params:
            opts.append(Adam(adam_params, lr=LR_MASK))
            opts[-1].emit_packed_masks(net)      # the FC layers' Binarizer bits come out of the Adam kernel
        return opts
    opts = [torch.optim.SGD(sgd_params, lr=LR, weight_decay=0.0, momentum=0.9, nesterov=True,
                            **({'fused': True} if on_gpu else {}))]
    if adam_params:
        opts.append(torch.optim.Adam(adam_params, lr=LR_MASK, capturable=capturable,
                                     **({'fused': True} if on_gpu else {})))
    return opts


def synth_batches(n, batch, seed):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(batch, 3, 32, 32, generator=g), torch.randint(0, 5, (batch,), generator=g))
            for _ in range(n)]


# --------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile('w', suffix='.csv', delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                f = [c.strip() for c in line.split(',')]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                   f[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # under load = samples in the upper half of observed power
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=max(power))
        return out


# --------------------------------------------------------------------------------------------
# context arm: the reference's expressions in stock PyTorch on the same GPU (cuDNN / cuBLAS, torch
# defaults: TF32 allowed for convs, fp32 matmul) -- what CPG runs today on a GPU ("the bar
# to beat is torch eager + cuDNN on the same B200").  Not the oracle, not the product.
# --------------------------------------------------------------------------------------------
class _TorchBinarizer(torch.autograd.Function):          # models/layers.py:11-23
    @staticmethod
    def forward(ctx, inputs, threshold):
        out = inputs.clone()
        out[inputs.le(threshold)] = 0
        out[inputs.gt(threshold)] = 1
        return out

    @staticmethod
    def backward(ctx, g):
        return g, None


class TorchSharableConv2d(nn.Module):                    # models/layers.py:43-109
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, **kw):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.stride, self.padding, self.piggymask = stride, padding, None

    def forward(self, x):
        w = self.weight if self.piggymask is None else _TorchBinarizer.apply(self.piggymask, 5e-3) * self.weight
        return torch.nn.functional.conv2d(x, w, self.bias, self.stride, self.padding)


class TorchSharableLinear(nn.Module):                    # models/layers.py:147-194
    def __init__(self, in_features, out_features, bias=True, **kw):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features)) if bias else None
        self.piggymask = None

    def forward(self, x):
        w = self.weight if self.piggymask is None else _TorchBinarizer.apply(self.piggymask, 5e-3) * self.weight
        return torch.nn.functional.linear(x, w, self.bias)


class TorchPruner:
    """do_weight_decay_and_make_grads_zero in stock torch ops (utils/prune.py:195-211), finetune mode."""

    def __init__(self, net, masks, cur):
        self.items = [(m, masks['module.' + n]) for n, m in net.module.named_modules()
                      if isinstance(m, (TorchSharableConv2d, TorchSharableLinear))]
        self.cur = cur

    def do_weight_decay_and_make_grads_zero(self):
        for m, t in self.items:
            if m.weight.grad is not None:
                m.weight.grad.add_(m.weight.data, alpha=WD)
                m.weight.grad[t.ne(self.cur)] = 0
            if m.piggymask is not None and m.piggymask.grad is not None:
                m.piggymask.grad[t.eq(0) | t.ge(self.cur)] = 0


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
class Trainer:
    """The training step of utils/manager.py:54-75 over the product layers/pruner, optionally
    captured into one CUDA graph (static input buffers)."""

    def __init__(self, regime, device, world, use_graph=True, impl='ours', width=1.0):
        from cpg_b200.ddp import GradAllReducer
        self.device, self.world = device, world
        if impl == 'ours':
            import cpg_b200
            import cpg_b200.layers as nl
            from cpg_b200.prune import SparsePruner
            self.model_source = 'cpg_b200.vgg_cifar.VGGCifar (harness mirroring models/vgg.py)'
            use_ref = have_reference() and os.environ.get('CPGB_BENCH_HARNESS_MODEL', '0') != '1'
            if use_ref:
                # the unmodified models/vgg.py of the staged reference checkout, built on the product layers
                if REF_DIR not in sys.path:
                    sys.path.insert(0, REF_DIR)
                cpg_b200.install()
                self.model_source = 'baseline/_ref/models/vgg.py custom_vgg_cifar100 (unmodified) after cpg_b200.install()'
            self.net, self.masks, datasets, self.cur = build_model(nl.SharableConv2d, nl.SharableLinear, regime, device,
                                                                   width=width, from_reference=use_ref)
            if FUSE_BN_RELU:
                from cpg_b200.fused_norm import fuse_bn_relu
                fuse_bn_relu(self.net)
            self.pruner = SparsePruner(self.net, self.masks, make_args(datasets), 0, 1, self.cur)
        else:
            self.model_source = 'stock torch modules'
            self.net, self.masks, datasets, self.cur = build_model(TorchSharableConv2d, TorchSharableLinear, regime,
                                                                   device, width=width)
            self.pruner = TorchPruner(self.net, self.masks, self.cur)
        self.opts = make_optimizers(self.net, capturable=use_graph, product=(impl == 'ours'))
        self.crit = nn.CrossEntropyLoss()
        self.reducer = GradAllReducer(self.net, world) if world > 1 else None
        self.net.train()
        self.x = torch.zeros(BATCH, 3, 32, 32, device=device)
        self.t = torch.zeros(BATCH, dtype=torch.int64, device=device)
        self.loss = torch.zeros((), device=device)
        self.graph = None
        self.use_graph = use_graph

    def _step_body(self):
        for o in self.opts:
            o.zero_grad(set_to_none=True)
        out = self.net(self.x)
        loss = self.crit(out, self.t)
        loss.backward()
        if self.reducer is not None:
            self.reducer.reduce()
        self.pruner.do_weight_decay_and_make_grads_zero()
        for o in self.opts:
            o.step()
        self.loss.copy_(loss.detach())

    def prepare(self, eager_warmup=3):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(eager_warmup):
                self._step_body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if self.use_graph:
            from cpg_b200 import _lib
            lib = _lib.load()
            before = lib.cpgb_launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._step_body()
            self.launches_per_step = lib.cpgb_launch_count() - before
        else:
            from cpg_b200 import _lib
            lib = _lib.load()
            before = lib.cpgb_launch_count()
            self._step_body()
            self.launches_per_step = lib.cpgb_launch_count() - before
        torch.cuda.synchronize()

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()


def dist_setup(gpus):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        import torch.distributed as dist
        from cpg_b200.ddp import tune_env
        tune_env(world)
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return world, rank, local


class HostFeed:
    """The end-to-end input / result path of a training loop around the captured step: every step's batch comes from
    pinned host memory and every step's loss goes back to the host, but neither stalls the GPU.  Batch i + 1 is copied
    host -> device on a copy stream into one of two staging buffers while step i runs; at the top of step i + 1 the
    main stream waits for that copy and moves the batch into the graph's static input (device -> device, 1.5 MB); the
    loss of step i is copied into a pinned slot behind the step and read by the host one step later, when it has
    long landed (asynchronous logging).  Bytes per step: the same H2D and D2H as a blocking loop."""

    def __init__(self, tr, host_batches):
        self.tr, self.host = tr, host_batches
        dev = tr.device
        self.copy = torch.cuda.Stream(device=dev)
        self.stage = [(torch.empty_like(tr.x), torch.empty_like(tr.t)) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.loss_ev = [torch.cuda.Event() for _ in range(2)]
        self.issued = -1          # last batch index handed to the copy stream
        self.last = 0.0

    def prefetch(self, i):
        if i <= self.issued:
            return
        slot = i % 2
        xb, tb = self.host[i % len(self.host)]
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(self.free[slot])          # the staging buffer's previous batch has been consumed
            self.stage[slot][0].copy_(xb, non_blocking=True)
            self.stage[slot][1].copy_(tb, non_blocking=True)
            self.ready[slot].record(self.copy)
        self.issued = i

    def step(self, i):
        tr, slot = self.tr, i % 2
        main = torch.cuda.current_stream(tr.device)
        self.prefetch(i)
        main.wait_event(self.ready[slot])
        tr.x.copy_(self.stage[slot][0]); tr.t.copy_(self.stage[slot][1])
        self.free[slot].record(main)
        tr.step()
        self.loss_host[slot].copy_(tr.loss, non_blocking=True)     # D2H of this step's result
        self.loss_ev[slot].record(main)
        self.prefetch(i + 1)                                       # overlaps the step just launched
        if i > 0:                                                  # read the previous step's loss: no stall
            self.loss_ev[1 - slot].synchronize()
            self.last = float(self.loss_host[1 - slot])

    def drain(self, i_last):
        self.loss_ev[i_last % 2].synchronize()
        self.last = float(self.loss_host[i_last % 2])
        return self.last


def timed_region(tr, batches_dev, steps, warmup, world, e2e_host=None):
    """Returns (ms_total_max_over_ranks, last_loss).  e2e_host: list of pinned (x, t) host
    batches -> per step H2D of the inputs + D2H of the loss inside the timed region (HostFeed)."""
    import torch.distributed as dist
    n = len(batches_dev) if e2e_host is None else len(e2e_host)
    feed_host = HostFeed(tr, e2e_host) if e2e_host is not None else None

    def feed(i):
        xb, tb = batches_dev[i % n]
        tr.x.copy_(xb); tr.t.copy_(tb)

    last = 0.0
    for i in range(warmup):
        if feed_host is not None:
            feed_host.step(i)
        else:
            feed(i); tr.step()
    if feed_host is not None and warmup > 0:
        feed_host.drain(warmup - 1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        if feed_host is not None:
            feed_host.step(warmup + i)     # H2D of the batch, the step, D2H of its loss
        else:
            feed(warmup + i); tr.step()
    if feed_host is not None:
        last = feed_host.drain(warmup + steps - 1)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=tr.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if e2e_host is None:
        last = tr.loss.item()
    return ms, last


def measure_tf32_peak(device):
    """TF32 dense peak is not in MEASURED_PEAKS.json: measure it the way that file measures bf16
    (torch.matmul 8192^3, best of 10, CUDA events)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device=device)
    b = torch.randn(8192, 8192, device=device)
    best = 1e9
    for i in range(13):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return 2 * 8192 ** 3 / (best * 1e-3) / 1e12


def kernel_table(device, regime_has_piggy, iters=5, width=1.0, only=None):
    """Per-layer, per-pass device time of OUR conv/linear kernels at the bench workload, timed
    with CUDA events on the launching stream, L2 flushed between launches.  Returns rows and
    the dominant pass (largest summed time) with its algorithmic FLOPs."""
    import cpg_b200.layers as nl
    from cpg_b200 import _lib
    lib = _lib.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    ch = lambda v: int(v * width)
    shapes = [(3, ch(64), 32), (ch(64), ch(64), 32), (ch(64), ch(128), 16), (ch(128), ch(128), 16), (ch(128), ch(256), 8),
              (ch(256), ch(256), 8), (ch(256), ch(256), 8), (ch(256), ch(512), 4), (ch(512), ch(512), 4),
              (ch(512), ch(512), 4), (ch(512), ch(512), 2), (ch(512), ch(512), 2), (ch(512), ch(512), 2)]
    from cpg_b200.functional import empty_nhwc
    rows = []

    def time_call(fn):
        ts = []
        for i in range(iters + 2):
            flush.zero_()
            # a ~100 us device-side spin lets the host enqueue (tensor-map encode + launch) ahead of
            # the GPU, so the events bracket device time only
            torch.cuda._sleep(200000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    def bench_layer(kind, d, x, w, p, y, dy, t, flops, first):
        if only and not any(f in kind for f in only):
            return
        # as in the real step, the activation operands arrive TF32-exact from their producers (fused BN kernels)
        for tns in (x, dy):
            base = tns if tns._base is None else tns._base
            if tns._base is not None:
                base.zero_()
                tns.copy_(torch.randn(tns.shape, device=device))
            _lib.check(lib.cpgb_round_tf32(_lib.ptr(base), _lib.ptr(base), base.numel(), _lib.stream_ptr()), 'round')
        d.flags = _lib.FLAG_X_TF32 | _lib.FLAG_DY_TF32
        ws = torch.empty(lib.cpgb_workspace_bytes(d), dtype=torch.uint8, device=device)
        dW, dP = torch.empty_like(w), (torch.empty_like(w) if p is not None else None)
        from cpg_b200.functional import empty_like_padded
        dx = empty_like_padded(x)
        st = _lib.stream_ptr()
        P = _lib.ptr
        nst = lib.cpgb_staged_weight_bytes(d)
        staged = torch.empty(max(nst, 256), dtype=torch.uint8, device=device)
        stg = 0.0
        if nst and lib.cpgb_intile_eligible(d):
            # in-tile masking: no staged copy; with a piggymask the only preparation is packing its bits
            d.flags |= _lib.FLAG_W_INTILE
            staged = torch.empty((w.numel() + 31) // 32, dtype=torch.int64, device=device)
            if p is not None:
                stg = time_call(lambda: _lib.check(lib.cpgb_pack_mask(P(p), None, w.numel(), 5e-3, 255, P(staged), st), 'pack'))
        elif nst:
            stg = time_call(lambda: _lib.check(lib.cpgb_stage_weights(d, P(w), P(p), 5e-3, P(staged), nst, st), 'stage'))
        sp = P(staged) if nst else None
        f = time_call(lambda: _lib.check(lib.cpgb_conv2d_fprop(d, P(x), P(w), P(p), None, P(y), 5e-3, sp, P(ws),
                                                                ws.numel(), st), 'fprop'))
        dg = 0.0 if first else time_call(lambda: _lib.check(lib.cpgb_conv2d_dgrad(
            d, P(dy), P(w), P(p), P(dx), 5e-3, sp, P(ws), ws.numel(), st), 'dgrad'))
        wg = time_call(lambda: _lib.check(lib.cpgb_conv2d_wgrad_fused(
            d, P(x), P(dy), P(w), P(p), P(t), 1, WD, _lib.GRAD_FINETUNE, P(dW), P(dP), None, 5e-3, P(ws),
            ws.numel(), st), 'wgrad'))
        rows.append({'layer': kind, 'flop': flops, 'stage_ms': stg, 'fprop_ms': f, 'dgrad_ms': dg, 'wgrad_ms': wg,
                     'n_weights': w.numel()})

    for i, (C, K, HW) in enumerate(shapes):
        # NHWC with the pixel stride padded to 4 where the channel count is not a multiple of 4, as
        # cpg_b200.functional stores such activations (the stem's input, the grown widths)
        x = empty_nhwc((BATCH, C, HW, HW), device)
        x.copy_(torch.randn(BATCH, C, HW, HW, device=device))
        w = torch.randn(K, C, 3, 3, device=device) * 0.05
        p = torch.rand_like(w) * 0.01 if regime_has_piggy else None
        y = empty_nhwc((BATCH, K, HW, HW), device)
        dy = empty_nhwc((BATCH, K, HW, HW), device)
        dy.copy_(torch.randn(BATCH, K, HW, HW, device=device))
        t = torch.ones(w.shape, dtype=torch.uint8, device=device)
        d = _lib.conv_desc(x.shape, x.stride(), w.shape, y.shape, y.stride(), (1, 1), (1, 1), (1, 1), 1)
        bench_layer(f'conv{C}x{K}@{HW}', d, x, w, p, y, dy, t, 2.0 * BATCH * K * C * 9 * HW * HW, i == 0)
    for (I, O) in ((ch(512), ch(4096)), (ch(4096), ch(4096))):
        if I % 4:
            continue          # row-padded matrices: timed inside the whole step only
        x = torch.randn(BATCH, I, device=device)
        w = torch.randn(O, I, device=device) * 0.01
        p = torch.rand_like(w) * 0.01 if regime_has_piggy else None
        y = torch.empty(BATCH, O, device=device)
        dy = torch.randn_like(y)
        t = torch.ones(w.shape, dtype=torch.uint8, device=device)
        d = _lib.ConvDesc()
        lib.cpgb_linear_desc(d, BATCH, I, O)
        bench_layer(f'fc{I}x{O}', d, x, w, p, y, dy, t, 2.0 * BATCH * I * O, False)
    tot = {k: sum(r[k + '_ms'] for r in rows) for k in ('fprop', 'dgrad', 'wgrad')}
    flops = {'fprop': sum(r['flop'] for r in rows), 'dgrad': sum(r['flop'] for r in rows[1:]),
             'wgrad': sum(r['flop'] for r in rows)}
    dom = max(tot, key=tot.get)
    return rows, tot, flops, dom


def prune_table(device, iters=5):
    """a7 on the device: one prune event (utils/prune.py:78-92 -> _pruning_mask per layer) over the 15
    sharable layers of VGG16 (33.6 M weights), CUDA events, L2 flushed.  Algorithmic bytes = one pass:
    read W 4 B + T 1 B, write T 1 B per element; `passes` = how often the select streams W and T
    (two: cpgb_prune_select_sampled; four: the 3-digit radix select, CPGB_PRUNE_SAMPLED=0)."""
    import cpg_b200.layers as nl
    from cpg_b200.prune import SparsePruner
    net, masks, datasets, cur = build_model(nl.SharableConv2d, nl.SharableLinear, 'task1', device)
    args = make_args(datasets, mode='prune')
    args.target_sparsity, args.initial_sparsity, args.pruning_frequency = 0.5, 0.0, 1
    pruner = SparsePruner(net, masks, args, 0, 10 ** 6, cur)
    n = sum(m.numel() for m in masks.values())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    ts = []
    for i in range(iters + 2):
        for m in masks.values():
            m.fill_(cur)
        flush.zero_()
        torch.cuda._sleep(4000000)     # ~2 ms head start: 105 launches are enqueued before the GPU gets to them
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pruner.last_prune_step = -10
        pruner.gradually_prune(1000 + i)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    ms = statistics.median(ts)
    # the select alone (what the roofline fraction is about): events around the batched C call, no read-back inside
    layers = list(pruner._sharable())
    infos = torch.zeros(len(layers), 4, dtype=torch.int64, device=device)
    tk = []
    for i in range(iters + 2):
        for m in masks.values():
            m.fill_(cur)
        flush.zero_()
        torch.cuda._sleep(4000000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pruner._launch_prune_batched(layers, 0.0015, infos)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            tk.append(e0.elapsed_time(e1))
    ms_k = statistics.median(tk)
    assert int(infos[:, 0].abs().sum()) == 0
    zeros = sum(int((m == 0).sum()) for m in masks.values())
    sampled = os.environ.get('CPGB_PRUNE_SAMPLED', '1') != '0'
    return {'ms_per_prune_event': ms, 'elements': n, 'algorithmic_bytes': 6 * n, 'passes': 2 if sampled else 4,
            'select': ('two-pass: sorted sample brackets the k-th magnitude, exact count + select inside the bracket '
                       '(cpgb_prune_select_sampled)' if sampled else 'four-pass 11/11/9-bit radix select'),
            'ms_select_kernels': ms_k, 'achieved_gbs_algorithmic': 6 * n / (ms_k * 1e-3) / 1e9,
            'achieved_gbs_event_incl_readback': 6 * n / (ms * 1e-3) / 1e9, 'pruned_fraction_check': zeros / n}


def cpu_port_step_time(regime, batch, steps, warmup, threads):
    """Fallback when baseline/_ref is not staged: the oracle port of the reference's CPU path (OracleSharable*
    layers + OraclePruner), same step sequence, on `threads` host cores."""
    from oracle import cpg_oracle as O
    torch.set_num_threads(threads)
    net, masks, datasets, cur = build_model(O.OracleSharableConv2d, O.OracleSharableLinear, regime, 'cpu')
    masks = {k[len('module.'):]: v for k, v in masks.items()}
    pruner = O.OraclePruner(net.module, masks, mode='finetune', weight_decay=WD, cur=cur, inference_idx=cur)
    opts = make_optimizers(net, capturable=False)
    crit = nn.CrossEntropyLoss()
    net.train()
    data = synth_batches(2, batch, seed=11)
    times = []
    for i in range(warmup + steps):
        x, t = data[i % 2]
        t0 = time.perf_counter()
        for o in opts:
            o.zero_grad()
        loss = crit(net(x), t)
        loss.backward()
        pruner.do_weight_decay_and_make_grads_zero()
        for o in opts:
            o.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times), float(loss.item())


def cpu_reference_step_time(regime, batch, steps, warmup, threads):
    """The UNMODIFIED reference (baseline/_ref, staged by tools/stage_reference.py) on the host cores, through its
    own public API and stock code path: models.custom_vgg_cifar100 (models/vgg.py:276) built on models.layers,
    utils.prune.SparsePruner, and utils.manager.Manager.train (utils/manager.py:39-100) looping over a list of
    synthetic (data, target) batches -- zero_grad, forward, loss, backward, do_weight_decay_and_make_grads_zero,
    optimizers.step, plus the per-batch accuracy / sparsity bookkeeping the reference does.  Returns (seconds for
    `steps` batches, last training accuracy)."""
    torch.set_num_threads(threads)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import models.layers as ref_nl
    assert ref_nl.__file__.startswith(REF_DIR), 'models.layers is not the staged reference (cpg_b200.install() ran?)'
    from utils import Optimizers
    from utils.manager import Manager
    net, masks, datasets, cur = build_model(ref_nl.SharableConv2d, ref_nl.SharableLinear, regime, 'cpu',
                                            from_reference=True)
    args = make_args(datasets)
    args.cuda, args.checkpoint_format = False, ''
    data = synth_batches(2, batch, seed=11)

    def run(n):
        loader = [data[i % 2] for i in range(n)]
        shared = {args.dataset: {'bias': {}, 'bn_layer_running_mean': {}, 'bn_layer_running_var': {},
                                 'bn_layer_weight': {}, 'bn_layer_bias': {}, 'piggymask': {}}}
        mgr = Manager(args, net, shared, masks, loader, loader, 0, 1)
        opts = Optimizers()
        for o, lr in zip(make_optimizers(net, capturable=False), (LR, LR_MASK)):
            opts.add(o, lr)
        t0 = time.perf_counter()
        acc, _ = mgr.train(opts, 0, [LR], 0)
        return time.perf_counter() - t0, acc

    if warmup:
        run(warmup)
    return run(steps)


def reference_arm_time(regime, batch, steps, warmup, threads):
    """(seconds, kind, what) of the reference's CPU implementation of the path."""
    if have_reference():
        total, _ = cpu_reference_step_time(regime, batch, steps, warmup, threads)
        return total, 'reference', ('unmodified ivclab/CPG from baseline/_ref: models.custom_vgg_cifar100 + '
                                    'utils.prune.SparsePruner driven by utils.manager.Manager.train')
    total, _ = cpu_port_step_time(regime, batch, steps, warmup, threads)
    return total, 'port', 'oracle port of the reference step (baseline/_ref not staged)'


def make_config(world, batch_per_gpu=BATCH):
    """The workload description both arms print (same keys, same values): BASELINE.json configs[1]."""
    return {'workload': WORKLOAD, 'regime': 'task1 (R1: no piggymask, T==1, cur=1)',
            'global_batch': batch_per_gpu * world, 'parallelism': f'dp{world}',
            'step': 'utils/manager.py:54-75 sequence: zero_grad, forward, CrossEntropyLoss, backward, '
                    'do_weight_decay_and_make_grads_zero, SGD-nesterov step (CPG_cifar100_main_normal.py:339-346)'}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores (rank 0 only)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # calibrate the per-step sample so the whole run ends within a few minutes
    t_cal, _, _ = reference_arm_time('task1', 16, 1, 1, cores)
    per_img = t_cal / 16
    budget = 150.0
    sample = BATCH
    while sample > 8 and per_img * sample * (args.steps + args.warmup) > budget:
        sample //= 2
    total, kind, what = reference_arm_time('task1', sample, args.steps, args.warmup, cores)
    value = sample * args.steps / total
    cfg = make_config(max(1, args.gpus))      # the product arm's config, verbatim; the bounded sample is in cpu_baseline
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': cfg,
        'cpu_baseline': {'value': value, 'unit': 'images/s', 'cores': cores, 'kind': kind,
                         'sample': f'{args.steps} steps x {sample} images after {args.warmup} warm-up ({what}; torch '
                                   f'{torch.__version__} CPU, {torch.get_num_threads()} threads)'},
        'e2e': {'value': value, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def cpu_baseline_subprocess(steps=3, warmup=1):
    """The product process has aliased models.layers (cpg_b200.install()), so the unmodified reference is timed in a
    child process: `bench.py --impl reference`, whose JSON line is parsed."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', str(steps),
                            '--warmup', str(warmup)], capture_output=True, text=True, timeout=600,
                           env={k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK')})
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith('{'):
                return json.loads(ln)['cpu_baseline']
        return {'error': (r.stderr or r.stdout)[-300:]}
    except Exception as ex:  # noqa: BLE001
        return {'error': f'{type(ex).__name__}: {ex}'[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip kernel table / task-2 regime / cpu baseline')
    ap.add_argument('--path', default='auto', choices=['auto', 'simt'])
    ap.add_argument('--workload', default='vgg16', choices=['vgg16', 'resnet50', 'spherenet20', 'vgg16_prune_cycle'],
                    help='vgg16 = the headline (BASELINE.json configs[1]); the others are configs[3], [4], [2]')
    ap.add_argument('--piggy', action='store_true', help='--layer-table: with piggymasks (task-2 regime)')
    ap.add_argument('--layer-filter', default='', help='--layer-table: comma-separated substrings of layer names')
    ap.add_argument('--layer-table', type=float, default=0.0,
                    help='print only the per-layer kernel table at this area width multiplier (1.0, 1.5) and exit')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        return run_reference_arm(args)

    world, rank, local = dist_setup(args.gpus)
    device = torch.device('cuda', local)
    torch.cuda.set_device(device)
    from cpg_b200 import _lib
    lib = _lib.load()
    if args.path == 'simt':
        _lib.set_path(_lib.PATH_SIMT)

    if args.workload != 'vgg16':
        import bench_workloads
        bench_workloads.main(args, device, world, rank)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return
    if args.layer_table:
        rows, tot, flops, dom = kernel_table(device, regime_has_piggy=args.piggy, width=args.layer_table ** 0.5,
                                             only=[f for f in args.layer_filter.split(',') if f])
        for r in rows:
            f = r['flop']
            print('%-18s stage %6.1f  fprop %6.1f us (%4.0f TF)  dgrad %6.1f us (%4.0f TF)  wgrad %6.1f us (%4.0f TF)' % (
                r['layer'], r['stage_ms'] * 1e3, r['fprop_ms'] * 1e3, f / r['fprop_ms'] / 1e9, r['dgrad_ms'] * 1e3,
                (f / r['dgrad_ms'] / 1e9) if r['dgrad_ms'] else 0, r['wgrad_ms'] * 1e3, f / r['wgrad_ms'] / 1e9))
        print('totals ms', tot)
        return

    def run_regime(regime, want_e2e, width=1.0):
        tr = Trainer(regime, device, world, use_graph=not args.no_graph, width=width)
        tr.prepare()
        batches = synth_batches(8, BATCH, seed=100 + rank)
        dev_batches = [(x.to(device), t.to(device)) for x, t in batches]
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ms, loss = timed_region(tr, dev_batches, args.steps, args.warmup, world)
        clocks = sampler.stop() if rank == 0 else None
        res = {'ms': ms, 'loss': loss, 'clocks': clocks, 'launches_per_step': tr.launches_per_step,
               'model_source': tr.model_source}
        if want_e2e:
            host = [(x.pin_memory(), t.pin_memory()) for x, t in batches]
            ms2, loss2 = timed_region(tr, None, args.steps, args.warmup, world, e2e_host=host)
            res.update(e2e_ms=ms2, e2e_loss=loss2,
                       h2d=host[0][0].numel() * 4 + host[0][1].numel() * 8, d2h=4)
        del tr
        torch.cuda.empty_cache()
        return res

    r1 = run_regime('task1', True)
    extras = not args.no_extras
    r2 = run_regime('task2', True) if extras else None
    # the grown network of experiment1/CPG_cifar100_scratch_mul_1.5.sh:89-94 (--network_width_multiplier 1.5 ->
    # channel counts * sqrt(1.5) = 78 / 156 / 313 / 627 / 5016, CPG_cifar100_main_normal.py:115)
    r15 = run_regime('task1', False, width=1.5 ** 0.5) if (extras and world == 1) else None

    def run_torch_gpu(regime, use_graph):
        """stock-PyTorch arm on the same GPU (context only)."""
        try:
            tr = Trainer(regime, device, 1, use_graph=use_graph, impl='torch')
            tr.prepare()
            batches = [(x.to(device), t.to(device)) for x, t in synth_batches(8, BATCH, seed=100)]
            ms, loss = timed_region(tr, batches, args.steps, args.warmup, 1)
            del tr
            torch.cuda.empty_cache()
            return {'images_per_s': BATCH * args.steps / (ms * 1e-3), 'ms_per_step': ms / args.steps, 'loss': loss}
        except Exception as ex:  # noqa: BLE001 -- context arm must never break the bench line
            return {'error': f'{type(ex).__name__}: {ex}'[:200]}

    line = None
    if rank == 0:
        imgs = BATCH * world * args.steps
        value = imgs / (r1['ms'] * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        line = {
            'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': r1['ms'] / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32 (fp32 in/out, fp32 accumulate)' if args.path == 'auto' else 'f32',
            'data': 'synthetic',
            'config': make_config(world),
            'impl_detail': {'optimizer': ('cpg_b200.optim.SGD (+ cpg_b200.optim.Adam on the piggymasks in task2): '
                                          'cpgb_sgd_nesterov_step / cpgb_adam_step, one launch each, bit-identical to '
                                          'the torch.optim calls of CPG_cifar100_main_normal.py:339-346'
                                          if OPTIM == 'cpgb' else
                                          'torch.optim.SGD(fused=True) (+Adam(fused=True) on piggymasks in task2): same '
                                          'update rule as CPG_cifar100_main_normal.py:339-346'),
                            'cuda_graph': not args.no_graph,
                            'model': r1.get('model_source'),
                            'bn_relu_pool': ('cpg_b200.fused_norm (BatchNorm2d+ReLU(+MaxPool2d) kernels, SURVEY 8f N4), '
                                             'outputs rounded to TF32 for the tcgen05 convolutions'
                                             if FUSE_BN_RELU else 'stock torch modules'),
                            'not_in_step': 'the per-batch host reads of utils/manager.py:60 (accuracy .cpu()) and :77-88 '
                                           '(calculate_sparsity for the progress bar)',
                            'l2': 'per-step working set (weights+grads+momentum 400 MB, activations 280 MB) exceeds '
                                  'the 126 MB L2; 8 distinct input batches rotate'},
            'clocks': {k: r1['clocks'].get(k) for k in ('sm_mhz', 'sm_max_mhz', 'reasons')} if r1['clocks'] else None,
            'e2e': {'value': imgs / (r1['e2e_ms'] * 1e-3), 'unit': 'images/s', 'h2d_bytes_per_step': r1['h2d'],
                    'd2h_bytes_per_step': r1['d2h'], 'ms_per_step': r1['e2e_ms'] / args.steps,
                    'how': 'every step: its batch pinned host -> device (copy stream, double-buffered staging, issued '
                           'while the previous step runs), the captured step, its loss device -> pinned host (read by '
                           'the host one step later); all inside the timed region (bench.HostFeed)'},
            'gpu_launches': int(r1['launches_per_step'] * args.steps),
            'loss': r1['loss'],
        }
        if r2 is not None:
            line['regime_task2'] = {
                'regime': 'task2 (R2: piggymask on all 15 sharable layers, 50% old weights, half picked)',
                'value': imgs / (r2['ms'] * 1e-3), 'ms_per_step': r2['ms'] / args.steps,
                'e2e_value': imgs / (r2['e2e_ms'] * 1e-3), 'gpu_launches': int(r2['launches_per_step'] * args.steps),
                'loss': r2['loss']}
    if r15 is not None and rank == 0:
        flop10, flop15 = VGG_TRAIN_FLOP_PER_IMG, 3 * 0.9911e9 - 2.0 * 78 * 3 * 9 * 32 * 32
        rate10 = flop10 * BATCH / (r1['ms'] / args.steps * 1e-3) / 1e12
        rate15 = flop15 * BATCH / (r15['ms'] / args.steps * 1e-3) / 1e12
        line['regime_width_1p5'] = {
            'regime': 'task1, --network_width_multiplier 1.5: channels 78/156/313/627, FC 627->5016->5016 (padded-NHWC '
                      'activations, zero-padded weight blocks, all layers but the stem on the tcgen05 kernels)',
            'value': imgs / (r15['ms'] * 1e-3), 'ms_per_step': r15['ms'] / args.steps,
            'algorithmic_tflops': rate15, 'algorithmic_tflops_width_1p0': rate10, 'flop_rate_vs_width_1p0': rate15 / rate10,
            'gpu_launches': int(r15['launches_per_step'] * args.steps), 'loss': r15['loss']}
    if extras and world == 1:
        tf32_cublas = measure_tf32_peak(device)
        # the roofline denominator is in MEASURED_PEAKS.json's terms: tcgen05 kind::tf32 issues at exactly half the
        # kind::f16 rate (tools/mma_rate.py: 2047 MAC/clk/SM), so TF32 dense peak = bf16_tflops / 2 (burst figure:
        # each kernel is timed alone); fallback 1590 / 2 when the file is absent (B200_PROFILING.md)
        tf32_peak = (peaks.get('bf16_tflops') or 1590.0) / 2.0
        rows, tot, flops, dom = kernel_table(device, regime_has_piggy=False)
        step_ms = r1['ms'] / args.steps
        achieved = flops[dom] / (tot[dom] * 1e-3) / 1e12
        kname = {'fprop': 'conv_gemm_kernel<BN,false>', 'dgrad': 'conv_gemm_kernel<BN,true>',
                 'wgrad': 'wgrad_gemm_kernel<BN,TG,HALO> + wgrad epilogue'}[dom]
        traffic, traffic_file = None, 'profiles/r2_dram_traffic.json'
        try:   # DRAM bytes of the same launches from an ncu capture (profiles/dram_traffic_from_ncu.py)
            for fn in ('r2_dram_traffic.json', 'r1_dram_traffic.json'):
                path = os.path.join(ROOT, 'profiles', fn)
                if os.path.exists(path):
                    traffic = json.load(open(path))[dom]['dram_bytes']
                    traffic_file = 'profiles/' + fn
                    break
        except Exception:
            pass
        line['roofline'] = {
            'bound': 'tensor', 'kernel': f'{kname}: masked implicit-GEMM {dom}, the 15 sharable layers of one step',
            'achieved': achieved, 'peak': tf32_peak, 'unit': 'TFLOP/s', 'frac': achieved / tf32_peak,
            'traffic': traffic,
            'traffic_note': 'dram__bytes_read+write summed over the launches of this pass for the 15 layers (ncu, cold '
                            'caches; %s); the pass is tensor/shared-memory bound, not HBM bound' % traffic_file,
            'peak_source': ('MEASURED_PEAKS.json bf16_tflops (burst) / 2' if peaks.get('bf16_tflops') else
                            'fallback 1590 / 2 (B200_PROFILING.md)') + ': tcgen05 kind::tf32 issues at half the '
                           'kind::f16 rate (tools/mma_rate.py); cuBLAS TF32 8192^3 measured in this run: %.1f TF/s'
                           % tf32_cublas,
            'share_of_step': tot[dom] / step_ms,
            'by_pass': {k: {'achieved_tflops': flops[k] / (tot[k] * 1e-3) / 1e12,
                            'frac': flops[k] / (tot[k] * 1e-3) / 1e12 / tf32_peak, 'ms': tot[k]} for k in tot},
            'best_layer': max(({'layer': r['layer'], 'pass': k, 'tflops': r['flop'] / (r[k + '_ms'] * 1e-3) / 1e12,
                                'frac': r['flop'] / (r[k + '_ms'] * 1e-3) / 1e12 / tf32_peak}
                               for r in rows for k in ('fprop', 'dgrad', 'wgrad') if r[k + '_ms'] > 0),
                              key=lambda e: e['tflops']),
            'how': 'CUDA events on the launching stream around each launch (a device-side spin hides host enqueue '
                   'time), L2 flushed (256 MiB write) between launches, median of 5; algorithmic FLOPs = '
                   '2*N*P*Q*K*C*R*S per layer and pass (SURVEY 8d)',
        }
        line['kernels'] = {'total_ms': tot, 'algorithmic_gflop': {k: v / 1e9 for k, v in flops.items()},
                           'conv_linear_share_of_step': sum(tot.values()) / step_ms,
                           'per_layer': [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}
                                         for r in rows]}
        line['peaks'] = {'hbm_gbs': peaks.get('hbm_gbs'), 'bf16_tflops': peaks.get('bf16_tflops'),
                         'tf32_tflops_peak_used': tf32_peak, 'tf32_tflops_cublas_measured_here': tf32_cublas}
        try:
            pt = prune_table(device)
            pt['frac_of_hbm_peak'] = pt['achieved_gbs_algorithmic'] / peaks['hbm_gbs'] if peaks.get('hbm_gbs') else None
            line['prune_event'] = pt
        except Exception as ex:  # noqa: BLE001
            line['prune_event'] = {'error': f'{type(ex).__name__}: {ex}'[:200]}
        line['torch_cudnn_same_gpu'] = {
            'what': 'the reference expressions (Binarizer*W -> F.conv2d/F.linear, utils/prune.py:195-211 in torch ops) '
                    'in stock PyTorch on this GPU: cuDNN/cuBLAS, torch default TF32 flags, NCHW; context, not the '
                    'graded reference arm',
            'task1_cuda_graph': run_torch_gpu('task1', True), 'task1_eager': run_torch_gpu('task1', False),
            'task2_cuda_graph': run_torch_gpu('task2', True)}
        line['cpu_baseline'] = cpu_baseline_subprocess(3, 1)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
