"""Fused optimizers for the step right after the masked-convolution path.

The reference builds ``optim.SGD(params, lr, weight_decay=0.0, momentum=0.9, nesterov=True)`` for the weights and
``optim.Adam(piggymasks, lr=lr_mask)`` for the piggymasks (CPG_cifar100_main_normal.py:339-346).  ``SGD`` and ``Adam``
below are ``torch.optim.Optimizer`` subclasses with the same constructor arguments, ``param_groups`` and per-parameter
state names (``momentum_buffer``; ``exp_avg``, ``exp_avg_sq``, ``step``), so the reference's ``Optimizers`` wrapper,
its learning-rate schedule (it writes ``param_group['lr']``) and optimizer checkpoints work unchanged; ``step()`` is one
launch of ``cpgb_sgd_nesterov_step`` / ``cpgb_adam_step`` per 40 tensors instead of torch's one kernel per operation,
and reproduces torch's multi-tensor arithmetic bit for bit (tests/test_optim_gpu.py).

Both are capturable into CUDA graphs: Adam's step count lives on the device and is advanced by the kernel.  A learning
rate that changes between replays must live on the device as well: ``lr_tensor=True`` keeps a device copy that
``sync_lr()`` (or any ``step()`` outside capture) refreshes from ``param_group['lr']``.

There is no CPU path: parameters must be CUDA fp32 tensors with dense gradients.
"""
import ctypes

import torch

from . import _lib

_MAX_TENSORS = 40      # per cpgb_adam_step call (one device step counter per call)


def _check(p):
    if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
        raise _lib.CpgbError('cpg_b200.optim needs contiguous CUDA fp32 parameters (there is no CPU path)')
    g = p.grad
    if g.is_sparse or g.dtype != torch.float32 or g.device != p.device:
        raise _lib.CpgbError('cpg_b200.optim needs dense fp32 gradients on the parameter\'s device')
    if not g.is_contiguous():
        # the raw pointer walk needs the parameter's element order
        p.grad = g = g.contiguous()
    return g


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() if t is not None else None for t in tensors])


class _LrMixin:
    def _lr_pointer(self, gi, group):
        """Device copy of group['lr'] when the optimizer was built with lr_tensor=True (else NULL)."""
        if not group.get('lr_tensor', False):
            return None
        t = getattr(self, '_lr_devs', {}).get(gi)
        if t is None:
            raise _lib.CpgbError('lr_tensor=True: call sync_lr() once before the first step()')
        if not torch.cuda.is_current_stream_capturing():
            t.fill_(float(group['lr']))
        return t.data_ptr()

    def sync_lr(self):
        """Refresh the device copies of the learning rates from param_groups (call outside graph capture, after the
        schedule changed ``param_group['lr']``)."""
        if not hasattr(self, '_lr_devs'):
            self._lr_devs = {}
        for gi, group in enumerate(self.param_groups):
            if not group.get('lr_tensor', False) or not group['params']:
                continue
            t = self._lr_devs.get(gi)
            if t is None:
                t = torch.zeros((), dtype=self._lr_dtype, device=group['params'][0].device)
                self._lr_devs[gi] = t
            t.fill_(float(group['lr']))


class SGD(_LrMixin, torch.optim.Optimizer):
    """``torch.optim.SGD(params, lr, momentum, nesterov=True, weight_decay=0, dampening=0)`` in one launch."""
    _lr_dtype = torch.float32

    def __init__(self, params, lr=1e-3, momentum=0.9, dampening=0, weight_decay=0.0, nesterov=True, lr_tensor=False):
        if lr < 0.0:
            raise ValueError(f'Invalid learning rate: {lr}')
        if momentum <= 0.0 or not nesterov or dampening != 0 or weight_decay != 0:
            raise ValueError('cpg_b200.optim.SGD implements the configuration the reference uses: momentum > 0, '
                             'nesterov=True, dampening=0, weight_decay=0 (weight decay is part of the gradient '
                             'epilogue, utils/prune.py:203)')
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov, lr_tensor=lr_tensor))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            by_dev = {}
            for p in group['params']:
                if p.grad is None:
                    continue
                g = _check(p)
                st = self.state[p]
                buf = st.get('momentum_buffer')
                if buf is None:
                    # zeros: buf * mu + g == g exactly, torch's first-step `buf = clone(g)`
                    buf = st['momentum_buffer'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                by_dev.setdefault(p.device, []).append((p, g, buf))
            for dev, items in by_dev.items():
                n = len(items)
                N = (ctypes.c_int64 * n)(*[it[0].numel() for it in items])
                with torch.cuda.device(dev):
                    _lib.check(lib.cpgb_sgd_nesterov_step(
                        n, _ptr_array([it[0] for it in items]), _ptr_array([it[1] for it in items]),
                        _ptr_array([it[2] for it in items]), N, float(group['lr']), float(group['momentum']),
                        self._lr_pointer(gi, group), _lib.stream_ptr()), 'cpgb_sgd_nesterov_step')
        return loss


class Adam(_LrMixin, torch.optim.Optimizer):
    """``torch.optim.Adam(params, lr, betas, eps, weight_decay=0, amsgrad=False)`` in one launch per 40 tensors.

    ``pack``: optional ``{parameter: (int64 tensor of ceil(numel / 32) words, task mask or None)}`` -- the kernel
    then also writes the cpgb_pack_mask words of the updated parameter (threshold `pack_threshold`)."""
    _lr_dtype = torch.float64

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, lr_tensor=False,
                 pack=None, pack_threshold=5e-3, pack_inference_idx=255):
        if lr < 0.0:
            raise ValueError(f'Invalid learning rate: {lr}')
        if weight_decay != 0 or amsgrad:
            raise ValueError('cpg_b200.optim.Adam implements the configuration the reference uses: weight_decay=0, '
                             'amsgrad=False')
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad,
                                      lr_tensor=lr_tensor))
        self.pack = pack or {}
        self.pack_threshold = float(pack_threshold)
        self.pack_inference_idx = int(pack_inference_idx)
        self._counters = {}          # (group index, device, chunk) -> int64[2] on the device

    def emit_packed_masks(self, model, threshold=5e-3):
        """Let step() also write the packed Binarizer bits (cpgb_pack_mask words) of every piggymask of `model` that
        this optimizer owns and whose layer can mask its weight tiles in shared memory (CPGB_FLAG_W_INTILE: the FC
        layers): the next forward pass then needs no pack pass.  The words are left on the parameter
        (``p._cpgb_bits``) together with its version counter; cpg_b200.functional consumes them once and ignores
        them if anything else touched the piggymask in between.  Returns the number of layers hooked up."""
        from . import layers as nl
        lib = _lib.load()
        mine = {id(p) for g in self.param_groups for p in g['params']}
        self.pack_threshold = float(threshold)
        self.pack_inference_idx = 255
        n = 0
        for m in model.modules():
            if not isinstance(m, (nl.SharableConv2d, nl.SharableLinear)):
                continue
            p = getattr(m, 'piggymask', None)
            if p is None or id(p) not in mine or not p.is_cuda or m.info.get('threshold_fn') != 'binarizer':
                continue
            if float(m.info.get('threshold', threshold)) != float(threshold):
                continue
            w = m.weight
            if w.dim() == 4:
                (K, C, R, S), (sh, sw), groups = w.shape, m.stride, m.groups
                C = C * groups
            else:
                (K, C), R, S, sh, sw, groups = w.shape, 1, 1, 1, 1, 1
            if not lib.cpgb_intile_weight_shape(K, C, R, S, sh, sw, groups):
                continue
            self.pack[p] = (torch.zeros((p.numel() + 31) // 32, dtype=torch.int64, device=p.device), None)
            n += 1
        return n

    def _counter(self, gi, dev, chunk, start):
        key = (gi, str(dev), chunk)
        c = self._counters.get(key)
        if c is None:
            c = torch.zeros(2, dtype=torch.int64, device=dev)
            c[0] = int(start)
            self._counters[key] = c
        return c

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            by_dev = {}
            for p in group['params']:
                if p.grad is None:
                    continue
                g = _check(p)
                st = self.state[p]
                if 'exp_avg' not in st:
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                by_dev.setdefault(p.device, []).append((p, g, st))
            b1, b2 = group['betas']
            for dev, items in by_dev.items():
                for chunk, lo in enumerate(range(0, len(items), _MAX_TENSORS)):
                    its = items[lo:lo + _MAX_TENSORS]
                    n = len(its)
                    # a loaded checkpoint carries per-parameter step counts: they seed the device counter
                    start = max([int(it[2].pop('step', 0)) for it in its] + [0])
                    counter = self._counter(gi, dev, chunk, start)
                    for it in its:
                        it[2]['_counter'] = counter
                    N = (ctypes.c_int64 * n)(*[it[0].numel() for it in its])
                    packs = [self.pack.get(it[0]) for it in its]
                    P = _ptr_array([pk[0] if pk else None for pk in packs]) if any(packs) else None
                    T = _ptr_array([pk[1] if pk else None for pk in packs]) if any(packs) else None
                    with torch.cuda.device(dev):
                        _lib.check(lib.cpgb_adam_step(
                            n, _ptr_array([it[0] for it in its]), _ptr_array([it[1] for it in its]),
                            _ptr_array([it[2]['exp_avg'] for it in its]), _ptr_array([it[2]['exp_avg_sq'] for it in its]),
                            N, float(group['lr']), float(b1), float(b2), float(group['eps']), counter.data_ptr(),
                            self._lr_pointer(gi, group), P, T, self.pack_threshold, self.pack_inference_idx,
                            _lib.stream_ptr()), 'cpgb_adam_step')
                    for it, pk in zip(its, packs):
                        if pk:      # (words, version of the parameter they describe, threshold)
                            it[0]._cpgb_bits = (pk[0], it[0]._version, self.pack_threshold)
        return loss

    def state_dict(self):
        """torch's layout: every parameter's state carries its own ``step`` (read back from the device counters)."""
        steps = {}
        for st in self.state.values():
            c = st.get('_counter')
            if c is not None and id(c) not in steps:
                steps[id(c)] = float(c[0].item())
        sd = super().state_dict()
        state = {}
        for pid, st in sd['state'].items():          # the base class hands out the live per-parameter dicts: copy
            c = st.get('_counter')
            st = {k: v for k, v in st.items() if k != '_counter'}
            if c is not None:
                st['step'] = torch.tensor(steps[id(c)], dtype=torch.float32)
            state[pid] = st
        sd['state'] = state
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._counters = {}          # re-seeded from the loaded ``step`` entries at the next step()
        for st in self.state.values():
            st.pop('_counter', None)
            if 'step' in st and torch.is_tensor(st['step']):
                st['step'] = int(st['step'].item())
