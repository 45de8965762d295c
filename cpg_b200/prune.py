"""Drop-in replacement of the reference's ``utils/prune.py`` SparsePruner: same constructor,
attributes and method names (utils/prune.py:6-243); the five hot methods run as CUDA kernels
behind include/cpgb200.h, the schedule arithmetic stays the reference's python doubles.

Differences that are deliberate and invisible to ``Manager``:
  * ``_pruning_mask`` never leaves the device (the reference does a D2H copy + CPU kthvalue,
    utils/prune.py:39); the exit-2 condition is read back once per prune event;
  * when ``fuse_grad_epilogue`` is on (default) the layers' backward already applies weight
    decay and gradient masking in the wgrad epilogue, and
    ``do_weight_decay_and_make_grads_zero`` only clears the per-layer "finalised" flag.
"""
import os
import sys
import weakref

import torch

from . import _lib
from . import layers as nl
from .functional import FuseCtx, _side_stream

# Build the staged weight operands next to the first layers of the forward pass (side stream + event)
# instead of in front of them.  Measured on B200 (VGG16 @ batch 128, A/B on one box): 1.958 ms/step with,
# 1.925 ms without -- the 118 MB transposing copy slows the stem and the first BN more than the overlap
# saves.  Off by default; CPGB_STAGE_SIDE=1 turns it on.
STAGE_ON_SIDE_STREAM = os.environ.get('CPGB_STAGE_SIDE', '0') == '1'

# `apply_mask()` (utils/prune.py:223-231) destroys the weights of the tasks after
# `inference_dataset_idx` in memory, so evaluating task j of a model that holds t > j tasks means reloading the
# checkpoint per task.  With this switch on (CPGB_NONDESTRUCTIVE_EVAL=1) `apply_mask()` leaves `weight.data` alone and
# calls `select_task()` instead: evaluation-mode forward passes read W * [1 <= T <= idx] from a resident copy.  Off by
# default: during training the reference relies on the destructive call to zero freshly pruned weights at every
# validation (utils/manager.py:105), and a drop-in must reproduce that trajectory.
NONDESTRUCTIVE_APPLY_MASK = os.environ.get('CPGB_NONDESTRUCTIVE_EVAL', '0') == '1'

_MODES = {'finetune': _lib.GRAD_FINETUNE, 'prune': _lib.GRAD_PRUNE}


class _TaskView(object):
    def __init__(self, pruner, idx):
        self.pruner, self.idx = pruner, idx

    def __enter__(self):
        self.pruner.select_task(self.idx)
        return self.pruner

    def __exit__(self, *exc):
        self.pruner.clear_task_view()
        return False


class SparsePruner(object):
    """Performs pruning on the given model."""

    def __init__(self, model, masks, args, begin_prune_step, end_prune_step, inference_dataset_idx):
        self.model = model
        self.args = args
        self.sparsity_func_exponent = 3
        self.begin_prune_step = begin_prune_step
        self.end_prune_step = end_prune_step
        self.last_prune_step = begin_prune_step
        self.masks = masks

        finetune_again = False if not hasattr(args, 'finetune_again') else args.finetune_again
        if args.mode == 'prune' or args.mode == 'inference' or (args.mode == 'finetune' and finetune_again):
            self.current_dataset_idx = self.model.module.datasets.index(args.dataset) + 1
        elif args.mode == 'finetune':
            self.current_dataset_idx = len(self.model.module.datasets) - 1
        else:
            print('We do not support \'{}\' mode'.format(args.mode))
            sys.exit(-1)

        self.inference_dataset_idx = inference_dataset_idx
        # data parallel: a cpg_b200.ddp.GradAllReducer whose reduce() runs at the top of
        # do_weight_decay_and_make_grads_zero -- the call the unmodified Manager.train makes right after backward()
        self.grad_reducer = None
        self.fuse_grad_epilogue = True
        self.batched_staging = True     # build every layer's tensor-core weight operand in one launch
        self._prune_ws = {}
        self._mask_epoch = 0          # bumped whenever one of OUR kernels rewrites a task mask (they bypass torch's
        self._stats_cache = None      # version counters); see _stats
        self._stage_bufs = {}
        self._stage_events = {}
        self._stage_hook = None
        self._stage_post_hook = None
        self._task_view_idx = None      # task of the resident evaluation view (select_task), None = no view
        self.attach()
        return

    # ------------------------------------------------------------------ helpers
    def _sharable(self):
        for name, module in self.model.named_modules():
            if isinstance(module, nl.SharableConv2d) or isinstance(module, nl.SharableLinear):
                yield name, module

    def attach(self):
        """Tell every sharable layer which pruner owns its task mask (enables the fused epilogue)."""
        ref = weakref.ref(self)
        for name, module in self._sharable():
            module._cpg_pruner = ref
            module._cpg_name = name
            module._cpg_grads_final = False

        if self._stage_hook is None:
            # The hooks hold the pruner weakly (the model must not keep a discarded pruner and its staging
            # buffers alive) and there is one pair per model: a new pruner on the same model replaces the
            # previous one's hooks instead of stacking a second staging pass on every forward.
            for h in getattr(self.model, '_cpg_stage_hooks', ()):
                h.remove()

            def pre(model, inputs, _ref=ref):
                pr = _ref()
                return pr._prestage_hook(model, inputs) if pr is not None else None

            def post(model, inputs, output, _ref=ref):
                pr = _ref()
                return pr._poststage_hook(model, inputs, output) if pr is not None else None

            self._stage_hook = self.model.register_forward_pre_hook(pre)
            self._stage_post_hook = self.model.register_forward_hook(post)
            self.model._cpg_stage_hooks = (self._stage_hook, self._stage_post_hook)

    def detach(self):
        for name, module in self._sharable():
            module._cpg_pruner = None
            module._cpg_grads_final = False
            module._cpg_prestaged = None
        if self._stage_hook is not None:
            self._stage_hook.remove()
            self._stage_hook = None
            self._stage_post_hook.remove()
            self._stage_post_hook = None
            if getattr(self.model, '_cpg_stage_hooks', None) is not None:
                self.model._cpg_stage_hooks = ()

    def _prestage_hook(self, model, inputs):
        """Forward pre-hook of the whole model: the masked TF32 weight operands of ALL sharable layers
        in one launch (models/layers.py:101-103 for every layer at once).  Each layer consumes its
        operand once; anything not covered here is staged by the layer itself."""
        if not self.batched_staging:
            return None
        import ctypes
        lib = _lib.load()
        items = []
        for name, m in self._sharable():
            w = m._task_weight()        # the resident task view in evaluation mode after select_task(), else m.weight
            if not w.is_cuda or w.dtype != torch.float32 or not w.is_contiguous() or m.info['threshold_fn'] != 'binarizer':
                continue
            if w.dim() == 4:
                K, C, R, S = w.shape
                sh, sw = m.stride
                groups = m.groups
                C = C * groups
            else:
                (K, C), R, S, sh, sw, groups = w.shape, 1, 1, 1, 1, 1
            nbytes = lib.cpgb_staged_weight_bytes_for(K, C, R, S, sh, sw, groups)
            p = m.piggymask
            if lib.cpgb_weights_usable_raw_for(K, C, R, S, sh, sw, groups, 1 if p is not None else 0):
                continue        # consumed as is, nothing to stage
            if nbytes == 0 or (p is not None and (not p.is_contiguous() or p.device != w.device)):
                continue
            if lib.cpgb_intile_weight_shape(K, C, R, S, sh, sw, groups):
                continue        # masked in shared memory by the GEMM itself (CPGB_FLAG_W_INTILE); stages itself otherwise
            key = (name, str(w.device))
            buf = self._stage_bufs.get(key)
            if buf is None or buf.numel() != nbytes:
                buf = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
                self._stage_bufs[key] = buf
            items.append((m, w, p, buf, K, C, R, S, sh, sw, float(m.info['threshold'])))
        by_dev = {}
        for it in items:
            by_dev.setdefault(it[1].device, []).append(it)
        for dev, its in by_dev.items():
            n = len(its)
            vp, i32 = ctypes.c_void_p, ctypes.c_int32
            W = (vp * n)(*[it[1].data_ptr() for it in its])
            P = (vp * n)(*[(it[2].data_ptr() if it[2] is not None else None) for it in its])
            O = (vp * n)(*[it[3].data_ptr() for it in its])
            arr = lambda j: (i32 * n)(*[int(it[j]) for it in its])
            T = (ctypes.c_float * n)(*[it[10] for it in its])
            with torch.cuda.device(dev):
                # staged on the side stream: the first layers of the forward pass (the stem takes no staged
                # operand) run next to it; every consumer waits on the event before its first use
                main = torch.cuda.current_stream()
                side = _side_stream(dev) if STAGE_ON_SIDE_STREAM else main
                if side is not main:
                    side.wait_stream(main)
                _lib.check(lib.cpgb_stage_weights_batched(n, W, P, O, arr(4), arr(5), arr(6), arr(7), arr(8), arr(9), T,
                                                          side.cuda_stream), 'cpgb_stage_weights_batched')
                ev = None
                if side is not main:
                    ev = torch.cuda.Event()
                    ev.record(side)
                    self._stage_events[dev] = ev
            for it in its:
                it[0]._cpg_prestaged = (it[3], it[1].data_ptr(), it[2].data_ptr() if it[2] is not None else 0, ev)
        return None

    def _poststage_hook(self, model, inputs, output):
        """Forward hook of the whole model: rejoin the staging stream even if no layer consumed an operand
        (a forked stream must not be left dangling, e.g. under CUDA-graph capture)."""
        for dev, ev in self._stage_events.items():
            with torch.cuda.device(dev):
                torch.cuda.current_stream().wait_event(ev)
        self._stage_events = {}
        return None

    def _fuse_ctx_for(self, name):
        if not self.fuse_grad_epilogue or self.args.mode not in _MODES:
            return None
        mask = self.masks.get(name) if self.masks else None
        if mask is None or not mask.is_cuda:
            return None
        return FuseCtx(self._mask(name), self.current_dataset_idx, self.args.weight_decay,
                       _MODES[self.args.mode])

    def _mask(self, name):
        m = self.masks[name]
        if m.dtype != torch.uint8:
            raise _lib.CpgbError('task masks must be uint8 (torch.ByteTensor), as in the reference checkpoints')
        if not m.is_contiguous():
            raise _lib.CpgbError('task masks must be contiguous')
        return m

    @staticmethod
    def _dense(t, what):
        if not t.is_contiguous():
            raise _lib.CpgbError(f'{what} must be contiguous for the in-place kernels')
        return t

    def _scratch(self, device):
        key = str(device)
        if key not in self._prune_ws:
            lib = _lib.load()
            self._prune_ws[key] = torch.empty(lib.cpgb_prune_workspace_bytes(), dtype=torch.uint8, device=device)
        return self._prune_ws[key]

    # ------------------------------------------------------------------ a7
    def _launch_prune(self, weights, mask, pruning_ratio, info):
        lib = _lib.load()
        self._mask_epoch += 1
        ws = self._scratch(weights.device)
        with torch.cuda.device(weights.device):
            _lib.check(lib.cpgb_prune_select(_lib.ptr(self._dense(weights, 'weight')), _lib.ptr(mask),
                                             weights.numel(), self.current_dataset_idx, float(pruning_ratio),
                                             _lib.ptr(info), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                       'cpgb_prune_select')

    def _launch_prune_batched(self, layers, pruning_ratio, infos, sampled=None):
        """a7 for up to 64 layers in one call.  `sampled` (default: on unless CPGB_PRUNE_SAMPLED=0) picks the two-pass
        select (cpgb_prune_select_sampled); layers it reports as status 3 are finished by the caller with the
        four-pass radix select."""
        import ctypes
        import os
        lib = _lib.load()
        self._mask_epoch += 1
        nl_ = len(layers)
        dev = layers[0][1].weight.device
        W = (ctypes.c_void_p * nl_)(*[_lib.ptr(self._dense(m.weight.data, 'weight')) for _, m in layers])
        T = (ctypes.c_void_p * nl_)(*[_lib.ptr(self._mask(n)) for n, _ in layers])
        N = (ctypes.c_int64 * nl_)(*[m.weight.numel() for _, m in layers])
        if sampled is None:
            sampled = os.environ.get('CPGB_PRUNE_SAMPLED', '1') != '0'
        sampled = sampled and all(m.weight.numel() < 2 ** 32 for _, m in layers)
        nbytes = (lib.cpgb_prune_sampled_workspace_bytes if sampled else lib.cpgb_prune_batched_workspace_bytes)(nl_)
        key = ('sampled' if sampled else 'batched', str(dev), nbytes)
        if key not in self._prune_ws:
            self._prune_ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ws = self._prune_ws[key]
        fn, what = ((lib.cpgb_prune_select_sampled, 'cpgb_prune_select_sampled') if sampled else
                    (lib.cpgb_prune_select_batched, 'cpgb_prune_select_batched'))
        with torch.cuda.device(dev):
            _lib.check(fn(nl_, W, T, N, self.current_dataset_idx, float(pruning_ratio), _lib.ptr(infos), _lib.ptr(ws),
                          ws.numel(), _lib.stream_ptr()), what)

    @staticmethod
    def _exit_not_enough():
        print("Not enough weights for pruning, that is to say, too little space for new task, need expand the network.")
        sys.exit(2)

    def _pruning_mask(self, weights, mask, layer_name, pruning_ratio):
        """Ranks weights by magnitude. Sets all below kth to 0.
           Returns pruned mask.  (utils/prune.py:30-53)"""
        info = torch.zeros(4, dtype=torch.int64, device=weights.device)
        self._launch_prune(weights, mask, pruning_ratio, info)
        if int(info[0].item()) != 0:
            self._exit_not_enough()
        return mask

    # ------------------------------------------------------------------ a8 (python doubles, verbatim semantics)
    def _adjust_sparsity(self, curr_prune_step):
        p = min(1.0,
                max(0.0,
                    ((curr_prune_step - self.begin_prune_step)
                     / (self.end_prune_step - self.begin_prune_step))
                    ))
        sparsity = self.args.target_sparsity + \
            (self.args.initial_sparsity - self.args.target_sparsity) * pow(1 - p, self.sparsity_func_exponent)
        return sparsity

    def _time_to_update_masks(self, curr_prune_step):
        is_step_within_pruning_range = \
            (curr_prune_step >= self.begin_prune_step) and \
            (curr_prune_step <= self.end_prune_step)
        is_pruning_step = (
            self.last_prune_step + self.args.pruning_frequency) <= curr_prune_step
        return is_step_within_pruning_range and is_pruning_step

    def gradually_prune(self, curr_prune_step):
        if self._time_to_update_masks(curr_prune_step):
            self.last_prune_step = curr_prune_step
            curr_pruning_ratio = self._adjust_sparsity(curr_prune_step)
            layers = list(self._sharable())
            if layers:
                dev = layers[0][1].weight.device
                infos = torch.zeros(len(layers), 4, dtype=torch.int64, device=dev)
                # all layers in one batched call (7 launches per 64 layers); pruned weights are NOT
                # zeroed here (utils/prune.py:88 is commented out)
                for lo in range(0, len(layers), 64):
                    self._launch_prune_batched(layers[lo:lo + 64], curr_pruning_ratio, infos[lo:lo + 64])
                status = infos[:, 0].cpu()                  # one read-back per prune event
                if bool((status == 3).any()):
                    # the two-pass select could not bracket these layers (adversarial value distributions):
                    # finish them with the four-pass radix select
                    redo = [layers[i] for i in range(len(layers)) if int(status[i]) == 3]
                    infos2 = torch.zeros(len(redo), 4, dtype=torch.int64, device=dev)
                    for lo in range(0, len(redo), 64):
                        self._launch_prune_batched(redo[lo:lo + 64], curr_pruning_ratio, infos2[lo:lo + 64], sampled=False)
                    status = torch.cat([status[status != 3], infos2[:, 0].cpu()])
                if bool((status != 0).any()):
                    self._exit_not_enough()
        else:
            curr_pruning_ratio = self._adjust_sparsity(self.last_prune_step)
        return curr_pruning_ratio

    def one_shot_prune(self, one_shot_prune_perc):
        """utils/prune.py:94-109."""
        print('Pruning for dataset idx: %d' % (self.current_dataset_idx))
        print('Pruning each layer by removing %.2f%% of values' % (100 * one_shot_prune_perc))
        for name, module in self._sharable():
            self.masks[name] = self._pruning_mask(module.weight.data, self._mask(name), name,
                                                  pruning_ratio=one_shot_prune_perc)
        self.make_pruned_zero()
        return

    # ------------------------------------------------------------------ K12 statistics
    def _stats(self, with_piggy=False):
        """[#(T==0), #(T==idx), #(0<T<idx), #(0<T<idx and piggymask>0.005), numel] over all sharable
        layers: one launch per device and one read-back (the reference does a .sum() + .cpu() per layer
        and statistic, utils/prune.py:111-193)."""
        import ctypes
        lib = _lib.load()
        # Manager.train asks for the sparsity after EVERY batch (utils/manager.py:77-88) although the masks only change
        # at prune events: the counters are cached against (our own mutation counter, every mask's identity and
        # in-place version), so the per-batch call costs neither a kernel nor a host synchronisation.
        # Piggymask-dependent counters change with every optimizer step and are never cached.
        key = None
        if not with_piggy:
            key = (self._mask_epoch, self.inference_dataset_idx,
                   tuple((id(m), m._version) for m in (self.masks[n] for n, _ in self._sharable())))
            if self._stats_cache is not None and self._stats_cache[0] == key:
                return list(self._stats_cache[1])
        by_dev = {}
        for name, module in self._sharable():
            mask = self._mask(name)
            piggy = self._dense(module.piggymask.data, 'piggymask') if with_piggy else None
            by_dev.setdefault(mask.device, []).append((mask, piggy))
        total = [0, 0, 0, 0, 0]
        outs = []
        for dev, items in by_dev.items():
            n = len(items)
            out = torch.zeros(5, dtype=torch.int64, device=dev)
            T = (ctypes.c_void_p * n)(*[_lib.ptr(m) for m, _ in items])
            P = (ctypes.c_void_p * n)(*[(_lib.ptr(p) if p is not None else None) for _, p in items])
            N = (ctypes.c_int64 * n)(*[m.numel() for m, _ in items])
            with torch.cuda.device(dev):
                _lib.check(lib.cpgb_mask_stats_batched(n, T, P, N, self.inference_dataset_idx, _lib.ptr(out),
                                                       _lib.stream_ptr()), 'cpgb_mask_stats_batched')
            outs.append(out)
        for out in outs:
            for i, v in enumerate(out.cpu().tolist()):
                total[i] += int(v)
        if key is not None:
            self._stats_cache = (key, list(total))
        return total

    def calculate_sparsity(self):
        zero, cur, _, _, _ = self._stats()
        total = zero + cur
        return float(zero) / float(total) if total != 0 else 0.0

    def calculate_curr_task_ratio(self):
        _, cur, _, _, numel = self._stats()
        return float(cur) / numel * (self.args.network_width_multiplier ** 2)

    def calculate_zero_ratio(self):
        zero, _, _, _, numel = self._stats()
        return float(zero) / numel * (self.args.network_width_multiplier ** 2)

    def calculate_shared_part_ratio(self):
        _, _, shared, picked, _ = self._stats(with_piggy=True)
        return float(picked) / float(shared) if shared != 0 else 0.0

    # ------------------------------------------------------------------ a6
    def do_weight_decay_and_make_grads_zero(self):
        """Sets grads of fixed weights to 0.  (utils/prune.py:195-211)"""
        assert self.masks
        if self.grad_reducer is not None:
            self.grad_reducer.reduce()          # masks are rank-invariant: masking commutes with the mean over ranks
        lib = _lib.load()
        mode = _MODES.get(self.args.mode)
        for name, module in self._sharable():
            if module._cpg_grads_final:
                # the fused wgrad epilogue already produced (g*b + wd*W)[T==cur] / (g*W)[1<=T<cur]
                module._cpg_grads_final = False
                continue
            mask = self._mask(name)
            dW = module.weight.grad
            dP = module.piggymask.grad if module.piggymask is not None else None
            if mode is None:
                # reference: weight grads are always decayed+masked; piggymask grads only in finetune/prune
                dP, kmode = None, _lib.GRAD_PRUNE
            else:
                kmode = mode
            if dW is None and dP is None:
                continue
            with torch.cuda.device(mask.device):
                _lib.check(lib.cpgb_grad_epilogue(
                    _lib.ptr(self._dense(dW.data, 'weight.grad')) if dW is not None else None,
                    _lib.ptr(self._dense(dP.data, 'piggymask.grad')) if dP is not None else None,
                    _lib.ptr(self._dense(module.weight.data, 'weight')), _lib.ptr(mask), mask.numel(),
                    self.current_dataset_idx, float(self.args.weight_decay), kmode, _lib.stream_ptr()),
                    'cpgb_grad_epilogue')
        return

    # ------------------------------------------------------------------ a9 / a10
    def _zero_weights(self, inference_idx):
        lib = _lib.load()
        for name, module in self._sharable():
            weight = self._dense(module.weight.data, 'weight')
            mask = self._mask(name)
            if mask.device != weight.device:
                mask = mask.to(weight.device)   # the reference's `.cuda()` at utils/prune.py:228
            with torch.cuda.device(weight.device):
                _lib.check(lib.cpgb_apply_mask(_lib.ptr(weight), _lib.ptr(mask), weight.numel(), inference_idx,
                                               _lib.stream_ptr()), 'cpgb_apply_mask')

    def make_pruned_zero(self):
        """Makes pruned weights 0.  (utils/prune.py:213-221)"""
        assert self.masks
        self._zero_weights(255)
        return

    def apply_mask(self):
        """To be done to retrieve weights just for a particular dataset.  (utils/prune.py:223-231)"""
        if NONDESTRUCTIVE_APPLY_MASK:
            self.select_task(self.inference_dataset_idx)
            return
        if self._task_view_idx is not None:
            self.clear_task_view()  # a view taken before this call would show weights that no longer exist
        self._zero_weights(self.inference_dataset_idx)
        return

    # ------------------------------------------------------------------ N4: non-destructive evaluation predicate
    def select_task(self, inference_dataset_idx=None):
        """Non-destructive twin of `apply_mask()`: every sharable layer gets a resident copy
        W * [1 <= T <= inference_dataset_idx] of its weights (the tensor utils/prune.py:229-230 would leave in
        `weight.data`) which its forward pass reads in evaluation mode; `weight.data` itself -- and with it the
        weights of the later tasks -- stays as it is, so one resident model serves every task it holds:
        `select_task(j)` per task instead of a checkpoint reload.  Training-mode forward passes ignore the view.
        The copy is a snapshot: call again after the weights or the masks have changed (the per-task batch-norm /
        bias / PReLU tensors are the caller's to re-bind, as in utils/manager.py:266-320).  One copy + one
        `cpgb_apply_mask` launch per layer per call, nothing per forward pass."""
        assert self.masks
        idx = self.inference_dataset_idx if inference_dataset_idx is None else int(inference_dataset_idx)
        lib = _lib.load()
        for name, module in self._sharable():
            weight = self._dense(module.weight.data, 'weight')
            mask = self._mask(name)
            if mask.device != weight.device:
                mask = mask.to(weight.device)
            view = module._cpg_task_view
            if view is None or view.shape != weight.shape or view.device != weight.device:
                view = torch.empty_like(weight)
            view.copy_(weight)
            with torch.cuda.device(weight.device):
                _lib.check(lib.cpgb_apply_mask(_lib.ptr(view), _lib.ptr(mask), view.numel(), idx, _lib.stream_ptr()),
                           'cpgb_apply_mask')
            module._cpg_task_view = view
            module._cpg_prestaged = None
        self._task_view_idx = idx
        return idx

    def serve_task(self, dataset, shared_layer_info):
        """Switch ONE resident model to `dataset` (any task it holds) for evaluation, without touching `weight.data`:
        the task's classifier (`set_dataset`), its per-task tensors out of the checkpoint's `shared_layer_info` -- conv /
        linear biases, piggymasks, batch-norm weight / bias / running statistics, PReLU slopes: the re-binding of
        utils/manager.py:305-325 and CPG_cifar100_main_normal.py:282-289 -- and `select_task(index + 1)`.  The reference
        does the same per process: load_checkpoint_only_for_evaluate + apply_mask, one checkpoint load per task.  A task
        trained at another network width (its tensors are sub-blocks of the resident ones,
        utils/manager.py:281-297) needs a model built at that width and is refused here."""
        import torch.nn as nn
        inner = self.model.module
        task_id = inner.datasets.index(dataset) + 1
        info = shared_layer_info[dataset]

        def bound(new, like, what, param):
            if new is None:
                raise _lib.CpgbError(f'shared_layer_info[{dataset!r}] has no {what}')
            if tuple(new.shape) != tuple(like.shape):
                raise _lib.CpgbError(f'{what} of task {dataset!r} has shape {tuple(new.shape)}, the resident model '
                                     f'{tuple(like.shape)}: the task was trained at another network width')
            if new.device != like.device:
                new = new.detach().to(like.device)
            if param and not isinstance(new, nn.Parameter):
                new = nn.Parameter(new.detach(), requires_grad=False)
            return new

        def entry(key, name):
            return info.get(key, {}).get(name)

        for name, module in inner.named_modules():
            if isinstance(module, nl.SharableConv2d) or isinstance(module, nl.SharableLinear):
                if module.bias is not None:
                    module.bias = bound(entry('bias', name), module.bias, f'bias[{name}]', True)
                if task_id > 1:
                    module.piggymask = bound(entry('piggymask', name), module.weight, f'piggymask[{name}]', True)
                else:
                    module.piggymask = None             # the first task has none (models/layers.py:104-105)
            elif isinstance(module, nn.BatchNorm2d):
                for attr, key, param in (('running_mean', 'bn_layer_running_mean', False),
                                         ('running_var', 'bn_layer_running_var', False),
                                         ('weight', 'bn_layer_weight', True), ('bias', 'bn_layer_bias', True)):
                    cur = getattr(module, attr)
                    if cur is not None:
                        setattr(module, attr, bound(entry(key, name), cur, f'{key}[{name}]', param))
            elif isinstance(module, nn.PReLU):
                module.weight = bound(entry('prelu_layer_weight', name), module.weight, f'prelu_layer_weight[{name}]', True)
        inner.set_dataset(dataset)
        self.inference_dataset_idx = task_id
        return self.select_task(task_id)

    def clear_task_view(self):
        """Drop the copies `select_task()` made: evaluation reads `weight.data` again."""
        for name, module in self._sharable():
            module._cpg_task_view = None
            module._cpg_prestaged = None
        self._task_view_idx = None

    def task_view(self, inference_dataset_idx=None):
        """`with pruner.task_view(j): validate(...)` -- select_task(j) on entry, clear_task_view() on exit."""
        return _TaskView(self, inference_dataset_idx)

    def make_finetuning_mask(self):
        """Turns previously pruned weights into trainable weights for
           current dataset.  (utils/prune.py:233-243)"""
        assert self.masks
        self.current_dataset_idx += 1
        self._mask_epoch += 1
        lib = _lib.load()
        for name, module in self._sharable():
            mask = self._mask(name)
            with torch.cuda.device(mask.device):
                _lib.check(lib.cpgb_make_finetuning_mask(_lib.ptr(mask), mask.numel(), self.current_dataset_idx,
                                                         _lib.stream_ptr()), 'cpgb_make_finetuning_mask')
