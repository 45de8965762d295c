"""Drop-in replacement of the reference's ``models/layers.py`` (same names, constructor
arguments, attributes and state_dict keys), with the forward/backward running on the
hand-written sm_100a kernels behind include/cpgb200.h.

Usage in a CPG checkout:  ``import cpg_b200.layers as nl``  instead of
``import models.layers as nl`` (or ``cpg_b200.install()`` which aliases the module), see
INTEGRATION.md.  Reference interface: models/layers.py:9-218.
"""
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair
from torch.nn.parameter import Parameter

from . import _lib
from .functional import (DEFAULT_THRESHOLD, Binarizer, FuseCtx, MaskedConv2dFn, MaskedLinearFn,
                         Ternarizer, is_tf32)

__all__ = ['DEFAULT_THRESHOLD', 'Binarizer', 'Ternarizer', 'SharableConv2d', 'SharableLinear']

# Conv outputs are produced physically NHWC (torch.channels_last; logically still NCHW) so
# that the next masked conv can TMA-load them; BatchNorm/ReLU/MaxPool keep the format.
OUTPUT_CHANNELS_LAST = True


class _SharableBase(nn.Module):
    def _init_common(self, mask_init, mask_scale, threshold_fn, threshold):
        self.mask_scale = mask_scale
        self.mask_init = mask_init
        if threshold is None:
            threshold = DEFAULT_THRESHOLD
        self.info = {'threshold_fn': threshold_fn, 'threshold': threshold}
        self._cpg_pruner = None       # set by cpg_b200.prune.SparsePruner
        self._cpg_name = None
        self._cpg_grads_final = False
        self._cpg_prestaged = None    # (buffer, weight ptr, piggymask ptr): one-shot, set by the model-level hook
        self._cpg_grad_slot = None    # set by cpg_b200.ddp.GradAllReducer: where the wgrad epilogue writes
        self._cpg_task_view = None    # set by cpg_b200.prune.SparsePruner.select_task: W * [1 <= T <= task], eval only

    def _finish_init(self, threshold_fn, threshold):
        # Give real-valued mask weights per task to manage the shared part from previous tasks:
        # a plain None attribute that becomes a registered Parameter when the training script
        # assigns one (CPG_cifar100_main_normal.py:262-270).
        self.piggymask = None
        if threshold_fn == 'binarizer':
            self.threshold_fn = Binarizer.apply
        elif threshold_fn == 'ternarizer':
            print('Calling ternarizer with threshold:', threshold)
            self.threshold_fn = Ternarizer.apply

    def _fuse_ctx(self):
        """FuseCtx when an attached SparsePruner wants weight decay + grad masking folded
        into the wgrad epilogue, else None."""
        ref = self._cpg_pruner          # weakref.ref set by SparsePruner.attach()
        pr = ref() if ref is not None else None
        if pr is None or not torch.is_grad_enabled():
            return None
        # nn.DataParallel replicas (CPG_cifar100_main_normal.py:199) copy __dict__, so they see the pruner too,
        # but their weights are non-leaf views on other devices, their gradients are summed by autograd's
        # ReduceAddCoalesced afterwards, and the task mask lives on the pruner's device: a replica hands out the
        # plain autograd gradients (immediate stream join) and the pruner finishes them once on the original.
        w = self.weight
        if getattr(self, '_is_replica', False) or not isinstance(w, Parameter) or not w.is_leaf:
            return None
        fc = pr._fuse_ctx_for(self._cpg_name)
        if fc is not None and fc.tmask.device != w.device:
            return None
        return fc

    def _owner(self, weight):
        """`self` when this module owns leaf parameters the backward pass may finalise (fused epilogue flag,
        gradient slots, deferred stream join); None for nn.DataParallel replicas and rebound weights."""
        if weight is not self.weight or getattr(self, '_is_replica', False):
            return None
        return self if (isinstance(weight, Parameter) and weight.is_leaf) else None

    def _take_prestaged(self, weight, piggy):
        """The operand the model-level batched staging built for THIS forward pass, if any (consumed:
        a layer called again without the hook re-stages itself)."""
        pre, self._cpg_prestaged = self._cpg_prestaged, None
        if pre is None:
            return None
        buf, wptr, pptr, ev = pre
        if wptr != weight.data_ptr() or pptr != (piggy.data_ptr() if piggy is not None else 0):
            return None          # e.g. a DataParallel replica on another device
        if ev is not None:       # the operand was built on the staging stream
            with torch.cuda.device(weight.device):
                torch.cuda.current_stream().wait_event(ev)
        return buf

    def _task_weight(self):
        """The weight tensor of this forward pass: `self.weight`, or -- in evaluation mode, after
        SparsePruner.select_task(t) -- the resident copy W * [1 <= T <= t] of it (what utils/prune.py:223-231 leaves
        in `weight.data`, without destroying the later tasks' weights)."""
        view = getattr(self, '_cpg_task_view', None)
        if view is None or self.training:
            return self.weight
        if view.shape != self.weight.shape:
            raise _lib.CpgbError('the task view of this layer was built for another weight shape: call '
                                 'SparsePruner.select_task() again')
        if view.device != self.weight.device:
            if getattr(self, '_is_replica', False):
                # nn.DataParallel replica (CPG_cifar100_main_normal.py:199): its weight was broadcast from the original's
                # device for this forward pass, the view follows the same way
                return view.to(self.weight.device)
            raise _lib.CpgbError('the task view of this layer lives on another device than its weight: call '
                                 'SparsePruner.select_task() again')
        return view

    def _effective(self):
        """(weight, piggymask) to hand to the fused kernels.  The ternarizer (dead code in the
        reference) is applied with torch ops and then treated as 'no piggymask'."""
        weight = self._task_weight()
        if self.piggymask is not None and self.info['threshold_fn'] != 'binarizer':
            return self.threshold_fn(self.piggymask, self.info['threshold']) * weight, None
        return weight, self.piggymask

    def _staged_for(self, weight, piggy):
        """The batched staging's operand for `weight` if the model-level hook built one (it stages the tensor
        _task_weight() returns)."""
        if weight is self.weight or weight is getattr(self, '_cpg_task_view', None):
            return self._take_prestaged(weight, piggy)
        self._cpg_prestaged = None      # built for another tensor (e.g. a replica's copy of the task view)
        return None


class SharableConv2d(_SharableBase):
    """Modified conv with masks for weights (models/layers.py:43-145)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1,
                 padding=0, dilation=1, groups=1, bias=True,
                 mask_init='1s', mask_scale=1e-2,
                 threshold_fn='binarizer', threshold=None):
        super(SharableConv2d, self).__init__()
        kernel_size = _pair(kernel_size)
        stride = _pair(stride)
        padding = _pair(padding)
        dilation = _pair(dilation)
        self._init_common(mask_init, mask_scale, threshold_fn, threshold)
        if in_channels % groups != 0:
            raise ValueError('in_channels must be divisible by groups')
        if out_channels % groups != 0:
            raise ValueError('out_channels must be divisible by groups')
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.transposed = False
        self.output_padding = _pair(0)
        self.groups = groups
        # uninitialised storage, like the reference (models initialise it)
        self.weight = Parameter(torch.empty(out_channels, in_channels // groups, *kernel_size),
                                requires_grad=True)
        if bias:
            self.bias = Parameter(torch.empty(out_channels), requires_grad=True)
        else:
            self.register_parameter('bias', None)
        self._finish_init(threshold_fn, self.info['threshold'])

    def forward(self, input, layer_info=None, name=None):
        weight, piggy = self._effective()
        fuse = self._fuse_ctx() if piggy is self.piggymask and weight is self.weight else None
        pre = self._staged_for(weight, piggy)
        return MaskedConv2dFn.apply(input, weight, piggy, self.bias, self.stride, self.padding,
                                    self.dilation, self.groups, float(self.info['threshold']), fuse,
                                    self._owner(weight), OUTPUT_CHANNELS_LAST, pre, is_tf32(input),
                                    bool(getattr(self, '_cpg_emit_colstats', False)) and self.training)

    def __repr__(self):
        s = ('{name} ({in_channels}, {out_channels}, kernel_size={kernel_size}'
             ', stride={stride}')
        if self.padding != (0,) * len(self.padding):
            s += ', padding={padding}'
        if self.dilation != (1,) * len(self.dilation):
            s += ', dilation={dilation}'
        if self.output_padding != (0,) * len(self.output_padding):
            s += ', output_padding={output_padding}'
        if self.groups != 1:
            s += ', groups={groups}'
        if self.bias is None:
            s += ', bias=False'
        s += ')'
        return s.format(name=self.__class__.__name__, **self.__dict__)


class SharableLinear(_SharableBase):
    """Modified linear layer (models/layers.py:147-218)."""

    def __init__(self, in_features, out_features, bias=True,
                 mask_init='1s', mask_scale=1e-2,
                 threshold_fn='binarizer', threshold=None):
        super(SharableLinear, self).__init__()
        self.in_features = in_features
        self.out_features = out_features
        self._init_common(mask_init, mask_scale, threshold_fn, threshold)
        self.weight = Parameter(torch.empty(out_features, in_features), requires_grad=True)
        if bias:
            self.bias = Parameter(torch.empty(out_features), requires_grad=True)
        else:
            self.register_parameter('bias', None)
        self._finish_init(threshold_fn, self.info['threshold'])

    def forward(self, input):
        weight, piggy = self._effective()
        fuse = self._fuse_ctx() if piggy is self.piggymask and weight is self.weight else None
        pre = self._staged_for(weight, piggy)
        return MaskedLinearFn.apply(input, weight, piggy, self.bias, float(self.info['threshold']), fuse,
                                    self._owner(weight), pre, is_tf32(input))

    def __repr__(self):
        return self.__class__.__name__ + '(' \
            + 'in_features=' + str(self.in_features) \
            + ', out_features=' + str(self.out_features) + ')'
