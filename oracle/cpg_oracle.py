"""CPU oracle for the CPG masked-conv train/prune hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cpg_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs use it, and only as the checker / the CPU arm.

It restates, in plain torch-CPU / numpy, the algorithm of the reference
(ivclab/CPG) for SURVEY.md section 8(a) rows a1..a11.  Every function cites the
reference file:line it follows (paths relative to /root/reference).

Pinning: the reference ships no tests, golden vectors or fixtures, so
the oracle is pinned against outputs of the *live* reference imported in the build
container: ``tests/golden/make_golden.py`` runs the unmodified reference modules
(``models.layers``, ``utils.prune``, ``utils.manager``) on seeded inputs and stores
their outputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays
the same inputs through this file and demands bit-exact masks / cut values and
bit-exact fp32 tensors (same torch CPU kernels underneath).

The convolution / GEMM arithmetic itself is not in the reference tree: it lives in
PyTorch ("PyTorch>=1.0", reference README.md:33-36; no lockfile), call sites
models/layers.py:108 and :194.  The oracle therefore calls the same third-party
``torch.nn.functional.conv2d`` / ``linear`` on CPU (torch 2.11.0 in this image) and
additionally offers ``conv2d_f64`` (float64 accumulate) as the exact yardstick for
tolerance statements.

The next row built in round 1 (nn.BatchNorm2d -> nn.ReLU -> nn.MaxPool2d(2, 2) after
every convolution, models/vgg.py:95-122) is restated at the end of this file in float64 numpy
(``bn_relu_pool_forward`` / ``bn_relu_pool_backward``) and pinned against the stock torch modules on the
CPU by tests/test_oracle_norm.py.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

DEFAULT_THRESHOLD = 5e-3  # models/layers.py:9


class NotEnoughWeights(Exception):
    """Raised where the reference prints its message and calls sys.exit(2)
    (utils/prune.py:38-42): kthvalue(k) with k outside 1..len(pool)."""


# ----------------------------------------------------------------------------
# a1/a2  Binarizer                                   models/layers.py:11-23
# ----------------------------------------------------------------------------
def binarize(p: torch.Tensor, threshold: float = DEFAULT_THRESHOLD) -> torch.Tensor:
    """b = 0 where p <= thr, 1 where p > thr, NaN stays NaN (models/layers.py:16-18).

    The comparison is done by torch between an fp32 tensor and a python double;
    torch casts the scalar to the tensor dtype, so the effective threshold is
    fp32(5e-3) = 0.004999999888241291.
    """
    thr = np.float32(threshold)
    a = p.detach().cpu().numpy().astype(np.float32, copy=True)
    out = a.copy()
    out[a <= thr] = 0.0
    out[a > thr] = 1.0
    return torch.from_numpy(out)


def binarize_backward(grad_out: torch.Tensor) -> torch.Tensor:
    """Straight-through estimator: identity (models/layers.py:21-23)."""
    return grad_out


# ----------------------------------------------------------------------------
# a3/a4  SharableConv2d forward + autograd            models/layers.py:98-109
# ----------------------------------------------------------------------------
def effective_weight(w: torch.Tensor, p: Optional[torch.Tensor],
                     threshold: float = DEFAULT_THRESHOLD) -> torch.Tensor:
    """W_eff = Binarizer(P) * W if a piggymask exists, else W
    (models/layers.py:99-105, 185-192).  Note: the task mask T is NOT applied in
    the forward."""
    if p is None:
        return w
    return binarize(p, threshold) * w


def conv2d_forward(x, w, p, bias, stride=1, padding=0, dilation=1, groups=1,
                   threshold=DEFAULT_THRESHOLD):
    """models/layers.py:98-109."""
    return F.conv2d(x, effective_weight(w, p, threshold), bias, stride, padding, dilation, groups)


def _autograd_backward(fwd, x, w, p, bias, dy, need_dx):
    """The reference has no backward source: it is torch autograd over the forward
    expression (MulBackward + ConvolutionBackward/AddmmBackward).  The oracle
    therefore differentiates its own restated forward with the same autograd."""
    xs = x.detach().clone().requires_grad_(need_dx)
    ws = w.detach().clone().requires_grad_(True)
    ps = p.detach().clone().requires_grad_(True) if p is not None else None
    bs = bias.detach().clone().requires_grad_(True) if bias is not None else None

    class _B(torch.autograd.Function):  # models/layers.py:11-23
        @staticmethod
        def forward(ctx, inputs, threshold):
            out = inputs.clone()
            out[inputs.le(threshold)] = 0
            out[inputs.gt(threshold)] = 1
            return out

        @staticmethod
        def backward(ctx, grad_out):
            return grad_out, None

    weff = _B.apply(ps, DEFAULT_THRESHOLD) * ws if ps is not None else ws
    y = fwd(xs, weff, bs)
    y.backward(dy)
    return (xs.grad if need_dx else None, ws.grad, ps.grad if ps is not None else None,
            bs.grad if bs is not None else None)


def conv2d_backward(x, w, p, bias, dy, stride=1, padding=0, dilation=1, groups=1,
                    threshold=DEFAULT_THRESHOLD, need_dx=True):
    """Autograd of a3: g = wgrad(x, dy); dW = g*b; dP = g*W;
    dX = dgrad(dy, W_eff); dbias = sum(dy).  Returns (dx, dW, dP, dbias, g) where g is the
    raw weight gradient (dL/dW_eff)."""
    dx, dW, dP, dbias = _autograd_backward(
        lambda xs, we, bs: F.conv2d(xs, we, bs, stride, padding, dilation, groups),
        x, w, p, bias, dy, need_dx)
    g = torch.nn.grad.conv2d_weight(x, w.shape, dy, stride, padding, dilation, groups)
    return dx, dW, dP, dbias, g


def conv2d_f64(x, w_eff, bias, stride=1, padding=0, dilation=1, groups=1):
    """float64 yardstick for tolerance statements (not a reference restatement)."""
    y = F.conv2d(x.double(), w_eff.double(), None if bias is None else bias.double(),
                 stride, padding, dilation, groups)
    return y


# ----------------------------------------------------------------------------
# a5  SharableLinear                                  models/layers.py:184-194
# ----------------------------------------------------------------------------
def linear_forward(x, w, p, bias, threshold=DEFAULT_THRESHOLD):
    return F.linear(x, effective_weight(w, p, threshold), bias)


def linear_backward(x, w, p, bias, dy, threshold=DEFAULT_THRESHOLD, need_dx=True):
    dx, dW, dP, dbias = _autograd_backward(lambda xs, we, bs: F.linear(xs, we, bs),
                                           x, w, p, bias, dy, need_dx)
    g = dy.reshape(-1, dy.shape[-1]).t() @ x.reshape(-1, x.shape[-1])
    return dx, dW, dP, dbias, g


# ----------------------------------------------------------------------------
# a6  do_weight_decay_and_make_grads_zero             utils/prune.py:195-211
# ----------------------------------------------------------------------------
def weight_decay_and_mask_grads(dW, dP, w, t, cur: int, weight_decay: float, mode: str):
    """In-place on dW / dP like the reference:
    dW += wd*W (utils/prune.py:203; fp32 axpy), dW[T != cur] = 0 (:204-205);
    finetune: dP[T == 0 or T >= cur] = 0 (:207-208); prune: dP = 0 (:209-210)."""
    if dW is not None:
        dW.add_(w, alpha=weight_decay)
        dW[t.ne(cur)] = 0
    if dP is not None:
        if mode == 'finetune':
            dP[t.eq(0) | t.ge(cur)] = 0
        elif mode == 'prune':
            dP.fill_(0)
    return dW, dP


def fused_weight_grads(g, w, p, t, cur: int, weight_decay: float, mode: str,
                       threshold=DEFAULT_THRESHOLD):
    """What optimizers.step() sees after a4 followed by a6, written as one
    expression (this is what the fused wgrad epilogue must equal):
    dW = (g*b + wd*W) * [T==cur];  dP = g*W*[1<=T<cur] (finetune) / 0 (prune)."""
    b = binarize(p, threshold) if p is not None else None
    dW = g * b if b is not None else g.clone()
    dP = g * w if p is not None else None
    return weight_decay_and_mask_grads(dW, dP, w, t, cur, weight_decay, mode)


# ----------------------------------------------------------------------------
# a7  _pruning_mask                                    utils/prune.py:30-53
# ----------------------------------------------------------------------------
def pruning_cutoff(w: torch.Tensor, t: torch.Tensor, cur: int, ratio: float):
    """pool = W[T==cur | T==0] (utils/prune.py:35); k = round(ratio*len(pool))
    with python-3 banker's rounding (:37); cut = k-th smallest |pool| (1-indexed,
    :39).  Returns (cut, k, pool_size).  k outside 1..len(pool) is the exit-2
    path."""
    wf = w.detach().cpu().numpy().reshape(-1)
    tf = t.detach().cpu().numpy().reshape(-1)
    pool = np.abs(wf[(tf == cur) | (tf == 0)])
    k = round(ratio * pool.size)
    if k < 1 or k > pool.size:
        raise NotEnoughWeights(f'k={k} pool={pool.size}')
    cut = np.partition(pool, k - 1)[k - 1]
    return np.float32(cut), int(k), int(pool.size)


def pruning_mask(w: torch.Tensor, t: torch.Tensor, cur: int, ratio: float):
    """utils/prune.py:30-53: T[(|W| <= cut) & (T == cur)] = 0, in place; returns
    (T, cut, k, pool_size)."""
    cut, k, pool = pruning_cutoff(w, t, cur, ratio)
    remove = (w.abs() <= float(cut)) & t.eq(cur)
    t[remove] = 0
    return t, cut, k, pool


# ----------------------------------------------------------------------------
# a11  one_shot_prune                                   utils/prune.py:94-109
# ----------------------------------------------------------------------------
def one_shot_prune(weights, masks, cur: int, ratio: float):
    """utils/prune.py:94-109 over parallel lists of layer weights / task masks (named_modules order): a7 at the
    fixed ratio (:102-105), then W[T == 0] = 0 (:108) -- every free element, not only the freshly pruned ones."""
    for w, t in zip(weights, masks):
        pruning_mask(w, t, cur, ratio)
        w[t.eq(0)] = 0.0
    return weights, masks


# ----------------------------------------------------------------------------
# a8  schedule                                          utils/prune.py:55-92
# ----------------------------------------------------------------------------
def adjust_sparsity(step, begin, end, initial_sparsity, target_sparsity, exponent=3):
    """utils/prune.py:55-66 (python doubles)."""
    p = min(1.0, max(0.0, ((step - begin) / (end - begin))))
    return target_sparsity + (initial_sparsity - target_sparsity) * pow(1 - p, exponent)


def time_to_update_masks(step, begin, end, last_prune_step, pruning_frequency):
    """utils/prune.py:68-76."""
    in_range = (step >= begin) and (step <= end)
    return in_range and (last_prune_step + pruning_frequency) <= step


# ----------------------------------------------------------------------------
# a9/a10  weight zeroing and finetuning mask           utils/prune.py:213-243
# ----------------------------------------------------------------------------
def make_pruned_zero(w, t):
    w[t.eq(0)] = 0.0  # utils/prune.py:220
    return w


def apply_mask(w, t, inference_idx: int):
    w[t.eq(0)] = 0.0                    # utils/prune.py:229
    w[t.gt(inference_idx)] = 0.0        # utils/prune.py:230
    return w


def make_finetuning_mask(t, cur: int):
    """utils/prune.py:233-243: cur += 1; T[T==0] = cur.  Returns new cur."""
    cur += 1
    t[t.eq(0)] = cur
    return cur


# ----------------------------------------------------------------------------
# K12 statistics                                        utils/prune.py:111-193
# ----------------------------------------------------------------------------
def mask_stats(t, inference_idx: int, p=None, threshold=DEFAULT_THRESHOLD):
    """Counts used by calculate_{sparsity,curr_task_ratio,zero_ratio,shared_part_ratio}:
    returns dict(zero, cur, shared, shared_picked, numel)."""
    tn = t.detach().cpu().numpy().reshape(-1)
    out = dict(zero=int((tn == 0).sum()), cur=int((tn == inference_idx).sum()),
               shared=int(((tn > 0) & (tn < inference_idx)).sum()), numel=int(tn.size),
               shared_picked=0)
    if p is not None:
        pn = p.detach().cpu().numpy().reshape(-1)
        out['shared_picked'] = int(((tn > 0) & (tn < inference_idx) & (pn > np.float32(0.005))).sum())
    return out


# ----------------------------------------------------------------------------
# Oracle modules: same construction/forward as models/layers.py:43-109,147-194,
# used to build the CPU arm of bench.py (cpu_baseline kind "port").
# ----------------------------------------------------------------------------
class _STE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, threshold):
        out = inputs.clone()
        out[inputs.le(threshold)] = 0
        out[inputs.gt(threshold)] = 1
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return grad_out, None


class OracleSharableConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 dilation=1, groups=1, bias=True, mask_init='1s', mask_scale=1e-2,
                 threshold_fn='binarizer', threshold=None):
        super().__init__()
        pair = lambda v: v if isinstance(v, tuple) else (v, v)
        if in_channels % groups != 0:
            raise ValueError('in_channels must be divisible by groups')
        if out_channels % groups != 0:
            raise ValueError('out_channels must be divisible by groups')
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = pair(kernel_size), pair(stride)
        self.padding, self.dilation, self.groups = pair(padding), pair(dilation), groups
        self.info = {'threshold_fn': threshold_fn,
                     'threshold': DEFAULT_THRESHOLD if threshold is None else threshold}
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.piggymask = None

    def forward(self, input, layer_info=None, name=None):
        if self.piggymask is not None:
            weight = _STE.apply(self.piggymask, self.info['threshold']) * self.weight
        else:
            weight = self.weight
        return F.conv2d(input, weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


class OracleSharableLinear(nn.Module):
    def __init__(self, in_features, out_features, bias=True, mask_init='1s', mask_scale=1e-2,
                 threshold_fn='binarizer', threshold=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.info = {'threshold_fn': threshold_fn,
                     'threshold': DEFAULT_THRESHOLD if threshold is None else threshold}
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter('bias', None)
        self.piggymask = None

    def forward(self, input):
        if self.piggymask is not None:
            weight = _STE.apply(self.piggymask, self.info['threshold']) * self.weight
        else:
            weight = self.weight
        return F.linear(input, weight, self.bias)


class OraclePruner:
    """Restatement of utils/prune.py SparsePruner's hot methods over a model built
    from Oracle* layers (module order = named_modules() order, utils/prune.py:84)."""

    def __init__(self, model, masks, *, mode, weight_decay, cur, inference_idx,
                 begin_prune_step=0, end_prune_step=1, initial_sparsity=0.0,
                 target_sparsity=0.0, pruning_frequency=10):
        self.model, self.masks, self.mode, self.weight_decay = model, masks, mode, weight_decay
        self.current_dataset_idx, self.inference_dataset_idx = cur, inference_idx
        self.begin_prune_step, self.end_prune_step = begin_prune_step, end_prune_step
        self.last_prune_step = begin_prune_step
        self.initial_sparsity, self.target_sparsity = initial_sparsity, target_sparsity
        self.pruning_frequency = pruning_frequency

    def _layers(self):
        for name, m in self.model.named_modules():
            if isinstance(m, (OracleSharableConv2d, OracleSharableLinear)):
                yield name, m

    def do_weight_decay_and_make_grads_zero(self):
        for name, m in self._layers():
            t = self.masks[name]
            dW = m.weight.grad
            dP = m.piggymask.grad if m.piggymask is not None else None
            weight_decay_and_mask_grads(dW, dP, m.weight.data, t, self.current_dataset_idx,
                                        self.weight_decay, self.mode)

    def gradually_prune(self, step):
        if time_to_update_masks(step, self.begin_prune_step, self.end_prune_step,
                                self.last_prune_step, self.pruning_frequency):
            self.last_prune_step = step
            ratio = adjust_sparsity(step, self.begin_prune_step, self.end_prune_step,
                                    self.initial_sparsity, self.target_sparsity)
            for name, m in self._layers():
                pruning_mask(m.weight.data, self.masks[name], self.current_dataset_idx, ratio)
        else:
            ratio = adjust_sparsity(self.last_prune_step, self.begin_prune_step,
                                    self.end_prune_step, self.initial_sparsity, self.target_sparsity)
        return ratio

    def apply_mask(self):
        for name, m in self._layers():
            apply_mask(m.weight.data, self.masks[name], self.inference_dataset_idx)

    def make_finetuning_mask(self):
        cur = self.current_dataset_idx
        for name, m in self._layers():
            make_finetuning_mask(self.masks[name], cur)
        self.current_dataset_idx = cur + 1


# ----------------------------------------------------------------------------
# the consumer of every masked convolution
#   nn.BatchNorm2d -> nn.ReLU(inplace=True) (-> nn.MaxPool2d(kernel_size=2, stride=2))
#   models/vgg.py:95-122 (make_layers), models/resnet.py:60-100
# The arithmetic lives in PyTorch (same pinned dependency as the convolution); this is its published
# definition restated in float64 numpy, pinned against torch CPU by tests/test_oracle_norm.py.
# ----------------------------------------------------------------------------
def bn_relu_pool_forward(x, gamma, beta, running_mean, running_var, training: bool, momentum: float = 0.1,
                         eps: float = 1e-5, relu: bool = True, pool: bool = False):
    """x [N, C, H, W] (numpy).  Returns (y, new_running_mean, new_running_var, mean, rstd) in float64.
    Training: batch statistics with the BIASED variance normalise; the running statistics are updated
    with `momentum` and the UNBIASED variance.  Evaluation: the running statistics normalise."""
    import numpy as np
    x = np.asarray(x, dtype=np.float64)
    C = x.shape[1]
    g = np.ones(C) if gamma is None else np.asarray(gamma, dtype=np.float64)
    b = np.zeros(C) if beta is None else np.asarray(beta, dtype=np.float64)
    rm = None if running_mean is None else np.asarray(running_mean, dtype=np.float64)
    rv = None if running_var is None else np.asarray(running_var, dtype=np.float64)
    if training:
        m = x.shape[0] * x.shape[2] * x.shape[3]
        mean = x.mean(axis=(0, 2, 3))
        var = x.var(axis=(0, 2, 3))                       # biased
        if rm is not None:
            rm = (1 - momentum) * rm + momentum * mean
            rv = (1 - momentum) * rv + momentum * var * m / max(m - 1, 1)
    else:
        mean, var = rm, rv
    rstd = 1.0 / np.sqrt(var + eps)
    y = (x - mean[None, :, None, None]) * (rstd * g)[None, :, None, None] + b[None, :, None, None]
    if relu:
        y = np.maximum(y, 0.0)
    if pool:
        n, c, h, w = y.shape
        y = y[:, :, :h // 2 * 2, :w // 2 * 2].reshape(n, c, h // 2, 2, w // 2, 2).max(axis=(3, 5))
    return y, rm, rv, mean, rstd


def bn_relu_pool_backward(x, dy, gamma, beta, mean, rstd, training: bool, relu: bool = True, pool: bool = False):
    """Gradients of bn_relu_pool_forward w.r.t. x, gamma, beta (float64 numpy).  With g the gradient that
    reaches the batch-norm output (dy routed to the FIRST maximum of each 2x2 window in (h, w) order, then
    masked by y > 0) and xhat = (x - mean) * rstd:
        dbeta = sum g;  dgamma = sum g * xhat;
        training: dx = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat));  evaluation: dx = gamma * rstd * g."""
    import numpy as np
    x = np.asarray(x, dtype=np.float64)
    dy = np.asarray(dy, dtype=np.float64)
    C = x.shape[1]
    gm = np.ones(C) if gamma is None else np.asarray(gamma, dtype=np.float64)
    bt = np.zeros(C) if beta is None else np.asarray(beta, dtype=np.float64)
    bc = lambda v: np.asarray(v, dtype=np.float64)[None, :, None, None]
    xhat = (x - bc(mean)) * bc(rstd)
    z = xhat * bc(gm) + bc(bt)
    if pool:
        n, c, h, w = z.shape
        win = z.reshape(n, c, h // 2, 2, w // 2, 2).transpose(0, 1, 2, 4, 3, 5).reshape(n, c, h // 2, w // 2, 4)
        first = win.argmax(axis=-1)                      # numpy argmax returns the first maximum
        g4 = np.zeros_like(win)
        np.put_along_axis(g4, first[..., None], dy[..., None], axis=-1)
        g = g4.reshape(n, c, h // 2, w // 2, 2, 2).transpose(0, 1, 2, 4, 3, 5).reshape(n, c, h, w)
    else:
        g = dy.copy()
    if relu:
        g = g * (z > 0)
    dbeta = g.sum(axis=(0, 2, 3))
    dgamma = (g * xhat).sum(axis=(0, 2, 3))
    if training:
        m = x.shape[0] * x.shape[2] * x.shape[3]
        dx = bc(gm * np.asarray(rstd, dtype=np.float64)) * (g - bc(dbeta / m) - xhat * bc(dgamma / m))
    else:
        dx = bc(gm * np.asarray(rstd, dtype=np.float64)) * g
    return dx, dgamma, dbeta


# ----------------------------------------------------------------------------
# SURVEY 8(f) N4, residual blocks: out = bn(out); out += identity; out = relu(out)
#   models/resnet.py:50-55 (BasicBlock.forward), :92-98 (Bottleneck.forward)
# Restated on top of the batch-norm above; pinned against the stock torch modules by tests/test_oracle_norm.py.
# ----------------------------------------------------------------------------
def bn_add_relu_forward(x, res, gamma, beta, running_mean, running_var, training: bool, momentum: float = 0.1,
                        eps: float = 1e-5):
    """y = max(0, batch_norm(x) + res).  Returns (y, new_running_mean, new_running_var, mean, rstd), float64."""
    import numpy as np
    z, rm, rv, mean, rstd = bn_relu_pool_forward(x, gamma, beta, running_mean, running_var, training, momentum, eps,
                                                 relu=False, pool=False)
    return np.maximum(z + np.asarray(res, dtype=np.float64), 0.0), rm, rv, mean, rstd


def bn_add_relu_backward(x, y, dy, gamma, beta, mean, rstd, training: bool):
    """Gradients of bn_add_relu_forward: g = dy * [y > 0] is the gradient of the identity branch (dres) and enters the
    batch-norm backward as its output gradient.  Returns (dx, dres, dgamma, dbeta)."""
    import numpy as np
    g = np.asarray(dy, dtype=np.float64) * (np.asarray(y, dtype=np.float64) > 0)
    dx, dgamma, dbeta = bn_relu_pool_backward(x, g, gamma, beta, mean, rstd, training, relu=False, pool=False)
    return dx, g, dgamma, dbeta
