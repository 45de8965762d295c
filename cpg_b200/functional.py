"""autograd.Functions over the C ABI (include/cpgb200.h).

One Function per masked layer type replaces the reference's three-node autograd graph
``Binarizer.apply -> mul -> F.conv2d / F.linear`` (models/layers.py:101-108, 187-194): the
forward evaluates the piggyback predicate inside the convolution kernels, the backward
produces dX, dW, dP (straight-through: dP = g * W, models/layers.py:21-23) and dbias, with
the pruner's weight-decay / grad-mask step (utils/prune.py:195-211) optionally folded into
the same epilogue.
"""
import os

import torch
from torch.autograd.function import once_differentiable

from . import _lib

DEFAULT_THRESHOLD = 5e-3  # models/layers.py:9
CL = torch.channels_last


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _up4(c):
    return (c + 3) // 4 * 4


def nhwc_pixel_stride(t):
    """Elements between consecutive pixels if the logical-NCHW tensor `t` is stored channels-innermost with its
    pixels in (n, h, w) order and no gaps other than channel padding, else None.  Dense torch.channels_last gives C;
    cpg_b200 stores activations whose channel count is not a multiple of 4 with the stride rounded up to 4
    ("padded NHWC": TMA and 16-byte accesses need 16-byte aligned pixels; the pad lanes carry no information)."""
    n, c, h, w = t.shape
    s = t.stride()
    if c > 1 and s[1] != 1:
        return None
    ps = s[3] if w > 1 else s[2] if h > 1 else s[0] if n > 1 else _up4(c)
    if ps < c or (w > 1 and h > 1 and s[2] != w * ps) or (n > 1 and s[0] != h * w * ps):
        return None
    return ps


def empty_nhwc(shape, device, ps=None):
    """Uninitialised logical-NCHW fp32 tensor in (padded) NHWC storage with pixel stride `ps` (default: C rounded
    up to 4)."""
    n, c, h, w = shape
    ps = _up4(c) if ps is None else ps
    buf = torch.empty((n, ps, h, w), dtype=torch.float32, device=device, memory_format=CL)
    return buf if ps == c else buf[:, :c]


def empty_like_padded(t):
    """Uninitialised tensor with `t`'s shape and strides whose storage covers the pad lanes of the LAST pixel / row
    too (torch.empty_strided stops at the last logical element; the kernels store whole 16-byte groups)."""
    if t.dim() == 4:
        ps = nhwc_pixel_stride(t)
        if ps is not None and ps != t.shape[1]:
            return empty_nhwc(t.shape, t.device, ps)
    elif t.dim() == 2 and t.stride(1) == 1 and t.stride(0) > t.shape[1]:
        return torch.empty((t.shape[0], t.stride(0)), dtype=t.dtype, device=t.device)[:, :t.shape[1]]
    return torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, device=t.device)


def to_nhwc_aligned(t):
    """`t` itself if its pixels are channels-innermost and 16-byte aligned, else a (padded) NHWC copy."""
    ps = nhwc_pixel_stride(t)
    if ps is not None and ps % 4 == 0:
        return t
    out = empty_nhwc(t.shape, t.device)
    out.copy_(t)
    return out


def _dense4(t):
    """Return t if it is dense NCHW, dense NHWC or padded NHWC, else a dense NCHW copy."""
    if t.is_contiguous() or t.is_contiguous(memory_format=CL):
        return t
    ps = nhwc_pixel_stride(t)
    if ps is not None and ps % 4 == 0:
        return t
    return t.contiguous()


# ---- TF32 exactness of activation operands --------------------------------------------------------------
# The tensor core truncates fp32 operands to TF32.  Tensors produced by our own kernels with round-to-nearest
# (cpg_b200.fused_norm with tf32_out, round_tf32 below) are tagged; anything else is rounded once per pass
# before it is handed to the tcgen05 kernels (include/cpgb200.h, CPGB_FLAG_X_TF32 / CPGB_FLAG_DY_TF32).
_TF32_PRESERVING = frozenset((
    'ViewBackward0', 'ReshapeAliasBackward0', 'UnsafeViewBackward0', 'AliasBackward0', 'ReluBackward0',
    'MaxPool2DWithIndicesBackward0', 'SqueezeBackward0', 'SqueezeBackward1', 'UnsqueezeBackward0',
    'TransposeBackward0', 'PermuteBackward0', 'TBackward0', 'CloneBackward0', 'SliceBackward0', 'SelectBackward0'))


def mark_tf32(t):
    """Tag `t` as holding TF32-representable values (valid until the next in-place modification)."""
    t._cpgb_tf32 = t._version
    return t


def is_tf32(t):
    """True when every element of `t` is known to be TF32-representable: tagged by its producer and not
    modified in place since, or derived from such a tensor through value-preserving autograd nodes (view,
    reshape, ReLU, max-pool, ...)."""
    tag = getattr(t, '_cpgb_tf32', None)
    if tag is not None and tag == t._version:
        return True
    fn = t.grad_fn
    for _ in range(8):
        if fn is None:
            return False
        if getattr(fn, 'cpgb_tf32_out', False):
            return True
        if type(fn).__name__ not in _TF32_PRESERVING or not fn.next_functions:
            return False
        fn = fn.next_functions[0][0]
    return False


def _is_dense(t):
    if t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=CL)):
        return True
    if t.dim() == 4:                      # padded NHWC: dense but for the pad lanes
        ps = nhwc_pixel_stride(t)
        return ps is not None and ps % 4 == 0
    return t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]      # row-padded matrix


def _span(t):
    return 1 + sum((n - 1) * st for n, st in zip(t.shape, t.stride()))


def round_tf32(lib, t):
    """A tagged copy of `t` (same strides; dense, or dense but for channel / row padding) rounded to the nearest
    TF32 value."""
    out = empty_like_padded(t)
    with torch.cuda.device(t.device):
        _lib.check(lib.cpgb_round_tf32(_lib.ptr(t), _lib.ptr(out), _span(t), _lib.stream_ptr()), 'cpgb_round_tf32')
    return mark_tf32(out)


def _prepare_operand(lib, t, exact, needed):
    """(tensor to hand to the kernels, exact?) -- rounds `t` when a tcgen05 pass will read it and it is not
    known to be exact.  Non-dense views are left to the library's own rounding pre-pass."""
    if exact or not needed:
        return t, exact
    if _is_dense(t):
        return round_tf32(lib, t), True
    return t, False


def _stage(lib, d, w, p, threshold, prestaged=None, owner=None):
    """Build the tensor-core weight operand (masked, TF32, [K][RS][Cp]) for descriptor d, or
    None when d takes the CUDA-core path (which evaluates the mask while loading tiles).
    `prestaged`: operand already built for this forward pass by the model-level batched staging
    (cpg_b200.prune.SparsePruner pre-forward hook).  `owner`: the piggymask Parameter object `p` was detached from
    (cpg_b200.optim.Adam leaves the packed mask words there).
    Returns (staged, scratch): scratch is the split-K workspace of the fprop/dgrad call."""
    nbytes = lib.cpgb_staged_weight_bytes(d)
    if nbytes == 0:
        return None, None
    if lib.cpgb_intile_eligible(d):
        # in-tile masking (CPGB_FLAG_W_INTILE): the kernels TMA-load tiles of `w` itself and mask / round them in
        # shared memory; all they need is the Binarizer's output packed to one bit per element (0.125 B instead of the
        # 12 B per element a staged copy costs), shared by this step's fprop and dgrad
        d.flags |= _lib.FLAG_W_INTILE
        bits = None
        if p is not None:
            # cpg_b200.optim.Adam leaves the packed words of the piggymask it just updated on the parameter
            # (bits, parameter version at that moment, threshold): consumed once, anything else packs here
            holder = owner if owner is not None else p
            emitted = getattr(holder, '_cpgb_bits', None)
            if (emitted is not None and emitted[1] == holder._version and emitted[2] == float(threshold) and
                    emitted[0].device == w.device and emitted[0].numel() == (w.numel() + 31) // 32):
                bits = emitted[0]
            else:
                bits = torch.empty((w.numel() + 31) // 32, dtype=torch.int64, device=w.device)
                _lib.check(lib.cpgb_pack_mask(_lib.ptr(p), None, w.numel(), threshold, 255, _lib.ptr(bits),
                                              _lib.stream_ptr()), 'cpgb_pack_mask')
            if emitted is not None:
                holder._cpgb_bits = None
        return bits, _ws(lib.cpgb_workspace_bytes(d) - nbytes, w.device)
    if lib.cpgb_weights_usable_raw(d, 1 if p is not None else 0):
        # linear / 1x1 layer without a piggymask: the weight tensor itself is the operand
        return w, _ws(lib.cpgb_workspace_bytes(d) - nbytes, w.device)
    if prestaged is not None and prestaged.numel() >= nbytes and prestaged.device == w.device:
        staged = prestaged
    else:
        staged = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
        _lib.check(lib.cpgb_stage_weights(d, _lib.ptr(w), _lib.ptr(p), threshold, _lib.ptr(staged), nbytes,
                                          _lib.stream_ptr()), 'cpgb_stage_weights')
    # cpgb_workspace_bytes = [staged operand][partial sums]; the operand is supplied separately
    return staged, _ws(lib.cpgb_workspace_bytes(d) - nbytes, w.device)


_SIDE_STREAMS = {}
_EPI_STREAMS = {}
# the fused wgrad epilogues (sum of the split partial sums + weight decay + masks) of a deferred backward pass run
# on a third stream, under the wgrad GEMM of the next layer (cpgb_conv2d_wgrad_fused_async); CPGB_EPILOGUE_STREAM=0
# keeps them on the wgrad stream
EPILOGUE_STREAM = os.environ.get('CPGB_EPILOGUE_STREAM', '1') != '0'
# issue dgrad (the critical chain of the backward pass) before the forked wgrad, so that the block scheduler hands
# free SMs to it first.  Measured: no difference (1.502 vs 1.499 ms per step), off by default (CPGB_DGRAD_FIRST=1)
DGRAD_FIRST = os.environ.get('CPGB_DGRAD_FIRST', '0') != '0'
# a convolution that feeds a fused batch-norm (cpg_b200.fused_norm.fuse_bn_relu marks it) accumulates the per-tile
# column statistics of y in its epilogue (cpgb_conv2d_fprop_stats) and the batch-norm skips its statistics pass
COLSTATS = os.environ.get('CPGB_COLSTATS', '1') != '0'
OVERLAP_BACKWARD = True   # run wgrad on a side stream (it only reads x and dy; nothing on the main chain needs it)
DEFER_JOIN = True         # join the side stream once, at the end of the backward pass, instead of per layer
_PENDING = {}             # device index -> (graph task id, tensors the side stream may still be reading)


def _dev_index(device):
    d = torch.device(device)
    return d.index if d.index is not None else torch.cuda.current_device()


def _side_stream(device):
    key = _dev_index(device)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _SIDE_STREAMS[key] = st
    return st


def _epilogue_stream(device):
    key = _dev_index(device)
    st = _EPI_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _EPI_STREAMS[key] = st
    return st


def join_epilogue_stream(device, into=None):
    """Make `into` (default: the wgrad side stream) wait for the epilogue kernels queued so far -- what has to happen
    before a collective launched from that stream may read the gradients."""
    est = _EPI_STREAMS.get(_dev_index(device))
    if est is not None:
        (into if into is not None else _side_stream(device)).wait_stream(est)


def pending_side_stream(device):
    """The side stream if weight-gradient kernels of the running backward pass are still queued on it
    (their outputs must not be consumed on another stream before `join_side_stream`), else None."""
    key = _dev_index(device)
    return _SIDE_STREAMS.get(key) if key in _PENDING else None


def join_side_stream(device=None):
    """Make the current stream wait for the weight-gradient kernels of this backward pass and release
    the tensors held for them.  Runs automatically at the end of every backward pass (autograd engine
    callback); idempotent."""
    keys = [_dev_index(device)] if device is not None else list(_PENDING.keys())
    for key in keys:
        if key in _PENDING:
            with torch.cuda.device(key):
                torch.cuda.current_stream().wait_stream(_SIDE_STREAMS[key])
                if key in _EPI_STREAMS:
                    torch.cuda.current_stream().wait_stream(_EPI_STREAMS[key])
            del _PENDING[key]


def _defer(device, tensors):
    key = _dev_index(device)
    task = torch._C._current_graph_task_id()
    entry = _PENDING.get(key)
    if entry is not None and entry[0] != task:
        # left over from a backward pass that never reached its callback (an exception in a later node):
        # join now, then start afresh for the running pass
        join_side_stream(key)
        entry = None
    if entry is None:
        entry = _PENDING[key] = (task, [])
        # end-of-backward callback of the autograd engine (the mechanism DDP uses to finalise buckets)
        torch.autograd.Variable._execution_engine.queue_callback(lambda: join_side_stream(key))
    entry[1].extend(t for t in tensors if t is not None)


def _backward_operands(lib, d, dy, dy_exact, x_exact, need_dx, need_w, w_intile=False):
    """Round dy once for dgrad + wgrad when a tcgen05 pass reads it; set the descriptor flags.  Returns
    (dy for the kernels, the caller's dy if a rounded copy replaced it else None)."""
    tc_d = need_dx and lib.cpgb_uses_tensor_cores(d, 1)
    tc_w = need_w and lib.cpgb_uses_tensor_cores(d, 2)
    dy_k, dy_exact = _prepare_operand(lib, dy, dy_exact, bool(tc_d or tc_w))
    d.flags = (_lib.FLAG_X_TF32 if x_exact else 0) | (_lib.FLAG_DY_TF32 if dy_exact else 0) | \
              (_lib.FLAG_W_INTILE if w_intile else 0)
    return dy_k, (dy if dy_k is not dy else None)


def _backward_kernels(lib, ctx, d, x, dy, w, p, threshold, staged, need_dx, need_w, has_bias, dx, dy_raw=None):
    """dgrad (+) wgrad with the fused epilogue for descriptor d.  Everything is allocated on the
    current stream.  The wgrad launch is forked onto a side stream: wgrad(l) only reads x(l) and dy(l),
    and nothing before optimizer.step() / the all-reduce reads dW, so the side stream is joined once at
    the end of the backward pass -- wgrad then overlaps dgrad *and* the BN / ReLU / pooling backward
    kernels of the following layers.  Capturable: fork and join become branches of the CUDA graph."""
    dW = dP = db = None
    device = x.device
    with torch.cuda.device(device):
        main = torch.cuda.current_stream()
        wbytes = lib.cpgb_workspace_bytes(d)
        ws_d = _ws(wbytes, device) if need_dx else None
        fuse = ctx.fuse
        in_backward = torch._C._current_graph_task_id() != -1
        # a parameter that already holds a gradient gets `grad += dW` on the main stream as soon as this
        # node returns (gradient accumulation): join immediately in that case
        mod = ctx.module
        fresh = mod is not None and all(getattr(mod, n, None) is None or getattr(mod, n).grad is None
                                        for n in ('weight', 'piggymask', 'bias'))
        wd = fuse.weight_decay if fuse is not None else 0.0
        if fuse is not None and mod is not None and not fresh:
            # gradient accumulation: the reference adds wd*W ONCE, after the last backward pass
            # (utils/prune.py:203).  If an earlier pass of this window already went through the fused epilogue
            # the term is in .grad: mask only.  Otherwise (.grad did not come from us) hand out the raw
            # autograd values and let do_weight_decay_and_make_grads_zero finish them.
            if mod._cpg_grads_final:
                wd = 0.0
            else:
                fuse = None
        defer = OVERLAP_BACKWARD and DEFER_JOIN and need_w and in_backward and fresh
        fork = OVERLAP_BACKWARD and need_w and (need_dx or defer)
        dgrad_done = False
        if need_dx and fork and DGRAD_FIRST:
            # fork point first (wgrad must not wait for dgrad), then dgrad on the main stream, then wgrad on the side
            _side_stream(device).wait_stream(main)
            _lib.check(lib.cpgb_conv2d_dgrad(d, _lib.ptr(dy), _lib.ptr(w), _lib.ptr(p), _lib.ptr(dx),
                                             threshold, _lib.ptr(staged), _lib.ptr(ws_d), ws_d.numel(),
                                             main.cuda_stream), 'cpgb_conv2d_dgrad')
            dgrad_done = True
        if need_w:
            mode = fuse.mode if fuse is not None else _lib.GRAD_RAW
            # data parallel (cpg_b200.ddp.GradAllReducer): the epilogue writes straight into the layer's slot of a
            # flat gradient bucket -- .weight.grad becomes a view of it -- and, with a piggymask, writes the MERGED
            # dW + dP (disjoint supports after the masking) so that one buffer travels through the all-reduce
            slot = getattr(mod, '_cpg_grad_slot', None) if (mod is not None and fresh and in_backward) else None
            if slot is not None and (slot.n != w.numel() or slot.reducer.flat[slot.bucket].device != w.device):
                slot = None
            ws_w = _ws(wbytes, device)
            dW = slot.view(w) if slot is not None else torch.empty_like(w)
            dP = torch.empty_like(w) if p is not None else None
            db = torch.empty(w.shape[0], dtype=torch.float32, device=device) if has_bias else None
            merged = slot is not None and p is not None and fuse is not None
            wstream = main
            if fork:
                wstream = _side_stream(device)
                if not dgrad_done:
                    wstream.wait_stream(main)
            # deferred join: the epilogue goes to its own stream and overlaps the next layer's wgrad GEMM; both
            # streams are joined by the end-of-backward callback
            estream = _epilogue_stream(device) if (defer and fork and EPILOGUE_STREAM) else wstream
            _lib.check(lib.cpgb_conv2d_wgrad_fused_async(
                d, _lib.ptr(x), _lib.ptr(dy), _lib.ptr(w), _lib.ptr(p),
                _lib.ptr(fuse.tmask) if fuse is not None else None,
                fuse.cur if fuse is not None else 0, wd, (mode | _lib.GRAD_MERGED) if merged else mode,
                _lib.ptr(dW), None if merged else _lib.ptr(dP), _lib.ptr(db) if dy_raw is None else None, threshold,
                _lib.ptr(ws_w), ws_w.numel(), wstream.cuda_stream, estream.cuda_stream), 'cpgb_conv2d_wgrad_fused')
            if db is not None and dy_raw is not None:
                # the bias gradient is a plain fp32 sum: take it from the caller's dy, not from the TF32-rounded copy
                _lib.check(lib.cpgb_conv2d_bias_grad_ws(d, _lib.ptr(dy_raw), _lib.ptr(db), _lib.ptr(ws_w), ws_w.numel(),
                                                        wstream.cuda_stream), 'cpgb_conv2d_bias_grad')
            if fuse is not None and mod is not None:
                mod._cpg_grads_final = True
            if slot is not None:
                # dP (merged case) is filled by the reducer's split after the all-reduce; it must not be kept
                # referenced here (AccumulateGrad would clone instead of adopting it)
                slot.state, slot.fuse = (2 if merged else 1), fuse
                slot.reducer.layer_done(slot, device)
        if need_dx and not dgrad_done:
            _lib.check(lib.cpgb_conv2d_dgrad(d, _lib.ptr(dy), _lib.ptr(w), _lib.ptr(p), _lib.ptr(dx),
                                             threshold, _lib.ptr(staged), _lib.ptr(ws_d), ws_d.numel(),
                                             main.cuda_stream), 'cpgb_conv2d_dgrad')
        if fork:
            if defer:
                # inputs and scratch only: holding the outputs too would raise their use count and make
                # AccumulateGrad clone them on the main stream instead of adopting them
                _defer(device, (x, dy, dy_raw, w, p, ws_w, fuse.tmask if fuse is not None else None))
            else:
                main.wait_stream(wstream)
    return dW, dP, db


class FuseCtx:
    """What the fused wgrad epilogue needs from the pruner (utils/prune.py:195-211)."""
    __slots__ = ('tmask', 'cur', 'weight_decay', 'mode')

    def __init__(self, tmask, cur, weight_decay, mode):
        self.tmask, self.cur, self.weight_decay, self.mode = tmask, int(cur), float(weight_decay), int(mode)


def _check_params(weight, piggymask, bias):
    for name, t in (('weight', weight), ('piggymask', piggymask), ('bias', bias)):
        if t is None:
            continue
        if t.dtype != torch.float32:
            raise _lib.CpgbError(f'{name} must be float32 (the reference path is fp32), got {t.dtype}')
        if not t.is_cuda:
            raise _lib.CpgbError(f'{name} is on {t.device}: cpg_b200 has no CPU path')
    if piggymask is not None and piggymask.shape != weight.shape:
        raise _lib.CpgbError('piggymask must have the weight\'s shape')


class MaskedConv2dFn(torch.autograd.Function):
    """y = conv2d(x, (piggymask > thr) * weight, bias, ...)  -- models/layers.py:98-109."""

    @staticmethod
    def forward(ctx, x, weight, piggymask, bias, stride, padding, dilation, groups, threshold, fuse,
                module, channels_last_out, prestaged=None, x_exact=False, emit_colstats=False):
        lib = _lib.load()
        _check_params(weight, piggymask, bias)
        if x.dim() != 4:
            raise _lib.CpgbError('SharableConv2d expects a 4-D input')
        if x.dtype != torch.float32 or not x.is_cuda:
            raise _lib.CpgbError(f'input must be a float32 CUDA tensor, got {x.dtype} on {x.device}')
        if groups == 1 and tuple(stride) == (1, 1):
            # the tcgen05 kernels TMA-load NHWC activations whose pixel stride is a multiple of 16 B: dense
            # channels_last when C % 4 == 0, else NHWC with the pixel stride padded to 4 (the 3-channel stem, the
            # 78 / 313 / 627-channel grown networks); the pad lanes are never read (TMA bounds the channel
            # dimension at C)
            x = to_nhwc_aligned(x)
        else:
            x = _dense4(x)
        w = weight.detach().contiguous()
        p = piggymask.detach().contiguous() if piggymask is not None else None
        b = bias.detach().contiguous() if bias is not None else None
        N, C, H, W = x.shape
        K, Cg, R, S = w.shape
        if C != Cg * groups:
            raise RuntimeError(f'expected input with {Cg * groups} channels, got {C}')
        P = (H + 2 * padding[0] - dilation[0] * (R - 1) - 1) // stride[0] + 1
        Q = (W + 2 * padding[1] - dilation[1] * (S - 1) - 1) // stride[1] + 1
        if P <= 0 or Q <= 0:
            raise RuntimeError('kernel size larger than the (padded) input')
        if channels_last_out or (nhwc_pixel_stride(x) is not None and not x.is_contiguous()):
            y = empty_nhwc((N, K, P, Q), x.device)        # padded NHWC when K % 4 != 0
        else:
            y = torch.empty((N, K, P, Q), dtype=torch.float32, device=x.device)
        d = _lib.conv_desc(x.shape, x.stride(), w.shape, y.shape, y.stride(), stride, padding, dilation, groups)
        # TF32 operands are rounded to nearest, once, and the rounded copy is what the backward pass reuses
        need_w = weight.requires_grad or (piggymask is not None and piggymask.requires_grad)
        uses_tc = lib.cpgb_uses_tensor_cores(d, 0) or (need_w and lib.cpgb_uses_tensor_cores(d, 2))
        x, x_exact = _prepare_operand(lib, x, bool(x_exact), bool(uses_tc))
        d.flags = _lib.FLAG_X_TF32 if x_exact else 0
        with torch.cuda.device(x.device):
            staged, ws = _stage(lib, d, w, p, threshold, prestaged, owner=piggymask)
            colstats, nparts = None, 0
            if emit_colstats and COLSTATS:
                # per pixel tile: column sums / sums of squares of y for the batch-norm behind this layer, when the
                # plan of this shape can deliver them and y has the pixel stride that kernel family expects
                nparts = lib.cpgb_fprop_colstats_parts(d)
                if nparts > 0 and nhwc_pixel_stride(y) == _up4(K):
                    colstats = torch.empty(nparts * _up4(K) * 2, dtype=torch.float32, device=x.device)
                else:
                    nparts = 0
            _lib.check(lib.cpgb_conv2d_fprop_stats(d, _lib.ptr(x), _lib.ptr(w), _lib.ptr(p), _lib.ptr(b),
                                                   _lib.ptr(y), threshold, _lib.ptr(staged), _lib.ptr(ws),
                                                   ws.numel() if ws is not None else 0, _lib.ptr(colstats),
                                                   _lib.stream_ptr()), 'cpgb_conv2d_fprop')
        # read by FusedBatchNormReLU2d off y.grad_fn: (partial pairs, how many, channels)
        ctx.cpgb_colstats = (colstats, nparts, K) if colstats is not None else None
        ctx.save_for_backward(x, w, p)
        ctx.staged = staged     # masked TF32 operand, shared with this step's dgrad
        ctx.has_bias = bias is not None
        ctx.geom = (stride, padding, dilation, groups, threshold)
        ctx.fuse, ctx.module, ctx.x_exact = fuse, module, x_exact
        ctx.w_intile = bool(d.flags & _lib.FLAG_W_INTILE)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        lib = _lib.load()
        x, w, p = ctx.saved_tensors
        stride, padding, dilation, groups, threshold = ctx.geom
        dy_exact = is_tf32(dy)        # layout copies below keep the values
        if nhwc_pixel_stride(x) is not None and not x.is_contiguous():
            dy = to_nhwc_aligned(dy)
        else:
            dy = _dense4(dy)
        d = _lib.conv_desc(x.shape, x.stride(), w.shape, dy.shape, dy.stride(), stride, padding, dilation, groups)
        need_dx = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        dy, dy_raw = _backward_operands(lib, d, dy, dy_exact, ctx.x_exact, need_dx, need_w, ctx.w_intile)
        dx = empty_like_padded(x) if need_dx else None
        dW, dP, db = _backward_kernels(lib, ctx, d, x, dy, w, p, threshold, ctx.staged, need_dx, need_w,
                                       ctx.has_bias, dx, dy_raw if ctx.has_bias else None)
        return dx, dW, dP, db, None, None, None, None, None, None, None, None, None, None, None


def _rows16(m):
    """The 2-D matrix `m` with unit column stride and 16-byte aligned rows: itself, or a copy whose row stride is
    rounded up to 4 floats (feature counts like 627 that are not a multiple of 4)."""
    if m.stride(1) == 1 and m.stride(0) % 4 == 0 and m.stride(0) >= m.shape[1] and m.data_ptr() % 16 == 0:
        return m
    if m.shape[1] % 4 == 0:
        return m.contiguous()
    out = torch.empty((m.shape[0], _up4(m.shape[1])), dtype=m.dtype, device=m.device)[:, :m.shape[1]]
    out.copy_(m)
    return out


def _linear_desc(lib, M, I, O, ldx, ldy):
    d = _lib.ConvDesc()
    lib.cpgb_linear_desc(d, M, I, O)
    d.xs[0] = d.xs[2] = d.xs[3] = ldx if M > 1 or ldx >= I else I
    d.ys[0] = d.ys[2] = d.ys[3] = ldy if M > 1 or ldy >= O else O
    return d


class MaskedLinearFn(torch.autograd.Function):
    """y = linear(x, (piggymask > thr) * weight, bias)  -- models/layers.py:184-194."""

    @staticmethod
    def forward(ctx, x, weight, piggymask, bias, threshold, fuse, module, prestaged=None, x_exact=False):
        lib = _lib.load()
        _check_params(weight, piggymask, bias)
        if x.dtype != torch.float32 or not x.is_cuda:
            raise _lib.CpgbError(f'input must be a float32 CUDA tensor, got {x.dtype} on {x.device}')
        w = weight.detach().contiguous()
        p = piggymask.detach().contiguous() if piggymask is not None else None
        b = bias.detach().contiguous() if bias is not None else None
        O, I = w.shape
        if x.shape[-1] != I:
            raise RuntimeError(f'size mismatch: input features {x.shape[-1]} vs weight {tuple(w.shape)}')
        x2 = _rows16(x.reshape(-1, I))
        M = x2.shape[0]
        # allocated in its final shape: returning a view of a custom Function's output would
        # forbid the in-place ReLU that follows it in models/vgg.py:116-118
        if O % 4 == 0 or x.dim() != 2:
            y = torch.empty((*x.shape[:-1], O), dtype=torch.float32, device=x.device)
            ldy = O
        else:                     # rows padded to 16 bytes (an odd O with ldy = O takes the CUDA-core kernels)
            ldy = _up4(O)
            y = torch.empty((M, ldy), dtype=torch.float32, device=x.device)[:, :O]
        d = _linear_desc(lib, M, I, O, x2.stride(0), ldy)
        need_w = weight.requires_grad or (piggymask is not None and piggymask.requires_grad)
        uses_tc = lib.cpgb_uses_tensor_cores(d, 0) or (need_w and lib.cpgb_uses_tensor_cores(d, 2))
        x2, x_exact = _prepare_operand(lib, x2, bool(x_exact), bool(uses_tc))
        d.flags = _lib.FLAG_X_TF32 if x_exact else 0
        with torch.cuda.device(x.device):
            staged, ws = _stage(lib, d, w, p, threshold, prestaged, owner=piggymask)
            _lib.check(lib.cpgb_conv2d_fprop(d, _lib.ptr(x2), _lib.ptr(w), _lib.ptr(p), _lib.ptr(b), _lib.ptr(y),
                                             threshold, _lib.ptr(staged), _lib.ptr(ws),
                                             ws.numel() if ws is not None else 0, _lib.stream_ptr()),
                       'cpgb_conv2d_fprop(linear)')
        ctx.save_for_backward(x2, w, p)
        ctx.staged = staged
        ctx.has_bias, ctx.threshold, ctx.fuse, ctx.module = bias is not None, threshold, fuse, module
        ctx.x_shape, ctx.x_exact = x.shape, x_exact
        ctx.w_intile = bool(d.flags & _lib.FLAG_W_INTILE)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        lib = _lib.load()
        x2, w, p = ctx.saved_tensors
        O, I = w.shape
        M = x2.shape[0]
        dy2 = _rows16(dy.reshape(M, O))
        d = _linear_desc(lib, M, I, O, x2.stride(0), dy2.stride(0))
        need_dx = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        dy2, dy_raw = _backward_operands(lib, d, dy2, is_tf32(dy), ctx.x_exact, need_dx, need_w, ctx.w_intile)
        dx2 = empty_like_padded(x2) if need_dx else None
        dW, dP, db = _backward_kernels(lib, ctx, d, x2, dy2, w, p, ctx.threshold, ctx.staged, need_dx, need_w,
                                       ctx.has_bias, dx2, dy_raw if ctx.has_bias else None)
        dx = dx2.reshape(ctx.x_shape) if need_dx else None
        return dx, dW, dP, db, None, None, None, None, None


class Binarizer(torch.autograd.Function):
    """Binarizes {0, 1} a real valued tensor; straight-through backward.
    Same call signature as the reference: ``Binarizer.apply(inputs, threshold)``
    (models/layers.py:11-23)."""

    @staticmethod
    def forward(ctx, inputs, threshold):
        lib = _lib.load()
        if not inputs.is_cuda or inputs.dtype != torch.float32:
            raise _lib.CpgbError('Binarizer needs a float32 CUDA tensor (no CPU path)')
        src = inputs.detach().contiguous()
        out = torch.empty_like(src)
        with torch.cuda.device(src.device):
            _lib.check(lib.cpgb_binarize(_lib.ptr(src), _lib.ptr(out), src.numel(), float(threshold),
                                         _lib.stream_ptr()), 'cpgb_binarize')
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return grad_out, None


class Ternarizer(torch.autograd.Function):
    """Ternarizes {-1, 0, 1}: -1 where x < 0, 1 where x > thr, else 0; straight-through backward.
    The reference's class (models/layers.py:25-40) is a legacy non-static Function that modern
    torch refuses to run and nothing selects it; kept for API completeness only,
    implemented with torch ops (not on the accelerated path)."""

    @staticmethod
    def forward(ctx, inputs, threshold=DEFAULT_THRESHOLD):
        out = torch.zeros_like(inputs)
        out[inputs < 0] = -1
        out[inputs > threshold] = 1
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return grad_out, None
