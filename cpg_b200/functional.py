"""autograd.Functions over the C ABI (include/cpgb200.h).

One Function per masked layer type replaces the reference's three-node autograd graph
``Binarizer.apply -> mul -> F.conv2d / F.linear`` (models/layers.py:101-108, 187-194): the
forward evaluates the piggyback predicate inside the convolution kernels, the backward
produces dX, dW, dP (straight-through: dP = g * W, models/layers.py:21-23) and dbias, with
the pruner's weight-decay / grad-mask step (utils/prune.py:195-211) optionally folded into
the same epilogue.
"""
import torch
from torch.autograd.function import once_differentiable

from . import _lib

DEFAULT_THRESHOLD = 5e-3  # models/layers.py:9
CL = torch.channels_last


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _dense4(t):
    """Return t if it is dense NCHW or dense NHWC, else a dense NCHW copy."""
    if t.is_contiguous() or t.is_contiguous(memory_format=CL):
        return t
    return t.contiguous()


def _stage(lib, d, w, p, threshold, prestaged=None):
    """Build the tensor-core weight operand (masked, TF32, [K][RS][Cp]) for descriptor d, or
    None when d takes the CUDA-core path (which evaluates the mask while loading tiles).
    `prestaged`: operand already built for this forward pass by the model-level batched staging
    (cpg_b200.prune.SparsePruner pre-forward hook).
    Returns (staged, scratch): scratch is the split-K workspace of the fprop/dgrad call."""
    nbytes = lib.cpgb_staged_weight_bytes(d)
    if nbytes == 0:
        return None, None
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


def _backward_kernels(lib, ctx, d, x, dy, w, p, threshold, staged, need_dx, need_w, has_bias, dx):
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
        if need_w:
            ws_w = _ws(wbytes, device)
            dW = torch.empty_like(w)
            dP = torch.empty_like(w) if p is not None else None
            db = torch.empty(w.shape[0], dtype=torch.float32, device=device) if has_bias else None
        fuse = ctx.fuse
        in_backward = torch._C._current_graph_task_id() != -1
        # a parameter that already holds a gradient gets `grad += dW` on the main stream as soon as this
        # node returns (gradient accumulation): join immediately in that case
        mod = ctx.module
        fresh = mod is not None and all(getattr(mod, n, None) is None or getattr(mod, n).grad is None
                                        for n in ('weight', 'piggymask', 'bias'))
        defer = OVERLAP_BACKWARD and DEFER_JOIN and need_w and in_backward and fresh
        fork = OVERLAP_BACKWARD and need_w and (need_dx or defer)
        if need_w:
            wstream = main
            if fork:
                wstream = _side_stream(device)
                wstream.wait_stream(main)
            mode = fuse.mode if fuse is not None else _lib.GRAD_RAW
            _lib.check(lib.cpgb_conv2d_wgrad_fused(
                d, _lib.ptr(x), _lib.ptr(dy), _lib.ptr(w), _lib.ptr(p),
                _lib.ptr(fuse.tmask) if fuse is not None else None,
                fuse.cur if fuse is not None else 0, fuse.weight_decay if fuse is not None else 0.0, mode,
                _lib.ptr(dW), _lib.ptr(dP), _lib.ptr(db), threshold, _lib.ptr(ws_w), ws_w.numel(),
                wstream.cuda_stream), 'cpgb_conv2d_wgrad_fused')
            if fuse is not None and ctx.module is not None:
                ctx.module._cpg_grads_final = True
        if need_dx:
            _lib.check(lib.cpgb_conv2d_dgrad(d, _lib.ptr(dy), _lib.ptr(w), _lib.ptr(p), _lib.ptr(dx),
                                             threshold, _lib.ptr(staged), _lib.ptr(ws_d), ws_d.numel(),
                                             main.cuda_stream), 'cpgb_conv2d_dgrad')
        if fork:
            if defer:
                # inputs and scratch only: holding the outputs too would raise their use count and make
                # AccumulateGrad clone them on the main stream instead of adopting them
                _defer(device, (x, dy, w, p, ws_w, fuse.tmask if fuse is not None else None))
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
                module, channels_last_out, prestaged=None):
        lib = _lib.load()
        _check_params(weight, piggymask, bias)
        if x.dim() != 4:
            raise _lib.CpgbError('SharableConv2d expects a 4-D input')
        if x.dtype != torch.float32 or not x.is_cuda:
            raise _lib.CpgbError(f'input must be a float32 CUDA tensor, got {x.dtype} on {x.device}')
        x = _dense4(x)
        if groups == 1 and tuple(stride) == (1, 1) and weight.shape[0] % 4 == 0:
            # the tcgen05 kernels TMA-load NHWC activations whose pixel stride is a multiple of 16 B
            if x.shape[1] % 4 == 0:
                if not x.is_contiguous(memory_format=CL):
                    x = x.contiguous(memory_format=CL)
            else:
                # odd channel count (the 3-channel stem): NHWC with the pixel stride padded to 4; the
                # pad lanes are never read (TMA bounds the channel dimension at C)
                n_, c_, h_, w_ = x.shape
                xp = torch.empty((n_, (c_ + 3) // 4 * 4, h_, w_), dtype=x.dtype, device=x.device, memory_format=CL)
                xv = xp[:, :c_]
                xv.copy_(x)
                x = xv
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
        fmt = CL if (channels_last_out or (x.is_contiguous(memory_format=CL) and not x.is_contiguous())) \
            else torch.contiguous_format
        y = torch.empty((N, K, P, Q), dtype=torch.float32, device=x.device, memory_format=fmt)
        d = _lib.conv_desc(x.shape, x.stride(), w.shape, y.shape, y.stride(), stride, padding, dilation, groups)
        with torch.cuda.device(x.device):
            staged, ws = _stage(lib, d, w, p, threshold, prestaged)
            _lib.check(lib.cpgb_conv2d_fprop(d, _lib.ptr(x), _lib.ptr(w), _lib.ptr(p), _lib.ptr(b),
                                             _lib.ptr(y), threshold, _lib.ptr(staged), _lib.ptr(ws),
                                             ws.numel() if ws is not None else 0,
                                             _lib.stream_ptr()), 'cpgb_conv2d_fprop')
        ctx.save_for_backward(x, w, p)
        ctx.staged = staged     # masked TF32 operand, shared with this step's dgrad
        ctx.has_bias = bias is not None
        ctx.geom = (stride, padding, dilation, groups, threshold)
        ctx.fuse, ctx.module = fuse, module
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        lib = _lib.load()
        x, w, p = ctx.saved_tensors
        stride, padding, dilation, groups, threshold = ctx.geom
        dy = _dense4(dy)
        if x.is_contiguous(memory_format=CL) and not dy.is_contiguous(memory_format=CL):
            dy = dy.contiguous(memory_format=CL)
        d = _lib.conv_desc(x.shape, x.stride(), w.shape, dy.shape, dy.stride(), stride, padding, dilation, groups)
        need_dx = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        dx = torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=x.device) if need_dx else None
        dW, dP, db = _backward_kernels(lib, ctx, d, x, dy, w, p, threshold, ctx.staged, need_dx, need_w,
                                       ctx.has_bias, dx)
        return dx, dW, dP, db, None, None, None, None, None, None, None, None, None


class MaskedLinearFn(torch.autograd.Function):
    """y = linear(x, (piggymask > thr) * weight, bias)  -- models/layers.py:184-194."""

    @staticmethod
    def forward(ctx, x, weight, piggymask, bias, threshold, fuse, module, prestaged=None):
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
        x2 = x.reshape(-1, I).contiguous()
        M = x2.shape[0]
        # allocated in its final shape: returning a view of a custom Function's output would
        # forbid the in-place ReLU that follows it in models/vgg.py:116-118
        y = torch.empty((*x.shape[:-1], O), dtype=torch.float32, device=x.device)
        d = _lib.ConvDesc()
        lib.cpgb_linear_desc(d, M, I, O)
        with torch.cuda.device(x.device):
            staged, ws = _stage(lib, d, w, p, threshold, prestaged)
            _lib.check(lib.cpgb_conv2d_fprop(d, _lib.ptr(x2), _lib.ptr(w), _lib.ptr(p), _lib.ptr(b), _lib.ptr(y),
                                             threshold, _lib.ptr(staged), _lib.ptr(ws),
                                             ws.numel() if ws is not None else 0, _lib.stream_ptr()),
                       'cpgb_conv2d_fprop(linear)')
        ctx.save_for_backward(x2, w, p)
        ctx.staged = staged
        ctx.has_bias, ctx.threshold, ctx.fuse, ctx.module = bias is not None, threshold, fuse, module
        ctx.x_shape = x.shape
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        lib = _lib.load()
        x2, w, p = ctx.saved_tensors
        O, I = w.shape
        M = x2.shape[0]
        dy2 = dy.reshape(M, O).contiguous()
        d = _lib.ConvDesc()
        lib.cpgb_linear_desc(d, M, I, O)
        need_dx = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        dx2 = torch.empty_like(x2) if need_dx else None
        dW, dP, db = _backward_kernels(lib, ctx, d, x2, dy2, w, p, ctx.threshold, ctx.staged, need_dx, need_w,
                                       ctx.has_bias, dx2)
        dx = dx2.reshape(ctx.x_shape) if need_dx else None
        return dx, dW, dP, db, None, None, None, None


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
