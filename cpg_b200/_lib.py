"""ctypes binding of libcpgb200.so (include/cpgb200.h).  There is no CPU or PyTorch
fallback: a missing library, or a call with CPU tensors, raises."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libcpgb200.so')

GRAD_RAW, GRAD_FINETUNE, GRAD_PRUNE = 0, 1, 2
GRAD_MERGED = 4                       # or'ed into FINETUNE / PRUNE: dW + dP in one buffer (data parallel)
FLAG_X_TF32, FLAG_DY_TF32, FLAG_W_INTILE = 1, 2, 4      # cpgb_conv_desc.flags
PATH_AUTO, PATH_SIMT, PATH_TCGEN05 = 0, 1, 2

EXPORTS = [
    'cpgb_version', 'cpgb_last_error', 'cpgb_set_path', 'cpgb_get_path', 'cpgb_launch_count', 'cpgb_linear_desc',
    'cpgb_workspace_bytes', 'cpgb_staged_weight_bytes', 'cpgb_stage_weights', 'cpgb_staged_weight_bytes_for',
    'cpgb_stage_weights_batched', 'cpgb_weights_usable_raw', 'cpgb_weights_usable_raw_for', 'cpgb_binarize', 'cpgb_conv2d_fprop', 'cpgb_conv2d_dgrad',
    'cpgb_conv2d_wgrad_fused', 'cpgb_conv2d_wgrad_fused_async', 'cpgb_grad_epilogue', 'cpgb_prune_workspace_bytes', 'cpgb_prune_select',
    'cpgb_prune_batched_workspace_bytes', 'cpgb_prune_select_batched', 'cpgb_prune_sampled_workspace_bytes',
    'cpgb_prune_select_sampled',
    'cpgb_apply_mask', 'cpgb_make_finetuning_mask', 'cpgb_mask_stats', 'cpgb_mask_stats_batched', 'cpgb_merge_grads',
    'cpgb_split_merged_grad', 'cpgb_bn_workspace_bytes', 'cpgb_bn_relu_fwd', 'cpgb_bn_relu_bwd',
    'cpgb_uses_tensor_cores', 'cpgb_round_tf32', 'cpgb_conv2d_bias_grad', 'cpgb_conv2d_bias_grad_ws', 'cpgb_pack_mask', 'cpgb_intile_eligible',
    'cpgb_intile_weight_shape', 'cpgb_sgd_nesterov_step', 'cpgb_adam_step', 'cpgb_prelu_workspace_bytes',
    'cpgb_prelu_fwd', 'cpgb_prelu_bwd', 'cpgb_fprop_colstats_parts', 'cpgb_conv2d_fprop_stats',
    'cpgb_bn_relu_fwd_stats',
]


class ConvDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ('N', 'C', 'H', 'W', 'K', 'R', 'S', 'P', 'Q', 'stride_h', 'stride_w', 'pad_h', 'pad_w',
                 'dil_h', 'dil_w', 'groups', 'flags')] + [('xs', ctypes.c_int64 * 4), ('ys', ctypes.c_int64 * 4)]


class CpgbError(RuntimeError):
    pass


_lib = None


def load():
    """Load libcpgb200.so (built by ``__graft_entry__.build()`` / ``make -C cpg_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CpgbError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; '
                        f'g.build()"` (there is no CPU fallback for the masked-conv path)')
    lib = ctypes.CDLL(LIB_PATH)
    vp, f32, i32, i64, sz, dbl = (ctypes.c_void_p, ctypes.c_float, ctypes.c_int32, ctypes.c_int64,
                                  ctypes.c_size_t, ctypes.c_double)
    dp = ctypes.POINTER(ConvDesc)
    sig = {
        'cpgb_version': (ctypes.c_int, []),
        'cpgb_last_error': (ctypes.c_char_p, []),
        'cpgb_set_path': (ctypes.c_int, [ctypes.c_int]),
        'cpgb_get_path': (ctypes.c_int, []),
        'cpgb_launch_count': (ctypes.c_int64, []),
        'cpgb_linear_desc': (None, [dp, i32, i32, i32]),
        'cpgb_workspace_bytes': (sz, [dp]),
        'cpgb_uses_tensor_cores': (ctypes.c_int, [dp, i32]),
        'cpgb_round_tf32': (ctypes.c_int, [vp, vp, i64, vp]),
        'cpgb_conv2d_bias_grad': (ctypes.c_int, [dp, vp, vp, vp]),
        'cpgb_conv2d_bias_grad_ws': (ctypes.c_int, [dp, vp, vp, vp, sz, vp]),
        'cpgb_pack_mask': (ctypes.c_int, [vp, vp, i64, f32, i32, vp, vp]),
        'cpgb_intile_eligible': (ctypes.c_int, [dp]),
        'cpgb_intile_weight_shape': (ctypes.c_int, [i32] * 7),
        'cpgb_binarize': (ctypes.c_int, [vp, vp, i64, f32, vp]),
        'cpgb_staged_weight_bytes': (sz, [dp]),
        'cpgb_stage_weights': (ctypes.c_int, [dp, vp, vp, f32, vp, sz, vp]),
        'cpgb_staged_weight_bytes_for': (sz, [i32] * 7),
        'cpgb_weights_usable_raw': (ctypes.c_int, [dp, i32]),
        'cpgb_weights_usable_raw_for': (ctypes.c_int, [i32] * 8),
        'cpgb_stage_weights_batched': (ctypes.c_int, [i32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp)]
                                       + [ctypes.POINTER(i32)] * 6 + [ctypes.POINTER(f32), vp]),
        'cpgb_conv2d_fprop': (ctypes.c_int, [dp, vp, vp, vp, vp, vp, f32, vp, vp, sz, vp]),
        'cpgb_fprop_colstats_parts': (i32, [dp]),
        'cpgb_conv2d_fprop_stats': (ctypes.c_int, [dp, vp, vp, vp, vp, vp, f32, vp, vp, sz, vp, vp]),
        'cpgb_bn_relu_fwd_stats': (ctypes.c_int, [vp, i64, i32, i32, vp, i32, vp, vp, vp, vp, vp, i32, f32, f32, i32, i32,
                                                  i32, i32, vp, vp, vp, vp, sz, vp]),
        'cpgb_conv2d_dgrad': (ctypes.c_int, [dp, vp, vp, vp, vp, f32, vp, vp, sz, vp]),
        'cpgb_conv2d_wgrad_fused': (ctypes.c_int, [dp, vp, vp, vp, vp, vp, i32, f32, i32, vp, vp, vp, f32,
                                                  vp, sz, vp]),
        'cpgb_conv2d_wgrad_fused_async': (ctypes.c_int, [dp, vp, vp, vp, vp, vp, i32, f32, i32, vp, vp, vp, f32,
                                                        vp, sz, vp, vp]),
        'cpgb_grad_epilogue': (ctypes.c_int, [vp, vp, vp, vp, i64, i32, f32, i32, vp]),
        'cpgb_prune_workspace_bytes': (sz, []),
        'cpgb_prune_select': (ctypes.c_int, [vp, vp, i64, i32, dbl, vp, vp, sz, vp]),
        'cpgb_prune_batched_workspace_bytes': (sz, [i32]),
        'cpgb_prune_select_batched': (ctypes.c_int, [i32, ctypes.POINTER(vp), ctypes.POINTER(vp),
                                                     ctypes.POINTER(i64), i32, dbl, vp, vp, sz, vp]),
        'cpgb_prune_sampled_workspace_bytes': (sz, [i32]),
        'cpgb_prune_select_sampled': (ctypes.c_int, [i32, ctypes.POINTER(vp), ctypes.POINTER(vp),
                                                     ctypes.POINTER(i64), i32, dbl, vp, vp, sz, vp]),
        'cpgb_sgd_nesterov_step': (ctypes.c_int, [i32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                                  ctypes.POINTER(i64), f32, f32, vp, vp]),
        'cpgb_adam_step': (ctypes.c_int, [i32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                          ctypes.POINTER(vp), ctypes.POINTER(i64), dbl, dbl, dbl, dbl, vp, vp,
                                          ctypes.POINTER(vp), ctypes.POINTER(vp), f32, i32, vp]),
        'cpgb_apply_mask': (ctypes.c_int, [vp, vp, i64, i32, vp]),
        'cpgb_make_finetuning_mask': (ctypes.c_int, [vp, i64, i32, vp]),
        'cpgb_mask_stats': (ctypes.c_int, [vp, vp, i64, i32, vp, vp]),
        'cpgb_mask_stats_batched': (ctypes.c_int, [i32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(i64),
                                                   i32, vp, vp]),
        'cpgb_merge_grads': (ctypes.c_int, [vp, vp, vp, i64, vp]),
        'cpgb_split_merged_grad': (ctypes.c_int, [vp, vp, i64, i32, vp, vp, vp]),
        'cpgb_prelu_workspace_bytes': (sz, [i64, i32]),
        'cpgb_prelu_fwd': (ctypes.c_int, [vp, i64, i32, i32, vp, i32, vp, vp]),
        'cpgb_prelu_bwd': (ctypes.c_int, [vp, vp, i64, i32, i32, vp, i32, vp, vp, vp, sz, vp]),
        'cpgb_bn_workspace_bytes': (sz, [i64, i32]),
        'cpgb_bn_relu_fwd': (ctypes.c_int, [vp, i64, i32, i32, vp, vp, vp, vp, vp, i32, f32, f32, i32, i32, i32, i32, vp, vp,
                                            vp, vp, sz, vp]),
        'cpgb_bn_relu_bwd': (ctypes.c_int, [vp, vp, i64, i32, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp,
                                            vp, sz, vp]),
        'cpgb_bn_add_relu_fwd': (ctypes.c_int, [vp, vp, i64, i32, i32, vp, vp, vp, vp, vp, i32, f32, f32, i32, vp, vp, vp,
                                                vp, sz, vp]),
        'cpgb_bn_add_relu_bwd': (ctypes.c_int, [vp, vp, vp, i64, i32, i32, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp,
                                                sz, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().cpgb_last_error()
        raise CpgbError(f'{what} failed with code {rc}: {msg.decode() if msg else ""}')


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CpgbError('cpg_b200 kernels need CUDA tensors; there is no CPU path '
                        '(the CPU restatement lives in oracle/ and is test-only)')
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def set_path(path):
    return load().cpgb_set_path(int(path))


def conv_desc(x_shape, x_strides, w_shape, y_shape, y_strides, stride, padding, dilation, groups, flags=0):
    d = ConvDesc()
    d.flags = int(flags)
    d.N, d.C, d.H, d.W = x_shape
    d.K, _, d.R, d.S = w_shape
    d.P, d.Q = y_shape[2], y_shape[3]
    d.stride_h, d.stride_w = stride
    d.pad_h, d.pad_w = padding
    d.dil_h, d.dil_w = dilation
    d.groups = groups
    for i in range(4):
        d.xs[i] = x_strides[i]
        d.ys[i] = y_strides[i]
    return d
