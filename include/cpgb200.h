/*
 * cpgb200.h -- C ABI of libcpgb200.so: the B200 (sm_100a) implementation of CPG's
 * masked-convolution train/prune hot path.
 *
 * The reference (ivclab/CPG) is pure Python and has no FFI; the functions below replace
 * the *Python expressions* cited next to each of them (paths relative to the reference
 * tree).  INTEGRATION.md shows the ctypes binding a CPG maintainer would add.
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE pointer (PyTorch's allocator); the library
 *     never allocates, frees or synchronises; scratch comes from the caller (`ws`);
 *   - every call takes the cudaStream_t to launch on (as void*), is re-entrant and
 *     capturable into a CUDA graph;
 *   - returns 0 on success, a negative CPGB_E* code otherwise; cpgb_last_error() gives a
 *     thread-local message;
 *   - tensors are fp32; task masks are uint8 (0 = free/pruned, k = owned by task k), the
 *     checkpoint layout of the reference (CPG_cifar100_main_normal.py:201-207);
 *   - activations are addressed through explicit element strides (n,c,h,w), so NCHW and
 *     channels_last (NHWC) both work; weights are dense [K, C/groups, R, S] as stored by
 *     the reference's modules (models/layers.py:80-81, 169-170).
 */
#ifndef CPGB200_H_
#define CPGB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPGB_VERSION 100

/* error codes */
#define CPGB_OK 0
#define CPGB_EINVAL (-1)    /* bad argument / unsupported shape               */
#define CPGB_EWORKSPACE (-2) /* workspace too small (see cpgb_workspace_bytes) */
#define CPGB_ECUDA (-3)     /* a CUDA runtime/driver call failed              */
#define CPGB_ENOTELIGIBLE (-4) /* tcgen05 path forced but shape not eligible  */

/* gradient-epilogue modes (SURVEY a4 + a6) */
#define CPGB_GRAD_RAW 0      /* autograd contract: dW = g*b, dP = g*W  (models/layers.py:21-23,103) */
#define CPGB_GRAD_FINETUNE 1 /* + utils/prune.py:203-208: dW=(g*b+wd*W)[T==cur]; dP=g*W[1<=T<cur]    */
#define CPGB_GRAD_PRUNE 2    /* + utils/prune.py:203-205,209-210: same dW; dP = 0                     */
/* or'ed into FINETUNE / PRUNE for cpgb_conv2d_wgrad_fused: after a6 dW and dP have disjoint support (T==cur vs
 * 1<=T<cur), so ONE buffer m = dW + dP is written to `dW` (pass dP = NULL) and travels through the data-parallel
 * all-reduce; cpgb_split_merged_grad restores the pair afterwards (SURVEY 8e "Collective"). */
#define CPGB_GRAD_MERGED 4

/* kernel-path selection (process-wide, for tests/benchmarks) */
#define CPGB_PATH_AUTO 0     /* tcgen05 implicit GEMM when eligible, else CUDA-core kernels */
#define CPGB_PATH_SIMT 1     /* CUDA-core (fp32 FFMA) kernels only                         */
#define CPGB_PATH_TCGEN05 2  /* tcgen05 only; CPGB_ENOTELIGIBLE when the shape is not      */

/* cpgb_conv_desc.flags.  The tcgen05 kernels multiply TF32 operands (fp32 storage, 10 explicit mantissa
 * bits) and the tensor core TRUNCATES whatever fp32 bits it is handed.  To get round-to-nearest TF32 -- an
 * unbiased result -- the library rounds x (fprop, wgrad) and dy (dgrad, wgrad) into `ws` before the GEMM
 * unless the caller promises that the tensor already holds TF32-representable values, e.g. because it was
 * produced by cpgb_bn_relu_fwd / _bwd with tf32_out != 0 or by cpgb_round_tf32.  The CUDA-core and stem
 * kernels compute in fp32 and ignore the flags. */
#define CPGB_FLAG_X_TF32 1   /* x  holds TF32-representable values: no rounding pre-pass */
#define CPGB_FLAG_DY_TF32 2  /* dy holds TF32-representable values: no rounding pre-pass */
/* In-tile weight masking (north_star: "weight tiles staged via TMA, ANDed in shared memory with the packed bitmask").
 * With this flag cpgb_conv2d_fprop / _dgrad TMA-load tiles of the module's own fp32 weight tensor `w`, and the CTA's
 * epilogue warps -- idle during the main loop -- multiply each landed tile by the Binarizer bits and round it to TF32
 * in shared memory before the MMAs read it: neither the masked weight (models/layers.py:101-103) nor a staged copy is
 * ever written to global memory.  `staged` must then point to the layer's cpgb_pack_mask words (ignored when
 * piggy == NULL: round only).  Only for layers cpgb_intile_eligible() accepts: linear / 1x1, C % 32 == 0, at most 256
 * output pixels (the FC layers of VGG16: one pixel tile, every weight tile is consumed exactly once per pass); for
 * convolutions with many pixel tiles the in-shared-memory work would be repeated per tile and the staged operand wins. */
#define CPGB_FLAG_W_INTILE 4

/* Geometry of one SharableConv2d call: F.conv2d(input, weight, bias, stride, padding,
 * dilation, groups) at models/layers.py:108.  SharableLinear (models/layers.py:194) is the
 * same descriptor with H=W=R=S=P=Q=1 (see cpgb_linear_desc). */
typedef struct cpgb_conv_desc {
  int32_t N, C, H, W;          /* input  [N, C, H, W]                      */
  int32_t K, R, S;             /* weight [K, C/groups, R, S]               */
  int32_t P, Q;                /* output [N, K, P, Q]                      */
  int32_t stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, groups;
  int32_t flags;               /* CPGB_FLAG_* (0 is always valid)          */
  int64_t xs[4];               /* element strides of input  (n, c, h, w)   */
  int64_t ys[4];               /* element strides of output (n, k, p, q)   */
} cpgb_conv_desc;

int cpgb_version(void);
const char *cpgb_last_error(void);
int cpgb_set_path(int path);  /* CPGB_PATH_*; returns previous */
int cpgb_get_path(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches claim) */
int64_t cpgb_launch_count(void);

/* Fill a descriptor for F.linear(x[M, I], W[O, I], bias) -- models/layers.py:194. */
void cpgb_linear_desc(cpgb_conv_desc *d, int32_t M, int32_t I, int32_t O);

/* Scratch bytes the conv/linear entry points need for descriptor `d` (max over
 * fprop/dgrad/wgrad; depends on d->flags: a tensor the library must round first needs room for the copy). */
size_t cpgb_workspace_bytes(const cpgb_conv_desc *d);

/* 1 when pass `op` (0 fprop, 1 dgrad, 2 wgrad) of descriptor `d` runs on the tcgen05 kernels under the
 * current path selection (so that a caller can pre-round its operands once and share them between passes). */
int cpgb_uses_tensor_cores(const cpgb_conv_desc *d, int32_t op);

/* out[i] = round-to-nearest-TF32(in[i]) (cvt.rna.tf32.f32: ties away from zero, NaN/Inf kept); in == out is
 * allowed.  What the masked weight operand gets inside cpgb_stage_weights, for activations. */
int cpgb_round_tf32(const float *in, float *out, int64_t n, void *stream);

/* a1: Binarizer.forward, models/layers.py:15-19.  b = (p > thr) ? 1 : 0, NaN -> NaN. */
int cpgb_binarize(const float *piggy, float *out, int64_t n, float thr, void *stream);

/* the {piggyback, task} bit masks of a layer, packed.  packed[g] (g = i / 32, n rounded up to 32 elements)
 * holds for elements 32g .. 32g+31: low 32 bits, bit j = (piggy[32g+j] > thr) -- Binarizer.forward, models/layers.py:
 * 15-19 (a NaN piggymask packs as 0; the unpacked kernels propagate it) -- all ones when piggy == NULL; high 32 bits,
 * bit j = (1 <= T[32g+j] <= inference_idx) -- the elements apply_mask keeps, utils/prune.py:229-230 -- all ones when
 * tmask == NULL.  4.1 B/element read, 0.25 B/element written.  Consumers: cpgb_conv2d_fprop / _dgrad with
 * CPGB_FLAG_W_INTILE (the weight tile is masked in shared memory from these bits). */
int cpgb_pack_mask(const float *piggy, const uint8_t *tmask, int64_t n, float thr, int32_t inference_idx, uint64_t *packed,
                   void *stream);

int cpgb_intile_eligible(const cpgb_conv_desc *d);
/* weight-only view of the same decision (a model-level staging pass skips such layers; the batch size decides later) */
int cpgb_intile_weight_shape(int32_t K, int32_t C, int32_t R, int32_t S, int32_t stride_h, int32_t stride_w,
                             int32_t groups);

/* Staged weight operand of the tcgen05 path: tf32_rna((piggy > thr ? 1 : 0) * w) reordered from
 * the module's [K][C/g][R][S] to [K][R*S][Cp] (Cp = C rounded up to 32, zero padded) -- the
 * expression models/layers.py:101-103 evaluated once per layer per step.  fprop and dgrad of the
 * same step share it: build it once with cpgb_stage_weights() and pass it as `staged`, or pass
 * staged == NULL and let each call build it in `ws`.  cpgb_staged_weight_bytes() returns 0 when
 * the descriptor takes the CUDA-core path (then `staged` is ignored). */
size_t cpgb_staged_weight_bytes(const cpgb_conv_desc *d);
int cpgb_stage_weights(const cpgb_conv_desc *d, const float *w, const float *piggy, float thr, void *staged,
                       size_t staged_bytes, void *stream);

/* Linear / 1x1 layers without a piggymask (task 1) need no staging: the fp32 weight tensor has the
 * operand's layout and the tensor core truncates it to TF32.  When this returns 1 the caller may pass
 * `staged = w` (or NULL) to cpgb_conv2d_fprop / _dgrad and skip cpgb_stage_weights. */
int cpgb_weights_usable_raw(const cpgb_conv_desc *d, int32_t has_piggymask);
int cpgb_weights_usable_raw_for(int32_t K, int32_t C, int32_t R, int32_t S, int32_t stride_h, int32_t stride_w,
                                int32_t groups, int32_t has_piggymask);

/* The same staging for every sharable layer of a model in ONE launch (weights do not depend on the
 * activations, so a model-level pre-forward hook can build all operands at once).  All arrays are HOST
 * arrays of n entries; piggy[i] may be NULL; staged[i] must hold
 * cpgb_staged_weight_bytes_for(K, C, R, S, stride_h, stride_w, groups) bytes (0 = this weight never takes
 * the tensor-core path: skip it).  The layout produced is the one cpgb_conv2d_fprop / _dgrad expect for a
 * descriptor with that weight shape and stride. */
size_t cpgb_staged_weight_bytes_for(int32_t K, int32_t C, int32_t R, int32_t S, int32_t stride_h, int32_t stride_w,
                                    int32_t groups);
int cpgb_stage_weights_batched(int32_t n, const float *const *w, const float *const *piggy, void *const *staged,
                               const int32_t *K, const int32_t *C, const int32_t *R, const int32_t *S,
                               const int32_t *stride_h, const int32_t *stride_w, const float *thr, void *stream);

/* a3/a5 forward: y = conv2d(x, (piggy > thr ? 1 : 0) * w, bias).  models/layers.py:98-109,
 * 184-194.  piggy == NULL means "no piggymask" (task 1, models/layers.py:104-105).  The
 * CUDA-core path evaluates the predicate while loading weight tiles; the tcgen05 path reads
 * the staged operand described above. */
int cpgb_conv2d_fprop(const cpgb_conv_desc *d, const float *x, const float *w, const float *piggy,
                      const float *bias, float *y, float thr, const void *staged, void *ws, size_t ws_bytes,
                      void *stream);
/* fprop that also hands the batch-norm behind it (conv -> nn.BatchNorm2d, models/vgg.py:109-118) its statistics: while
 * the epilogue stores y it accumulates, per pixel tile, the column sums and sums of squares of exactly the stored values
 * into colstats[cpgb_fprop_colstats_parts(d)][K rounded up to 4][2] -- the layout of the batch-norm kernels' partial
 * sums -- and cpgb_bn_relu_fwd_stats then skips its statistics pass over y.  cpgb_fprop_colstats_parts(d) == 0: this
 * layer / path cannot (stem, im2col tier, CUDA-core path, split-K plans, in-tile masked layers); colstats must then be
 * NULL.  colstats == NULL is cpgb_conv2d_fprop. */
int32_t cpgb_fprop_colstats_parts(const cpgb_conv_desc *d);
int cpgb_conv2d_fprop_stats(const cpgb_conv_desc *d, const float *x, const float *w, const float *piggy,
                            const float *bias, float *y, float thr, const void *staged, void *ws, size_t ws_bytes,
                            float *colstats, void *stream);

/* a4 dgrad: dx = conv_transpose(dy, W_eff).  dy uses d->ys strides, dx uses d->xs. */
int cpgb_conv2d_dgrad(const cpgb_conv_desc *d, const float *dy, const float *w, const float *piggy,
                      float *dx, float thr, const void *staged, void *ws, size_t ws_bytes, void *stream);

/* a4 wgrad with the fused epilogue.  g = wgrad(x, dy) is reduced in `ws`
 * and never returned; outputs:
 *   dW [K,C/g,R,S] (always), dP (same shape; NULL iff piggy == NULL), dbias [K] (NULL ok).
 * mode = CPGB_GRAD_RAW needs no tmask; the two fused modes read the uint8 task mask and
 * apply utils/prune.py:195-211 in the same pass. */
int cpgb_conv2d_wgrad_fused(const cpgb_conv_desc *d, const float *x, const float *dy, const float *w,
                            const float *piggy, const uint8_t *tmask, int32_t cur, float weight_decay,
                            int32_t mode, float *dW, float *dP, float *dbias, float thr, void *ws,
                            size_t ws_bytes, void *stream);
/* The same with the fused epilogue (the sum of the split partial sums + K6-K8) queued on `epilogue_stream` behind an
 * event recorded on `stream` after the GEMM: the latency-bound epilogue of one layer then runs under the GEMM of the
 * next layer queued on `stream`.  dW / dP are complete once `epilogue_stream` has drained: the caller joins it before
 * the optimizer / all-reduce reads them (cpg_b200/functional.py joins once, at the end of the backward pass).  Kernels
 * that finish the gradient themselves (stem, im2col tier, CUDA-core path) ignore epilogue_stream; dbias is written on
 * `stream`.  epilogue_stream == stream is cpgb_conv2d_wgrad_fused. */
int cpgb_conv2d_wgrad_fused_async(const cpgb_conv_desc *d, const float *x, const float *dy, const float *w,
                                  const float *piggy, const uint8_t *tmask, int32_t cur, float weight_decay,
                                  int32_t mode, float *dW, float *dP, float *dbias, float thr, void *ws,
                                  size_t ws_bytes, void *stream, void *epilogue_stream);

/* dbias[k] = sum over (n, p, q) of dy -- the bias gradient of F.conv2d / F.linear (models/layers.py:108,194) on its
 * own, in fp32.  cpgb_conv2d_wgrad_fused computes it from the dy it is given; a caller that hands that function a
 * TF32-rounded copy of dy (CPGB_FLAG_DY_TF32) passes dbias = NULL there and calls this with the unrounded tensor. */
int cpgb_conv2d_bias_grad(const cpgb_conv_desc *d, const float *dy, float *dbias, void *stream);
/* The same with scratch: ws = a cpgb_workspace_bytes(d) buffer (the one handed to cpgb_conv2d_wgrad_fused will do: the
 * partial column sums use a region of their own at its end).  NHWC dy is then summed by row-streaming blocks in two
 * deterministic phases instead of one strided walk per channel. */
int cpgb_conv2d_bias_grad_ws(const cpgb_conv_desc *d, const float *dy, float *dbias, void *ws, size_t ws_bytes,
                             void *stream);

/* a6 standalone, in place on existing gradients: utils/prune.py:195-211.
 * mode is CPGB_GRAD_FINETUNE or CPGB_GRAD_PRUNE; dW / dP may each be NULL
 * ("if module.weight.grad is not None", "if module.piggymask is not None"). */
int cpgb_grad_epilogue(float *dW, float *dP, const float *w, const uint8_t *tmask, int64_t n,
                       int32_t cur, float weight_decay, int32_t mode, void *stream);

/* a7: SparsePruner._pruning_mask, utils/prune.py:30-53, entirely on the device.
 *   pool = {i : T[i]==cur or T[i]==0};  k = round_half_even(ratio * |pool|)  (python round());
 *   cut = k-th smallest |w[pool]| (exact radix select on the fp32 bit pattern);
 *   T[i] = 0 where |w[i]| <= cut and T[i] == cur.
 * info (device, 4 x int64): [0]=status (0 ok, 2 = k outside 1..|pool| -> the reference's
 * sys.exit(2) path, T untouched), [1]=|pool|, [2]=k, [3]=bit pattern of cut (low 32 bits).
 * ws: cpgb_prune_workspace_bytes() bytes. */
size_t cpgb_prune_workspace_bytes(void);
int cpgb_prune_select(const float *w, uint8_t *tmask, int64_t n, int32_t cur, double ratio,
                      int64_t *info, void *ws, size_t ws_bytes, void *stream);

/* a7 for every sharable layer of the model in one call (what gradually_prune does at
 * utils/prune.py:84-90): seven launches in total instead of seven per layer.  w / tmask / n are HOST
 * arrays of nlayers (<= 64) device pointers / element counts; info is device [nlayers][4] with the
 * per-layer meaning of cpgb_prune_select; ws: cpgb_prune_batched_workspace_bytes(nlayers). */
size_t cpgb_prune_batched_workspace_bytes(int32_t nlayers);
int cpgb_prune_select_batched(int32_t nlayers, const float *const *w, uint8_t *const *tmask, const int64_t *n,
                              int32_t cur, double ratio, int64_t *info, void *ws, size_t ws_bytes, void *stream);

/* The same result as cpgb_prune_select_batched with TWO streaming passes over W / T instead of four (a7 is HBM-bound:
 * 5 B/element per pass): a sorted sample of 16384 pool keys per layer brackets the k-th magnitude, one pass counts
 * |pool|, the keys below the bracket and a 2048-bin histogram inside it, a second pass prunes everything below the
 * k-th element's bin and collects that bin's few elements, which one block per layer then selects from exactly.  The
 * sample only places the bracket; counting and selection are exact.  info[l][0] == 3 reports a layer whose bracket
 * missed the k-th element or whose bin overflowed the candidate list (possible only for adversarial value
 * distributions): its mask holds a correct partial result and the caller must run cpgb_prune_select_batched on that
 * layer to finish it (cpg_b200/prune.py does).  Layers of up to 2^32 - 1 elements.
 * ws: cpgb_prune_sampled_workspace_bytes(nlayers), 8-byte aligned. */
size_t cpgb_prune_sampled_workspace_bytes(int32_t nlayers);
int cpgb_prune_select_sampled(int32_t nlayers, const float *const *w, uint8_t *const *tmask, const int64_t *n,
                              int32_t cur, double ratio, int64_t *info, void *ws, size_t ws_bytes, void *stream);

/* a9: apply_mask (utils/prune.py:223-231): w[T==0]=0; w[T>inference_idx]=0.
 *     make_pruned_zero (utils/prune.py:213-221): pass inference_idx = 255. */
int cpgb_apply_mask(float *w, const uint8_t *tmask, int64_t n, int32_t inference_idx, void *stream);

/* a10: make_finetuning_mask (utils/prune.py:233-243): T[T==0] = new_cur. */
int cpgb_make_finetuning_mask(uint8_t *tmask, int64_t n, int32_t new_cur, void *stream);

/* K12: counts behind calculate_{sparsity,curr_task_ratio,zero_ratio,shared_part_ratio}
 * (utils/prune.py:111-193).  Accumulates (+=) into out[5] (device int64):
 * [0]=#(T==0) [1]=#(T==idx) [2]=#(0<T<idx) [3]=#(0<T<idx and piggy>0.005) [4]=n.
 * piggy may be NULL. */
int cpgb_mask_stats(const uint8_t *tmask, const float *piggy, int64_t n, int32_t inference_idx,
                    int64_t *out, void *stream);

/* The same counters accumulated over nlayers masks in one launch; tmask / piggy / n are HOST arrays
 * (piggy may be NULL, or hold NULL entries). */
int cpgb_mask_stats_batched(int32_t nlayers, const uint8_t *const *tmask, const float *const *piggy, const int64_t *n,
                            int32_t inference_idx, int64_t *out, void *stream);

/* Data-parallel helpers: dW and dP have disjoint support after a6, so one fp32
 * buffer m = dW + dP travels through the all-reduce; cpgb_split_merged_grad restores
 * dW = m[T==cur], dP = m[1<=T<cur]. */
int cpgb_merge_grads(const float *dW, const float *dP, float *merged, int64_t n, void *stream);
int cpgb_split_merged_grad(const float *merged, const uint8_t *tmask, int64_t n, int32_t cur,
                           float *dW, float *dP, void *stream);

/* SURVEY 8(f) N2 -- the optimizer step right after the path (CPG_cifar100_main_normal.py:339-346), one launch per
 * optimizer instead of one ATen kernel per operation and parameter list.  param / grad / state arrays are HOST arrays of
 * `ntensors` device pointers, n the element counts.  Every element is updated, also where the masked gradient is zero
 * (momentum drift of pruned weights).  The arithmetic reproduces torch's multi-tensor implementation
 * operation by operation (same fp32 roundings); state buffers start as zeros, as torch's do.
 *
 * cpgb_sgd_nesterov_step: optim.SGD(lr, momentum, nesterov=True, weight_decay=0, dampening=0):
 *   buf = momentum * buf + g;  p -= lr * (g + momentum * buf).   lr_dev (may be NULL): device fp32 that overrides lr
 *   (a learning-rate schedule inside a captured CUDA graph). */
int cpgb_sgd_nesterov_step(int32_t ntensors, float *const *param, const float *const *grad, float *const *momentum_buf,
                           const int64_t *n, float lr, float momentum, const float *lr_dev, void *stream);
/* cpgb_adam_step: optim.Adam(lr, betas, eps, weight_decay=0, amsgrad=False) on at most 40 tensors per call.
 *   step_dev: device int64[2], zero-initialised by the caller once: [0] = steps taken so far (incremented by the
 *   kernel: capturable), [1] = scratch.  lr_dev (may be NULL): device fp64 override of lr.
 *   packed (may be NULL, entries may be NULL): per tensor, receives the cpgb_pack_mask words of the UPDATED parameter
 *   (low half bit j = (p > thr), high half bit j = (1 <= tmask <= inference_idx), all ones when tmask[i] == NULL), so
 *   the next forward pass of an in-tile masked layer (CPGB_FLAG_W_INTILE) needs no pack pass. */
int cpgb_adam_step(int32_t ntensors, float *const *param, const float *const *grad, float *const *exp_avg,
                   float *const *exp_avg_sq, const int64_t *n, double lr, double beta1, double beta2, double eps,
                   int64_t *step_dev, const double *lr_dev, uint64_t *const *packed, const uint8_t *const *tmask,
                   float thr, int32_t inference_idx, void *stream);

/* SURVEY 8(f) N4 -- the consumer of every masked convolution: nn.BatchNorm2d followed by
 * nn.ReLU(inplace=True) (models/vgg.py:109-118, models/resnet.py:60-100), on NHWC fp32 activations seen as
 * [M = N*H*W][C] (pixel stride `ldc`).  Semantics of torch.nn.functional.batch_norm (+ relu):
 *   training != 0: batch statistics (biased variance) normalise; running_mean / running_var (may be NULL) are
 *                  updated in place with `momentum` and the unbiased variance, *num_batches_tracked (int64, may
 *                  be NULL) is incremented as nn.BatchNorm2d.forward does; save_mean / save_rstd [C] receive
 *                  the batch mean and 1/sqrt(var + eps) for the backward pass;
 *   training == 0: running statistics normalise, save_* are not written.
 * gamma / beta may be NULL (affine=False).  relu != 0 applies max(0, .) in the same pass.
 * pool_h = H, pool_w = W (both even, M = N*H*W) additionally folds the nn.MaxPool2d(kernel_size=2, stride=2) that
 * follows conv -> BN -> ReLU at the 'M' entries of models/vgg.py:95-122 into the same pass: y (and dy of the
 * backward call) are then [N*(H/2)*(W/2)][C]; the window gradient goes to the first maximum in (h, w) order as
 * in torch.  pool_h = pool_w = 0: no pooling.
 * ldc: floats between consecutive pixels of x / y (and dy / dx): 0 = dense (C, then a multiple of 4), or any multiple
 * of 4 that is >= C -- how cpg_b200 stores activations whose channel count is not a multiple of 4 (the grown networks'
 * 78 / 313 / 627 channels).  The lanes C..ldc-1 of the inputs are ignored, those of the outputs are written as zeros.
 * tf32_out != 0: y (and dx of the backward call) are rounded to the nearest TF32-representable value as they are
 * stored, so the masked convolution that consumes them can set CPGB_FLAG_X_TF32 / CPGB_FLAG_DY_TF32 (one ALU
 * operation here instead of a rounding pass there; relative change of the output <= 2^-11).
 * ws: cpgb_bn_workspace_bytes(M, C) bytes of scratch (per-block partial sums, deterministic). */
size_t cpgb_bn_workspace_bytes(int64_t M, int32_t C);
int cpgb_bn_relu_fwd(const float *x, int64_t M, int32_t C, int32_t ldc, const float *gamma, const float *beta,
                     float *running_mean, float *running_var, int64_t *num_batches_tracked, int32_t training,
                     float momentum, float eps, int32_t relu, int32_t pool_h, int32_t pool_w, int32_t tf32_out, float *y,
                     float *save_mean, float *save_rstd, void *ws, size_t ws_bytes, void *stream);
/* The same with the statistics of x supplied by its producer: colstats[nparts][ldc or C][2] partial (sum, sum of squares)
 * pairs over disjoint pixel sets that together cover all M pixels (cpgb_conv2d_fprop_stats).  Used in training mode for
 * tensors that take the three-kernel path; otherwise (evaluation mode, colstats == NULL, small tensors on the
 * single-launch kernels) identical to cpgb_bn_relu_fwd. */
int cpgb_bn_relu_fwd_stats(const float *x, int64_t M, int32_t C, int32_t ldc, const float *colstats, int32_t nparts,
                           const float *gamma, const float *beta, float *running_mean, float *running_var,
                           int64_t *num_batches_tracked, int32_t training, float momentum, float eps, int32_t relu,
                           int32_t pool_h, int32_t pool_w, int32_t tf32_out, float *y, float *save_mean, float *save_rstd,
                           void *ws, size_t ws_bytes, void *stream);
/* Backward of the above: with g = dy * [y > 0] (relu) or dy,  xhat = (x - mean) * rstd:
 *   dbeta = sum g;  dgamma = sum g * xhat;
 *   training: dx = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat));   evaluation: dx = gamma * rstd * g.
 * mean / rstd: save_mean / save_rstd of the forward call (training) or running_mean / 1/sqrt(running_var + eps).
 * dgamma / dbeta may be NULL. */
int cpgb_bn_relu_bwd(const float *x, const float *dy, int64_t M, int32_t C, int32_t ldc, const float *gamma,
                     const float *beta, const float *mean, const float *rstd, int32_t training, int32_t relu,
                     int32_t pool_h, int32_t pool_w, int32_t tf32_out, float *dx, float *dgamma, float *dbeta, void *ws,
                     size_t ws_bytes, void *stream);

/* SURVEY 8(f) N4, ResNet blocks: the last batch-norm of a block is followed by the residual add and the ReLU
 * (models/resnet.py:50-55 BasicBlock, :92-98 Bottleneck: out = bn(out); out += identity; out = relu(out)).
 *   forward : y = max(0, batch_norm(x) + res)          -- statistics / running statistics / save_* / ldc / tf32_out
 *                                                          exactly as cpgb_bn_relu_fwd; res has x's layout
 *   backward: g = dy * [y > 0] (y: the stored output of the forward call -- what torch's ReLU backward gates on);
 *             dres = g (gradient of the identity branch);  dbeta = sum g;  dgamma = sum g * xhat;
 *             dx = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat))  (training; evaluation: gamma * rstd * g).
 * Three launches per direction (stats -> finalize -> apply); the three separate passes of the stock modules (batch-norm,
 * add, ReLU) read and write the tensor 5 + 3 times in the forward direction, this reads it 3 times and writes it once.
 * ws: cpgb_bn_workspace_bytes(M, C). */
int cpgb_bn_add_relu_fwd(const float *x, const float *res, int64_t M, int32_t C, int32_t ldc, const float *gamma,
                         const float *beta, float *running_mean, float *running_var, int64_t *num_batches_tracked,
                         int32_t training, float momentum, float eps, int32_t tf32_out, float *y, float *save_mean,
                         float *save_rstd, void *ws, size_t ws_bytes, void *stream);
int cpgb_bn_add_relu_bwd(const float *x, const float *y, const float *dy, int64_t M, int32_t C, int32_t ldc,
                         const float *gamma, const float *beta, const float *mean, const float *rstd, int32_t training,
                         int32_t tf32_out, float *dx, float *dres, float *dgamma, float *dbeta, void *ws, size_t ws_bytes,
                         void *stream);

/* SURVEY 8(f) N4, SphereNet-20: every masked convolution feeds nn.PReLU(channels) (models/spherenet.py:204-249).
 * NHWC fp32 activations as [M][C] with pixel stride ldc (0 = dense), alpha[C]:
 *   y = x > 0 ? x : alpha[c] * x;   dx = x > 0 ? dy : alpha[c] * dy;   dalpha[c] = sum_pixels (x > 0 ? 0 : x * dy)
 * (torch.nn.functional.prelu and its backward, deterministic partial sums).  tf32_out as in cpgb_bn_relu_fwd: the
 * outputs feed the tcgen05 convolutions without a rounding pass.  ws: cpgb_prelu_workspace_bytes(M, C). */
size_t cpgb_prelu_workspace_bytes(int64_t M, int32_t C);
int cpgb_prelu_fwd(const float *x, int64_t M, int32_t C, int32_t ldc, const float *alpha, int32_t tf32_out, float *y,
                   void *stream);
int cpgb_prelu_bwd(const float *x, const float *dy, int64_t M, int32_t C, int32_t ldc, const float *alpha,
                   int32_t tf32_out, float *dx, float *dalpha, void *ws, size_t ws_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CPGB200_H_ */
