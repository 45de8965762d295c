// Shared host/device helpers for libcpgb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/cpgb200.h"

namespace cpgb {

// thread-local error text behind cpgb_last_error()
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
void count_launches(int n);

#define CPGB_CUDA_OK(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return cpgb::cuda_fail(_e, #expr); \
  } while (0)

#define CPGB_LAUNCH_OK_N(what, n)                             \
  do {                                                        \
    cudaError_t _e = cudaGetLastError();                      \
    if (_e != cudaSuccess) return cpgb::cuda_fail(_e, what);  \
    cpgb::count_launches(n);                                  \
  } while (0)
#define CPGB_LAUNCH_OK(what) CPGB_LAUNCH_OK_N(what, 1)

// cudaFuncSetAttribute (dynamic shared memory above 48 KB) is a per-DEVICE setting: a process that drives several GPUs
// (nn.DataParallel replicas, one thread per device) has to apply it on each of them.  One flag per call site and device.
struct PerDeviceOnce {
  bool done[64] = {};
  bool need() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) d = 0;
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// Launch with programmatic dependent launch enabled: the kernel's blocks are scheduled while the previous kernel on
// the stream drains; every kernel launched through here executes griddepcontrol.wait before its first global access.
// CPGB_NO_PDL=1 turns the attribute off.
inline bool pdl_enabled() {
  static const bool on = getenv("CPGB_NO_PDL") == nullptr;
  return on;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// the same with the grid's x extent grouped into thread-block clusters of (cluster_x, 1, 1)
template <class... KArgs, class... Args>
inline cudaError_t launch_dependent_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                            int cluster_x, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster_x; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Device-side copy of cpgb_conv_desc with the derived per-group sizes.
struct Geom {
  int N, C, H, W, K, R, S, P, Q;
  int sh, sw, ph, pw, dh, dw, groups;
  int Cg, Kg, RS, PQ, HW;
  long long xs0, xs1, xs2, xs3, ys0, ys1, ys2, ys3;
};

inline Geom make_geom(const cpgb_conv_desc &d) {
  Geom g;
  g.N = d.N; g.C = d.C; g.H = d.H; g.W = d.W; g.K = d.K; g.R = d.R; g.S = d.S; g.P = d.P; g.Q = d.Q;
  g.sh = d.stride_h; g.sw = d.stride_w; g.ph = d.pad_h; g.pw = d.pad_w; g.dh = d.dil_h; g.dw = d.dil_w;
  g.groups = d.groups;
  g.Cg = d.C / d.groups; g.Kg = d.K / d.groups; g.RS = d.R * d.S; g.PQ = d.P * d.Q; g.HW = d.H * d.W;
  g.xs0 = d.xs[0]; g.xs1 = d.xs[1]; g.xs2 = d.xs[2]; g.xs3 = d.xs[3];
  g.ys0 = d.ys[0]; g.ys1 = d.ys[1]; g.ys2 = d.ys[2]; g.ys3 = d.ys[3];
  return g;
}

int validate_desc(const cpgb_conv_desc *d);

// Binarizer applied to a weight: (piggy > thr ? 1 : piggy <= thr ? 0 : piggy[NaN]) * w,
// a true multiply like models/layers.py:103 so NaN/Inf propagate exactly as in the reference.
__device__ __forceinline__ float binarize_val(float p, float thr) {
  return p > thr ? 1.0f : (p <= thr ? 0.0f : p);
}
__device__ __forceinline__ float masked_weight(float w, const float *piggy, long long idx, float thr) {
  return piggy ? binarize_val(__ldg(piggy + idx), thr) * w : w;
}

// The fused gradient epilogue for one element: g is the raw weight gradient dL/dW_eff.
//   RAW      : dW = g*b                     dP = g*W                      (models/layers.py:21-23,103)
//   FINETUNE : dW = (g*b + wd*W)[T==cur]    dP = (g*W)[1<=T<cur]          (utils/prune.py:203-208)
//   PRUNE    : dW = (g*b + wd*W)[T==cur]    dP = 0                        (utils/prune.py:203-205,210)
// CPGB_GRAD_MERGED (or'ed into FINETUNE / PRUNE): the two results have disjoint support, so their sum is one
// buffer that carries both through the data-parallel all-reduce; it is returned in dw.
__device__ __forceinline__ void grad_epilogue_elem(float g, float w, float p, bool has_p, unsigned t, int cur, float wd,
                                                   int mode, float thr, float &dw, float &dp) {
  const float gb = has_p ? g * binarize_val(p, thr) : g;
  const int m = mode & 3;
  if (m == CPGB_GRAD_RAW) { dw = gb; dp = g * w; return; }
  dw = (t == (unsigned)cur) ? fmaf(wd, w, gb) : 0.f;
  dp = (m == CPGB_GRAD_FINETUNE && has_p && t != 0u && t < (unsigned)cur) ? g * w : 0.f;
  if (mode & CPGB_GRAD_MERGED) dw += dp;
}

// ---- entry points implemented in the individual .cu files ----
// CUDA-core (fp32 FFMA) implicit GEMM, any geometry.
int simt_fprop(const Geom &g, const float *x, const float *w, const float *piggy, const float *bias,
               float *y, float thr, cudaStream_t st);
int simt_dgrad(const Geom &g, const float *dy, const float *w, const float *piggy, float *dx, float thr,
               cudaStream_t st);
// raw weight-gradient partial sums gbuf[splits][K*Cg*R*S] (fp32, overwritten; fixed reduction order)
int simt_wgrad_splits(const Geom &g);
int simt_wgrad_raw(const Geom &g, const float *x, const float *dy, float *gbuf, int *splits_out, cudaStream_t st);
// scratch: bias_grad_scratch_bytes(g) bytes (0: none needed) for the two-phase NHWC column sum; without it (or for other
// layouts) a slow one-block-per-channel kernel runs
size_t bias_grad_scratch_bytes(const Geom &g);
int bias_grad(const Geom &g, const float *dy, float *dbias, void *scratch, size_t scratch_bytes, cudaStream_t st);

// out = rna_tf32(in), elementwise (elementwise.cu)
int round_tf32(const float *in, float *out, long long n, cudaStream_t st);

// fused epilogue g -> (dW, dP)  
int wgrad_epilogue(const float *gbuf, int splits, const float *w, const float *piggy, const uint8_t *tmask,
                   long long n, int cur, float wd, int mode, float thr, float *dW, float *dP, cudaStream_t st);

// direct fp32 kernels for the 3-channel 3x3 stem (stem_conv.cu); y / dy must be NHWC
bool stem_eligible(const cpgb_conv_desc &d);
bool stem_weight_shape(int K, int C, int R, int S, int groups);
size_t stem_workspace_bytes(const cpgb_conv_desc &d);
int stem_fprop(const cpgb_conv_desc &d, const float *x, const float *w, const float *piggy, const float *bias, float *y,
               float thr, cudaStream_t st);
int stem_wgrad_fused(const cpgb_conv_desc &d, const float *x, const float *dy, const float *w, const float *piggy,
                     const uint8_t *tmask, int cur, float wd, int mode, float thr, float *dW, float *dP, void *ws,
                     size_t ws_bytes, cudaStream_t st);

// tcgen05 implicit GEMM (tc_conv.cu).  tc_eligible() says whether the shape is supported.
bool tc_eligible(const cpgb_conv_desc &d, int op);  // op: 0 fprop, 1 dgrad, 2 wgrad
size_t tc_workspace_bytes(const cpgb_conv_desc &d);
size_t tc_staged_bytes(const cpgb_conv_desc &d);
void debug_set_mn(int layout, int lbo, int sbo, int kadv, int tma_swizzle);
// staged = tf32((piggy > thr) * w) reordered to [K][R*S][Cp]
int tc_stage_weights(const cpgb_conv_desc &d, const float *w, const float *piggy, float thr, void *staged,
                     size_t bytes, cudaStream_t st);
size_t tc_staged_bytes_for_weight(int K, int C, int R, int S, int stride_h, int stride_w, int groups);
int tc_stage_weights_batched(int n, const float *const *w, const float *const *piggy, void *const *staged,
                             const int *K, const int *C, const int *R, const int *S, const int *stride_h,
                             const int *stride_w, const float *thr, cudaStream_t st);
// part: scratch for split-K partial sums (tc_workspace_bytes covers staged operand + partials)
// raw: `staged` is the module's fp32 weight tensor itself (tc_weights_usable_raw)
bool tc_weights_usable_raw(const cpgb_conv_desc &d);
// colstats (may be NULL): [tc_fprop_colstats_parts(d)][up4(K)][2] per-tile column sums / sums of squares of y
int tc_fprop(const cpgb_conv_desc &d, const float *x, const float *staged, const float *bias, float *y, void *part,
             size_t part_bytes, cudaStream_t st, bool raw, float *colstats = nullptr);
int tc_fprop_colstats_parts(const cpgb_conv_desc &d);
int tc_dgrad(const cpgb_conv_desc &d, const float *dy, const float *staged, float *dx, void *part, size_t part_bytes,
             cudaStream_t st, bool raw);
// in-tile weight masking (CPGB_FLAG_W_INTILE): B operand = the raw weight tensor, masked from packed bits in shared memory
bool tc_intile_eligible(const cpgb_conv_desc &d);
bool tc_intile_weight_shape(int K, int C, int R, int S, int stride_h, int stride_w, int groups);
int tc_fprop_intile(const cpgb_conv_desc &d, const float *x, const float *w, const void *bits, const float *bias, float *y,
                    void *part, size_t part_bytes, cudaStream_t st);
int tc_dgrad_intile(const cpgb_conv_desc &d, const float *dy, const float *w, const void *bits, float *dx, void *part,
                    size_t part_bytes, cudaStream_t st);
// wgrad + fused epilogue (dW, dP); partial sums live in ws.  est: stream of the epilogue kernels (st, or a second
// stream the caller joins before dW / dP are consumed)
int tc_wgrad_fused(const cpgb_conv_desc &d, const float *x, const float *dy, const float *w, const float *piggy,
                   const uint8_t *tmask, int cur, float wd, int mode, float thr, float *dW, float *dP, void *ws,
                   size_t ws_bytes, cudaStream_t st, cudaStream_t est);

}  // namespace cpgb
