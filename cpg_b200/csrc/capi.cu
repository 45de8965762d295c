// extern "C" surface of libcpgb200.so: argument validation, path dispatch (tcgen05 implicit
// GEMM vs CUDA-core kernels), workspace accounting.  See include/cpgb200.h for the contract.
#include <atomic>
#include <string.h>

#include "common.cuh"

namespace cpgb {

static thread_local char g_err[512] = "";
static std::atomic<int> g_path{CPGB_PATH_AUTO};
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return CPGB_ECUDA;
}

int validate_desc(const cpgb_conv_desc *d) {
  if (!d) { set_error("null descriptor"); return CPGB_EINVAL; }
  if (d->N < 0 || d->C <= 0 || d->H <= 0 || d->W <= 0 || d->K <= 0 || d->R <= 0 || d->S <= 0) {
    set_error("bad tensor extents"); return CPGB_EINVAL;
  }
  if (d->groups <= 0 || d->C % d->groups || d->K % d->groups) {
    // models/layers.py:65-68 raises ValueError for these
    set_error("in/out channels must be divisible by groups"); return CPGB_EINVAL;
  }
  if (d->stride_h <= 0 || d->stride_w <= 0 || d->dil_h <= 0 || d->dil_w <= 0 || d->pad_h < 0 || d->pad_w < 0) {
    set_error("bad stride/dilation/padding"); return CPGB_EINVAL;
  }
  int P = (d->H + 2 * d->pad_h - d->dil_h * (d->R - 1) - 1) / d->stride_h + 1;
  int Q = (d->W + 2 * d->pad_w - d->dil_w * (d->S - 1) - 1) / d->stride_w + 1;
  if (P != d->P || Q != d->Q || P <= 0 || Q <= 0) {
    set_error("output extent mismatch: expected P=%d Q=%d, got P=%d Q=%d", P, Q, d->P, d->Q);
    return CPGB_EINVAL;
  }
  if ((long long)d->N * d->P * d->Q >= (1ll << 31) || (long long)d->N * d->H * d->W >= (1ll << 31) ||
      (long long)(d->C / d->groups) * d->R * d->S >= (1ll << 31)) {
    set_error("GEMM extent exceeds int32"); return CPGB_EINVAL;
  }
  return CPGB_OK;
}

static inline size_t weight_elems(const cpgb_conv_desc *d) {
  return (size_t)d->K * (size_t)(d->C / d->groups) * (size_t)d->R * (size_t)d->S;
}

}  // namespace cpgb

using namespace cpgb;

extern "C" {

int cpgb_version(void) { return CPGB_VERSION; }
const char *cpgb_last_error(void) { return g_err; }
int cpgb_set_path(int path) {
  if (path < CPGB_PATH_AUTO || path > CPGB_PATH_TCGEN05) return g_path.load();
  return g_path.exchange(path);
}
int cpgb_get_path(void) { return g_path.load(); }
int64_t cpgb_launch_count(void) { return (int64_t)g_launches.load(); }

void cpgb_linear_desc(cpgb_conv_desc *d, int32_t M, int32_t I, int32_t O) {
  memset(d, 0, sizeof(*d));
  d->N = M; d->C = I; d->H = 1; d->W = 1; d->K = O; d->R = 1; d->S = 1; d->P = 1; d->Q = 1;
  d->stride_h = d->stride_w = 1; d->dil_h = d->dil_w = 1; d->pad_h = d->pad_w = 0; d->groups = 1;
  d->xs[0] = I; d->xs[1] = 1; d->xs[2] = I; d->xs[3] = I;
  d->ys[0] = O; d->ys[1] = 1; d->ys[2] = O; d->ys[3] = O;
}

static size_t bias_scratch_bytes(const cpgb_conv_desc *d) {
  return (bias_grad_scratch_bytes(make_geom(*d)) + 255) & ~(size_t)255;
}

size_t cpgb_workspace_bytes(const cpgb_conv_desc *d) {
  if (!d || d->groups <= 0) return 0;
  // raw weight-gradient partial sums of the CUDA-core wgrad: one tensor per split of the pixel reduction
  size_t g_bytes = weight_elems(d) * sizeof(float) * (size_t)simt_wgrad_splits(make_geom(*d));
  size_t tc = tc_workspace_bytes(*d);                          // staged operand / split-K partial sums
  size_t stem = stem_workspace_bytes(*d);                      // per-block partial sums of the stem wgrad
  g_bytes = (g_bytes + 255) & ~(size_t)255;
  if (stem > g_bytes) g_bytes = stem;
  // the last bias_scratch_bytes(d) bytes are the partial column sums of the bias gradient (their own region: the wgrad
  // epilogue may still be reading the partial sums in front of it on another stream)
  return (((g_bytes > tc ? g_bytes : tc) + 255) & ~(size_t)255) + bias_scratch_bytes(d);
}

size_t cpgb_staged_weight_bytes(const cpgb_conv_desc *d) {
  if (!d || validate_desc(d)) return 0;
  if (g_path.load() == CPGB_PATH_SIMT) return 0;
  if (g_path.load() == CPGB_PATH_AUTO && stem_eligible(*d)) return 0;   // stem kernels mask while loading W
  if (!tc_eligible(*d, 0) && !tc_eligible(*d, 1)) return 0;
  return tc_staged_bytes(*d);
}

// The 3-channel 3x3 stem has its own direct fp32 kernels (stem_conv.cu) on the default path; the two
// forced paths (CPGB_PATH_SIMT / CPGB_PATH_TCGEN05) keep their meaning for tests and cross-checks.
static bool pick_stem(const cpgb_conv_desc *d, const void *y_or_dy) {
  return g_path.load() == CPGB_PATH_AUTO && stem_eligible(*d) && (reinterpret_cast<uintptr_t>(y_or_dy) & 15) == 0;
}

static int pick_tc(const cpgb_conv_desc *d, int op, bool *use_tc) {
  int path = g_path.load();
  bool ok = path != CPGB_PATH_SIMT && tc_eligible(*d, op);
  if (path == CPGB_PATH_TCGEN05 && !ok) {
    set_error("tcgen05 path forced but shape not eligible (op %d)", op);
    return CPGB_ENOTELIGIBLE;
  }
  *use_tc = ok;
  return CPGB_OK;
}

int cpgb_uses_tensor_cores(const cpgb_conv_desc *d, int32_t op) {
  if (!d || op < 0 || op > 2 || validate_desc(d) || d->N == 0) return 0;
  if (g_path.load() == CPGB_PATH_SIMT) return 0;
  if (g_path.load() == CPGB_PATH_AUTO && stem_eligible(*d) && op != 1) return 0;   // direct fp32 stem kernels
  return tc_eligible(*d, op) ? 1 : 0;
}

int cpgb_intile_eligible(const cpgb_conv_desc *d) {
  if (!d || validate_desc(d) || d->N == 0 || g_path.load() == CPGB_PATH_SIMT) return 0;
  return tc_eligible(*d, 0) && tc_eligible(*d, 1) && tc_intile_eligible(*d) ? 1 : 0;
}

int cpgb_intile_weight_shape(int32_t K, int32_t C, int32_t R, int32_t S, int32_t stride_h, int32_t stride_w,
                             int32_t groups) {
  if (g_path.load() == CPGB_PATH_SIMT) return 0;
  return tc_intile_weight_shape(K, C, R, S, stride_h, stride_w, groups) ? 1 : 0;
}

int cpgb_weights_usable_raw(const cpgb_conv_desc *d, int32_t has_piggymask) {
  if (!d || validate_desc(d) || has_piggymask || g_path.load() == CPGB_PATH_SIMT) return 0;
  return (tc_eligible(*d, 0) || tc_eligible(*d, 1)) && tc_weights_usable_raw(*d) ? 1 : 0;
}

int cpgb_weights_usable_raw_for(int32_t K, int32_t C, int32_t R, int32_t S, int32_t stride_h, int32_t stride_w,
                                int32_t groups, int32_t has_piggymask) {
  (void)K; (void)C; (void)R; (void)S; (void)stride_h; (void)stride_w; (void)groups; (void)has_piggymask;
  return 0;   // operands are always rounded to nearest TF32 now (see tc_weights_usable_raw)
}

size_t cpgb_staged_weight_bytes_for(int32_t K, int32_t C, int32_t R, int32_t S, int32_t stride_h, int32_t stride_w,
                                    int32_t groups) {
  if (g_path.load() == CPGB_PATH_SIMT) return 0;
  if (g_path.load() == CPGB_PATH_AUTO && stem_weight_shape(K, C, R, S, groups)) return 0;
  return tc_staged_bytes_for_weight(K, C, R, S, stride_h, stride_w, groups);
}

int cpgb_stage_weights_batched(int32_t n, const float *const *w, const float *const *piggy, void *const *staged,
                               const int32_t *K, const int32_t *C, const int32_t *R, const int32_t *S,
                               const int32_t *stride_h, const int32_t *stride_w, const float *thr, void *stream) {
  if (n < 0 || (n > 0 && (!w || !piggy || !staged || !K || !C || !R || !S || !stride_h || !stride_w || !thr))) {
    set_error("cpgb_stage_weights_batched: null pointer");
    return CPGB_EINVAL;
  }
  if (n == 0) return CPGB_OK;
  return tc_stage_weights_batched(n, w, piggy, staged, K, C, R, S, stride_h, stride_w, thr, (cudaStream_t)stream);
}

// bring-up hook (not part of the public header): MN-major operand descriptor fields
void cpgb_debug_set_mn(int layout, int lbo, int sbo, int kadv, int tma_swizzle) {
  debug_set_mn(layout, lbo, sbo, kadv, tma_swizzle);
}

int cpgb_stage_weights(const cpgb_conv_desc *d, const float *w, const float *piggy, float thr, void *staged,
                       size_t staged_bytes, void *stream) {
  int rc = validate_desc(d);
  if (rc) return rc;
  if (!w || !staged) { set_error("cpgb_stage_weights: null pointer"); return CPGB_EINVAL; }
  return tc_stage_weights(*d, w, piggy, thr, staged, staged_bytes, (cudaStream_t)stream);
}

// Operands of a tensor-core call out of (staged, ws): the staged weights are the caller's or are
// built at the front of ws; the split-K scratch is what remains of ws.
static int tc_operands(const cpgb_conv_desc *d, const float *w, const float *piggy, float thr, const void *staged,
                       void *ws, size_t ws_bytes, cudaStream_t st, const float **wt, void **part, size_t *part_bytes,
                       bool *raw) {
  char *base = reinterpret_cast<char *>(ws);
  size_t off = 0;
  *raw = false;
  if ((!staged || staged == w) && !piggy && tc_weights_usable_raw(*d) &&
      (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    *wt = w;              // linear / 1x1 layer without a piggymask: the weight tensor is the operand
    *raw = true;
  } else if (staged && staged != w) {
    *wt = reinterpret_cast<const float *>(staged);
  } else {
    const size_t sb = tc_staged_bytes(*d);
    if (!ws || ws_bytes < sb) { set_error("workspace %zu < %zu (staged weights)", ws_bytes, sb); return CPGB_EWORKSPACE; }
    int rc = tc_stage_weights(*d, w, piggy, thr, ws, ws_bytes, st);
    if (rc) return rc;
    *wt = reinterpret_cast<const float *>(ws);
    off = sb;
  }
  *part = ws && ws_bytes > off ? base + off : nullptr;
  *part_bytes = ws && ws_bytes > off ? ws_bytes - off : 0;
  return CPGB_OK;
}

int32_t cpgb_fprop_colstats_parts(const cpgb_conv_desc *d) {
  if (!d || validate_desc(d) || d->N == 0 || g_path.load() == CPGB_PATH_SIMT) return 0;
  if ((d->flags & CPGB_FLAG_W_INTILE) || pick_stem(d, nullptr)) return 0;
  return tc_fprop_colstats_parts(*d);
}

int cpgb_conv2d_fprop(const cpgb_conv_desc *d, const float *x, const float *w, const float *piggy,
                      const float *bias, float *y, float thr, const void *staged, void *ws, size_t ws_bytes,
                      void *stream) {
  return cpgb_conv2d_fprop_stats(d, x, w, piggy, bias, y, thr, staged, ws, ws_bytes, nullptr, stream);
}

int cpgb_conv2d_fprop_stats(const cpgb_conv_desc *d, const float *x, const float *w, const float *piggy,
                            const float *bias, float *y, float thr, const void *staged, void *ws, size_t ws_bytes,
                            float *colstats, void *stream) {
  int rc = validate_desc(d);
  if (rc) return rc;
  if (d->N == 0) return CPGB_OK;  // empty batch: nothing to compute, y is empty
  if (!x || !w || !y) { set_error("cpgb_conv2d_fprop: null pointer"); return CPGB_EINVAL; }
  if (colstats && ((reinterpret_cast<uintptr_t>(colstats) & 15) || cpgb_fprop_colstats_parts(d) == 0)) {
    set_error("cpgb_conv2d_fprop_stats: this layer cannot produce column statistics (cpgb_fprop_colstats_parts == 0)");
    return CPGB_EINVAL;
  }
  if (pick_stem(d, y) && (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0))
    return stem_fprop(*d, x, w, piggy, bias, y, thr, (cudaStream_t)stream);
  bool use_tc;
  if ((rc = pick_tc(d, 0, &use_tc))) return rc;
  if (use_tc && (d->flags & CPGB_FLAG_W_INTILE)) {
    if (!cpgb_intile_eligible(d) || (piggy && !staged)) {
      set_error("CPGB_FLAG_W_INTILE: layer not eligible (cpgb_intile_eligible) or packed mask missing"); return CPGB_EINVAL;
    }
    return tc_fprop_intile(*d, x, w, piggy ? staged : nullptr, bias, y, ws, ws_bytes, (cudaStream_t)stream);
  }
  if (use_tc) {
    const float *wt; void *part; size_t part_bytes; bool raw;
    if ((rc = tc_operands(d, w, piggy, thr, staged, ws, ws_bytes, (cudaStream_t)stream, &wt, &part, &part_bytes, &raw)))
      return rc;
    return tc_fprop(*d, x, wt, bias, y, part, part_bytes, (cudaStream_t)stream, raw, colstats);
  }
  return simt_fprop(make_geom(*d), x, w, piggy, bias, y, thr, (cudaStream_t)stream);
}

int cpgb_conv2d_dgrad(const cpgb_conv_desc *d, const float *dy, const float *w, const float *piggy, float *dx,
                      float thr, const void *staged, void *ws, size_t ws_bytes, void *stream) {
  int rc = validate_desc(d);
  if (rc) return rc;
  if (d->N == 0) return CPGB_OK;
  if (!dy || !w || !dx) { set_error("cpgb_conv2d_dgrad: null pointer"); return CPGB_EINVAL; }
  bool use_tc;
  if ((rc = pick_tc(d, 1, &use_tc))) return rc;
  if (use_tc && (d->flags & CPGB_FLAG_W_INTILE)) {
    if (!cpgb_intile_eligible(d) || (piggy && !staged)) {
      set_error("CPGB_FLAG_W_INTILE: layer not eligible (cpgb_intile_eligible) or packed mask missing"); return CPGB_EINVAL;
    }
    return tc_dgrad_intile(*d, dy, w, piggy ? staged : nullptr, dx, ws, ws_bytes, (cudaStream_t)stream);
  }
  if (use_tc) {
    const float *wt; void *part; size_t part_bytes; bool raw;
    if ((rc = tc_operands(d, w, piggy, thr, staged, ws, ws_bytes, (cudaStream_t)stream, &wt, &part, &part_bytes, &raw)))
      return rc;
    return tc_dgrad(*d, dy, wt, dx, part, part_bytes, (cudaStream_t)stream, raw);
  }
  return simt_dgrad(make_geom(*d), dy, w, piggy, dx, thr, (cudaStream_t)stream);
}

int cpgb_conv2d_bias_grad(const cpgb_conv_desc *d, const float *dy, float *dbias, void *stream) {
  return cpgb_conv2d_bias_grad_ws(d, dy, dbias, nullptr, 0, stream);
}

int cpgb_conv2d_bias_grad_ws(const cpgb_conv_desc *d, const float *dy, float *dbias, void *ws, size_t ws_bytes,
                             void *stream) {
  int rc = validate_desc(d);
  if (rc) return rc;
  if (!dbias || (d->N != 0 && !dy)) { set_error("cpgb_conv2d_bias_grad: null pointer"); return CPGB_EINVAL; }
  if (d->N == 0) { CPGB_CUDA_OK(cudaMemsetAsync(dbias, 0, d->K * sizeof(float), (cudaStream_t)stream)); return CPGB_OK; }
  // ws is a cpgb_workspace_bytes(d) buffer: the bias partial sums live in its last bias_scratch_bytes(d) bytes
  const size_t tail = bias_scratch_bytes(d), total = cpgb_workspace_bytes(d);
  void *scratch = (ws && tail && ws_bytes >= total) ? reinterpret_cast<char *>(ws) + (total - tail) : nullptr;
  return bias_grad(make_geom(*d), dy, dbias, scratch, scratch ? tail : 0, (cudaStream_t)stream);
}

int cpgb_conv2d_wgrad_fused(const cpgb_conv_desc *d, const float *x, const float *dy, const float *w,
                            const float *piggy, const uint8_t *tmask, int32_t cur, float weight_decay,
                            int32_t mode, float *dW, float *dP, float *dbias, float thr, void *ws,
                            size_t ws_bytes, void *stream) {
  return cpgb_conv2d_wgrad_fused_async(d, x, dy, w, piggy, tmask, cur, weight_decay, mode, dW, dP, dbias, thr, ws,
                                       ws_bytes, stream, stream);
}

int cpgb_conv2d_wgrad_fused_async(const cpgb_conv_desc *d, const float *x, const float *dy, const float *w,
                                  const float *piggy, const uint8_t *tmask, int32_t cur, float weight_decay,
                                  int32_t mode, float *dW, float *dP, float *dbias, float thr, void *ws,
                                  size_t ws_bytes, void *stream, void *epilogue_stream) {
  int rc = validate_desc(d);
  if (rc) return rc;
  if ((d->N != 0 && (!x || !dy)) || !w || !dW) { set_error("cpgb_conv2d_wgrad_fused: null pointer"); return CPGB_EINVAL; }
  const int base_mode = mode & 3;
  const bool merged = (mode & CPGB_GRAD_MERGED) != 0;
  if (mode < 0 || (mode & ~7) || base_mode > CPGB_GRAD_PRUNE || (merged && base_mode == CPGB_GRAD_RAW)) {
    set_error("bad grad mode %d", mode); return CPGB_EINVAL;
  }
  if (base_mode != CPGB_GRAD_RAW && !tmask) { set_error("fused grad modes need the task mask"); return CPGB_EINVAL; }
  if (merged ? dP != nullptr : (piggy == nullptr) != (dP == nullptr)) {
    set_error(merged ? "merged mode writes dW + dP to dW: pass dP = NULL" : "dP must be given iff piggy is");
    return CPGB_EINVAL;
  }
  const size_t n = weight_elems(d);
  if (!ws || ws_bytes < cpgb_workspace_bytes(d)) {
    set_error("workspace %zu < %zu", ws_bytes, cpgb_workspace_bytes(d));
    return CPGB_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  Geom g = make_geom(*d);
  float *gbuf = reinterpret_cast<float *>(ws);
  bool use_tc = false;
  const bool use_stem = d->N != 0 && pick_stem(d, dy);
  if (d->N != 0 && !use_stem && (rc = pick_tc(d, 2, &use_tc))) return rc;
  if (use_stem) {
    if ((rc = stem_wgrad_fused(*d, x, dy, w, piggy, tmask, cur, weight_decay, mode, thr, dW, dP, ws, ws_bytes, st)))
      return rc;
  } else if (use_tc) {
    // tensor-core wgrad: split-K partial sums in ws, summed inside the fused epilogue
    if ((rc = tc_wgrad_fused(*d, x, dy, w, piggy, tmask, cur, weight_decay, mode, thr, dW, dP, ws, ws_bytes, st,
                             (cudaStream_t)epilogue_stream)))
      return rc;
  } else {
    int splits = 1;
    if (d->N == 0) {
      CPGB_CUDA_OK(cudaMemsetAsync(gbuf, 0, n * sizeof(float), st));
    } else if ((rc = simt_wgrad_raw(g, x, dy, gbuf, &splits, st))) {
      return rc;
    }
    if ((rc = wgrad_epilogue(gbuf, splits, w, piggy, tmask, (long long)n, cur, weight_decay, mode, thr, dW, dP, st)))
      return rc;
  }
  if (dbias) {
    if (d->N == 0) CPGB_CUDA_OK(cudaMemsetAsync(dbias, 0, d->K * sizeof(float), st));
    else {
      const size_t tail = bias_scratch_bytes(d), total = cpgb_workspace_bytes(d);
      void *scratch = tail ? reinterpret_cast<char *>(ws) + (total - tail) : nullptr;      // ws_bytes >= total was checked
      if ((rc = bias_grad(g, dy, dbias, scratch, tail, st))) return rc;
    }
  }
  return CPGB_OK;
}

}  // extern "C"
