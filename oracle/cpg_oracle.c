/* Plain-C restatement of the integer/byte arithmetic of the CPG hot path, plus a naive
 * double-accumulate convolution.  TEST INFRASTRUCTURE ONLY (see oracle/cpg_oracle.py for the
 * pinning story): second, independent implementation used to cross-check the numpy/torch
 * oracle against tests/golden/.  Citations are reference paths (ivclab/CPG).
 *
 * Build: make -C oracle   ->  oracle/_build/libcpg_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* models/layers.py:15-19 : b = 0 where p <= thr, 1 where p > thr, NaN kept */
void orc_binarize(const float *p, float *out, int64_t n, float thr) {
  for (int64_t i = 0; i < n; ++i) {
    float v = p[i];
    if (v <= thr) v = 0.0f;
    else if (v > thr) v = 1.0f;
    out[i] = v;
  }
}

/* utils/prune.py:195-211 : mode 1 = finetune, 2 = prune */
void orc_grad_epilogue(float *dW, float *dP, const float *w, const uint8_t *t, int64_t n, int cur,
                       float wd, int mode) {
  for (int64_t i = 0; i < n; ++i) {
    if (dW) {
      float v = dW[i] + wd * w[i];
      dW[i] = (t[i] != cur) ? 0.0f : v;
    }
    if (dP) {
      if (mode == 2) dP[i] = 0.0f;
      else if (t[i] == 0 || t[i] >= cur) dP[i] = 0.0f;
    }
  }
}

/* Total order with NaN last -- the order torch.kthvalue (utils/prune.py:39) and numpy's partition use; found by
 * tests/test_prune_differential_cpu.py: with the plain (x > y) - (x < y) a NaN weight broke the sort. */
static int cmp_float(const void *a, const void *b) {
  float x = *(const float *)a, y = *(const float *)b;
  int nx = x != x, ny = y != y;
  if (nx || ny) return nx - ny;
  return (x > y) - (x < y);
}

/* utils/prune.py:30-53.  Returns 0, or 2 for the sys.exit(2) path; *cut_out = cutoff value.
 * k = round-half-even(ratio * pool) (python round(), utils/prune.py:37). */
int orc_pruning_mask(const float *w, uint8_t *t, int64_t n, int cur, double ratio, float *cut_out,
                     int64_t *k_out, int64_t *pool_out) {
  float *pool = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i)
    if (t[i] == cur || t[i] == 0) pool[m++] = fabsf(w[i]);
  int64_t k = (int64_t)nearbyint(ratio * (double)m);   /* FE_TONEAREST = half-to-even */
  if (k_out) *k_out = k;
  if (pool_out) *pool_out = m;
  if (k < 1 || k > m) { free(pool); return 2; }
  qsort(pool, (size_t)m, sizeof(float), cmp_float);
  float cut = pool[k - 1];
  free(pool);
  if (cut_out) *cut_out = cut;
  for (int64_t i = 0; i < n; ++i)
    if (fabsf(w[i]) <= cut && t[i] == cur) t[i] = 0;
  return 0;
}

/* utils/prune.py:223-231 (inference_idx = 255 gives make_pruned_zero, :213-221) */
void orc_apply_mask(float *w, const uint8_t *t, int64_t n, int inference_idx) {
  for (int64_t i = 0; i < n; ++i)
    if (t[i] == 0 || t[i] > inference_idx) w[i] = 0.0f;
}

/* utils/prune.py:233-243 */
void orc_make_finetuning_mask(uint8_t *t, int64_t n, int new_cur) {
  for (int64_t i = 0; i < n; ++i)
    if (t[i] == 0) t[i] = (uint8_t)new_cur;
}

/* models/layers.py:98-109 with double accumulation; NCHW dense; p may be NULL. */
void orc_conv2d_fwd(const float *x, const float *w, const float *p, const float *bias, float *y, int N,
                    int C, int H, int W, int K, int R, int S, int stride, int pad, int dil, int groups,
                    float thr) {
  int Cg = C / groups, Kg = K / groups;
  int P = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
  int Q = (W + 2 * pad - dil * (S - 1) - 1) / stride + 1;
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      int g = k / Kg;
      for (int pp = 0; pp < P; ++pp)
        for (int q = 0; q < Q; ++q) {
          double acc = bias ? (double)bias[k] : 0.0;
          for (int c = 0; c < Cg; ++c)
            for (int r = 0; r < R; ++r)
              for (int s = 0; s < S; ++s) {
                int h = pp * stride - pad + r * dil, ww = q * stride - pad + s * dil;
                if (h < 0 || h >= H || ww < 0 || ww >= W) continue;
                int64_t wi = (((int64_t)k * Cg + c) * R + r) * S + s;
                float b = 1.0f;
                if (p) b = p[wi] > thr ? 1.0f : (p[wi] <= thr ? 0.0f : p[wi]);
                acc += (double)(b * w[wi]) * (double)x[(((int64_t)n * C + g * Cg + c) * H + h) * W + ww];
              }
          y[(((int64_t)n * K + k) * P + pp) * Q + q] = (float)acc;
        }
    }
}
