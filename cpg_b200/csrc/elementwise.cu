// HBM-bound element kernels of the path: Binarizer (a1), the fused wgrad epilogue (K6-K8),
// the standalone grad mask (a6), apply_mask / make_finetuning_mask (a9/a10), mask statistics
// (K12) and the data-parallel merge/split helpers.  All are 16-byte vectorised grid-stride
// loops sized to a multiple of the SM count.
#include "common.cuh"

namespace cpgb {

static int g_num_sms = 0;
static inline int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    g_num_sms = n > 0 ? n : 148;
  }
  return g_num_sms;
}
static inline int grid_for(long long nvec, int threads = 256, int ctas_per_sm = 8) {
  long long want = (nvec + threads - 1) / threads;
  long long cap = (long long)num_sms() * ctas_per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}
static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------- a1 Binarizer.forward (models/layers.py:15-19) ----------------
__global__ void __launch_bounds__(256) binarize_kernel(const float *__restrict__ p, float *__restrict__ out,
                                                       long long n, float thr, bool vec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long v = i; v < n4; v += stride) {
      float4 a = __ldg(reinterpret_cast<const float4 *>(p) + v);
      float4 b = make_float4(binarize_val(a.x, thr), binarize_val(a.y, thr), binarize_val(a.z, thr),
                             binarize_val(a.w, thr));
      reinterpret_cast<float4 *>(out)[v] = b;
    }
    for (long long t = (n4 << 2) + i; t < n; t += stride) out[t] = binarize_val(p[t], thr);
  } else {
    for (; i < n; i += stride) out[i] = binarize_val(p[i], thr);
  }
}

// ---------------- round-to-nearest TF32 (operand preparation of the tcgen05 kernels) ----------------
__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}
__global__ void __launch_bounds__(256) round_tf32_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                         long long n, bool vec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long v = i; v < n4; v += stride) {
      const float4 a = __ldg(reinterpret_cast<const float4 *>(in) + v);
      reinterpret_cast<float4 *>(out)[v] = make_float4(rna_tf32(a.x), rna_tf32(a.y), rna_tf32(a.z), rna_tf32(a.w));
    }
    for (long long t = (n4 << 2) + i; t < n; t += stride) out[t] = rna_tf32(in[t]);
  } else {
    for (; i < n; i += stride) out[i] = rna_tf32(in[i]);
  }
}
int round_tf32(const float *in, float *out, long long n, cudaStream_t st) {
  if (n <= 0) return CPGB_OK;
  round_tf32_kernel<<<grid_for(n / 4 + 1), 256, 0, st>>>(in, out, n, aligned16(in) && aligned16(out));
  CPGB_LAUNCH_OK("round_tf32");
  return CPGB_OK;
}

// ---------------- packed {piggyback, task} bit masks (SURVEY 8b cpgb_pack_mask) ----------------
// One 64-bit word per 32 consecutive elements: low half bit i = (piggy[32g + i] > thr) -- the Binarizer of
// models/layers.py:15-19 -- high half bit i = (1 <= T[32g + i] <= inference_idx) -- the weights apply_mask keeps
// (utils/prune.py:229-230).  A warp handles 32 elements per step: coalesced 128-byte reads, one ballot each.
__global__ void __launch_bounds__(256)
pack_mask_kernel(const float *__restrict__ piggy, const uint8_t *__restrict__ tmask, long long n, float thr, int inf_idx,
                 unsigned long long *__restrict__ out, bool vec) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (vec) {
    // 128 elements per warp step: lane l holds elements 4l .. 4l+3 (one 16-byte load of the piggymask, one 4-byte
    // load of the task mask); its four bits move to position 4*(l % 8) of the word of its 8-lane group and the
    // group ORs them together with three shuffles.  Two steps in flight per warp.
    const long long groups128 = n >> 7;
    for (long long g = warp0; g < groups128; g += 2 * warps) {
      const long long g2 = g + warps;
      const bool two = g2 < groups128;
      float4 p0 = make_float4(1.f, 1.f, 1.f, 1.f), p1 = p0;
      uchar4 t0 = make_uchar4(1, 1, 1, 1), t1 = t0;
      if (piggy) p0 = __ldg(reinterpret_cast<const float4 *>(piggy) + (g << 5) + lane);
      if (tmask) t0 = __ldg(reinterpret_cast<const uchar4 *>(tmask) + (g << 5) + lane);
      if (two && piggy) p1 = __ldg(reinterpret_cast<const float4 *>(piggy) + (g2 << 5) + lane);
      if (two && tmask) t1 = __ldg(reinterpret_cast<const uchar4 *>(tmask) + (g2 << 5) + lane);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 pv = h ? p1 : p0;
        const uchar4 tv = h ? t1 : t0;
        const unsigned u = (unsigned)inf_idx;
        unsigned lo = piggy ? ((pv.x > thr) | ((pv.y > thr) << 1) | ((pv.z > thr) << 2) | ((pv.w > thr) << 3)) : 0xFu;
        unsigned hi = tmask ? ((tv.x != 0 && tv.x <= u) | ((tv.y != 0 && tv.y <= u) << 1) | ((tv.z != 0 && tv.z <= u) << 2) |
                               ((tv.w != 0 && tv.w <= u) << 3)) : 0xFu;
        lo <<= 4 * (lane & 7); hi <<= 4 * (lane & 7);
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          lo |= __shfl_xor_sync(0xffffffffu, lo, o);
          hi |= __shfl_xor_sync(0xffffffffu, hi, o);
        }
        if ((lane & 7) == 0 && (h == 0 || two))
          out[((h ? g2 : g) << 2) + (lane >> 3)] = ((unsigned long long)hi << 32) | lo;
      }
    }
  }
  // tail (or everything when the pointers are not 16 / 4-byte aligned): 32 elements per warp step, one ballot each
  const long long first = vec ? ((n >> 7) << 2) : 0;
  const long long groups = (n + 31) >> 5;
  for (long long g = first + warp0; g < groups; g += warps) {
    const long long i = (g << 5) + lane;
    bool pb = false, tb = false;
    if (i < n) {
      pb = piggy ? (__ldg(piggy + i) > thr) : true;
      const unsigned t = tmask ? (unsigned)__ldg(tmask + i) : 1u;
      tb = tmask ? (t != 0u && t <= (unsigned)inf_idx) : true;
    }
    const unsigned lo = __ballot_sync(0xffffffffu, pb), hi = __ballot_sync(0xffffffffu, tb);
    if (lane == 0) out[g] = ((unsigned long long)hi << 32) | lo;
  }
}

// ---------------- fused wgrad epilogue ----------------
// g: raw weight gradient dL/dW_eff.  One pass produces what optimizers.step() must see:
//   RAW      : dW = g*b                     dP = g*W                      (models/layers.py:21-23,103)
//   FINETUNE : dW = (g*b + wd*W)[T==cur]    dP = (g*W)[1<=T<cur]          (utils/prune.py:203-208)
//   PRUNE    : dW = (g*b + wd*W)[T==cur]    dP = 0                        (utils/prune.py:203-205,210)
__device__ __forceinline__ void epi_one(float g, float w, float p, bool has_p, unsigned t, int cur, float wd,
                                        int mode, float thr, float &dw, float &dp) {
  grad_epilogue_elem(g, w, p, has_p, t, cur, wd, mode, thr, dw, dp);
}

__global__ void __launch_bounds__(256)
wgrad_epilogue_kernel(const float *__restrict__ g, int splits, const float *__restrict__ w,
                      const float *__restrict__ piggy, const uint8_t *__restrict__ tmask, long long n, int cur, float wd,
                      int mode, float thr, float *__restrict__ dW, float *__restrict__ dP, bool vec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool has_p = piggy != nullptr;
  long long tail = 0;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long v = i0; v < n4; v += stride) {
      float4 gg = __ldg(reinterpret_cast<const float4 *>(g) + v);
      for (int sp = 1; sp < splits; ++sp) {     // fixed order: deterministic sums
        const float4 b = __ldg(reinterpret_cast<const float4 *>(g + sp * n) + v);
        gg.x += b.x; gg.y += b.y; gg.z += b.z; gg.w += b.w;
      }
      float4 ww = __ldg(reinterpret_cast<const float4 *>(w) + v);
      float4 pp = has_p ? __ldg(reinterpret_cast<const float4 *>(piggy) + v) : make_float4(0, 0, 0, 0);
      uchar4 tt = tmask ? __ldg(reinterpret_cast<const uchar4 *>(tmask) + v) : make_uchar4(0, 0, 0, 0);
      float4 ow, op;
      epi_one(gg.x, ww.x, pp.x, has_p, tt.x, cur, wd, mode, thr, ow.x, op.x);
      epi_one(gg.y, ww.y, pp.y, has_p, tt.y, cur, wd, mode, thr, ow.y, op.y);
      epi_one(gg.z, ww.z, pp.z, has_p, tt.z, cur, wd, mode, thr, ow.z, op.z);
      epi_one(gg.w, ww.w, pp.w, has_p, tt.w, cur, wd, mode, thr, ow.w, op.w);
      reinterpret_cast<float4 *>(dW)[v] = ow;
      if (dP) reinterpret_cast<float4 *>(dP)[v] = op;
    }
    tail = n4 << 2;
  }
  for (long long i = tail + i0; i < n; i += stride) {
    float ow, op, gs = g[i];
    for (int sp = 1; sp < splits; ++sp) gs += g[sp * n + i];
    epi_one(gs, w[i], has_p ? piggy[i] : 0.f, has_p, tmask ? tmask[i] : 0u, cur, wd, mode, thr, ow, op);
    dW[i] = ow;
    if (dP) dP[i] = op;
  }
}

int wgrad_epilogue(const float *gbuf, int splits, const float *w, const float *piggy, const uint8_t *tmask,
                   long long n, int cur, float wd, int mode, float thr, float *dW, float *dP, cudaStream_t st) {
  bool vec = aligned16(gbuf) && aligned16(w) && aligned16(dW) && (!piggy || aligned16(piggy)) &&
             (!dP || aligned16(dP)) && (!tmask || (reinterpret_cast<uintptr_t>(tmask) & 3) == 0) &&
             (splits == 1 || n % 4 == 0);
  wgrad_epilogue_kernel<<<grid_for(n / 4 + 1), 256, 0, st>>>(gbuf, splits, w, piggy, tmask, n, cur, wd, mode, thr, dW,
                                                            dP, vec);
  CPGB_LAUNCH_OK("wgrad_epilogue");
  return CPGB_OK;
}

// ---------------- a6 standalone, in place (utils/prune.py:195-211) ----------------
__global__ void __launch_bounds__(256)
grad_epilogue_kernel(float *__restrict__ dW, float *__restrict__ dP, const float *__restrict__ w,
                     const uint8_t *__restrict__ tmask, long long n, int cur, float wd, int mode) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned t = tmask[i];
    if (dW) dW[i] = (t == (unsigned)cur) ? fmaf(wd, w[i], dW[i]) : 0.f;
    if (dP) {
      bool keep = (mode == CPGB_GRAD_FINETUNE) && t != 0u && t < (unsigned)cur;
      if (!keep) dP[i] = 0.f;
    }
  }
}

// ---------------- a9 apply_mask / make_pruned_zero (utils/prune.py:213-231) ----------------
__global__ void __launch_bounds__(256)
apply_mask_kernel(float *__restrict__ w, const uint8_t *__restrict__ tmask, long long n, int inference_idx) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned t = tmask[i];
    if (t == 0u || t > (unsigned)inference_idx) w[i] = 0.f;
  }
}

// ---------------- a10 make_finetuning_mask (utils/prune.py:233-243) ----------------
__global__ void __launch_bounds__(256)
finetuning_mask_kernel(uint8_t *__restrict__ tmask, long long n, int new_cur) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (tmask[i] == 0) tmask[i] = (uint8_t)new_cur;
}

// ---------------- K12 statistics (utils/prune.py:111-193) ----------------
__global__ void __launch_bounds__(256)
mask_stats_kernel(const uint8_t *__restrict__ tmask, const float *__restrict__ piggy, long long n, int idx,
                  unsigned long long *__restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned t = tmask[i];
    c0 += (t == 0u);
    c1 += (t == (unsigned)idx);
    bool shared = t > 0u && t < (unsigned)idx;
    c2 += shared;
    if (piggy && shared) c3 += (piggy[i] > 0.005f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, o);
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    c3 += __shfl_xor_sync(0xffffffffu, c3, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c0) atomicAdd(out + 0, (unsigned long long)c0);
    if (c1) atomicAdd(out + 1, (unsigned long long)c1);
    if (c2) atomicAdd(out + 2, (unsigned long long)c2);
    if (c3) atomicAdd(out + 3, (unsigned long long)c3);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(out + 4, (unsigned long long)n);
}

// all layers in one launch (grid.y = layer): what Manager.train asks for after every batch
// (utils/manager.py:77-88 -> calculate_sparsity / calculate_curr_task_ratio / ...)
constexpr int STATS_MAX_LAYERS = 64;
struct StatsBatch {
  const uint8_t *t[STATS_MAX_LAYERS];
  const float *p[STATS_MAX_LAYERS];
  long long n[STATS_MAX_LAYERS];
};
__global__ void __launch_bounds__(256)
mask_stats_batched_kernel(const __grid_constant__ StatsBatch sb, int idx, unsigned long long *__restrict__ out) {
  const int layer = blockIdx.y;
  const uint8_t *__restrict__ tmask = sb.t[layer];
  const float *__restrict__ piggy = sb.p[layer];
  const long long n = sb.n[layer];
  const long long n4 = n >> 2;
  long long nblk = (n4 + blockDim.x - 1) / blockDim.x;
  if (nblk < 1) nblk = 1;
  const long long gx = nblk < (long long)gridDim.x ? nblk : (long long)gridDim.x;
  if ((long long)blockIdx.x >= gx) return;
  const long long stride = gx * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  auto visit = [&](unsigned t, long long i) {
    c0 += (t == 0u);
    c1 += (t == (unsigned)idx);
    const bool shared = t > 0u && t < (unsigned)idx;
    c2 += shared;
    if (piggy && shared) c3 += (__ldg(piggy + i) > 0.005f);
  };
  long long tail = 0;
  if ((reinterpret_cast<uintptr_t>(tmask) & 3) == 0) {
    for (long long v = i0; v < n4; v += stride) {
      const uchar4 t = __ldg(reinterpret_cast<const uchar4 *>(tmask) + v);
      visit(t.x, 4 * v); visit(t.y, 4 * v + 1); visit(t.z, 4 * v + 2); visit(t.w, 4 * v + 3);
    }
    tail = n4 << 2;
  }
  for (long long i = tail + i0; i < n; i += stride) visit(tmask[i], i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, o);
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    c3 += __shfl_xor_sync(0xffffffffu, c3, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c0) atomicAdd(out + 0, (unsigned long long)c0);
    if (c1) atomicAdd(out + 1, (unsigned long long)c1);
    if (c2) atomicAdd(out + 2, (unsigned long long)c2);
    if (c3) atomicAdd(out + 3, (unsigned long long)c3);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(out + 4, (unsigned long long)n);
}

// ---------------- data-parallel helpers ----------------
__global__ void __launch_bounds__(256)
merge_grads_kernel(const float *__restrict__ dW, const float *__restrict__ dP, float *__restrict__ m, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    m[i] = dW[i] + (dP ? dP[i] : 0.f);
}
__global__ void __launch_bounds__(256)
split_grads_kernel(const float *m, const uint8_t *__restrict__ tmask, long long n, int cur, float *dW,
                   float *__restrict__ dP, bool vec) {
  // dW may alias m (the data-parallel reducer splits in place inside its gradient bucket): every thread reads
  // its own elements before it writes them
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long tail = 0;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long v = i0; v < n4; v += stride) {
      const float4 a = reinterpret_cast<const float4 *>(m)[v];
      const uchar4 t = __ldg(reinterpret_cast<const uchar4 *>(tmask) + v);
      const unsigned c = (unsigned)cur;
      reinterpret_cast<float4 *>(dW)[v] = make_float4(t.x == c ? a.x : 0.f, t.y == c ? a.y : 0.f, t.z == c ? a.z : 0.f,
                                                      t.w == c ? a.w : 0.f);
      if (dP)
        reinterpret_cast<float4 *>(dP)[v] = make_float4((t.x != 0u && t.x < c) ? a.x : 0.f, (t.y != 0u && t.y < c) ? a.y : 0.f,
                                                        (t.z != 0u && t.z < c) ? a.z : 0.f, (t.w != 0u && t.w < c) ? a.w : 0.f);
    }
    tail = n4 << 2;
  }
  for (long long i = tail + i0; i < n; i += stride) {
    unsigned t = tmask[i];
    float v = m[i];
    dW[i] = (t == (unsigned)cur) ? v : 0.f;
    if (dP) dP[i] = (t != 0u && t < (unsigned)cur) ? v : 0.f;
  }
}

}  // namespace cpgb

using namespace cpgb;

extern "C" {

int cpgb_round_tf32(const float *in, float *out, int64_t n, void *stream) {
  if (n < 0 || (n > 0 && (!in || !out))) { set_error("cpgb_round_tf32: bad arguments"); return CPGB_EINVAL; }
  return round_tf32(in, out, (long long)n, (cudaStream_t)stream);
}

int cpgb_pack_mask(const float *piggy, const uint8_t *tmask, int64_t n, float thr, int32_t inference_idx, uint64_t *packed,
                   void *stream) {
  if (n < 0 || (n > 0 && !packed)) { set_error("cpgb_pack_mask: bad arguments"); return CPGB_EINVAL; }
  if (n == 0) return CPGB_OK;
  const bool vec = (!piggy || aligned16(piggy)) && (!tmask || (reinterpret_cast<uintptr_t>(tmask) & 3) == 0);
  pack_mask_kernel<<<grid_for(n / 8 + 1), 256, 0, (cudaStream_t)stream>>>(
      piggy, tmask, (long long)n, thr, inference_idx, reinterpret_cast<unsigned long long *>(packed), vec);
  CPGB_LAUNCH_OK("pack_mask");
  return CPGB_OK;
}

int cpgb_binarize(const float *piggy, float *out, int64_t n, float thr, void *stream) {
  if (n < 0 || (n > 0 && (!piggy || !out))) { set_error("cpgb_binarize: null pointer"); return CPGB_EINVAL; }
  if (n == 0) return CPGB_OK;
  bool vec = aligned16(piggy) && aligned16(out);
  binarize_kernel<<<grid_for(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(piggy, out, n, thr, vec);
  CPGB_LAUNCH_OK("cpgb_binarize");
  return CPGB_OK;
}

int cpgb_grad_epilogue(float *dW, float *dP, const float *w, const uint8_t *tmask, int64_t n, int32_t cur,
                       float weight_decay, int32_t mode, void *stream) {
  if (n < 0 || !tmask || (dW && !w)) { set_error("cpgb_grad_epilogue: null pointer"); return CPGB_EINVAL; }
  if (mode != CPGB_GRAD_FINETUNE && mode != CPGB_GRAD_PRUNE) {
    set_error("cpgb_grad_epilogue: mode must be FINETUNE or PRUNE");
    return CPGB_EINVAL;
  }
  if (n == 0 || (!dW && !dP)) return CPGB_OK;
  grad_epilogue_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(dW, dP, w, tmask, n, cur, weight_decay, mode);
  CPGB_LAUNCH_OK("cpgb_grad_epilogue");
  return CPGB_OK;
}

int cpgb_apply_mask(float *w, const uint8_t *tmask, int64_t n, int32_t inference_idx, void *stream) {
  if (n < 0 || (n > 0 && (!w || !tmask))) { set_error("cpgb_apply_mask: null pointer"); return CPGB_EINVAL; }
  if (n == 0) return CPGB_OK;
  apply_mask_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(w, tmask, n, inference_idx);
  CPGB_LAUNCH_OK("cpgb_apply_mask");
  return CPGB_OK;
}

int cpgb_make_finetuning_mask(uint8_t *tmask, int64_t n, int32_t new_cur, void *stream) {
  if (n < 0 || (n > 0 && !tmask) || new_cur < 1 || new_cur > 255) {
    set_error("cpgb_make_finetuning_mask: bad argument");
    return CPGB_EINVAL;
  }
  if (n == 0) return CPGB_OK;
  finetuning_mask_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(tmask, n, new_cur);
  CPGB_LAUNCH_OK("cpgb_make_finetuning_mask");
  return CPGB_OK;
}

int cpgb_mask_stats(const uint8_t *tmask, const float *piggy, int64_t n, int32_t inference_idx, int64_t *out,
                    void *stream) {
  if (n < 0 || !out || (n > 0 && !tmask)) { set_error("cpgb_mask_stats: null pointer"); return CPGB_EINVAL; }
  if (n == 0) return CPGB_OK;
  mask_stats_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(
      tmask, piggy, n, inference_idx, reinterpret_cast<unsigned long long *>(out));
  CPGB_LAUNCH_OK("cpgb_mask_stats");
  return CPGB_OK;
}

int cpgb_mask_stats_batched(int32_t nlayers, const uint8_t *const *tmask, const float *const *piggy, const int64_t *n,
                            int32_t inference_idx, int64_t *out, void *stream) {
  if (nlayers < 0 || !out || (nlayers > 0 && (!tmask || !n))) { set_error("cpgb_mask_stats_batched: null pointer"); return CPGB_EINVAL; }
  for (int base = 0; base < nlayers; base += STATS_MAX_LAYERS) {
    StatsBatch sb;
    const int cnt = nlayers - base < STATS_MAX_LAYERS ? nlayers - base : STATS_MAX_LAYERS;
    long long nmax = 0;
    for (int j = 0; j < cnt; ++j) {
      const int i = base + j;
      if (n[i] < 0 || (n[i] > 0 && !tmask[i])) { set_error("cpgb_mask_stats_batched: bad layer %d", i); return CPGB_EINVAL; }
      sb.t[j] = tmask[i]; sb.p[j] = piggy ? piggy[i] : nullptr; sb.n[j] = n[i];
      if (n[i] > nmax) nmax = n[i];
    }
    dim3 grid(grid_for(nmax / 4 + 1, 256, 4), cnt);
    mask_stats_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sb, inference_idx,
                                                                      reinterpret_cast<unsigned long long *>(out));
    CPGB_LAUNCH_OK("cpgb_mask_stats_batched");
  }
  return CPGB_OK;
}

int cpgb_merge_grads(const float *dW, const float *dP, float *merged, int64_t n, void *stream) {
  if (n < 0 || (n > 0 && (!dW || !merged))) { set_error("cpgb_merge_grads: null pointer"); return CPGB_EINVAL; }
  if (n == 0) return CPGB_OK;
  merge_grads_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(dW, dP, merged, n);
  CPGB_LAUNCH_OK("cpgb_merge_grads");
  return CPGB_OK;
}

int cpgb_split_merged_grad(const float *merged, const uint8_t *tmask, int64_t n, int32_t cur, float *dW, float *dP,
                           void *stream) {
  if (n < 0 || (n > 0 && (!merged || !tmask || !dW))) { set_error("cpgb_split_merged_grad: null pointer"); return CPGB_EINVAL; }
  if (n == 0) return CPGB_OK;
  const bool vec = aligned16(merged) && aligned16(dW) && (!dP || aligned16(dP)) && (reinterpret_cast<uintptr_t>(tmask) & 3) == 0;
  split_grads_kernel<<<grid_for(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(merged, tmask, n, cur, dW, dP, vec);
  CPGB_LAUNCH_OK("cpgb_split_merged_grad");
  return CPGB_OK;
}

}  // extern "C"
