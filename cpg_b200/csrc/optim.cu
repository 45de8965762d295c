// the optimizer step that follows the masked-convolution path, one launch per optimizer.
//
//   cpgb_sgd_nesterov_step : torch.optim.SGD(momentum, nesterov=True, weight_decay=0, dampening=0) on the weights /
//                            batch-norm parameters / heads      (CPG_cifar100_main_normal.py:339-340)
//   cpgb_adam_step         : torch.optim.Adam(lr_mask) on the piggymasks   (CPG_cifar100_main_normal.py:344-345),
//                            optionally emitting the packed Binarizer bits of the UPDATED piggymask (cpgb_pack_mask's
//                            words), so that the next forward pass needs no pack pass for the in-tile masked layers
//
// Every element is updated, also where the (masked) gradient is exactly zero: momentum keeps moving pruned weights
// until the next apply_mask and Adam's first moment keeps decaying -- the behaviour of the reference,
// not an optimisation opportunity.  The arithmetic follows torch's multi-tensor ("foreach") implementation operation
// by operation, with the same intermediate roundings (one fp32 rounding per ATen kernel it launches):
//   SGD : buf = buf * mu ; buf = buf + g ; g' = fma(mu, buf, g) ; p = fma(-lr, g', p)
//   Adam: m = fma(1 - b1, g - m, m) ; v = v * b2 ; v = fma(1 - b2, g * g, v) ;
//         d = sqrt(v) / sqrt(1 - b2^t) + eps ; p = fma(-lr / (1 - b1^t), m / d, p)
// (tests/test_optim_gpu.py compares bit patterns with torch.optim on the same GPU).  HBM-bound: SGD reads p, g, buf
// and writes p, buf (20 B/element); Adam reads p, g, m, v and writes p, m, v (28 B/element).
#include "common.cuh"

namespace cpgb {

namespace {

constexpr int OPT_MAX_TENSORS = 40;       // per launch (the pointer table travels as a kernel parameter)
constexpr int OPT_THREADS = 256;
constexpr int OPT_CHUNK = OPT_THREADS * 16;   // elements per block: four float4 per thread

struct OptTable {
  float *p[OPT_MAX_TENSORS];
  const float *g[OPT_MAX_TENSORS];
  float *s1[OPT_MAX_TENSORS];              // momentum buffer / exp_avg
  float *s2[OPT_MAX_TENSORS];              // exp_avg_sq
  unsigned long long *packed[OPT_MAX_TENSORS];
  const uint8_t *tmask[OPT_MAX_TENSORS];
  long long n[OPT_MAX_TENSORS];
  int blk_start[OPT_MAX_TENSORS + 1];
  int count;
};

__device__ __forceinline__ int opt_tensor_of(const OptTable &tb, int b) {
  int lo = 0, hi = tb.count - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tb.blk_start[mid] <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ void sgd_elem(float &p, float g, float &buf, float mu, float neg_lr) {
  buf = __fadd_rn(__fmul_rn(buf, mu), g);          // _foreach_mul_(bufs, mu); _foreach_add_(bufs, grads, alpha=1)
  const float gn = __fmaf_rn(mu, buf, g);          // _foreach_add_(grads, bufs, alpha=mu)
  p = __fmaf_rn(neg_lr, gn, p);                    // _foreach_add_(params, grads, alpha=-lr)
}

__global__ void __launch_bounds__(OPT_THREADS)
sgd_nesterov_kernel(const __grid_constant__ OptTable tb, float mu, float lr, const float *__restrict__ lr_dev) {
  const int ti = opt_tensor_of(tb, blockIdx.x);
  float *__restrict__ p = tb.p[ti];
  const float *__restrict__ g = tb.g[ti];
  float *__restrict__ buf = tb.s1[ti];
  const long long n = tb.n[ti];
  const float neg_lr = -(lr_dev ? *lr_dev : lr);
  const long long base = (long long)(blockIdx.x - tb.blk_start[ti]) * OPT_CHUNK;
  const long long end = base + OPT_CHUNK < n ? base + OPT_CHUNK : n;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(buf)) & 15) == 0;
  if (vec && end - base == OPT_CHUNK) {
    float4 pv[4], gv[4], bv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i4 = (base >> 2) + u * OPT_THREADS + threadIdx.x;
      pv[u] = reinterpret_cast<const float4 *>(p)[i4];
      gv[u] = __ldg(reinterpret_cast<const float4 *>(g) + i4);
      bv[u] = reinterpret_cast<const float4 *>(buf)[i4];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i4 = (base >> 2) + u * OPT_THREADS + threadIdx.x;
      sgd_elem(pv[u].x, gv[u].x, bv[u].x, mu, neg_lr); sgd_elem(pv[u].y, gv[u].y, bv[u].y, mu, neg_lr);
      sgd_elem(pv[u].z, gv[u].z, bv[u].z, mu, neg_lr); sgd_elem(pv[u].w, gv[u].w, bv[u].w, mu, neg_lr);
      reinterpret_cast<float4 *>(p)[i4] = pv[u];
      reinterpret_cast<float4 *>(buf)[i4] = bv[u];
    }
    return;
  }
  for (long long i = base + threadIdx.x; i < end; i += OPT_THREADS) {
    float pe = p[i], be = buf[i];
    sgd_elem(pe, g[i], be, mu, neg_lr);
    p[i] = pe; buf[i] = be;
  }
}

struct AdamCoef { float w1, b2, w2, bc2_sqrt, eps, neg_step; };
// step counter: step_dev[0] = number of steps taken so far (int64), step_dev[1] = ticket of finished blocks
__device__ __forceinline__ AdamCoef adam_coef(double lr, double beta1, double beta2, double eps, long long step) {
  AdamCoef c;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  c.w1 = (float)(1.0 - beta1);
  c.b2 = (float)beta2;
  c.w2 = (float)(1.0 - beta2);
  c.bc2_sqrt = (float)sqrt(bc2);
  c.eps = (float)eps;
  c.neg_step = (float)((lr / bc1) * -1.0);
  return c;
}
__device__ __forceinline__ void adam_elem(float &p, float g, float &m, float &v, const AdamCoef &c) {
  m = __fmaf_rn(c.w1, __fsub_rn(g, m), m);                          // _foreach_lerp_(exp_avgs, grads, 1 - beta1)
  v = __fmul_rn(v, c.b2);                                           // _foreach_mul_(exp_avg_sqs, beta2)
  v = __fmaf_rn(c.w2, __fmul_rn(g, g), v);                          // _foreach_addcmul_(.., grads, grads, 1 - beta2)
  const float d = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), c.bc2_sqrt), c.eps);   // sqrt, div_(bc2_sqrt), add_(eps)
  p = __fmaf_rn(c.neg_step, __fdiv_rn(m, d), p);                    // _foreach_addcdiv_(params, m, d, -lr / bc1)
}

// one warp = 128 consecutive elements per step (lane l: elements 4l .. 4l+3), so that the packed words of
// cpgb_pack_mask (32 elements each) are assembled with three shuffles per 8-lane group
__global__ void __launch_bounds__(OPT_THREADS)
adam_kernel(const __grid_constant__ OptTable tb, double lr, double beta1, double beta2, double eps,
            long long *__restrict__ step_dev, const double *__restrict__ lr_dev, float thr, int inf_idx, int total_blocks) {
  const int ti = opt_tensor_of(tb, blockIdx.x);
  float *__restrict__ p = tb.p[ti];
  const float *__restrict__ g = tb.g[ti];
  float *__restrict__ m = tb.s1[ti];
  float *__restrict__ v = tb.s2[ti];
  unsigned long long *__restrict__ packed = tb.packed[ti];
  const uint8_t *__restrict__ tmask = tb.tmask[ti];
  const long long n = tb.n[ti];
  const long long step = step_dev[0] + 1;
  const AdamCoef c = adam_coef(lr_dev ? *lr_dev : lr, beta1, beta2, eps, step);
  const long long base = (long long)(blockIdx.x - tb.blk_start[ti]) * OPT_CHUNK;
  const long long end = base + OPT_CHUNK < n ? base + OPT_CHUNK : n;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (!tmask || (reinterpret_cast<uintptr_t>(tmask) & 3) == 0);
  const int lane = threadIdx.x & 31;
  if (vec && end - base == OPT_CHUNK) {
#pragma unroll 2
    for (int u = 0; u < 4; ++u) {
      const long long i4 = (base >> 2) + u * OPT_THREADS + threadIdx.x;
      float4 pv = reinterpret_cast<const float4 *>(p)[i4];
      const float4 gv = __ldg(reinterpret_cast<const float4 *>(g) + i4);
      float4 mv = reinterpret_cast<const float4 *>(m)[i4], vv = reinterpret_cast<const float4 *>(v)[i4];
      adam_elem(pv.x, gv.x, mv.x, vv.x, c); adam_elem(pv.y, gv.y, mv.y, vv.y, c);
      adam_elem(pv.z, gv.z, mv.z, vv.z, c); adam_elem(pv.w, gv.w, mv.w, vv.w, c);
      reinterpret_cast<float4 *>(p)[i4] = pv;
      reinterpret_cast<float4 *>(m)[i4] = mv;
      reinterpret_cast<float4 *>(v)[i4] = vv;
      if (packed) {
        // cpgb_pack_mask word of 32 elements: low half bit j = (piggy > thr), high half bit j = (1 <= T <= inf_idx)
        unsigned lo = (pv.x > thr) | ((pv.y > thr) << 1) | ((pv.z > thr) << 2) | ((pv.w > thr) << 3);
        unsigned hi = 0xFu;
        if (tmask) {
          const uchar4 tv = __ldg(reinterpret_cast<const uchar4 *>(tmask) + i4);
          const unsigned uu = (unsigned)inf_idx;
          hi = (tv.x != 0 && tv.x <= uu) | ((tv.y != 0 && tv.y <= uu) << 1) | ((tv.z != 0 && tv.z <= uu) << 2) |
               ((tv.w != 0 && tv.w <= uu) << 3);
        }
        lo <<= 4 * (lane & 7); hi <<= 4 * (lane & 7);
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          lo |= __shfl_xor_sync(0xffffffffu, lo, o);
          hi |= __shfl_xor_sync(0xffffffffu, hi, o);
        }
        if ((lane & 7) == 0) packed[i4 >> 3] = ((unsigned long long)hi << 32) | lo;
      }
    }
  } else {
    // ragged tail of a tensor (or unaligned pointers): scalar; the packed words of this range bit by bit
    for (long long i0 = base; i0 < end; i0 += OPT_THREADS) {
      const long long i = i0 + threadIdx.x;
      bool keep = false, old = true;
      if (i < end) {
        float pe = p[i], me = m[i], ve = v[i];
        adam_elem(pe, g[i], me, ve, c);
        p[i] = pe; m[i] = me; v[i] = ve;
        keep = pe > thr;
        old = tmask ? (tmask[i] != 0 && tmask[i] <= (unsigned)inf_idx) : true;
      }
      if (packed) {
        // i0 is a multiple of 32 (OPT_CHUNK and OPT_THREADS are): a warp covers one word; lanes beyond n pack as the
        // bits cpgb_pack_mask leaves there (zero)
        const unsigned lo = __ballot_sync(0xffffffffu, i < end && keep);
        const unsigned hi = __ballot_sync(0xffffffffu, i < end && old);
        if (lane == 0 && i < end) packed[i >> 5] = ((unsigned long long)hi << 32) | lo;
      }
    }
  }
  // the last block to finish advances the step counter
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long *>(step_dev + 1), 1ull);
    if (done == (unsigned long long)total_blocks - 1ull) {
      step_dev[1] = 0;
      step_dev[0] = step;
    }
  }
}

int fill_table(OptTable &tb, int first, int count, float *const *param, const float *const *grad, float *const *s1,
               float *const *s2, const int64_t *n, uint64_t *const *packed, const uint8_t *const *tmask) {
  tb.count = count;
  int acc = 0;
  for (int i = 0; i < count; ++i) {
    const int j = first + i;
    tb.p[i] = param[j]; tb.g[i] = grad[j]; tb.s1[i] = s1[j]; tb.s2[i] = s2 ? s2[j] : nullptr;
    tb.packed[i] = packed ? reinterpret_cast<unsigned long long *>(packed[j]) : nullptr;
    tb.tmask[i] = tmask ? tmask[j] : nullptr;
    tb.n[i] = n[j];
    tb.blk_start[i] = acc;
    acc += (int)((n[j] + OPT_CHUNK - 1) / OPT_CHUNK);
  }
  tb.blk_start[count] = acc;
  return acc;
}

}  // namespace

}  // namespace cpgb

using namespace cpgb;

extern "C" {

int cpgb_sgd_nesterov_step(int32_t ntensors, float *const *param, const float *const *grad, float *const *momentum_buf,
                           const int64_t *n, float lr, float momentum, const float *lr_dev, void *stream) {
  if (ntensors < 0) { set_error("cpgb_sgd_nesterov_step: negative tensor count"); return CPGB_EINVAL; }
  if (ntensors == 0) return CPGB_OK;
  if (!param || !grad || !momentum_buf || !n) { set_error("cpgb_sgd_nesterov_step: null pointer"); return CPGB_EINVAL; }
  for (int i = 0; i < ntensors; ++i)
    if (n[i] < 0 || (n[i] > 0 && (!param[i] || !grad[i] || !momentum_buf[i]))) {
      set_error("cpgb_sgd_nesterov_step: bad tensor %d", i); return CPGB_EINVAL;
    }
  cudaStream_t st = (cudaStream_t)stream;
  int launches = 0;
  for (int first = 0; first < ntensors; first += OPT_MAX_TENSORS) {
    OptTable tb;
    const int count = ntensors - first < OPT_MAX_TENSORS ? ntensors - first : OPT_MAX_TENSORS;
    const int blocks = fill_table(tb, first, count, param, grad, momentum_buf, nullptr, n, nullptr, nullptr);
    if (blocks == 0) continue;
    sgd_nesterov_kernel<<<blocks, OPT_THREADS, 0, st>>>(tb, momentum, lr, lr_dev);
    ++launches;
  }
  CPGB_LAUNCH_OK_N("cpgb_sgd_nesterov_step", launches);
  return CPGB_OK;
}

int cpgb_adam_step(int32_t ntensors, float *const *param, const float *const *grad, float *const *exp_avg,
                   float *const *exp_avg_sq, const int64_t *n, double lr, double beta1, double beta2, double eps,
                   int64_t *step_dev, const double *lr_dev, uint64_t *const *packed, const uint8_t *const *tmask,
                   float thr, int32_t inference_idx, void *stream) {
  if (ntensors < 0) { set_error("cpgb_adam_step: negative tensor count"); return CPGB_EINVAL; }
  if (!step_dev) { set_error("cpgb_adam_step: the device step counter is required"); return CPGB_EINVAL; }
  if (ntensors > OPT_MAX_TENSORS) {
    // one launch must own the step counter: split the parameter list into groups of <= 40 tensors with a counter each
    set_error("cpgb_adam_step: at most %d tensors per call", OPT_MAX_TENSORS); return CPGB_EINVAL;
  }
  if (ntensors == 0) return CPGB_OK;
  if (!param || !grad || !exp_avg || !exp_avg_sq || !n) { set_error("cpgb_adam_step: null pointer"); return CPGB_EINVAL; }
  for (int i = 0; i < ntensors; ++i)
    if (n[i] < 0 || (n[i] > 0 && (!param[i] || !grad[i] || !exp_avg[i] || !exp_avg_sq[i]))) {
      set_error("cpgb_adam_step: bad tensor %d", i); return CPGB_EINVAL;
    }
  OptTable tb;
  const int blocks = fill_table(tb, 0, ntensors, param, grad, exp_avg, exp_avg_sq, n, packed, tmask);
  if (blocks == 0) return CPGB_OK;
  adam_kernel<<<blocks, OPT_THREADS, 0, (cudaStream_t)stream>>>(tb, lr, beta1, beta2, eps,
                                                               reinterpret_cast<long long *>(step_dev), lr_dev, thr,
                                                               inference_idx, blocks);
  CPGB_LAUNCH_OK("cpgb_adam_step");
  return CPGB_OK;
}

}  // extern "C"
