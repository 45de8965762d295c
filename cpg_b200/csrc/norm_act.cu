// BatchNorm2d (+ ReLU) over NHWC fp32 activations: the consumer of every masked convolution in
// models/vgg.py:109-118 and models/resnet.py:60-100 (conv -> nn.BatchNorm2d -> nn.ReLU(inplace=True)).
// the step right after the hot path.  Every kernel is an HBM-bound streaming pass
// over an [M = N*H*W pixels][C channels] array (C % 4 == 0), 16 bytes per access:
//
//   training forward : stats    -- per-channel sum / sum of squares; ONE 1024-thread block per SM, so a
//                                  channel has <= 148 partial pairs
//                      finalize -- mean, biased variance, rstd, running statistics, num_batches_tracked,
//                                  a = gamma*rstd, b = beta - mean*a           (one warp per channel, double)
//                      apply    -- y = max(0, a*x + b), optionally the maximum over each 2x2 window
//                                  (nn.MaxPool2d(2, 2) folded in: y is written at pooled resolution)
//   backward         : stats    -- g = dy * [a*x + b > 0] (at the window's first maximum when pooled);
//                                  sum g, sum g*xhat; a and b are recomputed from gamma, beta, mean, rstd
//                      finalize -- dbeta, dgamma, c1 = sum g / M, c2 = sum g*xhat / M
//                      apply    -- dx = a * (g - c1 - xhat * c2)
//   evaluation mode  : coefficients from the running statistics, then the same apply / backward kernels
//                      with c1 = c2 = 0.
//
// A thread owns one float4 of channels for the whole kernel (its coefficients live in registers) and
// walks pixel rows, so there is no per-element index arithmetic.  The ReLU mask and the pooling argmax are
// recomputed from x in the backward kernels, nothing but mean / rstd is saved.  Partial sums are combined
// in a fixed order in double precision: results are deterministic.
#include "common.cuh"

namespace cpgb {

namespace {

constexpr int NA_THREADS = 256;         // streaming (apply) passes: 8 blocks per SM
constexpr int NA_STATS_THREADS = 1024;  // reduction passes: ONE fat block per SM, so that a channel has only
constexpr int NA_STATS_PER_SM = 1;      // ~148 partial pairs and the finalize step is a single round of loads

int na_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n = v > 0 ? v : 148;
  }
  return n;
}

struct NaGeom {
  long long M;      // pixels
  int C;            // channels
  int Cs;           // floats between consecutive pixels (>= C, a multiple of 4; lanes C..Cs-1 are padding)
  int lanes;        // float4 channel groups handled by one block (<= NA_THREADS)
  int slots;        // pixel rows in flight per block = NA_THREADS / lanes
  int cchunks;      // blockIdx.y extent
};

NaGeom na_geom(long long M, int C, int Cs, int threads = NA_THREADS) {
  NaGeom g;
  g.M = M; g.C = C; g.Cs = Cs;
  const int c4 = Cs / 4;
  g.lanes = c4 < threads ? c4 : threads;
  g.slots = threads / g.lanes;
  g.cchunks = (c4 + g.lanes - 1) / g.lanes;
  return g;
}

int na_blocks(const NaGeom &g, int per_sm) {
  long long want = (g.M + g.slots - 1) / g.slots;
  long long cap = (long long)na_sms() * per_sm / g.cchunks;
  if (cap < 1) cap = 1;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

// Reduce the per-thread float4 pairs (s, q) of the pixel slots of a block and write one partial pair
// per channel: part[(block * C + c) * 2 + {0, 1}].
__device__ __forceinline__ void block_reduce_pairs(const NaGeom &g, float4 s, float4 q, int lane, int slot, int c4,
                                                   bool active, float *__restrict__ part) {
  __shared__ float4 red[2][NA_STATS_THREADS];
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = q;
  __syncthreads();
  if (slot == 0 && active) {
    for (int k = 1; k < g.slots; ++k) {
      const float4 a = red[0][k * g.lanes + lane], b = red[1][k * g.lanes + lane];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
    }
    float *dst = part + ((long long)blockIdx.x * g.Cs + c4 * 4) * 2;
    reinterpret_cast<float4 *>(dst)[0] = make_float4(s.x, q.x, s.y, q.y);
    reinterpret_cast<float4 *>(dst)[1] = make_float4(s.z, q.z, s.w, q.w);
  }
}

// Four per-channel values of channel group c4 out of an array of C floats; lanes beyond C read as `fill`.
__device__ __forceinline__ float4 ldp4(const float *__restrict__ p, int c4, int C, float fill) {
  const int c = c4 * 4;
  if (c + 4 <= C) return __ldg(reinterpret_cast<const float4 *>(p) + c4);
  float4 v;
  v.x = c + 0 < C ? __ldg(p + c + 0) : fill;
  v.y = c + 1 < C ? __ldg(p + c + 1) : fill;
  v.z = c + 2 < C ? __ldg(p + c + 2) : fill;
  v.w = c + 3 < C ? __ldg(p + c + 3) : fill;
  return v;
}
// Activation values of the padding lanes are undefined (whoever produced the tensor may not have written them):
// force them to zero so that nothing non-finite can leak into the (zero-coefficient) pad outputs.
__device__ __forceinline__ float4 pad0(float4 v, int c4, int C) {
  const int c = c4 * 4;
  if (c + 4 > C) {
    if (c + 0 >= C) v.x = 0.f;
    if (c + 1 >= C) v.y = 0.f;
    if (c + 2 >= C) v.z = 0.f;
    if (c + 3 >= C) v.w = 0.f;
  }
  return v;
}

__global__ void __launch_bounds__(NA_STATS_THREADS)
bn_stats_kernel(const NaGeom g, const float *__restrict__ x, float *__restrict__ part) {
  pdl_wait();
  const int lane = threadIdx.x % g.lanes, slot = threadIdx.x / g.lanes;
  const int c4 = blockIdx.y * g.lanes + lane;
  const bool active = slot < g.slots && c4 * 4 < g.Cs;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  if (active) {
    const float4 *xp = reinterpret_cast<const float4 *>(x) + c4;
    const long long stride = (long long)gridDim.x * g.slots, cq = g.Cs / 4;
#pragma unroll 4
    for (long long r = (long long)blockIdx.x * g.slots + slot; r < g.M; r += stride) {
      const float4 v = pad0(__ldg(xp + r * cq), c4, g.C);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
  }
  block_reduce_pairs(g, s, q, lane, slot, c4, active, part);
}

// Sum the per-block partial pairs of channel c: one warp per channel, lanes stride over the blocks,
// fixed butterfly in double precision (deterministic for a given grid).
__device__ __forceinline__ void sum_partials(const float *__restrict__ part, int nblocks, int C, int c, double &s,
                                             double &q) {
  const int lane = threadIdx.x & 31;
  s = 0.0; q = 0.0;
  const float2 *p2 = reinterpret_cast<const float2 *>(part) + c;
  int b = lane;
  for (; b + 96 < nblocks; b += 128) {          // four independent loads in flight per lane
    const float2 v0 = __ldg(p2 + (long long)b * C), v1 = __ldg(p2 + (long long)(b + 32) * C);
    const float2 v2 = __ldg(p2 + (long long)(b + 64) * C), v3 = __ldg(p2 + (long long)(b + 96) * C);
    s += (double)v0.x; q += (double)v0.y;
    s += (double)v1.x; q += (double)v1.y;
    s += (double)v2.x; q += (double)v2.y;
    s += (double)v3.x; q += (double)v3.y;
  }
  for (; b < nblocks; b += 32) {
    const float2 v = __ldg(p2 + (long long)b * C);
    s += (double)v.x; q += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
}

// mean / biased variance / rstd, running statistics (torch.nn.BatchNorm2d: unbiased variance, momentum;
// momentum < 0 means cumulative average with factor 1 / num_batches_tracked passed as -momentum),
// coefficients a = gamma * rstd, b = beta - mean * a.
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float *__restrict__ part, int nblocks, int C, int Cs, long long M, const float *__restrict__ gamma,
                   const float *__restrict__ beta, float *__restrict__ running_mean, float *__restrict__ running_var,
                   float momentum, float eps, float *__restrict__ save_mean, float *__restrict__ save_rstd,
                   float *__restrict__ coef_a, float *__restrict__ coef_b, long long *__restrict__ num_batches_tracked) {
  pdl_wait();
  if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *num_batches_tracked += 1;   // nn.BatchNorm2d.forward
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= Cs) return;
  if (c >= C) {                       // padding lane: zero coefficients, so y = 0 there
    if ((threadIdx.x & 31) == 0) { coef_a[c] = 0.f; coef_b[c] = 0.f; }
    return;
  }
  double s, q;
  sum_partials(part, nblocks, Cs, c, s, q);
  if ((threadIdx.x & 31) != 0) return;
  const double mean = s / (double)M;
  double var = q / (double)M - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  save_mean[c] = (float)mean;
  save_rstd[c] = rstd;
  const float a = (gamma ? gamma[c] : 1.f) * rstd;
  coef_a[c] = a;
  coef_b[c] = (beta ? beta[c] : 0.f) - (float)mean * a;
  if (running_mean) {
    const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mean);
    running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unbiased);
  }
}

// evaluation mode: coefficients out of the running statistics
__global__ void __launch_bounds__(128)
bn_eval_coef_kernel(int C, int Cs, const float *__restrict__ gamma, const float *__restrict__ beta,
                    const float *__restrict__ running_mean, const float *__restrict__ running_var, float eps,
                    float *__restrict__ coef_a, float *__restrict__ coef_b) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cs) return;
  if (c >= C) { coef_a[c] = 0.f; coef_b[c] = 0.f; return; }
  const float rstd = 1.0f / sqrtf(running_var[c] + eps);
  const float a = (gamma ? gamma[c] : 1.f) * rstd;
  coef_a[c] = a;
  coef_b[c] = (beta ? beta[c] : 0.f) - running_mean[c] * a;
}

// a = gamma * rstd, b = beta - mean * a for the four channels of a thread (the same expressions as the
// forward finalize step, so the recomputed ReLU mask matches the forward pass bit for bit)
__device__ __forceinline__ void coef_from_stats(const float *__restrict__ gamma, const float *__restrict__ beta, int c4,
                                                int C, const float4 &mu, const float4 &rs, float4 &a, float4 &b) {
  const float4 gm = gamma ? ldp4(gamma, c4, C, 0.f) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 bt = beta ? ldp4(beta, c4, C, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
  a = make_float4(gm.x * rs.x, gm.y * rs.y, gm.z * rs.z, gm.w * rs.w);
  b = make_float4(bt.x - mu.x * a.x, bt.y - mu.y * a.y, bt.z - mu.z * a.z, bt.w - mu.w * a.w);
}

// round-to-nearest TF32 of the stored outputs (cpgb_bn_relu_*'s tf32_out): the consumer convolution's tensor-core
// operand is then exact instead of truncated
__device__ __forceinline__ float na_rna(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}
__device__ __forceinline__ float4 na_out(float4 o, int tf32) {
  return tf32 ? make_float4(na_rna(o.x), na_rna(o.y), na_rna(o.z), na_rna(o.w)) : o;
}

__global__ void __launch_bounds__(NA_THREADS)
bn_apply_kernel(const NaGeom g, const float *__restrict__ x, const float *__restrict__ coef_a,
                const float *__restrict__ coef_b, int relu, int tf32, float *__restrict__ y) {
  pdl_wait();
  const int lane = threadIdx.x % g.lanes, slot = threadIdx.x / g.lanes;
  const int c4 = blockIdx.y * g.lanes + lane;
  if (slot >= g.slots || c4 * 4 >= g.Cs) return;
  const float4 a = __ldg(reinterpret_cast<const float4 *>(coef_a) + c4);
  const float4 b = __ldg(reinterpret_cast<const float4 *>(coef_b) + c4);
  const float4 *xp = reinterpret_cast<const float4 *>(x) + c4;
  float4 *yp = reinterpret_cast<float4 *>(y) + c4;
  const long long stride = (long long)gridDim.x * g.slots, cq = g.Cs / 4;
#pragma unroll 4
  for (long long r = (long long)blockIdx.x * g.slots + slot; r < g.M; r += stride) {
    const float4 v = pad0(__ldg(xp + r * cq), c4, g.C);
    float4 o = make_float4(fmaf(v.x, a.x, b.x), fmaf(v.y, a.y, b.y), fmaf(v.z, a.z, b.z), fmaf(v.w, a.w, b.w));
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    yp[r * cq] = na_out(o, tf32);
  }
}

// g = dy * [relu ? a*x + b > 0 : 1];  xhat = (x - mean) * rstd
#define NA_GRAD_ELEM(G, XH, V, D, A, B, MU, RS)                     \
  {                                                                 \
    const float z_ = fmaf(V, A, B);                                 \
    G = (relu && !(z_ > 0.f)) ? 0.f : D;                            \
    XH = (V - MU) * RS;                                             \
  }

__global__ void __launch_bounds__(NA_STATS_THREADS)
bn_bwd_stats_kernel(const NaGeom g, const float *__restrict__ x, const float *__restrict__ dy,
                    const float *__restrict__ gamma, const float *__restrict__ beta,
                    const float *__restrict__ save_mean, const float *__restrict__ save_rstd, int relu,
                    float *__restrict__ part) {
  pdl_wait();
  const int lane = threadIdx.x % g.lanes, slot = threadIdx.x / g.lanes;
  const int c4 = blockIdx.y * g.lanes + lane;
  const bool active = slot < g.slots && c4 * 4 < g.Cs;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  if (active) {
    const float4 mu = ldp4(save_mean, c4, g.C, 0.f);
    const float4 rs = ldp4(save_rstd, c4, g.C, 0.f);
    float4 a, b;
    coef_from_stats(gamma, beta, c4, g.C, mu, rs, a, b);
    const float4 *xp = reinterpret_cast<const float4 *>(x) + c4;
    const float4 *dp = reinterpret_cast<const float4 *>(dy) + c4;
    const long long stride = (long long)gridDim.x * g.slots, cq = g.Cs / 4;
#pragma unroll 4
    for (long long r = (long long)blockIdx.x * g.slots + slot; r < g.M; r += stride) {
      const float4 v = pad0(__ldg(xp + r * cq), c4, g.C), d = pad0(__ldg(dp + r * cq), c4, g.C);
      float gx, gy, gz, gw, hx, hy, hz, hw;
      NA_GRAD_ELEM(gx, hx, v.x, d.x, a.x, b.x, mu.x, rs.x)
      NA_GRAD_ELEM(gy, hy, v.y, d.y, a.y, b.y, mu.y, rs.y)
      NA_GRAD_ELEM(gz, hz, v.z, d.z, a.z, b.z, mu.z, rs.z)
      NA_GRAD_ELEM(gw, hw, v.w, d.w, a.w, b.w, mu.w, rs.w)
      s.x += gx; s.y += gy; s.z += gz; s.w += gw;
      q.x = fmaf(gx, hx, q.x); q.y = fmaf(gy, hy, q.y); q.z = fmaf(gz, hz, q.z); q.w = fmaf(gw, hw, q.w);
    }
  }
  block_reduce_pairs(g, s, q, lane, slot, c4, active, part);
}

// dbeta = sum g, dgamma = sum g * xhat; c1 = dbeta / M, c2 = dgamma / M (training) or 0 (evaluation
// mode: the statistics are constants, dx = a * g)
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float *__restrict__ part, int nblocks, int C, int Cs, long long M, int training,
                       float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ c1,
                       float *__restrict__ c2) {
  pdl_wait();
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= Cs) return;
  if (c >= C) {
    if ((threadIdx.x & 31) == 0) { c1[c] = 0.f; c2[c] = 0.f; }
    return;
  }
  double s, q;
  sum_partials(part, nblocks, Cs, c, s, q);
  if ((threadIdx.x & 31) != 0) return;
  if (dbeta) dbeta[c] = (float)s;
  if (dgamma) dgamma[c] = (float)q;
  c1[c] = training ? (float)(s / (double)M) : 0.f;
  c2[c] = training ? (float)(q / (double)M) : 0.f;
}

__global__ void __launch_bounds__(NA_THREADS)
bn_bwd_apply_kernel(const NaGeom g, const float *__restrict__ x, const float *__restrict__ dy,
                    const float *__restrict__ gamma, const float *__restrict__ beta,
                    const float *__restrict__ save_mean, const float *__restrict__ save_rstd,
                    const float *__restrict__ c1, const float *__restrict__ c2, int relu, int tf32,
                    float *__restrict__ dx) {
  pdl_wait();
  const int lane = threadIdx.x % g.lanes, slot = threadIdx.x / g.lanes;
  const int c4 = blockIdx.y * g.lanes + lane;
  if (slot >= g.slots || c4 * 4 >= g.Cs) return;
  const float4 mu = ldp4(save_mean, c4, g.C, 0.f);
  const float4 rs = ldp4(save_rstd, c4, g.C, 0.f);
  float4 a, b;
  coef_from_stats(gamma, beta, c4, g.C, mu, rs, a, b);
  const float4 k1 = __ldg(reinterpret_cast<const float4 *>(c1) + c4);
  const float4 k2 = __ldg(reinterpret_cast<const float4 *>(c2) + c4);
  const float4 *xp = reinterpret_cast<const float4 *>(x) + c4;
  const float4 *dp = reinterpret_cast<const float4 *>(dy) + c4;
  float4 *op = reinterpret_cast<float4 *>(dx) + c4;
  const long long stride = (long long)gridDim.x * g.slots, cq = g.Cs / 4;
#pragma unroll 4
  for (long long r = (long long)blockIdx.x * g.slots + slot; r < g.M; r += stride) {
    const float4 v = pad0(__ldg(xp + r * cq), c4, g.C), d = pad0(__ldg(dp + r * cq), c4, g.C);
    float gx, gy, gz, gw, hx, hy, hz, hw;
    NA_GRAD_ELEM(gx, hx, v.x, d.x, a.x, b.x, mu.x, rs.x)
    NA_GRAD_ELEM(gy, hy, v.y, d.y, a.y, b.y, mu.y, rs.y)
    NA_GRAD_ELEM(gz, hz, v.z, d.z, a.z, b.z, mu.z, rs.z)
    NA_GRAD_ELEM(gw, hw, v.w, d.w, a.w, b.w, mu.w, rs.w)
    op[r * cq] = na_out(make_float4(a.x * (gx - k1.x - hx * k2.x), a.y * (gy - k1.y - hy * k2.y),
                                    a.z * (gz - k1.z - hz * k2.z), a.w * (gw - k1.w - hw * k2.w)), tf32);
  }
}

// ---- residual add folded in (models/resnet.py:52-55, 95-98: out = bn(conv(..)); out += identity; out = relu(out)) ----
// y = max(0, a*x + b + r).  The backward pass gates on the stored block output (y > 0, what torch's ReLU backward
// reads): g = dy * [y > 0] is at once the gradient of the identity branch and the input of the batch-norm backward, so
// the statistics kernel writes it out and bn_bwd_apply_kernel (relu = 0) finishes dx from it.
__global__ void __launch_bounds__(NA_THREADS)
bn_apply_res_kernel(const NaGeom g, const float *__restrict__ x, const float *__restrict__ res,
                    const float *__restrict__ coef_a, const float *__restrict__ coef_b, int tf32,
                    float *__restrict__ y) {
  pdl_wait();
  const int lane = threadIdx.x % g.lanes, slot = threadIdx.x / g.lanes;
  const int c4 = blockIdx.y * g.lanes + lane;
  if (slot >= g.slots || c4 * 4 >= g.Cs) return;
  const float4 a = __ldg(reinterpret_cast<const float4 *>(coef_a) + c4);
  const float4 b = __ldg(reinterpret_cast<const float4 *>(coef_b) + c4);
  const float4 *xp = reinterpret_cast<const float4 *>(x) + c4;
  const float4 *rp = reinterpret_cast<const float4 *>(res) + c4;
  float4 *yp = reinterpret_cast<float4 *>(y) + c4;
  const long long stride = (long long)gridDim.x * g.slots, cq = g.Cs / 4;
#pragma unroll 4
  for (long long r = (long long)blockIdx.x * g.slots + slot; r < g.M; r += stride) {
    const float4 v = pad0(__ldg(xp + r * cq), c4, g.C), w = pad0(__ldg(rp + r * cq), c4, g.C);
    float4 o = make_float4(fmaf(v.x, a.x, b.x) + w.x, fmaf(v.y, a.y, b.y) + w.y, fmaf(v.z, a.z, b.z) + w.z,
                           fmaf(v.w, a.w, b.w) + w.w);
    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    yp[r * cq] = na_out(o, tf32);
  }
}

__global__ void __launch_bounds__(NA_STATS_THREADS)
bn_bwd_stats_res_kernel(const NaGeom g, const float *__restrict__ x, const float *__restrict__ y,
                        const float *__restrict__ dy, const float *__restrict__ save_mean,
                        const float *__restrict__ save_rstd, float *__restrict__ gout, float *__restrict__ part) {
  pdl_wait();
  const int lane = threadIdx.x % g.lanes, slot = threadIdx.x / g.lanes;
  const int c4 = blockIdx.y * g.lanes + lane;
  const bool active = slot < g.slots && c4 * 4 < g.Cs;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  if (active) {
    const float4 mu = ldp4(save_mean, c4, g.C, 0.f);
    const float4 rs = ldp4(save_rstd, c4, g.C, 0.f);
    const float4 *xp = reinterpret_cast<const float4 *>(x) + c4;
    const float4 *yp = reinterpret_cast<const float4 *>(y) + c4;
    const float4 *dp = reinterpret_cast<const float4 *>(dy) + c4;
    float4 *gp = reinterpret_cast<float4 *>(gout) + c4;
    const long long stride = (long long)gridDim.x * g.slots, cq = g.Cs / 4;
#pragma unroll 4
    for (long long r = (long long)blockIdx.x * g.slots + slot; r < g.M; r += stride) {
      const float4 v = pad0(__ldg(xp + r * cq), c4, g.C), o = pad0(__ldg(yp + r * cq), c4, g.C),
                   d = pad0(__ldg(dp + r * cq), c4, g.C);
      float4 gg;
      gg.x = o.x > 0.f ? d.x : 0.f; gg.y = o.y > 0.f ? d.y : 0.f;
      gg.z = o.z > 0.f ? d.z : 0.f; gg.w = o.w > 0.f ? d.w : 0.f;
      gp[r * cq] = gg;
      s.x += gg.x; s.y += gg.y; s.z += gg.z; s.w += gg.w;
      q.x = fmaf(gg.x, (v.x - mu.x) * rs.x, q.x); q.y = fmaf(gg.y, (v.y - mu.y) * rs.y, q.y);
      q.z = fmaf(gg.z, (v.z - mu.z) * rs.z, q.z); q.w = fmaf(gg.w, (v.w - mu.w) * rs.w, q.w);
    }
  }
  block_reduce_pairs(g, s, q, lane, slot, c4, active, part);
}

// ---- 2x2 / stride-2 max-pool folded in (models/vgg.py 'M' entries follow conv -> BN -> ReLU directly) ----
// Rows are POOLED pixels; a thread reads the four window pixels of its channels.  pool(relu(z)) ==
// relu(max z), and the gradient of a window goes to its first maximum in (h, w) scan order -- the element
// torch's max_pool2d records (ties other than at z <= 0, where the ReLU zeroes the gradient anyway, are
// measure-zero for float activations).
struct PoolGeom { int H, W, Ho, Wo; long long Mo; };

__device__ __forceinline__ long long pool_base(const PoolGeom &pg, long long r, long long cq) {
  const int wo = (int)(r % pg.Wo);
  const long long t = r / pg.Wo;
  const int ho = (int)(t % pg.Ho);
  const long long n = t / pg.Ho;
  return ((n * pg.H + 2 * ho) * pg.W + 2 * wo) * cq;
}

__device__ __forceinline__ float max4_first(float z0, float z1, float z2, float z3, int &j) {
  float m = z0; j = 0;
  if (z1 > m) { m = z1; j = 1; }
  if (z2 > m) { m = z2; j = 2; }
  if (z3 > m) { m = z3; j = 3; }
  return m;
}

__global__ void __launch_bounds__(NA_THREADS)
bn_apply_pool_kernel(const NaGeom g, const PoolGeom pg, const float *__restrict__ x, const float *__restrict__ coef_a,
                     const float *__restrict__ coef_b, int relu, int tf32, float *__restrict__ y) {
  pdl_wait();
  const int lane = threadIdx.x % g.lanes, slot = threadIdx.x / g.lanes;
  const int c4 = blockIdx.y * g.lanes + lane;
  if (slot >= g.slots || c4 * 4 >= g.Cs) return;
  const float4 a = __ldg(reinterpret_cast<const float4 *>(coef_a) + c4);
  const float4 b = __ldg(reinterpret_cast<const float4 *>(coef_b) + c4);
  const float4 *xp = reinterpret_cast<const float4 *>(x) + c4;
  float4 *yp = reinterpret_cast<float4 *>(y) + c4;
  const long long stride = (long long)gridDim.x * g.slots, cq = g.Cs / 4, down = (long long)pg.W * cq;
#pragma unroll 2
  for (long long r = (long long)blockIdx.x * g.slots + slot; r < pg.Mo; r += stride) {
    const float4 *p = xp + pool_base(pg, r, cq);
    const float4 v0 = pad0(__ldg(p), c4, g.C), v1 = pad0(__ldg(p + cq), c4, g.C), v2 = pad0(__ldg(p + down), c4, g.C),
                 v3 = pad0(__ldg(p + down + cq), c4, g.C);
    float4 o;
    o.x = fmaxf(fmaxf(fmaf(v0.x, a.x, b.x), fmaf(v1.x, a.x, b.x)), fmaxf(fmaf(v2.x, a.x, b.x), fmaf(v3.x, a.x, b.x)));
    o.y = fmaxf(fmaxf(fmaf(v0.y, a.y, b.y), fmaf(v1.y, a.y, b.y)), fmaxf(fmaf(v2.y, a.y, b.y), fmaf(v3.y, a.y, b.y)));
    o.z = fmaxf(fmaxf(fmaf(v0.z, a.z, b.z), fmaf(v1.z, a.z, b.z)), fmaxf(fmaf(v2.z, a.z, b.z), fmaf(v3.z, a.z, b.z)));
    o.w = fmaxf(fmaxf(fmaf(v0.w, a.w, b.w), fmaf(v1.w, a.w, b.w)), fmaxf(fmaf(v2.w, a.w, b.w), fmaf(v3.w, a.w, b.w)));
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    yp[r * cq] = na_out(o, tf32);
  }
}

// one channel of one window: gradient g of the winning pixel, its index j and its xhat
#define NA_POOL_ELEM(G, J, XH, V0, V1, V2, V3, D, A, B, MU, RS)                                   \
  {                                                                                                \
    const float m_ = max4_first(fmaf(V0, A, B), fmaf(V1, A, B), fmaf(V2, A, B), fmaf(V3, A, B), J); \
    G = (relu && !(m_ > 0.f)) ? 0.f : D;                                                           \
    const float xv_ = J == 0 ? V0 : J == 1 ? V1 : J == 2 ? V2 : V3;                                \
    XH = (xv_ - MU) * RS;                                                                          \
  }

__global__ void __launch_bounds__(NA_STATS_THREADS)
bn_bwd_stats_pool_kernel(const NaGeom g, const PoolGeom pg, const float *__restrict__ x, const float *__restrict__ dy,
                         const float *__restrict__ gamma, const float *__restrict__ beta,
                         const float *__restrict__ save_mean, const float *__restrict__ save_rstd, int relu,
                         float *__restrict__ part) {
  pdl_wait();
  const int lane = threadIdx.x % g.lanes, slot = threadIdx.x / g.lanes;
  const int c4 = blockIdx.y * g.lanes + lane;
  const bool active = slot < g.slots && c4 * 4 < g.Cs;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  if (active) {
    const float4 mu = ldp4(save_mean, c4, g.C, 0.f);
    const float4 rs = ldp4(save_rstd, c4, g.C, 0.f);
    float4 a, b;
    coef_from_stats(gamma, beta, c4, g.C, mu, rs, a, b);
    const float4 *xp = reinterpret_cast<const float4 *>(x) + c4;
    const float4 *dp = reinterpret_cast<const float4 *>(dy) + c4;
    const long long stride = (long long)gridDim.x * g.slots, cq = g.Cs / 4, down = (long long)pg.W * cq;
#pragma unroll 2
    for (long long r = (long long)blockIdx.x * g.slots + slot; r < pg.Mo; r += stride) {
      const float4 *p = xp + pool_base(pg, r, cq);
      const float4 v0 = pad0(__ldg(p), c4, g.C), v1 = pad0(__ldg(p + cq), c4, g.C), v2 = pad0(__ldg(p + down), c4, g.C),
                 v3 = pad0(__ldg(p + down + cq), c4, g.C);
      const float4 d = pad0(__ldg(dp + r * cq), c4, g.C);
      float gx, gy, gz, gw, hx, hy, hz, hw;
      int jx, jy, jz, jw;
      NA_POOL_ELEM(gx, jx, hx, v0.x, v1.x, v2.x, v3.x, d.x, a.x, b.x, mu.x, rs.x)
      NA_POOL_ELEM(gy, jy, hy, v0.y, v1.y, v2.y, v3.y, d.y, a.y, b.y, mu.y, rs.y)
      NA_POOL_ELEM(gz, jz, hz, v0.z, v1.z, v2.z, v3.z, d.z, a.z, b.z, mu.z, rs.z)
      NA_POOL_ELEM(gw, jw, hw, v0.w, v1.w, v2.w, v3.w, d.w, a.w, b.w, mu.w, rs.w)
      s.x += gx; s.y += gy; s.z += gz; s.w += gw;
      q.x = fmaf(gx, hx, q.x); q.y = fmaf(gy, hy, q.y); q.z = fmaf(gz, hz, q.z); q.w = fmaf(gw, hw, q.w);
    }
  }
  block_reduce_pairs(g, s, q, lane, slot, c4, active, part);
}

// dx of the four window pixels of one channel: a * ([j == J] * g - c1 - xhat_j * c2)
#define NA_POOL_DX(O0, O1, O2, O3, V0, V1, V2, V3, D, A, B, MU, RS, K1, K2)                        \
  {                                                                                                \
    int j_;                                                                                        \
    const float m_ = max4_first(fmaf(V0, A, B), fmaf(V1, A, B), fmaf(V2, A, B), fmaf(V3, A, B), j_); \
    const float g_ = (relu && !(m_ > 0.f)) ? 0.f : D;                                              \
    O0 = A * ((j_ == 0 ? g_ : 0.f) - K1 - (V0 - MU) * RS * K2);                                    \
    O1 = A * ((j_ == 1 ? g_ : 0.f) - K1 - (V1 - MU) * RS * K2);                                    \
    O2 = A * ((j_ == 2 ? g_ : 0.f) - K1 - (V2 - MU) * RS * K2);                                    \
    O3 = A * ((j_ == 3 ? g_ : 0.f) - K1 - (V3 - MU) * RS * K2);                                    \
  }

__global__ void __launch_bounds__(NA_THREADS)
bn_bwd_apply_pool_kernel(const NaGeom g, const PoolGeom pg, const float *__restrict__ x, const float *__restrict__ dy,
                         const float *__restrict__ gamma, const float *__restrict__ beta,
                         const float *__restrict__ save_mean, const float *__restrict__ save_rstd,
                         const float *__restrict__ c1, const float *__restrict__ c2, int relu, int tf32,
                         float *__restrict__ dx) {
  pdl_wait();
  const int lane = threadIdx.x % g.lanes, slot = threadIdx.x / g.lanes;
  const int c4 = blockIdx.y * g.lanes + lane;
  if (slot >= g.slots || c4 * 4 >= g.Cs) return;
  const float4 mu = ldp4(save_mean, c4, g.C, 0.f);
  const float4 rs = ldp4(save_rstd, c4, g.C, 0.f);
  float4 a, b;
  coef_from_stats(gamma, beta, c4, g.C, mu, rs, a, b);
  const float4 k1 = __ldg(reinterpret_cast<const float4 *>(c1) + c4);
  const float4 k2 = __ldg(reinterpret_cast<const float4 *>(c2) + c4);
  const float4 *xp = reinterpret_cast<const float4 *>(x) + c4;
  const float4 *dp = reinterpret_cast<const float4 *>(dy) + c4;
  float4 *op = reinterpret_cast<float4 *>(dx) + c4;
  const long long stride = (long long)gridDim.x * g.slots, cq = g.Cs / 4, down = (long long)pg.W * cq;
#pragma unroll 2
  for (long long r = (long long)blockIdx.x * g.slots + slot; r < pg.Mo; r += stride) {
    const long long base = pool_base(pg, r, cq);
    const float4 *p = xp + base;
    const float4 v0 = pad0(__ldg(p), c4, g.C), v1 = pad0(__ldg(p + cq), c4, g.C), v2 = pad0(__ldg(p + down), c4, g.C),
                 v3 = pad0(__ldg(p + down + cq), c4, g.C);
    const float4 d = pad0(__ldg(dp + r * cq), c4, g.C);
    float4 o0, o1, o2, o3;
    NA_POOL_DX(o0.x, o1.x, o2.x, o3.x, v0.x, v1.x, v2.x, v3.x, d.x, a.x, b.x, mu.x, rs.x, k1.x, k2.x)
    NA_POOL_DX(o0.y, o1.y, o2.y, o3.y, v0.y, v1.y, v2.y, v3.y, d.y, a.y, b.y, mu.y, rs.y, k1.y, k2.y)
    NA_POOL_DX(o0.z, o1.z, o2.z, o3.z, v0.z, v1.z, v2.z, v3.z, d.z, a.z, b.z, mu.z, rs.z, k1.z, k2.z)
    NA_POOL_DX(o0.w, o1.w, o2.w, o3.w, v0.w, v1.w, v2.w, v3.w, d.w, a.w, b.w, mu.w, rs.w, k1.w, k2.w)
    float4 *o = op + base;
    o[0] = na_out(o0, tf32); o[cq] = na_out(o1, tf32); o[down] = na_out(o2, tf32); o[down + cq] = na_out(o3, tf32);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Single-kernel variants: one launch per direction instead of stats -> finalize -> apply.
//
// A CTA owns `lanes` float4 channel groups (blockIdx.y) for a share of the pixel rows; the `cs` CTAs of a thread-block
// cluster (blockIdx.x = cluster rank) split the rows of the same channels.  Pass 1 accumulates the per-channel sums,
// the CTA's partial pair goes to its own shared memory, ONE cluster barrier later every CTA adds the `cs` partials of
// its channels through distributed shared memory in rank order (so all of them compute bit-identical statistics, in
// double precision), derives the coefficients and walks its rows a second time (x / dy come out of L2 now) to write y
// or dx.  No partial sums in global memory, no finalize kernel, no second and third launch.  Deterministic: every reduction is a
// fixed tree (warp shuffles over the pixel slots of a warp, then the warps and the ranks in ascending order).
// ------------------------------------------------------------------------------------------------------------------
constexpr int NC_THREADS = 512;        // two CTAs per SM can co-reside (also next to a wgrad CTA of the side stream)
constexpr int NC_WARPS = NC_THREADS / 32;
constexpr int NC_MAXL = 32;            // channel groups per CTA: a power of two <= 32 (lanes of a warp)
constexpr int NC_MAX_CLUSTER = 8;      // portable cluster size

struct NcPlan { int lanes, slots, chunks, cs; };
NcPlan nc_plan(long long rows, int Cs) {
  NcPlan p;
  const int c4 = Cs / 4;
  // experiment knobs: CPGB_BN_TARGET_CHUNKS (channel chunks aimed at), CPGB_BN_CS (largest cluster)
  const char *et = getenv("CPGB_BN_TARGET_CHUNKS"), *ec = getenv("CPGB_BN_CS");
  const int target = et ? atoi(et) : 32, max_cs = ec ? atoi(ec) : 4;     // measured best for tensors <= 9 MB
  int L = 2;                                             // >= 32 contiguous bytes per pixel row and CTA: whole sectors
  while (L < NC_MAXL && c4 / (L * 2) >= target) L *= 2;  // about `target` channel chunks ...
  p.lanes = L; p.slots = NC_THREADS / L; p.chunks = (c4 + L - 1) / L;
  int cs = na_sms() / p.chunks;                          // ... times up to `max_cs` row shares: 64-148 CTAs
  if (cs > max_cs) cs = max_cs;
  if (cs > NC_MAX_CLUSTER) cs = NC_MAX_CLUSTER;
  if (cs < 1) cs = 1;
  while (cs > 1 && (long long)(cs - 1) * p.slots >= rows) --cs;   // tiny maps: no idle ranks
  p.cs = cs;
  return p;
}

__device__ __forceinline__ uint32_t nc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 nc_ld_peer(const void *p, uint32_t rank) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(nc_smem_u32(p)), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void nc_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void nc_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

struct NcShared {
  float4 wred[2][NC_WARPS][NC_MAXL];   // per-warp partial pairs
  float4 xch[2][NC_MAXL];              // this CTA's partial pair per channel group: read by the cluster peers
  float4 coef[4][NC_MAXL];             // what pass 2 needs, per channel group
};

// Sum the (s, q) pairs of all pixel slots of the CTA: shuffles over the slots of a warp, then the warps in order (double).
// The result lands in sh.xch[.][lane] (as fp32, like the per-block partials of the three-kernel path).
__device__ __forceinline__ void nc_block_reduce(NcShared &sh, int L, float4 s, float4 q) {
  for (int o = L; o < 32; o <<= 1) {
    s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
    s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
    q.x += __shfl_xor_sync(0xffffffffu, q.x, o); q.y += __shfl_xor_sync(0xffffffffu, q.y, o);
    q.z += __shfl_xor_sync(0xffffffffu, q.z, o); q.w += __shfl_xor_sync(0xffffffffu, q.w, o);
  }
  const int wl = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (wl < L) { sh.wred[0][w][wl] = s; sh.wred[1][w][wl] = q; }
  __syncthreads();
  if (threadIdx.x < 8 * L) {
    // one thread per (pair member, channel): the warps in ascending order, double precision
    const int which = threadIdx.x / (4 * L), ch = threadIdx.x % (4 * L);
    const float *src = reinterpret_cast<const float *>(&sh.wred[which][0][0]) + ch;
    double d = 0.0;
#pragma unroll
    for (int k = 0; k < NC_WARPS; ++k) d += (double)src[k * NC_MAXL * 4];
    reinterpret_cast<float *>(&sh.xch[which][0])[ch] = (float)d;
  }
}
__device__ __forceinline__ float nc_ld_peer_f(const float *p, uint32_t rank) {
  uint32_t a;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(nc_smem_u32(p)), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
// After the cluster barrier: totals of channel `ch` (0 .. 4 * lanes - 1 of this CTA's chunk) over the cluster, ranks in
// ascending order.  All loads are issued before the first add.
__device__ __forceinline__ void nc_cluster_total(const NcShared &sh, int ch, int cs, double &ds, double &dq) {
  float a[NC_MAX_CLUSTER], b[NC_MAX_CLUSTER];
#pragma unroll
  for (int r = 0; r < NC_MAX_CLUSTER; ++r) {
    a[r] = r < cs ? nc_ld_peer_f(reinterpret_cast<const float *>(&sh.xch[0][0]) + ch, (uint32_t)r) : 0.f;
    b[r] = r < cs ? nc_ld_peer_f(reinterpret_cast<const float *>(&sh.xch[1][0]) + ch, (uint32_t)r) : 0.f;
  }
  ds = 0.0; dq = 0.0;
#pragma unroll
  for (int r = 0; r < NC_MAX_CLUSTER; ++r) { ds += (double)a[r]; dq += (double)b[r]; }
}

struct NcFwdArgs {
  const float *x, *gamma, *beta;
  float *running_mean, *running_var;
  long long *nbt;
  float *y, *save_mean, *save_rstd;
  float momentum, eps;
  int training, relu, tf32, cs;
};

template <bool POOL>
__global__ void __launch_bounds__(NC_THREADS)
bn_fwd_cluster_kernel(const NaGeom g, const PoolGeom pg, const NcFwdArgs p) {
  __shared__ NcShared sh;
  pdl_wait();
  const int L = g.lanes;
  const int lane = threadIdx.x % L, slot = threadIdx.x / L;
  const int c4 = blockIdx.y * L + lane;
  const bool active = c4 * 4 < g.Cs;
  const long long cq = g.Cs / 4, stride = (long long)p.cs * g.slots;
  const long long r0 = (long long)blockIdx.x * g.slots + slot;       // blockIdx.x == cluster rank
  const float4 *xp = reinterpret_cast<const float4 *>(p.x) + c4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (p.training) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    if (active) {
#pragma unroll 4
      for (long long r = r0; r < g.M; r += stride) {
        const float4 v = pad0(__ldg(xp + r * cq), c4, g.C);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
      }
    }
    nc_block_reduce(sh, L, s, q);
    nc_cluster_arrive();
    nc_cluster_wait();
    if (threadIdx.x < 4 * L) {
      const int c = blockIdx.y * L * 4 + threadIdx.x;       // one thread per channel of this CTA's chunk
      float mu = 0.f, rs = 0.f;                             // padding lane: zero coefficients, y = 0 there
      if (c < g.C) {
        double ds, dq;
        nc_cluster_total(sh, threadIdx.x, p.cs, ds, dq);
        const double mean = ds / (double)g.M;
        double var = dq / (double)g.M - mean * mean;
        if (var < 0.0) var = 0.0;
        mu = (float)mean;
        rs = (float)(1.0 / sqrt(var + (double)p.eps));
        if (blockIdx.x == 0) {
          p.save_mean[c] = mu;
          p.save_rstd[c] = rs;
          if (p.running_mean) {
            const double unbiased = g.M > 1 ? var * (double)g.M / (double)(g.M - 1) : var;
            p.running_mean[c] = (float)((1.0 - (double)p.momentum) * (double)p.running_mean[c] + (double)p.momentum * mean);
            p.running_var[c] = (float)((1.0 - (double)p.momentum) * (double)p.running_var[c] + (double)p.momentum * unbiased);
          }
        }
      }
      reinterpret_cast<float *>(&sh.coef[2][0])[threadIdx.x] = mu;
      reinterpret_cast<float *>(&sh.coef[3][0])[threadIdx.x] = rs;
    }
    if (p.nbt && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *p.nbt += 1;     // nn.BatchNorm2d.forward
    nc_cluster_arrive();                                  // the peers' partials have been read; waited for at the end
    __syncthreads();
    if (active) coef_from_stats(p.gamma, p.beta, c4, g.C, sh.coef[2][lane], sh.coef[3][lane], a, b);
  } else if (active) {
    const float4 rm = ldp4(p.running_mean, c4, g.C, 0.f), rv = ldp4(p.running_var, c4, g.C, 1.f);
    const float4 rs = make_float4(1.0f / sqrtf(rv.x + p.eps), 1.0f / sqrtf(rv.y + p.eps), 1.0f / sqrtf(rv.z + p.eps),
                                  1.0f / sqrtf(rv.w + p.eps));
    coef_from_stats(p.gamma, p.beta, c4, g.C, rm, rs, a, b);
  }
  if (active) {
    // padding lanes (channels >= C of the last group): zero coefficients
    const int c = c4 * 4;
    if (c + 0 >= g.C) { a.x = 0.f; b.x = 0.f; }
    if (c + 1 >= g.C) { a.y = 0.f; b.y = 0.f; }
    if (c + 2 >= g.C) { a.z = 0.f; b.z = 0.f; }
    if (c + 3 >= g.C) { a.w = 0.f; b.w = 0.f; }
    float4 *yp = reinterpret_cast<float4 *>(p.y) + c4;
    const int relu = p.relu, tf32 = p.tf32;
    if (!POOL) {
#pragma unroll 4
      for (long long r = r0; r < g.M; r += stride) {
        const float4 v = pad0(__ldg(xp + r * cq), c4, g.C);
        float4 o = make_float4(fmaf(v.x, a.x, b.x), fmaf(v.y, a.y, b.y), fmaf(v.z, a.z, b.z), fmaf(v.w, a.w, b.w));
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        yp[r * cq] = na_out(o, tf32);
      }
    } else {
      const long long down = (long long)pg.W * cq;
#pragma unroll 2
      for (long long r = r0; r < pg.Mo; r += stride) {
        const float4 *w = xp + pool_base(pg, r, cq);
        const float4 v0 = pad0(__ldg(w), c4, g.C), v1 = pad0(__ldg(w + cq), c4, g.C), v2 = pad0(__ldg(w + down), c4, g.C),
                     v3 = pad0(__ldg(w + down + cq), c4, g.C);
        float4 o;
        o.x = fmaxf(fmaxf(fmaf(v0.x, a.x, b.x), fmaf(v1.x, a.x, b.x)), fmaxf(fmaf(v2.x, a.x, b.x), fmaf(v3.x, a.x, b.x)));
        o.y = fmaxf(fmaxf(fmaf(v0.y, a.y, b.y), fmaf(v1.y, a.y, b.y)), fmaxf(fmaf(v2.y, a.y, b.y), fmaf(v3.y, a.y, b.y)));
        o.z = fmaxf(fmaxf(fmaf(v0.z, a.z, b.z), fmaf(v1.z, a.z, b.z)), fmaxf(fmaf(v2.z, a.z, b.z), fmaf(v3.z, a.z, b.z)));
        o.w = fmaxf(fmaxf(fmaf(v0.w, a.w, b.w), fmaf(v1.w, a.w, b.w)), fmaxf(fmaf(v2.w, a.w, b.w), fmaf(v3.w, a.w, b.w)));
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        yp[r * cq] = na_out(o, tf32);
      }
    }
  }
  if (p.training) nc_cluster_wait();                     // nobody leaves while a peer may still read its partials
}

struct NcBwdArgs {
  const float *x, *dy, *gamma, *beta, *mean, *rstd;
  float *dx, *dgamma, *dbeta;
  int training, relu, tf32, cs;
};

template <bool POOL>
__global__ void __launch_bounds__(NC_THREADS)
bn_bwd_cluster_kernel(const NaGeom g, const PoolGeom pg, const NcBwdArgs p) {
  __shared__ NcShared sh;
  pdl_wait();
  const int L = g.lanes;
  const int lane = threadIdx.x % L, slot = threadIdx.x / L;
  const int c4 = blockIdx.y * L + lane;
  const bool active = c4 * 4 < g.Cs;
  const long long cq = g.Cs / 4, stride = (long long)p.cs * g.slots, down = (long long)pg.W * cq;
  const long long r0 = (long long)blockIdx.x * g.slots + slot;
  const long long rows = POOL ? pg.Mo : g.M;
  const int relu = p.relu, tf32 = p.tf32;
  const float4 *xp = reinterpret_cast<const float4 *>(p.x) + c4;
  const float4 *dp = reinterpret_cast<const float4 *>(p.dy) + c4;
  float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = mu, a = mu, b = mu;
  float4 s = mu, q = mu;
  if (active) {
    mu = ldp4(p.mean, c4, g.C, 0.f);
    rs = ldp4(p.rstd, c4, g.C, 0.f);
    coef_from_stats(p.gamma, p.beta, c4, g.C, mu, rs, a, b);
    if (!POOL) {
#pragma unroll 4
      for (long long r = r0; r < rows; r += stride) {
        const float4 v = pad0(__ldg(xp + r * cq), c4, g.C), d = pad0(__ldg(dp + r * cq), c4, g.C);
        float gx, gy, gz, gw, hx, hy, hz, hw;
        NA_GRAD_ELEM(gx, hx, v.x, d.x, a.x, b.x, mu.x, rs.x)
        NA_GRAD_ELEM(gy, hy, v.y, d.y, a.y, b.y, mu.y, rs.y)
        NA_GRAD_ELEM(gz, hz, v.z, d.z, a.z, b.z, mu.z, rs.z)
        NA_GRAD_ELEM(gw, hw, v.w, d.w, a.w, b.w, mu.w, rs.w)
        s.x += gx; s.y += gy; s.z += gz; s.w += gw;
        q.x = fmaf(gx, hx, q.x); q.y = fmaf(gy, hy, q.y); q.z = fmaf(gz, hz, q.z); q.w = fmaf(gw, hw, q.w);
      }
    } else {
#pragma unroll 2
      for (long long r = r0; r < rows; r += stride) {
        const float4 *w = xp + pool_base(pg, r, cq);
        const float4 v0 = pad0(__ldg(w), c4, g.C), v1 = pad0(__ldg(w + cq), c4, g.C), v2 = pad0(__ldg(w + down), c4, g.C),
                     v3 = pad0(__ldg(w + down + cq), c4, g.C);
        const float4 d = pad0(__ldg(dp + r * cq), c4, g.C);
        float gx, gy, gz, gw, hx, hy, hz, hw;
        int jx, jy, jz, jw;
        NA_POOL_ELEM(gx, jx, hx, v0.x, v1.x, v2.x, v3.x, d.x, a.x, b.x, mu.x, rs.x)
        NA_POOL_ELEM(gy, jy, hy, v0.y, v1.y, v2.y, v3.y, d.y, a.y, b.y, mu.y, rs.y)
        NA_POOL_ELEM(gz, jz, hz, v0.z, v1.z, v2.z, v3.z, d.z, a.z, b.z, mu.z, rs.z)
        NA_POOL_ELEM(gw, jw, hw, v0.w, v1.w, v2.w, v3.w, d.w, a.w, b.w, mu.w, rs.w)
        s.x += gx; s.y += gy; s.z += gz; s.w += gw;
        q.x = fmaf(gx, hx, q.x); q.y = fmaf(gy, hy, q.y); q.z = fmaf(gz, hz, q.z); q.w = fmaf(gw, hw, q.w);
      }
    }
  }
  nc_block_reduce(sh, L, s, q);
  nc_cluster_arrive();
  nc_cluster_wait();
  if (threadIdx.x < 4 * L) {
    const int c = blockIdx.y * L * 4 + threadIdx.x;
    float k1 = 0.f, k2 = 0.f;
    if (c < g.C) {
      double ds, dq;
      nc_cluster_total(sh, threadIdx.x, p.cs, ds, dq);
      if (blockIdx.x == 0) {
        if (p.dbeta) p.dbeta[c] = (float)ds;
        if (p.dgamma) p.dgamma[c] = (float)dq;
      }
      // statistics are constants in evaluation mode: dx = a * g
      k1 = p.training ? (float)(ds / (double)g.M) : 0.f;
      k2 = p.training ? (float)(dq / (double)g.M) : 0.f;
    }
    reinterpret_cast<float *>(&sh.coef[0][0])[threadIdx.x] = k1;
    reinterpret_cast<float *>(&sh.coef[1][0])[threadIdx.x] = k2;
  }
  nc_cluster_arrive();
  __syncthreads();
  if (active) {
    const float4 k1 = sh.coef[0][lane], k2 = sh.coef[1][lane];
    float4 *op = reinterpret_cast<float4 *>(p.dx) + c4;
    if (!POOL) {
#pragma unroll 4
      for (long long r = r0; r < rows; r += stride) {
        const float4 v = pad0(__ldg(xp + r * cq), c4, g.C), d = pad0(__ldg(dp + r * cq), c4, g.C);
        float gx, gy, gz, gw, hx, hy, hz, hw;
        NA_GRAD_ELEM(gx, hx, v.x, d.x, a.x, b.x, mu.x, rs.x)
        NA_GRAD_ELEM(gy, hy, v.y, d.y, a.y, b.y, mu.y, rs.y)
        NA_GRAD_ELEM(gz, hz, v.z, d.z, a.z, b.z, mu.z, rs.z)
        NA_GRAD_ELEM(gw, hw, v.w, d.w, a.w, b.w, mu.w, rs.w)
        op[r * cq] = na_out(make_float4(a.x * (gx - k1.x - hx * k2.x), a.y * (gy - k1.y - hy * k2.y),
                                        a.z * (gz - k1.z - hz * k2.z), a.w * (gw - k1.w - hw * k2.w)), tf32);
      }
    } else {
#pragma unroll 2
      for (long long r = r0; r < rows; r += stride) {
        const long long base = pool_base(pg, r, cq);
        const float4 *w = xp + base;
        const float4 v0 = pad0(__ldg(w), c4, g.C), v1 = pad0(__ldg(w + cq), c4, g.C), v2 = pad0(__ldg(w + down), c4, g.C),
                     v3 = pad0(__ldg(w + down + cq), c4, g.C);
        const float4 d = pad0(__ldg(dp + r * cq), c4, g.C);
        float4 o0, o1, o2, o3;
        NA_POOL_DX(o0.x, o1.x, o2.x, o3.x, v0.x, v1.x, v2.x, v3.x, d.x, a.x, b.x, mu.x, rs.x, k1.x, k2.x)
        NA_POOL_DX(o0.y, o1.y, o2.y, o3.y, v0.y, v1.y, v2.y, v3.y, d.y, a.y, b.y, mu.y, rs.y, k1.y, k2.y)
        NA_POOL_DX(o0.z, o1.z, o2.z, o3.z, v0.z, v1.z, v2.z, v3.z, d.z, a.z, b.z, mu.z, rs.z, k1.z, k2.z)
        NA_POOL_DX(o0.w, o1.w, o2.w, o3.w, v0.w, v1.w, v2.w, v3.w, d.w, a.w, b.w, mu.w, rs.w, k1.w, k2.w)
        float4 *o = op + base;
        o[0] = na_out(o0, tf32); o[cq] = na_out(o1, tf32); o[down] = na_out(o2, tf32); o[down + cq] = na_out(o3, tf32);
      }
    }
  }
  nc_cluster_wait();
}

// Used for activation tensors up to CPGB_BN_CLUSTER_MB (default 5 MB: the 4x4 and 2x2 maps of VGG16, 6 of its 13
// batch-norm layers; inside the training step the 8 MB tensors do better on three kernels: 1.499 ms per step with a
// 5 MB limit, 1.517 with 9 MB, 1.541 without the cluster kernels); CPGB_BN_CLUSTER=0 turns it off (both read per call, so that the tests compare the two paths in
// one process).  Measured on B200 inside CUDA graphs with x left in L2 by a producer (tools/bn_ab.py,
// gpurun_out/r2_bn_ab3.txt; clusters of 4, ~32 channel chunks), single launch vs stats -> finalize -> apply, forward /
// backward in us: 512 ch @ 2x2 (1 MB) 5.3 / 7.3 vs 7.2 / 9.7; 512 @ 4x4 (4 MB) 7.2 / 9.5 vs 9.2 / 17.1; 256 @ 8x8 (8 MB)
// 11.4 / 14.7 vs 10.7 / 18.4 -- these layers are launch-latency bound, two launches fewer is what pays.  Above that the
// three-kernel path wins and keeps the work: 128 @ 16x16 (17 MB) 30.6 / 42.8 vs 13.1 / 27.8, 64 @ 32x32 (33 MB) 58.6 /
// 99.5 vs 23.6 / 42.3: a CTA that owns 8 of 64 channels touches 32 bytes of every 256-byte pixel row (16 L1
// wavefronts per warp load), and 64-128 CTAs of 16 warps keep too few bytes in flight.  (A first version that left the
// finalize arithmetic to `lanes` threads per CTA lost everywhere: the serial chain cost more than two launches.)
bool nc_enabled(long long M, int Cs) {
  const char *e = getenv("CPGB_BN_CLUSTER");
  if (e && atoi(e) == 0) return false;
  const char *t = getenv("CPGB_BN_CLUSTER_MB");
  const double limit_mb = t ? atof(t) : 5.0;
  return (double)M * Cs * 4.0 / 1e6 <= limit_mb;
}

// ---- PReLU (models/spherenet.py:204-249: every SharableConv2d of SphereNet-20 feeds nn.PReLU(channels)) ----
// y = x > 0 ? x : alpha[c] * x;  dx = x > 0 ? dy : alpha[c] * dy;  dalpha[c] = sum over pixels of (x > 0 ? 0 : x * dy)
// (torch's prelu_kernel / prelu_backward_kernel).  Same streaming layout as the batch-norm kernels.
__global__ void __launch_bounds__(NA_THREADS)
prelu_fwd_kernel(const NaGeom g, const float *__restrict__ x, const float *__restrict__ alpha, int tf32,
                 float *__restrict__ y) {
  pdl_wait();
  const int lane = threadIdx.x % g.lanes, slot = threadIdx.x / g.lanes;
  const int c4 = blockIdx.y * g.lanes + lane;
  if (slot >= g.slots || c4 * 4 >= g.Cs) return;
  const float4 a = ldp4(alpha, c4, g.C, 0.f);
  const float4 *xp = reinterpret_cast<const float4 *>(x) + c4;
  float4 *yp = reinterpret_cast<float4 *>(y) + c4;
  const long long stride = (long long)gridDim.x * g.slots, cq = g.Cs / 4;
#pragma unroll 4
  for (long long r = (long long)blockIdx.x * g.slots + slot; r < g.M; r += stride) {
    const float4 v = pad0(__ldg(xp + r * cq), c4, g.C);
    yp[r * cq] = na_out(make_float4(v.x > 0.f ? v.x : a.x * v.x, v.y > 0.f ? v.y : a.y * v.y,
                                    v.z > 0.f ? v.z : a.z * v.z, v.w > 0.f ? v.w : a.w * v.w), tf32);
  }
}

__global__ void __launch_bounds__(NA_STATS_THREADS)
prelu_bwd_kernel(const NaGeom g, const float *__restrict__ x, const float *__restrict__ dy,
                 const float *__restrict__ alpha, int tf32, float *__restrict__ dx, float *__restrict__ part) {
  pdl_wait();
  const int lane = threadIdx.x % g.lanes, slot = threadIdx.x / g.lanes;
  const int c4 = blockIdx.y * g.lanes + lane;
  const bool active = slot < g.slots && c4 * 4 < g.Cs;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
    const float4 a = ldp4(alpha, c4, g.C, 0.f);
    const float4 *xp = reinterpret_cast<const float4 *>(x) + c4;
    const float4 *dp = reinterpret_cast<const float4 *>(dy) + c4;
    float4 *op = reinterpret_cast<float4 *>(dx) + c4;
    const long long stride = (long long)gridDim.x * g.slots, cq = g.Cs / 4;
#pragma unroll 4
    for (long long r = (long long)blockIdx.x * g.slots + slot; r < g.M; r += stride) {
      const float4 v = pad0(__ldg(xp + r * cq), c4, g.C), d = pad0(__ldg(dp + r * cq), c4, g.C);
      op[r * cq] = na_out(make_float4(v.x > 0.f ? d.x : a.x * d.x, v.y > 0.f ? d.y : a.y * d.y,
                                      v.z > 0.f ? d.z : a.z * d.z, v.w > 0.f ? d.w : a.w * d.w), tf32);
      s.x += v.x > 0.f ? 0.f : v.x * d.x; s.y += v.y > 0.f ? 0.f : v.y * d.y;
      s.z += v.z > 0.f ? 0.f : v.z * d.z; s.w += v.w > 0.f ? 0.f : v.w * d.w;
    }
  }
  block_reduce_pairs(g, s, make_float4(0.f, 0.f, 0.f, 0.f), lane, slot, c4, active, part);
}

__global__ void __launch_bounds__(256)
prelu_bwd_finalize_kernel(const float *__restrict__ part, int nblocks, int C, int Cs, float *__restrict__ dalpha) {
  pdl_wait();
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  double s, q;
  sum_partials(part, nblocks, Cs, c, s, q);
  if ((threadIdx.x & 31) == 0) dalpha[c] = (float)s;
}

bool na_args_ok(const void *x, long long M, int C, int Cs) {
  return M > 0 && C > 0 && Cs >= C && Cs % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
}

}  // namespace

// scratch layout: [coef_a Cs][coef_b Cs][c1 Cs][c2 Cs][partials blocks*Cs*2]
static size_t na_ws_bytes(long long M, int Cs) {
  NaGeom g = na_geom(M, Cs, Cs, NA_STATS_THREADS);
  return ((size_t)4 * Cs + (size_t)na_blocks(g, NA_STATS_PER_SM) * Cs * 2) * sizeof(float) + 64;
}

}  // namespace cpgb

using namespace cpgb;

extern "C" {

size_t cpgb_bn_workspace_bytes(int64_t M, int32_t C) {
  if (M <= 0 || C <= 0) return 0;
  return na_ws_bytes(M, (C + 3) & ~3);
}

static bool pool_geom(int64_t M, int32_t pool_h, int32_t pool_w, PoolGeom *pg) {
  if (pool_h <= 0 || pool_w <= 0 || (pool_h & 1) || (pool_w & 1) || M % ((int64_t)pool_h * pool_w)) return false;
  pg->H = pool_h; pg->W = pool_w; pg->Ho = pool_h / 2; pg->Wo = pool_w / 2; pg->Mo = M / 4;
  return true;
}

// ldc: floats between consecutive pixels of x / y / dy / dx; 0 means dense (= C, which must then be a multiple of 4)
static int na_stride(int32_t C, int32_t ldc) { return ldc > 0 ? ldc : C; }

int cpgb_bn_relu_fwd(const float *x, int64_t M, int32_t C, int32_t ldc, const float *gamma, const float *beta,
                     float *running_mean, float *running_var, int64_t *num_batches_tracked, int32_t training,
                     float momentum, float eps, int32_t relu, int32_t pool_h, int32_t pool_w, int32_t tf32_out, float *y,
                     float *save_mean, float *save_rstd, void *ws, size_t ws_bytes, void *stream) {
  return cpgb_bn_relu_fwd_stats(x, M, C, ldc, nullptr, 0, gamma, beta, running_mean, running_var, num_batches_tracked,
                                training, momentum, eps, relu, pool_h, pool_w, tf32_out, y, save_mean, save_rstd, ws,
                                ws_bytes, stream);
}

int cpgb_bn_relu_fwd_stats(const float *x, int64_t M, int32_t C, int32_t ldc, const float *colstats, int32_t nparts,
                           const float *gamma, const float *beta, float *running_mean, float *running_var,
                           int64_t *num_batches_tracked, int32_t training, float momentum, float eps, int32_t relu,
                           int32_t pool_h, int32_t pool_w, int32_t tf32_out, float *y, float *save_mean, float *save_rstd,
                           void *ws, size_t ws_bytes, void *stream) {
  PoolGeom pg;
  const bool pool = pool_h != 0 || pool_w != 0;
  const int Cs = na_stride(C, ldc);
  if (pool && !pool_geom(M, pool_h, pool_w, &pg)) {
    set_error("cpgb_bn_relu_fwd: 2x2 pooling needs even H, W with M = N*H*W"); return CPGB_EINVAL;
  }
  if (!na_args_ok(x, M, C, Cs) || !y || (reinterpret_cast<uintptr_t>(y) & 15)) {
    set_error("cpgb_bn_relu_fwd: needs NHWC fp32 with a pixel stride that is a multiple of 4 and 16-byte aligned x / y");
    return CPGB_EINVAL;
  }
  if (!ws || ws_bytes < na_ws_bytes(M, Cs) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
    set_error("cpgb_bn_relu_fwd: workspace %zu < %zu", ws_bytes, na_ws_bytes(M, Cs)); return CPGB_EWORKSPACE;
  }
  if (training ? (!save_mean || !save_rstd) : (!running_mean || !running_var)) {
    set_error("cpgb_bn_relu_fwd: missing statistics buffers"); return CPGB_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // per-tile column statistics from the convolution that produced x (cpgb_conv2d_fprop_stats): the statistics pass over
  // x is not needed, finalize sums the producer's partial pairs instead (same [part][Cs][2] layout).  Tensors small
  // enough for the single-launch kernels keep using those: one launch beats finalize + apply.
  const bool have_stats = training && colstats != nullptr && nparts > 0 && !nc_enabled(M, Cs);
  if (nc_enabled(M, Cs)) {
    // one launch: clusters of CTAs split the pixel rows of a channel chunk and exchange their partial sums through
    // distributed shared memory
    const NcPlan pl = nc_plan(M, Cs);
    NaGeom gc;
    gc.M = M; gc.C = C; gc.Cs = Cs; gc.lanes = pl.lanes; gc.slots = pl.slots; gc.cchunks = pl.chunks;
    NcFwdArgs a;
    a.x = x; a.gamma = gamma; a.beta = beta; a.running_mean = running_mean; a.running_var = running_var;
    a.nbt = training ? reinterpret_cast<long long *>(num_batches_tracked) : nullptr;
    a.y = y; a.save_mean = save_mean; a.save_rstd = save_rstd; a.momentum = momentum; a.eps = eps;
    a.training = training ? 1 : 0; a.relu = relu; a.tf32 = tf32_out; a.cs = pl.cs;
    if (!pool) { pg.H = pg.W = pg.Ho = pg.Wo = 0; pg.Mo = 0; }
    const dim3 grid(pl.cs, pl.chunks);
    if (pool) CPGB_CUDA_OK(launch_dependent_cluster(bn_fwd_cluster_kernel<true>, grid, dim3(NC_THREADS), 0, st, pl.cs, gc, pg, a));
    else CPGB_CUDA_OK(launch_dependent_cluster(bn_fwd_cluster_kernel<false>, grid, dim3(NC_THREADS), 0, st, pl.cs, gc, pg, a));
    CPGB_LAUNCH_OK("bn_fwd_cluster");
    return CPGB_OK;
  }
  NaGeom g = na_geom(M, C, Cs);
  float *coef_a = reinterpret_cast<float *>(ws), *coef_b = coef_a + Cs, *part = coef_a + 4 * Cs;
  if (training) {
    const NaGeom gs = na_geom(M, C, Cs, NA_STATS_THREADS);
    int nb = na_blocks(gs, NA_STATS_PER_SM);
    const float *fin_part = part;
    if (have_stats) {
      fin_part = colstats; nb = nparts;
    } else {
      CPGB_CUDA_OK(launch_dependent(bn_stats_kernel, dim3(dim3(nb, gs.cchunks)), dim3(NA_STATS_THREADS), 0, st, gs, x, part));
      CPGB_LAUNCH_OK("bn_stats");
    }
    CPGB_CUDA_OK(launch_dependent(bn_finalize_kernel, dim3((Cs + 7) / 8), dim3(256), 0, st, fin_part, nb, C, Cs, M, gamma, beta, running_mean, running_var, momentum,
                                                         eps, save_mean, save_rstd, coef_a, coef_b,
                                                         reinterpret_cast<long long *>(num_batches_tracked)));
    CPGB_LAUNCH_OK("bn_finalize");
  } else {
    CPGB_CUDA_OK(launch_dependent(bn_eval_coef_kernel, dim3((Cs + 127) / 128), dim3(128), 0, st, C, Cs, gamma, beta, running_mean, running_var, eps, coef_a,
                                                          coef_b));
    CPGB_LAUNCH_OK("bn_eval_coef");
  }
  if (pool) {
    NaGeom gp = g;
    gp.M = pg.Mo;
    CPGB_CUDA_OK(launch_dependent(bn_apply_pool_kernel, dim3(dim3(na_blocks(gp, 8), g.cchunks)), dim3(NA_THREADS), 0, st, g, pg, x, coef_a, coef_b, relu, tf32_out, y));
  } else {
    CPGB_CUDA_OK(launch_dependent(bn_apply_kernel, dim3(dim3(na_blocks(g, 8), g.cchunks)), dim3(NA_THREADS), 0, st, g, x, coef_a, coef_b, relu, tf32_out, y));
  }
  CPGB_LAUNCH_OK("bn_apply");
  return CPGB_OK;
}

int cpgb_bn_relu_bwd(const float *x, const float *dy, int64_t M, int32_t C, int32_t ldc, const float *gamma,
                     const float *beta, const float *mean, const float *rstd, int32_t training, int32_t relu,
                     int32_t pool_h, int32_t pool_w, int32_t tf32_out, float *dx, float *dgamma, float *dbeta, void *ws,
                     size_t ws_bytes, void *stream) {
  PoolGeom pg;
  const bool pool = pool_h != 0 || pool_w != 0;
  const int Cs = na_stride(C, ldc);
  if (pool && !pool_geom(M, pool_h, pool_w, &pg)) {
    set_error("cpgb_bn_relu_bwd: 2x2 pooling needs even H, W with M = N*H*W"); return CPGB_EINVAL;
  }
  if (!na_args_ok(x, M, C, Cs) || !dy || !dx || !mean || !rstd || (reinterpret_cast<uintptr_t>(dy) & 15) ||
      (reinterpret_cast<uintptr_t>(dx) & 15)) {
    set_error("cpgb_bn_relu_bwd: needs NHWC fp32 with a pixel stride that is a multiple of 4 and 16-byte aligned tensors");
    return CPGB_EINVAL;
  }
  if (!ws || ws_bytes < na_ws_bytes(M, Cs) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
    set_error("cpgb_bn_relu_bwd: workspace %zu < %zu", ws_bytes, na_ws_bytes(M, Cs)); return CPGB_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (nc_enabled(M, Cs)) {
    const NcPlan pl = nc_plan(pool ? pg.Mo : M, Cs);
    NaGeom gc;
    gc.M = M; gc.C = C; gc.Cs = Cs; gc.lanes = pl.lanes; gc.slots = pl.slots; gc.cchunks = pl.chunks;
    NcBwdArgs a;
    a.x = x; a.dy = dy; a.gamma = gamma; a.beta = beta; a.mean = mean; a.rstd = rstd;
    a.dx = dx; a.dgamma = dgamma; a.dbeta = dbeta; a.training = training ? 1 : 0; a.relu = relu; a.tf32 = tf32_out;
    a.cs = pl.cs;
    if (!pool) { pg.H = pg.W = pg.Ho = pg.Wo = 0; pg.Mo = 0; }
    const dim3 grid(pl.cs, pl.chunks);
    if (pool) CPGB_CUDA_OK(launch_dependent_cluster(bn_bwd_cluster_kernel<true>, grid, dim3(NC_THREADS), 0, st, pl.cs, gc, pg, a));
    else CPGB_CUDA_OK(launch_dependent_cluster(bn_bwd_cluster_kernel<false>, grid, dim3(NC_THREADS), 0, st, pl.cs, gc, pg, a));
    CPGB_LAUNCH_OK("bn_bwd_cluster");
    return CPGB_OK;
  }
  NaGeom g = na_geom(M, C, Cs);
  float *c1 = reinterpret_cast<float *>(ws) + 2 * Cs, *c2 = c1 + Cs, *part = c2 + Cs;
  const NaGeom gs = na_geom(M, C, Cs, NA_STATS_THREADS);
  int nb;
  if (pool) {
    NaGeom gp = gs;
    gp.M = pg.Mo;
    nb = na_blocks(gp, NA_STATS_PER_SM);
    CPGB_CUDA_OK(launch_dependent(bn_bwd_stats_pool_kernel, dim3(dim3(nb, gs.cchunks)), dim3(NA_STATS_THREADS), 0, st, gs, pg, x, dy, gamma, beta, mean, rstd, relu,
                                                                               part));
  } else {
    nb = na_blocks(gs, NA_STATS_PER_SM);
    CPGB_CUDA_OK(launch_dependent(bn_bwd_stats_kernel, dim3(dim3(nb, gs.cchunks)), dim3(NA_STATS_THREADS), 0, st, gs, x, dy, gamma, beta, mean, rstd, relu, part));
  }
  CPGB_LAUNCH_OK("bn_bwd_stats");
  CPGB_CUDA_OK(launch_dependent(bn_bwd_finalize_kernel, dim3((Cs + 7) / 8), dim3(256), 0, st, part, nb, C, Cs, M, training, dgamma, dbeta, c1, c2));
  CPGB_LAUNCH_OK("bn_bwd_finalize");
  if (pool) {
    NaGeom gp = g;
    gp.M = pg.Mo;
    CPGB_CUDA_OK(launch_dependent(bn_bwd_apply_pool_kernel, dim3(dim3(na_blocks(gp, 8), g.cchunks)), dim3(NA_THREADS), 0, st, g, pg, x, dy, gamma, beta, mean, rstd,
                                                                                      c1, c2, relu, tf32_out, dx));
  } else {
    CPGB_CUDA_OK(launch_dependent(bn_bwd_apply_kernel, dim3(dim3(na_blocks(g, 8), g.cchunks)), dim3(NA_THREADS), 0, st, g, x, dy, gamma, beta, mean, rstd, c1, c2,
                                                                                relu, tf32_out, dx));
  }
  CPGB_LAUNCH_OK("bn_bwd_apply");
  return CPGB_OK;
}

int cpgb_bn_add_relu_fwd(const float *x, const float *res, int64_t M, int32_t C, int32_t ldc, const float *gamma,
                         const float *beta, float *running_mean, float *running_var, int64_t *num_batches_tracked,
                         int32_t training, float momentum, float eps, int32_t tf32_out, float *y, float *save_mean,
                         float *save_rstd, void *ws, size_t ws_bytes, void *stream) {
  const int Cs = na_stride(C, ldc);
  if (!na_args_ok(x, M, C, Cs) || !y || !res || (reinterpret_cast<uintptr_t>(y) & 15) ||
      (reinterpret_cast<uintptr_t>(res) & 15)) {
    set_error("cpgb_bn_add_relu_fwd: needs NHWC fp32 with a pixel stride that is a multiple of 4 and 16-byte aligned x / res / y");
    return CPGB_EINVAL;
  }
  if (!ws || ws_bytes < na_ws_bytes(M, Cs) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
    set_error("cpgb_bn_add_relu_fwd: workspace %zu < %zu", ws_bytes, na_ws_bytes(M, Cs)); return CPGB_EWORKSPACE;
  }
  if (training ? (!save_mean || !save_rstd) : (!running_mean || !running_var)) {
    set_error("cpgb_bn_add_relu_fwd: missing statistics buffers"); return CPGB_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  NaGeom g = na_geom(M, C, Cs);
  float *coef_a = reinterpret_cast<float *>(ws), *coef_b = coef_a + Cs, *part = coef_a + 4 * Cs;
  if (training) {
    const NaGeom gs = na_geom(M, C, Cs, NA_STATS_THREADS);
    const int nb = na_blocks(gs, NA_STATS_PER_SM);
    CPGB_CUDA_OK(launch_dependent(bn_stats_kernel, dim3(nb, gs.cchunks), dim3(NA_STATS_THREADS), 0, st, gs, x, part));
    CPGB_LAUNCH_OK("bn_stats");
    CPGB_CUDA_OK(launch_dependent(bn_finalize_kernel, dim3((Cs + 7) / 8), dim3(256), 0, st, (const float *)part, nb, C, Cs,
                                  (long long)M, gamma, beta, running_mean, running_var, momentum, eps, save_mean, save_rstd,
                                  coef_a, coef_b, reinterpret_cast<long long *>(num_batches_tracked)));
    CPGB_LAUNCH_OK("bn_finalize");
  } else {
    CPGB_CUDA_OK(launch_dependent(bn_eval_coef_kernel, dim3((Cs + 127) / 128), dim3(128), 0, st, C, Cs, gamma, beta,
                                  (const float *)running_mean, (const float *)running_var, eps, coef_a, coef_b));
    CPGB_LAUNCH_OK("bn_eval_coef");
  }
  CPGB_CUDA_OK(launch_dependent(bn_apply_res_kernel, dim3(na_blocks(g, 8), g.cchunks), dim3(NA_THREADS), 0, st, g, x, res,
                                (const float *)coef_a, (const float *)coef_b, tf32_out, y));
  CPGB_LAUNCH_OK("bn_apply_res");
  return CPGB_OK;
}

int cpgb_bn_add_relu_bwd(const float *x, const float *y, const float *dy, int64_t M, int32_t C, int32_t ldc,
                         const float *gamma, const float *beta, const float *mean, const float *rstd, int32_t training,
                         int32_t tf32_out, float *dx, float *dres, float *dgamma, float *dbeta, void *ws, size_t ws_bytes,
                         void *stream) {
  const int Cs = na_stride(C, ldc);
  if (!na_args_ok(x, M, C, Cs) || !y || !dy || !dx || !dres || !mean || !rstd || (reinterpret_cast<uintptr_t>(y) & 15) ||
      (reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(dx) & 15) ||
      (reinterpret_cast<uintptr_t>(dres) & 15)) {
    set_error("cpgb_bn_add_relu_bwd: needs NHWC fp32 with a pixel stride that is a multiple of 4 and 16-byte aligned tensors");
    return CPGB_EINVAL;
  }
  if (!ws || ws_bytes < na_ws_bytes(M, Cs) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
    set_error("cpgb_bn_add_relu_bwd: workspace %zu < %zu", ws_bytes, na_ws_bytes(M, Cs)); return CPGB_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  NaGeom g = na_geom(M, C, Cs);
  float *c1 = reinterpret_cast<float *>(ws) + 2 * Cs, *c2 = c1 + Cs, *part = c2 + Cs;
  const NaGeom gs = na_geom(M, C, Cs, NA_STATS_THREADS);
  const int nb = na_blocks(gs, NA_STATS_PER_SM);
  CPGB_CUDA_OK(launch_dependent(bn_bwd_stats_res_kernel, dim3(nb, gs.cchunks), dim3(NA_STATS_THREADS), 0, st, gs, x, y, dy,
                                mean, rstd, dres, part));
  CPGB_LAUNCH_OK("bn_bwd_stats_res");
  CPGB_CUDA_OK(launch_dependent(bn_bwd_finalize_kernel, dim3((Cs + 7) / 8), dim3(256), 0, st, (const float *)part, nb, C, Cs,
                                (long long)M, (int)training, dgamma, dbeta, c1, c2));
  CPGB_LAUNCH_OK("bn_bwd_finalize");
  // dx = a * (g - c1 - xhat * c2) from the gated gradient just written: the plain batch-norm backward, no ReLU gate
  CPGB_CUDA_OK(launch_dependent(bn_bwd_apply_kernel, dim3(na_blocks(g, 8), g.cchunks), dim3(NA_THREADS), 0, st, g, x,
                                (const float *)dres, gamma, beta, mean, rstd, (const float *)c1, (const float *)c2, 0,
                                tf32_out, dx));
  CPGB_LAUNCH_OK("bn_bwd_apply");
  return CPGB_OK;
}

size_t cpgb_prelu_workspace_bytes(int64_t M, int32_t C) {
  if (M <= 0 || C <= 0) return 0;
  return na_ws_bytes(M, (C + 3) & ~3);
}

int cpgb_prelu_fwd(const float *x, int64_t M, int32_t C, int32_t ldc, const float *alpha, int32_t tf32_out, float *y,
                   void *stream) {
  const int Cs = na_stride(C, ldc);
  if (!na_args_ok(x, M, C, Cs) || !y || !alpha || (reinterpret_cast<uintptr_t>(y) & 15)) {
    set_error("cpgb_prelu_fwd: needs NHWC fp32 with a pixel stride that is a multiple of 4, 16-byte aligned x / y, and alpha[C]");
    return CPGB_EINVAL;
  }
  NaGeom g = na_geom(M, C, Cs);
  CPGB_CUDA_OK(launch_dependent(prelu_fwd_kernel, dim3(na_blocks(g, 8), g.cchunks), dim3(NA_THREADS), 0,
                                (cudaStream_t)stream, g, x, alpha, tf32_out, y));
  CPGB_LAUNCH_OK("prelu_fwd");
  return CPGB_OK;
}

int cpgb_prelu_bwd(const float *x, const float *dy, int64_t M, int32_t C, int32_t ldc, const float *alpha,
                   int32_t tf32_out, float *dx, float *dalpha, void *ws, size_t ws_bytes, void *stream) {
  const int Cs = na_stride(C, ldc);
  if (!na_args_ok(x, M, C, Cs) || !dy || !dx || !alpha || !dalpha || (reinterpret_cast<uintptr_t>(dy) & 15) ||
      (reinterpret_cast<uintptr_t>(dx) & 15)) {
    set_error("cpgb_prelu_bwd: needs NHWC fp32 with a pixel stride that is a multiple of 4 and 16-byte aligned tensors");
    return CPGB_EINVAL;
  }
  if (!ws || ws_bytes < na_ws_bytes(M, Cs) || (reinterpret_cast<uintptr_t>(ws) & 15)) {
    set_error("cpgb_prelu_bwd: workspace %zu < %zu", ws_bytes, na_ws_bytes(M, Cs)); return CPGB_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const NaGeom gs = na_geom(M, C, Cs, NA_STATS_THREADS);
  const int nb = na_blocks(gs, NA_STATS_PER_SM);
  float *part = reinterpret_cast<float *>(ws) + 4 * Cs;
  CPGB_CUDA_OK(launch_dependent(prelu_bwd_kernel, dim3(nb, gs.cchunks), dim3(NA_STATS_THREADS), 0, st, gs, x, dy, alpha,
                                tf32_out, dx, part));
  CPGB_LAUNCH_OK("prelu_bwd");
  CPGB_CUDA_OK(launch_dependent(prelu_bwd_finalize_kernel, dim3((C + 7) / 8), dim3(256), 0, st, part, nb, C, Cs, dalpha));
  CPGB_LAUNCH_OK("prelu_bwd_finalize");
  return CPGB_OK;
}

}  // extern "C"
