// CUDA-core (fp32 FFMA) implicit-GEMM kernels for the masked convolution: any stride /
// padding / dilation / groups / memory format.  They are the exact-fp32 path for shapes the
// tcgen05 kernels do not cover (C=3 stems, odd grown widths, grouped convs) and the on-GPU
// cross-check for the tensor-core path.
//
// One generic 64x64x16 tile kernel; the three convolution passes differ only in how a GEMM
// coordinate maps to memory:
//   fprop : M = N*P*Q pixels,  cols = K/g,        Kd = C/g*R*S   (models/layers.py:108)
//   dgrad : M = N*H*W pixels,  cols = C/g,        Kd = K/g*R*S   (autograd of the above)
//   wgrad : M = K/g,           cols = C/g*R*S,    Kd = N*P*Q     (autograd of the above)
// The piggyback predicate (models/layers.py:101-103) is evaluated while the weight tile is
// loaded -- the masked weight is never written to memory.
#include "common.cuh"

namespace cpgb {

constexpr int BM = 64, BN = 64, BK = 16, NTHREADS = 256;

struct FpropProb {
  Geom g; const float *x, *w, *piggy, *bias; float *y; float thr;
  static constexpr bool A_M_FAST = true, B_N_FAST = false;
  __device__ int M() const { return g.N * g.PQ; }
  __device__ int Ncols() const { return g.Kg; }
  __device__ int Kd() const { return g.Cg * g.RS; }
  __device__ float loadA(int grp, int m, int kd) const {
    int n = m / g.PQ, pq = m - n * g.PQ, p = pq / g.Q, q = pq - p * g.Q;
    int c = kd / g.RS, rs = kd - c * g.RS, r = rs / g.S, s = rs - r * g.S;
    int h = p * g.sh - g.ph + r * g.dh, ww = q * g.sw - g.pw + s * g.dw;
    if ((unsigned)h >= (unsigned)g.H || (unsigned)ww >= (unsigned)g.W) return 0.f;
    return __ldg(x + n * g.xs0 + (long long)(grp * g.Cg + c) * g.xs1 + h * g.xs2 + ww * g.xs3);
  }
  __device__ float loadB(int grp, int kd, int j) const {
    long long idx = (long long)(grp * g.Kg + j) * g.Cg * g.RS + kd;
    return masked_weight(__ldg(w + idx), piggy, idx, thr);
  }
  __device__ void store(int grp, int m, int j, float acc, int) const {
    int n = m / g.PQ, pq = m - n * g.PQ, p = pq / g.Q, q = pq - p * g.Q;
    int k = grp * g.Kg + j;
    y[n * g.ys0 + k * g.ys1 + p * g.ys2 + q * g.ys3] = acc + (bias ? __ldg(bias + k) : 0.f);
  }
};

struct DgradProb {
  Geom g; const float *dy, *w, *piggy; float *dx; float thr;
  static constexpr bool A_M_FAST = true, B_N_FAST = false;
  __device__ int M() const { return g.N * g.HW; }
  __device__ int Ncols() const { return g.Cg; }
  __device__ int Kd() const { return g.Kg * g.RS; }
  __device__ float loadA(int grp, int m, int kd) const {
    int n = m / g.HW, hw = m - n * g.HW, h = hw / g.W, ww = hw - h * g.W;
    int kk = kd / g.RS, rs = kd - kk * g.RS, r = rs / g.S, s = rs - r * g.S;
    int hp = h + g.ph - r * g.dh, wq = ww + g.pw - s * g.dw;
    if (hp < 0 || wq < 0) return 0.f;
    int p = hp / g.sh, q = wq / g.sw;
    if (p * g.sh != hp || q * g.sw != wq || p >= g.P || q >= g.Q) return 0.f;
    return __ldg(dy + n * g.ys0 + (long long)(grp * g.Kg + kk) * g.ys1 + p * g.ys2 + q * g.ys3);
  }
  __device__ float loadB(int grp, int kd, int c) const {
    int kk = kd / g.RS, rs = kd - kk * g.RS;
    long long idx = ((long long)(grp * g.Kg + kk) * g.Cg + c) * g.RS + rs;
    return masked_weight(__ldg(w + idx), piggy, idx, thr);
  }
  __device__ void store(int grp, int m, int c, float acc, int) const {
    int n = m / g.HW, hw = m - n * g.HW, h = hw / g.W, ww = hw - h * g.W;
    dx[n * g.xs0 + (long long)(grp * g.Cg + c) * g.xs1 + h * g.xs2 + ww * g.xs3] = acc;
  }
};

struct WgradProb {
  Geom g; const float *x, *dy; float *gbuf; long long split_stride;   // gbuf[split][K*Cg*R*S]
  static constexpr bool A_M_FAST = false, B_N_FAST = false;
  __device__ int M() const { return g.Kg; }
  __device__ int Ncols() const { return g.Cg * g.RS; }
  __device__ int Kd() const { return g.N * g.PQ; }
  __device__ float loadA(int grp, int j, int pix) const {
    int n = pix / g.PQ, pq = pix - n * g.PQ, p = pq / g.Q, q = pq - p * g.Q;
    return __ldg(dy + n * g.ys0 + (long long)(grp * g.Kg + j) * g.ys1 + p * g.ys2 + q * g.ys3);
  }
  __device__ float loadB(int grp, int pix, int crs) const {
    int n = pix / g.PQ, pq = pix - n * g.PQ, p = pq / g.Q, q = pq - p * g.Q;
    int c = crs / g.RS, rs = crs - c * g.RS, r = rs / g.S, s = rs - r * g.S;
    int h = p * g.sh - g.ph + r * g.dh, ww = q * g.sw - g.pw + s * g.dw;
    if ((unsigned)h >= (unsigned)g.H || (unsigned)ww >= (unsigned)g.W) return 0.f;
    return __ldg(x + n * g.xs0 + (long long)(grp * g.Cg + c) * g.xs1 + h * g.xs2 + ww * g.xs3);
  }
  __device__ void store(int grp, int j, int crs, float acc, int split) const {
    // every split owns a full partial tensor: the reduction order is fixed (summed by the epilogue)
    gbuf[split * split_stride + (long long)(grp * g.Kg + j) * g.Cg * g.RS + crs] = acc;
  }
};

// grid: x = M tiles, y = column tiles, z = group * splits + split
template <class Prob>
__global__ void __launch_bounds__(NTHREADS) simt_gemm_kernel(const Prob pb, int splits, int kchunk) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int grp = blockIdx.z / splits, split = blockIdx.z - grp * splits;
  const int M = pb.M(), NC = pb.Ncols(), KD = pb.Kd();
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = split * kchunk;
  const int kend = min(KD, kbeg + kchunk);
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * NTHREADS;
      int mm, kk;
      if (Prob::A_M_FAST) { mm = e & (BM - 1); kk = e >> 6; } else { kk = e & (BK - 1); mm = e >> 4; }
      float v = 0.f;
      if (m0 + mm < M && k0 + kk < kend) v = pb.loadA(grp, m0 + mm, k0 + kk);
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * NTHREADS;
      int nn, kk;
      if (Prob::B_N_FAST) { nn = e & (BN - 1); kk = e >> 6; } else { kk = e & (BK - 1); nn = e >> 4; }
      float v = 0.f;
      if (n0 + nn < NC && k0 + kk < kend) v = pb.loadB(grp, k0 + kk, n0 + nn);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < NC) pb.store(grp, m, n, acc[i][j], split);
    }
  }
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int simt_fprop(const Geom &g, const float *x, const float *w, const float *piggy, const float *bias,
               float *y, float thr, cudaStream_t st) {
  FpropProb pb{g, x, w, piggy, bias, y, thr};
  dim3 grid(cdiv((long long)g.N * g.PQ, BM), cdiv(g.Kg, BN), g.groups);
  simt_gemm_kernel<FpropProb><<<grid, NTHREADS, 0, st>>>(pb, 1, g.Cg * g.RS);
  CPGB_LAUNCH_OK("simt_fprop");
  return CPGB_OK;
}

int simt_dgrad(const Geom &g, const float *dy, const float *w, const float *piggy, float *dx, float thr,
               cudaStream_t st) {
  DgradProb pb{g, dy, w, piggy, dx, thr};
  dim3 grid(cdiv((long long)g.N * g.HW, BM), cdiv(g.Cg, BN), g.groups);
  simt_gemm_kernel<DgradProb><<<grid, NTHREADS, 0, st>>>(pb, 1, g.Kg * g.RS);
  CPGB_LAUNCH_OK("simt_dgrad");
  return CPGB_OK;
}

static void simt_wgrad_plan(const Geom &g, int *splits_out, int *kchunk_out) {
  const long long kd = (long long)g.N * g.PQ;
  const int tiles = cdiv(g.Kg, BM) * cdiv((long long)g.Cg * g.RS, BN) * g.groups;
  // split the pixel reduction so that ~2 waves of CTAs exist (148 SMs)
  int splits = 1;
  if (tiles < 296) splits = (int)min((long long)cdiv(296, tiles), (kd + 4 * BK - 1) / (4 * BK));
  if (splits < 1) splits = 1;
  int kchunk = (int)((kd + splits - 1) / splits);
  kchunk = ((kchunk + BK - 1) / BK) * BK;
  *splits_out = (int)((kd + kchunk - 1) / kchunk);
  *kchunk_out = kchunk;
}

int simt_wgrad_splits(const Geom &g) {
  int splits, kchunk;
  simt_wgrad_plan(g, &splits, &kchunk);
  return splits;
}

// raw weight-gradient partial sums gbuf[split][K*Cg*R*S]; returns the number of splits
int simt_wgrad_raw(const Geom &g, const float *x, const float *dy, float *gbuf, int *splits_out, cudaStream_t st) {
  int splits, kchunk;
  simt_wgrad_plan(g, &splits, &kchunk);
  WgradProb pb{g, x, dy, gbuf, (long long)g.K * g.Cg * g.RS};
  dim3 grid(cdiv(g.Kg, BM), cdiv((long long)g.Cg * g.RS, BN), g.groups * splits);
  simt_gemm_kernel<WgradProb><<<grid, NTHREADS, 0, st>>>(pb, splits, kchunk);
  CPGB_LAUNCH_OK("simt_wgrad");
  *splits_out = splits;
  return CPGB_OK;
}

// dbias[k] = sum over n,p,q of dy  (autograd of the bias add, models/layers.py:108)
__global__ void __launch_bounds__(256) bias_grad_kernel(Geom g, const float *__restrict__ dy,
                                                        float *__restrict__ dbias) {
  const int k = blockIdx.x;
  const long long total = (long long)g.N * g.PQ;
  float s = 0.f;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    int n = (int)(i / g.PQ), pq = (int)(i - (long long)n * g.PQ), p = pq / g.Q, q = pq - p * g.Q;
    s += __ldg(dy + n * g.ys0 + k * g.ys1 + p * g.ys2 + q * g.ys3);
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) dbias[k] = s;
  }
}

// NHWC activations (channel stride 1, one uniform pixel stride): dbias = column sums of the [M pixels][ps] matrix.
// The kernel above gives every channel its own block and walks the pixels with a stride of `ps` floats -- 4 useful
// bytes per 32-byte sector; on SphereNet-20 (20 biased convolutions, up to 51 MB of dy each) it was 37 % of the device
// time of a step.  Two deterministic phases instead: fat blocks stream whole pixel rows (a thread owns one float4 of
// channels), per-block partial sums go to scratch, one thread per channel adds them in block order.
constexpr int BG_THREADS = 512;
struct BgGeom { long long M; int K, ps, lanes, slots, cchunks; };

static bool bg_geom(const Geom &g, const float *dy, BgGeom *out) {
  if (g.ys1 != 1 || (reinterpret_cast<uintptr_t>(dy) & 15)) return false;
  long long ps;
  if (g.P == 1 && g.Q == 1) ps = g.N > 1 ? g.ys0 : (long long)((g.K + 3) & ~3);
  else {
    ps = g.Q > 1 ? g.ys3 : g.ys2;
    if (g.Q > 1 && g.P > 1 && g.ys2 != (long long)g.Q * ps) return false;
    if (g.N > 1 && g.ys0 != (long long)g.P * g.Q * ps) return false;
  }
  if (ps < g.K || ps % 4 != 0 || ps > (1 << 20)) return false;
  BgGeom b;
  b.M = (long long)g.N * g.P * g.Q; b.K = g.K; b.ps = (int)ps;
  const int c4 = (g.K + 3) / 4;
  b.lanes = c4 < BG_THREADS ? c4 : BG_THREADS;
  b.slots = BG_THREADS / b.lanes;
  b.cchunks = (c4 + b.lanes - 1) / b.lanes;
  *out = b;
  return true;
}
static int bg_blocks(const BgGeom &b) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // at least 8 pixel rows per thread: few partial sums for short matrices (the FC layers: 128 rows)
  long long want = (b.M + 8LL * b.slots - 1) / (8LL * b.slots), cap = (long long)sms * 2 / b.cchunks;
  if (cap < 1) cap = 1;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}
size_t bias_grad_scratch_bytes(const Geom &g) {
  BgGeom b;
  // alignment is a property of the pointer; the size only depends on the geometry
  if (!bg_geom(g, nullptr, &b)) return 0;
  return (size_t)bg_blocks(b) * b.cchunks * b.lanes * 4 * sizeof(float) + 256;
}

__global__ void __launch_bounds__(BG_THREADS)
bias_grad_nhwc_kernel(const BgGeom b, const float *__restrict__ dy, float *__restrict__ part) {
  __shared__ float4 red[BG_THREADS];
  const int lane = threadIdx.x % b.lanes, slot = threadIdx.x / b.lanes;
  const int c4 = blockIdx.y * b.lanes + lane;
  const bool active = slot < b.slots && c4 * 4 < b.K;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
    const float4 *p = reinterpret_cast<const float4 *>(dy) + c4;
    const long long stride = (long long)gridDim.x * b.slots, cq = b.ps / 4;
    const int c = c4 * 4;
#pragma unroll 4
    for (long long r = (long long)blockIdx.x * b.slots + slot; r < b.M; r += stride) {
      const float4 v = __ldg(p + r * cq);
      // lanes beyond K inside the last group are padding of the activation layout: not part of any channel
      s.x += v.x; s.y += c + 1 < b.K ? v.y : 0.f; s.z += c + 2 < b.K ? v.z : 0.f; s.w += c + 3 < b.K ? v.w : 0.f;
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  if (slot == 0 && active) {
    for (int k = 1; k < b.slots; ++k) {
      const float4 a = red[k * b.lanes + lane];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
    const long long cols = (long long)b.cchunks * b.lanes * 4;
    *reinterpret_cast<float4 *>(part + (long long)blockIdx.x * cols + c4 * 4) = s;
  }
}
// 32 channels per block, 8 thread groups: group j adds the partials of blocks j, j + 8, ... (four loads in flight),
// the groups are then combined in order -- fixed summation order, double precision
__global__ void __launch_bounds__(256)
bias_grad_finish_kernel(const float *__restrict__ part, int nblk, int K, long long cols, float *__restrict__ dbias) {
  __shared__ double red[8][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (k < K) {
    const float *p = part + k;
    int b = grp;
    for (; b + 24 < nblk; b += 32) {
      const float v0 = __ldg(p + (long long)b * cols), v1 = __ldg(p + (long long)(b + 8) * cols);
      const float v2 = __ldg(p + (long long)(b + 16) * cols), v3 = __ldg(p + (long long)(b + 24) * cols);
      s += (double)v0; s += (double)v1; s += (double)v2; s += (double)v3;
    }
    for (; b < nblk; b += 8) s += (double)__ldg(p + (long long)b * cols);
  }
  red[grp][lane] = s;
  __syncthreads();
  if (grp == 0 && k < K) {
    for (int j = 1; j < 8; ++j) s += red[j][lane];
    dbias[k] = (float)s;
  }
}

int bias_grad(const Geom &g, const float *dy, float *dbias, void *scratch, size_t scratch_bytes, cudaStream_t st) {
  BgGeom b;
  if (scratch && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0 && bg_geom(g, dy, &b)) {
    const int nblk = bg_blocks(b);
    const long long cols = (long long)b.cchunks * b.lanes * 4;
    if ((size_t)nblk * cols * sizeof(float) <= scratch_bytes) {
      float *part = reinterpret_cast<float *>(scratch);
      bias_grad_nhwc_kernel<<<dim3(nblk, b.cchunks), BG_THREADS, 0, st>>>(b, dy, part);
      CPGB_LAUNCH_OK("bias_grad_nhwc");
      bias_grad_finish_kernel<<<(g.K + 31) / 32, 256, 0, st>>>(part, nblk, g.K, cols, dbias);
      CPGB_LAUNCH_OK("bias_grad_finish");
      return CPGB_OK;
    }
  }
  bias_grad_kernel<<<g.K, 256, 0, st>>>(g, dy, dbias);
  CPGB_LAUNCH_OK("bias_grad");
  return CPGB_OK;
}

}  // namespace cpgb
