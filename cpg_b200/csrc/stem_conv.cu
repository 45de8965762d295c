// Direct fp32 kernels for the few-input-channel stem of a masked network (models/vgg.py:97 /
// models/spherenet.py:201: SharableConv2d(3, 64, kernel_size=3)).
//
// With C*R*S = 27 the layer is not a GEMM worth staging for the tensor core: a 32-wide reduction
// block would be 16 % padding, the operand would need an explicit im2col (a second pass over the
// activations) and the whole layer is 0.45 GFLOP against 33.5 MB of output at batch 128 -- it is
// HBM bound by the write of Y (fprop) and the read of dY (wgrad).  These kernels keep the 27 x 4
// weights (fprop) or 27 x 4 accumulators (wgrad) of one thread in registers, read the activations
// through L1 in whatever layout the caller has (element strides, no padded copy) and touch the
// NHWC output / output-gradient exactly once with 16-byte accesses:
//
//   fprop : y[n,p,q,k]  = bias[k] + sum_{c,r,s} x[n,c,p*sh-ph+r*dh,q*sw-pw+s*dw] * b(P[k,c,r,s]) * W[k,c,r,s]
//   wgrad : g[k,c,r,s]  = sum_{n,p,q} dy[n,p,q,k] * x[n,c,...]      then the fused epilogue (a4 + a6)
//
// Thread mapping: K/KPT lanes share one output pixel (KPT = 2 output channels per thread for K <= 64, 4 for
// K = 128; lane j owns channels KPT*j ..), so a pixel row of Y is one contiguous 4*K-byte store.  The common
// case (x as NHWC with a pixel stride of 4 floats, unit stride / dilation along w, K <= 64) runs the
// pixel-PAIR kernels: a thread owns two horizontally adjacent pixels that share a 3 x 4 window of 16-byte
// loads.  All arithmetic is exact fp32 on packed FFMA2 (fma.rn.f32x2) -- no TF32.  wgrad streams dY
// through shared memory with 1-D TMA bulk copies and is deterministic: per-block partial sums in a fixed
// order, then a fixed-shape tree.
#include "common.cuh"
#include "ptx.cuh"

namespace cpgb {

namespace {

constexpr int STEM_THREADS = 128;
constexpr int STEM_WG_THREADS = 256;

__device__ __forceinline__ void epi_stem(float g, float w, float p, bool has_p, unsigned t, int cur, float wd,
                                        int mode, float thr, float &dw, float &dp) {
  grad_epilogue_elem(g, w, p, has_p, t, cur, wd, mode, thr, dw, dp);
}

struct StemGeom {
  int N, H, W, K, P, Q;
  int sh, sw, ph, pw, dh, dw;
  int xs0, xs2, xs3;           // element strides of x (fit in int32: checked by stem_eligible)
  int toff[27];                // element offset of tap (c, r, s) from the window origin (h0, w0)
  long long ys0, ys2, ys3;     // element strides of y / dy (channel stride is 1)
  int pixels;                  // N * P * Q  (< 2^31: validate_desc)
  int lpp_log2;                // log2(K / KPT): lanes per pixel
  int xdense4;                 // x is dense NHWC4 (xs3 = 4, xs2 = 4W, xs0 = 4WH): rows of a block are contiguous
};

// Output-pixel cursor: (n, p, q) advanced by a fixed number of pixels without divisions.
struct PixCursor {
  int n, p, q;
  __device__ __forceinline__ void seek(const StemGeom &g, int pix) {
    q = pix % g.Q;
    const int t = pix / g.Q;
    p = t % g.P;
    n = t / g.P;
  }
  __device__ __forceinline__ void advance(const StemGeom &g, int by) {
    q += by;
    while (q >= g.Q) {
      q -= g.Q;
      if (++p == g.P) { p = 0; ++n; }
    }
  }
};

// Gather the 27 inputs of the output pixel under the cursor; out-of-image taps are the zero padding.
// VEC4: x is NHWC with a pixel stride of 4 floats (channel 3 is padding): one 16-byte load per tap;
// `toff` are then the nine (r, s) offsets of channel 0, held in registers by the caller.
template <bool VEC4>
__device__ __forceinline__ void gather_taps(const StemGeom &g, const float *__restrict__ x, const PixCursor &c,
                                            const int (&toff)[9], float (&v)[27]) {
  const int h0 = c.p * g.sh - g.ph, w0 = c.q * g.sw - g.pw;
  const float *xb = x + (c.n * g.xs0 + h0 * g.xs2 + w0 * g.xs3);
  bool okh[3], okw[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) okh[r] = (unsigned)(h0 + r * g.dh) < (unsigned)g.H;
#pragma unroll
  for (int s = 0; s < 3; ++s) okw[s] = (unsigned)(w0 + s * g.dw) < (unsigned)g.W;
  if (VEC4) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (okh[r] && okw[s]) t = __ldg(reinterpret_cast<const float4 *>(xb + toff[r * 3 + s]));
        v[0 * 9 + r * 3 + s] = t.x; v[1 * 9 + r * 3 + s] = t.y; v[2 * 9 + r * 3 + s] = t.z;
      }
  } else {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int t = (ch * 3 + r) * 3 + s;
          v[t] = (okh[r] && okw[s]) ? __ldg(xb + g.toff[t]) : 0.f;
        }
  }
}

// KPT output channels per thread as KPT/2 channel pairs: packed FFMA2 (fma.rn.f32x2, sm_100), the
// scalar operand is broadcast by the instruction.
template <int KPT>
struct Pairs { float2 v[KPT / 2]; };

template <int KPT>
__device__ __forceinline__ void fma_pairs(Pairs<KPT> &acc, float a, const Pairs<KPT> &b) {
  const float2 aa = make_float2(a, a);
#pragma unroll
  for (int j = 0; j < KPT / 2; ++j) acc.v[j] = __ffma2_rn(aa, b.v[j], acc.v[j]);
}

template <int KPT>
__device__ __forceinline__ Pairs<KPT> load_pairs(const float *p) {   // 8- or 16-byte aligned
  Pairs<KPT> r;
  if (KPT == 4) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    r.v[0] = make_float2(t.x, t.y); r.v[KPT / 2 - 1] = make_float2(t.z, t.w);
  } else {
    r.v[0] = *reinterpret_cast<const float2 *>(p);
  }
  return r;
}

template <bool VEC4, int KPT>
__global__ void __launch_bounds__(STEM_THREADS)
stem_fprop_kernel(const __grid_constant__ StemGeom g, const float *__restrict__ x, const float *__restrict__ w,
                  const float *__restrict__ piggy, const float *__restrict__ bias, float *__restrict__ y, float thr,
                  int per_block) {
  constexpr int T = 27;
  const int lpp = 1 << g.lpp_log2;
  const int kq = threadIdx.x & (lpp - 1);            // channel group of this thread
  const int slot = threadIdx.x >> g.lpp_log2;        // pixel slot inside the block
  const int slots = STEM_THREADS >> g.lpp_log2;
  // the thread's KPT x 27 masked weights (module order [K][C][R][S]) as channel pairs
  Pairs<KPT> wr[T];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int j = 0; j < KPT / 2; ++j) {
      const long long i0 = (long long)(kq * KPT + 2 * j) * T + t;
      wr[t].v[j] = make_float2(masked_weight(__ldg(w + i0), piggy, i0, thr),
                               masked_weight(__ldg(w + i0 + T), piggy, i0 + T, thr));
    }
  Pairs<KPT> b;
#pragma unroll
  for (int j = 0; j < KPT / 2; ++j) b.v[j] = make_float2(0.f, 0.f);
  if (bias) b = load_pairs<KPT>(bias + kq * KPT);
  int toff[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) toff[i] = g.toff[i];
  const int beg = blockIdx.x * per_block;
  const int end = min(g.pixels, beg + per_block);
  int pix = beg + slot;
  if (pix >= end) return;
  PixCursor c;
  c.seek(g, pix);
  for (; pix < end; pix += slots, c.advance(g, slots)) {
    float v[T];
    gather_taps<VEC4>(g, x, c, toff, v);
    // one accumulator set per input channel: three short FMA chains instead of one of 27
    Pairs<KPT> a0 = b, a1, a2;
#pragma unroll
    for (int j = 0; j < KPT / 2; ++j) a1.v[j] = a2.v[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      fma_pairs<KPT>(a0, v[t], wr[t]);
      fma_pairs<KPT>(a1, v[9 + t], wr[9 + t]);
      fma_pairs<KPT>(a2, v[18 + t], wr[18 + t]);
    }
    float *yp = y + (c.n * g.ys0 + c.p * g.ys2 + c.q * g.ys3) + kq * KPT;
    if (KPT == 4) {
      __stcs(reinterpret_cast<float4 *>(yp),
             make_float4(a0.v[0].x + a1.v[0].x + a2.v[0].x, a0.v[0].y + a1.v[0].y + a2.v[0].y,
                         a0.v[KPT / 2 - 1].x + a1.v[KPT / 2 - 1].x + a2.v[KPT / 2 - 1].x,
                         a0.v[KPT / 2 - 1].y + a1.v[KPT / 2 - 1].y + a2.v[KPT / 2 - 1].y));
    } else {
      __stcs(reinterpret_cast<float2 *>(yp),
             make_float2(a0.v[0].x + a1.v[0].x + a2.v[0].x, a0.v[0].y + a1.v[0].y + a2.v[0].y));
    }
  }
}

// Block reduction shared by the two wgrad kernels: pixel slots that share a warp are folded with a
// fixed butterfly, the warps' sums go through `red` ([warps][K * 27]) and are added in warp order.
template <int KPT>
__device__ __forceinline__ void wgrad_block_reduce(const StemGeom &g, Pairs<KPT> (&acc)[27], float *red,
                                                   float *__restrict__ out) {
  constexpr int T = 27;
  const int lpp = 1 << g.lpp_log2;
  const int kq = threadIdx.x & (lpp - 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o >= lpp; o >>= 1) {
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int j = 0; j < KPT / 2; ++j) {
        acc[t].v[j].x += __shfl_xor_sync(0xffffffffu, acc[t].v[j].x, o);
        acc[t].v[j].y += __shfl_xor_sync(0xffffffffu, acc[t].v[j].y, o);
      }
  }
  const int KT = g.K * T;
  if (lane < lpp) {
    float *dst = red + warp * KT + kq * KPT * T;
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int j = 0; j < KPT / 2; ++j) {
        dst[(2 * j) * T + t] = acc[t].v[j].x;
        dst[(2 * j + 1) * T + t] = acc[t].v[j].y;
      }
  }
  __syncthreads();
  constexpr int NW = STEM_WG_THREADS / 32;
  for (int i = threadIdx.x; i < KT; i += STEM_WG_THREADS) {
    float s = red[i];
#pragma unroll
    for (int wq = 1; wq < NW; ++wq) s += red[wq * KT + i];
    out[i] = s;
  }
}

// per-block partial sums: part[block][K * 27] in the module's [K][C][R][S] order.  Plain-load variant
// for a dY that is NHWC but not pixel-dense.
template <bool VEC4, int KPT>
__global__ void __launch_bounds__(STEM_WG_THREADS)
stem_wgrad_kernel(const __grid_constant__ StemGeom g, const float *__restrict__ x, const float *__restrict__ dy,
                  float *__restrict__ part, int per_block) {
  constexpr int T = 27;
  extern __shared__ __align__(128) float ring[];     // [warps][K * 27]
  const int lpp = 1 << g.lpp_log2;
  const int kq = threadIdx.x & (lpp - 1);
  const int slot = threadIdx.x >> g.lpp_log2;
  const int slots = STEM_WG_THREADS >> g.lpp_log2;
  Pairs<KPT> acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int j = 0; j < KPT / 2; ++j) acc[t].v[j] = make_float2(0.f, 0.f);
  int toff[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) toff[i] = g.toff[i];
  const int beg = blockIdx.x * per_block;
  const int end = min(g.pixels, beg + per_block);
  int pix = beg + slot;
  if (pix < end) {
    PixCursor c;
    c.seek(g, pix);
    for (; pix < end; pix += slots, c.advance(g, slots)) {
      float v[T];
      gather_taps<VEC4>(g, x, c, toff, v);
      const Pairs<KPT> d = load_pairs<KPT>(dy + (c.n * g.ys0 + c.p * g.ys2 + c.q * g.ys3) + kq * KPT);
#pragma unroll
      for (int t = 0; t < T; ++t) fma_pairs<KPT>(acc[t], v[t], d);
    }
  }
  wgrad_block_reduce<KPT>(g, acc, ring, part + (long long)blockIdx.x * g.K * T);
}

// The same reduction with dY streamed through shared memory by the TMA unit: dense NHWC dY is one
// contiguous array of pixel rows, so a chunk of WG_CHUNK pixels is a single 1-D bulk copy
// (cp.async.bulk, mbarrier complete_tx).  WG_STAGES chunks are in flight per SM -- the accumulators
// (KPT x 27 per thread) leave room for 8-16 warps per SM, too few to cover HBM latency with plain
// loads.  After the loop the ring is reused for the block reduction.
constexpr int WG_CHUNK = 64;
constexpr int WG_STAGES = 4;

__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ptx::smem_u32(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

template <bool VEC4, int KPT>
__global__ void __launch_bounds__(STEM_WG_THREADS)
stem_wgrad_tma_kernel(const __grid_constant__ StemGeom g, const float *__restrict__ x, const float *__restrict__ dy,
                      float *__restrict__ part, int chunks_per_block) {
  constexpr int T = 27;
  extern __shared__ __align__(128) float ring[];     // [WG_STAGES][WG_CHUNK * K], later [warps][K * 27]
  __shared__ __align__(8) uint64_t full[WG_STAGES];
  const int lpp = 1 << g.lpp_log2;
  const int kq = threadIdx.x & (lpp - 1);
  const int slot = threadIdx.x >> g.lpp_log2;
  const int slots = STEM_WG_THREADS >> g.lpp_log2;
  const int chunk_elems = WG_CHUNK * g.K;
  const int nchunks = (g.pixels + WG_CHUNK - 1) / WG_CHUNK;
  const int ch0 = blockIdx.x * chunks_per_block;
  const int my = max(0, min(chunks_per_block, nchunks - ch0));

  auto issue = [&](int i) {                          // chunk i of this block -> stage i % WG_STAGES
    const int pix0 = (ch0 + i) * WG_CHUNK;
    const int npix = min(WG_CHUNK, g.pixels - pix0);
    const uint32_t bytes = (uint32_t)npix * g.K * 4u;
    uint64_t *bar = full + (i % WG_STAGES);
    ptx::mbar_arrive_expect_tx(bar, bytes);
    bulk_load_1d(ring + (i % WG_STAGES) * chunk_elems, dy + (long long)pix0 * g.K, bytes, bar);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) ptx::mbar_init(full + s, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 0; i < my && i < WG_STAGES; ++i) issue(i);

  Pairs<KPT> acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int j = 0; j < KPT / 2; ++j) acc[t].v[j] = make_float2(0.f, 0.f);
  int toff[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) toff[i] = g.toff[i];
  for (int i = 0; i < my; ++i) {
    const int st = i % WG_STAGES;
    const int pix0 = (ch0 + i) * WG_CHUNK;
    const int npix = min(WG_CHUNK, g.pixels - pix0);
    PixCursor c;
    if (slot < npix) c.seek(g, pix0 + slot);          // overlaps the wait below
    ptx::mbar_wait(full + st, (uint32_t)((i / WG_STAGES) & 1));
    const float *sd = ring + st * chunk_elems + kq * KPT;
    for (int j = slot; j < npix; j += slots, c.advance(g, slots)) {
      float v[T];
      gather_taps<VEC4>(g, x, c, toff, v);
      const Pairs<KPT> d = load_pairs<KPT>(sd + j * g.K);
#pragma unroll
      for (int t = 0; t < T; ++t) fma_pairs<KPT>(acc[t], v[t], d);
    }
    __syncthreads();                                  // every thread is done with stage st
    if (threadIdx.x == 0 && i + WG_STAGES < my) issue(i + WG_STAGES);
  }
  // every bulk copy has been consumed: the ring is free for the block reduction
  wgrad_block_reduce<KPT>(g, acc, ring, part + (long long)blockIdx.x * g.K * T);
}

// ---- pixel-pair kernels -----------------------------------------------------------------------
// The common stem (stride_w = dil_w = 1, x as NHWC4, K <= 64): one thread owns two horizontally
// adjacent output pixels and two output channels.  The pair shares a 3 x 4 window of 16-byte pixels
// (12 loads instead of 18, immediate column offsets from three row pointers), so the address and
// predicate work per pixel halves and 54 FFMA2 amortise it.
struct PairCursor {       // (n, p, pair-of-q) over N x P x Qh
  int n, p, qh;
  __device__ __forceinline__ void seek(const StemGeom &g, int Qh, int idx) {
    qh = idx % Qh;
    const int t = idx / Qh;
    p = t % g.P;
    n = t / g.P;
  }
  __device__ __forceinline__ void advance(const StemGeom &g, int Qh, int by) {
    qh += by;
    while (qh >= Qh) {
      qh -= Qh;
      if (++p == g.P) { p = 0; ++n; }
    }
  }
};

__device__ __forceinline__ void gather_window(const StemGeom &g, const float4 *__restrict__ xq, int n, int p, int q0,
                                              float4 (&win)[3][4]) {
  const int h0 = p * g.sh - g.ph, w0 = q0 - g.pw;
  const int xs2q = g.xs2 >> 2;
  const float4 *base = xq + ((long long)n * (g.xs0 >> 2) + w0);
  bool okw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) okw[j] = (unsigned)(w0 + j) < (unsigned)g.W;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int h = h0 + r * g.dh;
    const bool okh = (unsigned)h < (unsigned)g.H;
    const float4 *rowp = base + (long long)h * xs2q;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      win[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (okh && okw[j]) win[r][j] = __ldg(rowp + j);
    }
  }
}

// A block walks a contiguous range of output pixels, i.e. a few consecutive input rows.  Left to the
// main loop, every new row is a serialised cold miss (one per iteration, ~1 us each from HBM);
// touching the whole range up front turns them into one parallel burst.
__device__ __forceinline__ void prefetch_rows(const StemGeom &g, const float4 *__restrict__ xq, int pix_first,
                                              int pix_last) {
  if (!g.xdense4 || pix_last < pix_first) return;
  const int PQ = g.P * g.Q;
  const int n0 = pix_first / PQ, p0 = (pix_first - n0 * PQ) / g.Q;
  const int n1 = pix_last / PQ, p1 = (pix_last - n1 * PQ) / g.Q;
  const int h_lo = max(0, p0 * g.sh - g.ph), h_hi = min(g.H - 1, p1 * g.sh - g.ph + 2 * g.dh);
  const long long lo = ((long long)n0 * g.H + h_lo) * g.W, hi = ((long long)n1 * g.H + h_hi + 1) * g.W;
  for (long long i = lo + (long long)threadIdx.x * 8; i < hi; i += (long long)blockDim.x * 8)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(xq + i));
}

__device__ __forceinline__ float chan(const float4 &v, int c) { return c == 0 ? v.x : c == 1 ? v.y : v.z; }

__global__ void __launch_bounds__(STEM_THREADS)
stem_fprop_pair_kernel(const __grid_constant__ StemGeom g, const float *__restrict__ x, const float *__restrict__ w,
                       const float *__restrict__ piggy, const float *__restrict__ bias, float *__restrict__ y,
                       float thr, int Qh, int npairs, int per_block) {
  constexpr int T = 27;
  const int lpp = 1 << g.lpp_log2;
  const int kq = threadIdx.x & (lpp - 1);
  const int slot = threadIdx.x >> g.lpp_log2;
  const int slots = STEM_THREADS >> g.lpp_log2;
  float2 wr[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const long long i0 = (long long)(kq * 2) * T + t;
    wr[t] = make_float2(masked_weight(__ldg(w + i0), piggy, i0, thr),
                        masked_weight(__ldg(w + i0 + T), piggy, i0 + T, thr));
  }
  float2 b = make_float2(0.f, 0.f);
  if (bias) b = *reinterpret_cast<const float2 *>(bias + kq * 2);
  const int beg = blockIdx.x * per_block;
  const int end = min(npairs, beg + per_block);
  int idx = beg + slot;
  if (idx >= end) return;
  PairCursor c;
  c.seek(g, Qh, idx);
  const float4 *xq = reinterpret_cast<const float4 *>(x);
  if (Qh * 2 == g.Q) prefetch_rows(g, xq, beg * 2, end * 2 - 1);
  for (; idx < end; idx += slots, c.advance(g, Qh, slots)) {
    float4 win[3][4];
    const int q0 = c.qh * 2;
    gather_window(g, xq, c.n, c.p, q0, win);
    // per pixel one accumulator per input channel: three short FMA chains
    float2 a[3], bb[3];
    a[0] = bb[0] = b;
    a[1] = a[2] = bb[1] = bb[2] = make_float2(0.f, 0.f);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) {
          const float2 wv = wr[(ch * 3 + r) * 3 + s2];
          const float va = chan(win[r][s2], ch), vb = chan(win[r][s2 + 1], ch);
          a[ch] = __ffma2_rn(make_float2(va, va), wv, a[ch]);
          bb[ch] = __ffma2_rn(make_float2(vb, vb), wv, bb[ch]);
        }
    float *yp = y + (c.n * g.ys0 + c.p * g.ys2 + q0 * g.ys3) + kq * 2;
    __stcs(reinterpret_cast<float2 *>(yp), make_float2(a[0].x + a[1].x + a[2].x, a[0].y + a[1].y + a[2].y));
    if (q0 + 1 < g.Q)
      __stcs(reinterpret_cast<float2 *>(yp + g.ys3),
             make_float2(bb[0].x + bb[1].x + bb[2].x, bb[0].y + bb[1].y + bb[2].y));
  }
}

// wgrad over pixel pairs, dY through the TMA ring (dense NHWC dY, Q even so that a pair never
// straddles an image row and chunks of WG_CHUNK pixels hold whole pairs).
__global__ void __launch_bounds__(STEM_WG_THREADS)
stem_wgrad_pair_kernel(const __grid_constant__ StemGeom g, const float *__restrict__ x, const float *__restrict__ dy,
                       float *__restrict__ part, int chunks_per_block) {
  constexpr int T = 27;
  extern __shared__ __align__(128) float ring[];     // [WG_STAGES][WG_CHUNK * K], later [warps][K * 27]
  __shared__ __align__(8) uint64_t full[WG_STAGES];
  const int lpp = 1 << g.lpp_log2;
  const int kq = threadIdx.x & (lpp - 1);
  const int slot = threadIdx.x >> g.lpp_log2;
  const int slots = STEM_WG_THREADS >> g.lpp_log2;
  const int chunk_elems = WG_CHUNK * g.K;
  const int nchunks = (g.pixels + WG_CHUNK - 1) / WG_CHUNK;
  const int ch0 = blockIdx.x * chunks_per_block;
  const int my = max(0, min(chunks_per_block, nchunks - ch0));
  const int Qh = g.Q >> 1;

  auto issue = [&](int i) {
    const int pix0 = (ch0 + i) * WG_CHUNK;
    const int npix = min(WG_CHUNK, g.pixels - pix0);
    const uint32_t bytes = (uint32_t)npix * g.K * 4u;
    uint64_t *bar = full + (i % WG_STAGES);
    ptx::mbar_arrive_expect_tx(bar, bytes);
    bulk_load_1d(ring + (i % WG_STAGES) * chunk_elems, dy + (long long)pix0 * g.K, bytes, bar);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) ptx::mbar_init(full + s, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 0; i < my && i < WG_STAGES; ++i) issue(i);

  Pairs<2> acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t].v[0] = make_float2(0.f, 0.f);
  const float4 *xq = reinterpret_cast<const float4 *>(x);
  if (my > 0) prefetch_rows(g, xq, ch0 * WG_CHUNK, min(g.pixels, (ch0 + my) * WG_CHUNK) - 1);
  for (int i = 0; i < my; ++i) {
    const int st = i % WG_STAGES;
    const int pix0 = (ch0 + i) * WG_CHUNK;
    const int npairs = min(WG_CHUNK, g.pixels - pix0) >> 1;      // pixels is even (Q even)
    PairCursor c;
    if (slot < npairs) c.seek(g, Qh, (pix0 >> 1) + slot);
    ptx::mbar_wait(full + st, (uint32_t)((i / WG_STAGES) & 1));
    const float *sd = ring + st * chunk_elems + kq * 2;
    for (int j = slot; j < npairs; j += slots, c.advance(g, Qh, slots)) {
      float4 win[3][4];
      gather_window(g, xq, c.n, c.p, c.qh * 2, win);
      const float2 da = *reinterpret_cast<const float2 *>(sd + (2 * j) * g.K);
      const float2 db = *reinterpret_cast<const float2 *>(sd + (2 * j + 1) * g.K);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2) {
            const int t = (ch * 3 + r) * 3 + s2;
            const float va = chan(win[r][s2], ch), vb = chan(win[r][s2 + 1], ch);
            acc[t].v[0] = __ffma2_rn(make_float2(va, va), da, acc[t].v[0]);
            acc[t].v[0] = __ffma2_rn(make_float2(vb, vb), db, acc[t].v[0]);
          }
    }
    __syncthreads();
    if (threadIdx.x == 0 && i + WG_STAGES < my) issue(i + WG_STAGES);
  }
  wgrad_block_reduce<2>(g, acc, ring, part + (long long)blockIdx.x * g.K * T);
}

// g[i] = sum over blocks of part[b][i] (fixed order per lane, fixed shuffle tree), then the fused
// epilogue of SURVEY K6-K8 on element i.  One warp per weight element.
__global__ void __launch_bounds__(256)
stem_wgrad_finish_kernel(const float *__restrict__ part, int nblocks, int n, const float *__restrict__ w,
                         const float *__restrict__ piggy, const uint8_t *__restrict__ tmask, int cur, float wd, int mode,
                         float thr, float *__restrict__ dW, float *__restrict__ dP) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  float s = 0.f;
  for (int b = lane; b < nblocks; b += 32) s += __ldg(part + (long long)b * n + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    float ow, op;
    const bool has_p = piggy != nullptr;
    epi_stem(s, w[i], has_p ? piggy[i] : 0.f, has_p, tmask ? tmask[i] : 0u, cur, wd, mode, thr, ow, op);
    dW[i] = ow;
    if (dP) dP[i] = op;
  }
}

int stem_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n = v > 0 ? v : 148;
  }
  return n;
}

// channels per thread: 2 (K <= 64: more, lighter threads) or 4 (K = 128: 32 lanes per pixel)
int kpt_of(int K) { return K <= 64 ? 2 : 4; }

bool shape_ok(int K, int C, int R, int S, int groups) {
  if (groups != 1 || C != 3 || R != 3 || S != 3 || K % 4 != 0) return false;
  const int lpp = K / kpt_of(K);
  return lpp >= 1 && lpp <= 32 && (lpp & (lpp - 1)) == 0;
}

bool fill_geom(const cpgb_conv_desc &d, StemGeom *sg) {
  if (!shape_ok(d.K, d.C, d.R, d.S, d.groups)) return false;
  const int lpp = d.K / kpt_of(d.K);
  // y / dy: channels contiguous, pixel rows 16-byte aligned
  if (d.ys[1] != 1 || d.ys[0] % 4 || d.ys[2] % 4 || d.ys[3] % 4) return false;
  // x offsets in int32, halo included
  long long span = 0;
  const long long ext[4] = {d.N, d.C, (long long)d.H + 2 * d.pad_h + d.dil_h * d.R, (long long)d.W + 2 * d.pad_w + d.dil_w * d.S};
  for (int i = 0; i < 4; ++i) {
    if (d.xs[i] < 0) return false;
    span += ext[i] * d.xs[i];
  }
  if (span >= (1ll << 31)) return false;
  if (!sg) return true;
  sg->N = d.N; sg->H = d.H; sg->W = d.W; sg->K = d.K; sg->P = d.P; sg->Q = d.Q;
  sg->sh = d.stride_h; sg->sw = d.stride_w; sg->ph = d.pad_h; sg->pw = d.pad_w; sg->dh = d.dil_h; sg->dw = d.dil_w;
  sg->xs0 = (int)d.xs[0]; sg->xs2 = (int)d.xs[2]; sg->xs3 = (int)d.xs[3];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s)
        sg->toff[(c * 3 + r) * 3 + s] = (int)(c * d.xs[1] + r * d.dil_h * d.xs[2] + s * d.dil_w * d.xs[3]);
  sg->ys0 = d.ys[0]; sg->ys2 = d.ys[2]; sg->ys3 = d.ys[3];
  sg->pixels = (int)((long long)d.N * d.P * d.Q);
  int l = 0;
  while ((1 << l) < lpp) ++l;
  sg->lpp_log2 = l;
  sg->xdense4 = d.xs[1] == 1 && d.xs[3] == 4 && d.xs[2] == 4ll * d.W && d.xs[0] == 4ll * d.W * d.H;
  return true;
}

bool x_is_vec4(const cpgb_conv_desc &d, const float *x) {
  return d.xs[1] == 1 && d.xs[3] == 4 && d.xs[0] % 4 == 0 && d.xs[2] % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(x) & 15) == 0;
}

// pixel-pair kernels: x NHWC4, unit stride / dilation along w, two channels per thread
bool pair_ok(const cpgb_conv_desc &d, const float *x) {
  return x_is_vec4(d, x) && d.stride_w == 1 && d.dil_w == 1 && kpt_of(d.K) == 2;
}

constexpr int MAX_WG_SMEM = 131072;

// launch tables indexed by [vec4][kpt == 4]
typedef void (*FpropKern)(const StemGeom, const float *, const float *, const float *, const float *, float *, float, int);
typedef void (*WgradKern)(const StemGeom, const float *, const float *, float *, int);
const FpropKern kFprop[2][2] = {{stem_fprop_kernel<false, 2>, stem_fprop_kernel<false, 4>},
                                {stem_fprop_kernel<true, 2>, stem_fprop_kernel<true, 4>}};
const WgradKern kWgrad[2][2] = {{stem_wgrad_kernel<false, 2>, stem_wgrad_kernel<false, 4>},
                                {stem_wgrad_kernel<true, 2>, stem_wgrad_kernel<true, 4>}};
const WgradKern kWgradTma[2][2] = {{stem_wgrad_tma_kernel<false, 2>, stem_wgrad_tma_kernel<false, 4>},
                                   {stem_wgrad_tma_kernel<true, 2>, stem_wgrad_tma_kernel<true, 4>}};

// resident blocks of a kernel on the whole device (one wave)
template <class Kern>
int resident_blocks(Kern kern, int threads, size_t smem) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  return per_sm * stem_sms();
}

// upper bound of the wgrad grid (sizes the partial-sum workspace): two resident blocks per SM
int wgrad_blocks_max(const cpgb_conv_desc &d) {
  const long long slots = STEM_WG_THREADS / (d.K / kpt_of(d.K));
  const long long groups = ((long long)d.N * d.P * d.Q + slots - 1) / slots;
  const long long cap = 2ll * stem_sms();
  return (int)(groups < 1 ? 1 : groups < cap ? groups : cap);
}

}  // namespace

bool stem_eligible(const cpgb_conv_desc &d) { return fill_geom(d, nullptr); }
bool stem_weight_shape(int K, int C, int R, int S, int groups) { return shape_ok(K, C, R, S, groups); }

size_t stem_workspace_bytes(const cpgb_conv_desc &d) {
  if (!stem_eligible(d)) return 0;
  return (size_t)wgrad_blocks_max(d) * d.K * 27 * sizeof(float);
}

int stem_fprop(const cpgb_conv_desc &d, const float *x, const float *w, const float *piggy, const float *bias, float *y,
               float thr, cudaStream_t st) {
  StemGeom sg;
  if (!fill_geom(d, &sg)) { set_error("stem_fprop: shape not eligible"); return CPGB_EINVAL; }
  const int vec = x_is_vec4(d, x) ? 1 : 0, k4 = kpt_of(d.K) == 4 ? 1 : 0;
  const int slots = STEM_THREADS >> sg.lpp_log2;
  if (pair_ok(d, x) && (!bias || (reinterpret_cast<uintptr_t>(bias) & 7) == 0)) {
    const int Qh = (d.Q + 1) / 2;
    const long long npairs = (long long)d.N * d.P * Qh;
    const long long groups = (npairs + slots - 1) / slots;
    static int cap_pair = 0;
    if (!cap_pair) cap_pair = resident_blocks(stem_fprop_pair_kernel, STEM_THREADS, 0);
    const int grid = (int)(groups < 1 ? 1 : groups < cap_pair ? groups : cap_pair);
    int per_block = (int)((npairs + grid - 1) / grid);
    per_block = (per_block + slots - 1) / slots * slots;
    stem_fprop_pair_kernel<<<grid, STEM_THREADS, 0, st>>>(sg, x, w, piggy, bias, y, thr, Qh, (int)npairs, per_block);
    CPGB_LAUNCH_OK("stem_fprop_pair");
    return CPGB_OK;
  }
  const long long groups = ((long long)sg.pixels + slots - 1) / slots;
  // one wave of resident blocks, each walking a contiguous pixel range (the KPT x 27 weights of a
  // thread are loaded once per block)
  static int cap[2][2] = {{0, 0}, {0, 0}};
  if (!cap[vec][k4]) cap[vec][k4] = resident_blocks(kFprop[vec][k4], STEM_THREADS, 0);
  const int grid = (int)(groups < 1 ? 1 : groups < cap[vec][k4] ? groups : cap[vec][k4]);
  int per_block = (int)(((long long)sg.pixels + grid - 1) / grid);
  per_block = (per_block + slots - 1) / slots * slots;
  kFprop[vec][k4]<<<grid, STEM_THREADS, 0, st>>>(sg, x, w, piggy, bias, y, thr, per_block);
  CPGB_LAUNCH_OK("stem_fprop");
  return CPGB_OK;
}

int stem_wgrad_fused(const cpgb_conv_desc &d, const float *x, const float *dy, const float *w, const float *piggy,
                     const uint8_t *tmask, int cur, float wd, int mode, float thr, float *dW, float *dP, void *ws,
                     size_t ws_bytes, cudaStream_t st) {
  StemGeom sg;
  if (!fill_geom(d, &sg)) { set_error("stem_wgrad: shape not eligible"); return CPGB_EINVAL; }
  const int nb_max = wgrad_blocks_max(d);
  const int n = d.K * 27;
  if (!ws || ws_bytes < (size_t)nb_max * n * sizeof(float)) {
    set_error("workspace %zu < %zu (stem wgrad partial sums)", ws_bytes, (size_t)nb_max * n * sizeof(float));
    return CPGB_EWORKSPACE;
  }
  const int vec = x_is_vec4(d, x) ? 1 : 0, k4 = kpt_of(d.K) == 4 ? 1 : 0;
  const size_t red_bytes = (size_t)(STEM_WG_THREADS / 32) * n * sizeof(float);
  const size_t ring_bytes = (size_t)WG_STAGES * WG_CHUNK * d.K * sizeof(float);
  const bool dense = d.ys[3] == d.K && d.ys[2] == (long long)d.Q * d.K && d.ys[0] == (long long)d.P * d.Q * d.K &&
                     (reinterpret_cast<uintptr_t>(dy) & 15) == 0;
  const size_t smem = dense && ring_bytes > red_bytes ? ring_bytes : red_bytes;
  if (dense && pair_ok(d, x) && d.Q % 2 == 0) {
    static int cap_pair = 0;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
      CPGB_CUDA_OK(cudaFuncSetAttribute(stem_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_WG_SMEM));
      cap_pair = resident_blocks(stem_wgrad_pair_kernel, STEM_WG_THREADS, smem);
    }
    int nb = cap_pair < nb_max ? cap_pair : nb_max;
    const int nchunks = (sg.pixels + WG_CHUNK - 1) / WG_CHUNK;
    const int cpb = (nchunks + nb - 1) / nb;
    nb = (nchunks + cpb - 1) / cpb;
    stem_wgrad_pair_kernel<<<nb, STEM_WG_THREADS, smem, st>>>(sg, x, dy, reinterpret_cast<float *>(ws), cpb);
    CPGB_LAUNCH_OK("stem_wgrad_pair");
    stem_wgrad_finish_kernel<<<(n + 7) / 8, 256, 0, st>>>(reinterpret_cast<float *>(ws), nb, n, w, piggy, tmask, cur, wd,
                                                         mode, thr, dW, dP);
    CPGB_LAUNCH_OK("stem_wgrad_finish");
    return CPGB_OK;
  }
  const WgradKern kern = dense ? kWgradTma[vec][k4] : kWgrad[vec][k4];
  static int cap[2][2][2] = {};
  static PerDeviceOnce attr_once[2][2][2];
  if (attr_once[dense][vec][k4].need()) {
    CPGB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_WG_SMEM));
    int c = resident_blocks(kern, STEM_WG_THREADS, smem);
    cap[dense][vec][k4] = c < nb_max ? c : nb_max;
    if (cap[dense][vec][k4] < 1) cap[dense][vec][k4] = 1;
  }
  int nb = cap[dense][vec][k4] < nb_max ? cap[dense][vec][k4] : nb_max;
  float *part = reinterpret_cast<float *>(ws);
  if (dense) {
    const int nchunks = (sg.pixels + WG_CHUNK - 1) / WG_CHUNK;
    const int cpb = (nchunks + nb - 1) / nb;            // contiguous run of chunks per block
    nb = (nchunks + cpb - 1) / cpb;
    kern<<<nb, STEM_WG_THREADS, smem, st>>>(sg, x, dy, part, cpb);
  } else {
    const int slots = STEM_WG_THREADS >> sg.lpp_log2;
    int per_block = (sg.pixels + nb - 1) / nb;
    per_block = (per_block + slots - 1) / slots * slots;
    kern<<<nb, STEM_WG_THREADS, smem, st>>>(sg, x, dy, part, per_block);
  }
  CPGB_LAUNCH_OK("stem_wgrad");
  stem_wgrad_finish_kernel<<<(n + 7) / 8, 256, 0, st>>>(part, nb, n, w, piggy, tmask, cur, wd, mode, thr, dW, dP);
  CPGB_LAUNCH_OK("stem_wgrad_finish");
  return CPGB_OK;
}

}  // namespace cpgb
