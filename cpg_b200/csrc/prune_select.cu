// a7: SparsePruner._pruning_mask (utils/prune.py:30-53) on the device, no host round trip.
//
// The reference gathers the prunable pool, copies it to the host and calls kthvalue
// (utils/prune.py:39).  Here the exact k-th smallest |w| over the pool is found with a
// 3-digit (11/11/9-bit) radix select on the fp32 bit pattern of |w| (monotone for
// non-negative floats), every pass a coalesced streaming read of W (4 B) and T (1 B) with a
// shared-memory histogram per CTA and warp-aggregated flushes; a final pass applies
// T[|w| <= cut and T == cur] = 0.  k = round_half_even(ratio * |pool|) is computed on the
// device in double precision, which is what python's round() does at utils/prune.py:37.
#include "common.cuh"

namespace cpgb {

constexpr int RS_BINS = 2048;
struct PruneState {
  unsigned long long hist[3][RS_BINS];
  unsigned long long pool, k, kremain;
  unsigned int prefix;     // key bits selected so far (aligned to the low end)
  unsigned int cut_bits;
  int status;              // 0 ok, 2 = exit-2 path
  int pad;
};

__device__ __forceinline__ unsigned key_of(float w) { return __float_as_uint(fabsf(w)); }

// pass 0: digit = key >> 20 (11 bits; bit 31 is always 0)
// pass 1: digit = (key >> 9) & 0x7ff, restricted to key >> 20 == prefix
// pass 2: digit = key & 0x1ff,        restricted to key >> 9  == prefix
template <int PASS>
__global__ void __launch_bounds__(256)
prune_hist_kernel(const float *__restrict__ w, const uint8_t *__restrict__ tmask, long long n, int cur,
                  PruneState *__restrict__ stt) {
  __shared__ unsigned int sh[RS_BINS];
  for (int i = threadIdx.x; i < RS_BINS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  if (PASS > 0 && stt->status != 0) return;
  const unsigned prefix = PASS > 0 ? stt->prefix : 0u;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(w) & 15) == 0) && ((reinterpret_cast<uintptr_t>(tmask) & 3) == 0);
  // bin of one element, or 0xffffffff when it does not take part in this pass
  auto bin_of = [&](float wv, unsigned t) -> unsigned {
    if (t != (unsigned)cur && t != 0u) return 0xffffffffu;
    unsigned key = key_of(wv);
    if (PASS == 0) return key >> 20;
    if (PASS == 1) return (key >> 20) == prefix ? ((key >> 9) & 0x7ffu) : 0xffffffffu;
    return (key >> 9) == prefix ? (key & 0x1ffu) : 0xffffffffu;
  };
  // plain shared-memory atomics: ptxas already aggregates same-address atomics of a warp (REDUX), and an
  // explicit __match_any_sync aggregation measured 35 % slower on B200 (0.60 vs 0.44 ms per VGG16 event)
  auto add_warp = [&](unsigned bin) {
    if (bin != 0xffffffffu) atomicAdd(&sh[bin], 1u);
  };
  long long tail = 0;
  if (vec) {
    const long long n4 = n >> 2;
    const long long n4r = (n4 + stride - 1) / stride * stride;     // every lane of a warp takes part in match_any
    for (long long v = i0; v < n4r; v += stride) {
      unsigned b0 = 0xffffffffu, b1 = b0, b2 = b0, b3 = b0;
      if (v < n4) {
        float4 a = __ldg(reinterpret_cast<const float4 *>(w) + v);
        uchar4 t = __ldg(reinterpret_cast<const uchar4 *>(tmask) + v);
        b0 = bin_of(a.x, t.x); b1 = bin_of(a.y, t.y); b2 = bin_of(a.z, t.z); b3 = bin_of(a.w, t.w);
      }
      add_warp(b0); add_warp(b1); add_warp(b2); add_warp(b3);
    }
    tail = n4 << 2;
  }
  for (long long i = tail + i0; i < n; i += stride) {
    unsigned b = bin_of(w[i], tmask[i]);
    if (b != 0xffffffffu) atomicAdd(&sh[b], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RS_BINS; i += blockDim.x) {
    unsigned c = sh[i];
    if (c) atomicAdd(&stt->hist[PASS][i], (unsigned long long)c);
  }
}

// One warp walks the histogram: finds the bin holding the k-th element.
template <int PASS>
__global__ void prune_scan_kernel(PruneState *__restrict__ stt, double ratio, long long *__restrict__ info) {
  const int lane = threadIdx.x;
  if (PASS > 0 && stt->status != 0) return;
  const int nbins = PASS == 2 ? 512 : RS_BINS;
  unsigned long long k;
  if (PASS == 0) {
    unsigned long long s = 0;
    for (int i = lane; i < nbins; i += 32) s += stt->hist[0][i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    // python: round(ratio * pool) -> IEEE double product, round-half-even (utils/prune.py:37)
    double kd = rint(ratio * (double)s);
    long long kk = (long long)kd;
    if (lane == 0) {
      stt->pool = s;
      stt->k = kk > 0 ? (unsigned long long)kk : 0ull;
      info[1] = (long long)s;
      info[2] = kk;
      info[3] = 0;
    }
    if (kk < 1 || (unsigned long long)kk > s) {  // kthvalue raises -> sys.exit(2) (utils/prune.py:38-42)
      if (lane == 0) { stt->status = 2; info[0] = 2; }
      return;
    }
    if (lane == 0) info[0] = 0;
    k = (unsigned long long)kk;
  } else {
    k = stt->kremain;
  }
  // chunked inclusive scan, 32 bins at a time
  unsigned long long base = 0;
  for (int c = 0; c < nbins; c += 32) {
    unsigned long long v = stt->hist[PASS][c + lane], incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
    if (base + total >= k) {
      unsigned hit = __ballot_sync(0xffffffffu, base + incl >= k);
      int first = __ffs(hit) - 1;
      unsigned long long before = base + __shfl_sync(0xffffffffu, incl - v, first);
      if (lane == 0) {
        unsigned bin = (unsigned)(c + first);
        stt->kremain = k - before;
        if (PASS == 0) stt->prefix = bin;
        else if (PASS == 1) stt->prefix = (stt->prefix << 11) | bin;
        else {
          unsigned cut = (stt->prefix << 9) | bin;
          stt->cut_bits = cut;
          info[3] = (long long)cut;
        }
      }
      return;
    }
    base += total;
  }
}

__global__ void __launch_bounds__(256)
prune_update_kernel(const float *__restrict__ w, uint8_t *__restrict__ tmask, long long n, int cur,
                    const PruneState *__restrict__ stt) {
  if (stt->status != 0) return;
  const unsigned cut = stt->cut_bits;
  if (cut > 0x7f800000u) return;   // the k-th magnitude is a NaN: `abs(w) <= NaN` holds nowhere (utils/prune.py:44-47)
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    // |w| <= cut as floats == key <= cut_bits for non-NaN; NaN keys exceed every finite cut
    if (tmask[i] == (uint8_t)cur && key_of(w[i]) <= cut) tmask[i] = 0;
  }
}


// ---------------------------------------------------------------------------------------------
// Batched variant: every sharable layer of the model in ONE launch per pass (grid.y = layer).  A prune
// event touches 15 (VGG16) to 53 (ResNet-50) layers; per-layer launches are launch-bound (7 launches
// of a few microseconds each per layer), the batched passes stream all weights at HBM speed.
// ---------------------------------------------------------------------------------------------
constexpr int PRUNE_MAX_LAYERS = 64;
struct PruneBatch {
  const float *w[PRUNE_MAX_LAYERS];
  uint8_t *t[PRUNE_MAX_LAYERS];
  long long n[PRUNE_MAX_LAYERS];
};

template <int PASS>
__global__ void __launch_bounds__(256)
prune_hist_batched_kernel(const __grid_constant__ PruneBatch pb, int cur, PruneState *__restrict__ states) {
  __shared__ unsigned int sh[RS_BINS];
  const int layer = blockIdx.y;
  const float *__restrict__ w = pb.w[layer];
  const uint8_t *__restrict__ tmask = pb.t[layer];
  const long long n = pb.n[layer];
  PruneState *stt = states + layer;
  long long nblk = (n / 4 + blockDim.x - 1) / blockDim.x;
  if (nblk < 1) nblk = 1;
  const long long gx = nblk < (long long)gridDim.x ? nblk : (long long)gridDim.x;   // blocks working on this layer
  if ((long long)blockIdx.x >= gx) return;   // the grid is sized for the largest layer
  for (int i = threadIdx.x; i < RS_BINS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  if (PASS > 0 && stt->status != 0) return;
  const unsigned prefix = PASS > 0 ? stt->prefix : 0u;
  const long long stride = gx * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(w) & 15) == 0) && ((reinterpret_cast<uintptr_t>(tmask) & 3) == 0);
  // bin of one element, or 0xffffffff when it does not take part in this pass
  auto bin_of = [&](float wv, unsigned t) -> unsigned {
    if (t != (unsigned)cur && t != 0u) return 0xffffffffu;
    unsigned key = key_of(wv);
    if (PASS == 0) return key >> 20;
    if (PASS == 1) return (key >> 20) == prefix ? ((key >> 9) & 0x7ffu) : 0xffffffffu;
    return (key >> 9) == prefix ? (key & 0x1ffu) : 0xffffffffu;
  };
  // plain shared-memory atomics: ptxas already aggregates same-address atomics of a warp (REDUX), and an
  // explicit __match_any_sync aggregation measured 35 % slower on B200 (0.60 vs 0.44 ms per VGG16 event)
  auto add_warp = [&](unsigned bin) {
    if (bin != 0xffffffffu) atomicAdd(&sh[bin], 1u);
  };
  long long tail = 0;
  if (vec) {
    const long long n4 = n >> 2;
    const long long n4r = (n4 + stride - 1) / stride * stride;     // every lane of a warp takes part in match_any
    for (long long v = i0; v < n4r; v += stride) {
      unsigned b0 = 0xffffffffu, b1 = b0, b2 = b0, b3 = b0;
      if (v < n4) {
        float4 a = __ldg(reinterpret_cast<const float4 *>(w) + v);
        uchar4 t = __ldg(reinterpret_cast<const uchar4 *>(tmask) + v);
        b0 = bin_of(a.x, t.x); b1 = bin_of(a.y, t.y); b2 = bin_of(a.z, t.z); b3 = bin_of(a.w, t.w);
      }
      add_warp(b0); add_warp(b1); add_warp(b2); add_warp(b3);
    }
    tail = n4 << 2;
  }
  for (long long i = tail + i0; i < n; i += stride) {
    unsigned b = bin_of(w[i], tmask[i]);
    if (b != 0xffffffffu) atomicAdd(&sh[b], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RS_BINS; i += blockDim.x) {
    unsigned c = sh[i];
    if (c) atomicAdd(&stt->hist[PASS][i], (unsigned long long)c);
  }
}

template <int PASS>
__global__ void prune_scan_batched_kernel(PruneState *__restrict__ states, double ratio, long long *__restrict__ info) {
  // one warp per layer; same walk as prune_scan_kernel
  PruneState *stt = states + blockIdx.x;
  long long *inf = info + 4 * blockIdx.x;
  const int lane = threadIdx.x;
  if (PASS > 0 && stt->status != 0) return;
  const int nbins = PASS == 2 ? 512 : RS_BINS;
  unsigned long long k;
  if (PASS == 0) {
    unsigned long long s = 0;
    for (int i = lane; i < nbins; i += 32) s += stt->hist[0][i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    double kd = rint(ratio * (double)s);
    long long kk = (long long)kd;
    if (lane == 0) {
      stt->pool = s;
      stt->k = kk > 0 ? (unsigned long long)kk : 0ull;
      inf[1] = (long long)s; inf[2] = kk; inf[3] = 0;
    }
    if (kk < 1 || (unsigned long long)kk > s) {
      if (lane == 0) { stt->status = 2; inf[0] = 2; }
      return;
    }
    if (lane == 0) inf[0] = 0;
    k = (unsigned long long)kk;
  } else {
    k = stt->kremain;
  }
  unsigned long long base = 0;
  for (int c = 0; c < nbins; c += 32) {
    unsigned long long v = stt->hist[PASS][c + lane], incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
    if (base + total >= k) {
      unsigned hit = __ballot_sync(0xffffffffu, base + incl >= k);
      int first = __ffs(hit) - 1;
      unsigned long long before = base + __shfl_sync(0xffffffffu, incl - v, first);
      if (lane == 0) {
        unsigned bin = (unsigned)(c + first);
        stt->kremain = k - before;
        if (PASS == 0) stt->prefix = bin;
        else if (PASS == 1) stt->prefix = (stt->prefix << 11) | bin;
        else {
          unsigned cut = (stt->prefix << 9) | bin;
          stt->cut_bits = cut;
          inf[3] = (long long)cut;
        }
      }
      return;
    }
    base += total;
  }
}

__global__ void __launch_bounds__(256)
prune_update_batched_kernel(const __grid_constant__ PruneBatch pb, int cur, const PruneState *__restrict__ states) {
  const int layer = blockIdx.y;
  const PruneState *stt = states + layer;
  if (stt->status != 0) return;
  const float *__restrict__ w = pb.w[layer];
  uint8_t *__restrict__ tmask = pb.t[layer];
  const long long n = pb.n[layer];
  const unsigned cut = stt->cut_bits;
  if (cut > 0x7f800000u) return;   // NaN cut: nothing is pruned
  long long nblk = (n + blockDim.x - 1) / blockDim.x;
  const long long gx = nblk < (long long)gridDim.x ? nblk : (long long)gridDim.x;
  if ((long long)blockIdx.x >= gx) return;
  const long long stride = gx * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (tmask[i] == (uint8_t)cur && key_of(w[i]) <= cut) tmask[i] = 0;
}


// ---------------------------------------------------------------------------------------------
// Two-pass variant (cpgb_prune_select_sampled): the same exact k-th magnitude, found with TWO streaming
// passes over W / T instead of four and almost no shared-memory atomics.
//
//   1. sample   : SAMPLE keys per layer at evenly spaced positions (pool members only).
//   2. bracket  : one block per layer sorts its sample (bitonic, shared memory) and takes the sample quantiles
//                 ratio*m -/+ 4 sigma: a key interval [lo, hi] that holds the k-th pool element with overwhelming
//                 probability and only a few percent of the pool.
//   3. count    : streaming pass: |pool|, #(key < lo) and #(key == lo) in registers (lo is often a heavy tie: the
//                 zeros apply_mask left in the freed weights), the keys inside (lo, hi] into a 2048-bin histogram of
//                 (key - lo - 1) >> shift (a few percent of the elements: few atomics).
//   4. locate   : k = round_half_even(ratio * |pool|); if it does not fall inside the bracket -> status 3
//                 (caller falls back to the four-pass radix select; T untouched).  Otherwise the bin of the k-th.
//   5. update   : streaming pass: everything below that bin is pruned right away; the elements OF the bin
//                 (about |bracket| / 2048 of them) are collected as (key, index) candidates -- unless shift == 0,
//                 where a bin is one exact key and the cut is already known.
//   6. finish   : one block per layer: exact select among the candidates, prune the ones <= cut, write info.
// Exactness does not depend on the sample: steps 3-6 count and select exactly; the sample only decides how
// narrow the bracket is, and a bracket that misses (or a candidate list that overflows because one coarse bin
// holds > SEL_CAP elements) is reported, not guessed.
// ---------------------------------------------------------------------------------------------
constexpr int SEL_SAMPLE = 16384;
constexpr int SEL_CAP = 65536;          // candidate slots per layer
constexpr unsigned KEY_INF = 0x7f800000u;   // keys above this are NaNs
struct SelState {
  unsigned int hist[RS_BINS];
  unsigned long long pool, below, eqlo, inside, nans;   // pool size; keys < lo, == lo, in (lo, hi], NaNs
  unsigned long long k, kremain;
  unsigned int lo, hi, shift, first;     // bins cover (lo, hi]: bin = (key - lo - 1) >> shift; first key of the k-th's bin
  unsigned int exact;                    // the cut is known after step 4 (a tie at lo, or one-key bins)
  unsigned int ncand, cut_bits;
  int status;                            // 0 ok, 2 = exit-2 path, 3 = bracket missed / candidates overflowed
  int located;                           // step 4 found the bin of the k-th element
};
// the streaming passes run on ONE flat grid: block b works on layer l with blk_start[l] <= b < blk_start[l + 1]
struct SelGrid { int blk_start[PRUNE_MAX_LAYERS + 1]; int nlayers; };

__device__ __forceinline__ bool in_pool(unsigned t, int cur) { return t == (unsigned)cur || t == 0u; }
__device__ __forceinline__ int sel_layer_of(const SelGrid &sg, int b) {
  int lo = 0, hi = sg.nlayers - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (sg.blk_start[mid] <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// one thread per sample: evenly spaced positions, pool members only
__global__ void __launch_bounds__(256)
sel_sample_kernel(const __grid_constant__ PruneBatch pb, int cur, unsigned int *__restrict__ samples) {
  const int layer = blockIdx.y;
  const long long n = pb.n[layer];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned key = 0xffffffffu;            // not a pool member / beyond the layer
  if (j < SEL_SAMPLE && (long long)j < n) {
    const long long i = n <= SEL_SAMPLE ? j : (long long)(((unsigned long long)j * (unsigned long long)n) / SEL_SAMPLE);
    const float wv = __ldg(pb.w[layer] + i);
    const unsigned t = pb.t[layer][i];
    if (in_pool(t, cur)) key = key_of(wv);
  }
  if (j < SEL_SAMPLE) samples[(long long)layer * SEL_SAMPLE + j] = key;
}

// rank-th smallest (0-based) of the n keys in shared memory: three digit passes (11 / 11 / 10 bits)
__device__ unsigned block_kth(const unsigned *keys, int n, unsigned rank, unsigned *hist /*[2048]*/, unsigned *scratch /*[40]*/) {
  unsigned prefix = 0;
  const int shifts[3] = {21, 10, 0}, widths[3] = {11, 11, 10};
  for (int ps = 0; ps < 3; ++ps) {
    const int sh = shifts[ps], nb = 1 << widths[ps];
    for (int i = threadIdx.x; i < RS_BINS; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned kx = keys[i];
      if (ps == 0 || (kx >> (sh + widths[ps])) == prefix) atomicAdd(&hist[(kx >> sh) & (nb - 1)], 1u);
    }
    __syncthreads();
    // inclusive prefix sums of the bins: two bins per thread (blockDim.x == 1024)
    const unsigned b0 = hist[2 * threadIdx.x], b1 = hist[2 * threadIdx.x + 1];
    unsigned incl = b0 + b1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) scratch[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      unsigned v = scratch[lane], iv = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, iv, o);
        if (lane >= o) iv += up;
      }
      scratch[lane] = iv - v;             // exclusive offset of each warp
    }
    __syncthreads();
    incl += scratch[wid];
    const unsigned before = incl - b0 - b1;       // elements in bins below 2 * threadIdx.x
    __syncthreads();
    if (before <= rank && rank < incl) {          // exactly one thread
      const unsigned bin = rank < before + b0 ? 2 * threadIdx.x : 2 * threadIdx.x + 1;
      scratch[32] = bin;
      scratch[33] = rank - (bin == 2 * threadIdx.x ? before : before + b0);
    }
    __syncthreads();
    prefix = (prefix << widths[ps]) | scratch[32];
    rank = scratch[33];
    __syncthreads();
  }
  return prefix;
}

__global__ void __launch_bounds__(1024)
sel_bracket_kernel(const unsigned int *__restrict__ samples, double ratio, SelState *__restrict__ states) {
  extern __shared__ unsigned int sk[];   // SEL_SAMPLE keys
  __shared__ unsigned int hist[RS_BINS];
  __shared__ unsigned int scratch[40];
  __shared__ unsigned int nvalid;
  SelState *stt = states + blockIdx.x;
  const unsigned int *src = samples + (long long)blockIdx.x * SEL_SAMPLE;
  if (threadIdx.x == 0) nvalid = 0;
  __syncthreads();
  unsigned cnt = 0;
  for (int i = threadIdx.x; i < SEL_SAMPLE; i += blockDim.x) {
    const unsigned v = src[i];
    sk[i] = v;
    cnt += v != 0xffffffffu;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&nvalid, cnt);
  __syncthreads();
  const unsigned mvu = nvalid;
  const double mv = (double)mvu;
  unsigned lo = 0u, hi = KEY_INF;                           // everything that is not a NaN
  if (mvu >= 64) {
    // sample ranks ratio * m -/+ 4 sigma (binomial), one rank of slack on either side
    const double r = ratio * mv;
    const double q = ratio < 0.0 ? 0.0 : ratio > 1.0 ? 1.0 : ratio;
    const double margin = 4.0 * sqrt(mv * q * (1.0 - q)) + 2.0;
    const double rl = floor(r - margin) - 1.0, rh = ceil(r + margin);
    if (rh < mv) hi = block_kth(sk, SEL_SAMPLE, (unsigned)rh, hist, scratch);   // the invalid entries sort last
    if (rl >= 0.0) lo = block_kth(sk, SEL_SAMPLE, (unsigned)rl, hist, scratch);
    else if (block_kth(sk, SEL_SAMPLE, 0u, hist, scratch) == hi) lo = hi;        // one value up to the upper quantile
    if (hi > KEY_INF) hi = KEY_INF;
    if (lo > hi) lo = hi;
  }
  if (threadIdx.x == 0) {
    unsigned shift = 0;
    while (hi > lo && ((unsigned long long)(hi - lo - 1) >> shift) >= RS_BINS) ++shift;
    stt->lo = lo; stt->hi = hi; stt->shift = shift;
  }
}

__global__ void __launch_bounds__(256)
sel_count_kernel(const __grid_constant__ PruneBatch pb, const __grid_constant__ SelGrid sg, int cur,
                 SelState *__restrict__ states) {
  __shared__ unsigned int sh[RS_BINS];
  __shared__ unsigned long long red[4][8];
  const int layer = sel_layer_of(sg, blockIdx.x);
  const float *__restrict__ w = pb.w[layer];
  const uint8_t *__restrict__ tmask = pb.t[layer];
  const long long n = pb.n[layer];
  SelState *stt = states + layer;
  const long long gx = sg.blk_start[layer + 1] - sg.blk_start[layer];
  const long long bx = blockIdx.x - sg.blk_start[layer];
  for (int i = threadIdx.x; i < RS_BINS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const unsigned lo = stt->lo, hi = stt->hi, shift = stt->shift;
  unsigned pool = 0, below = 0, eqlo = 0, inside = 0, nans = 0;
  auto one = [&](float wv, unsigned t) {
    if (!in_pool(t, cur)) return;
    ++pool;
    const unsigned key = key_of(wv);
    if (key < lo) { ++below; return; }
    if (key == lo) { ++eqlo; return; }
    if (key <= hi) { ++inside; atomicAdd(&sh[(key - lo - 1) >> shift], 1u); return; }
    if (key > KEY_INF) ++nans;
  };
  const long long stride = gx * blockDim.x;
  const long long i0 = bx * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(w) & 15) == 0) && ((reinterpret_cast<uintptr_t>(tmask) & 3) == 0);
  long long tail = 0;
  if (vec) {
    const long long n4 = n >> 2;
    long long v = i0;
    for (; v + 3 * stride < n4; v += 4 * stride) {       // four 16-byte + four 4-byte loads in flight per thread
      float4 a[4]; uchar4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = __ldg(reinterpret_cast<const float4 *>(w) + v + u * stride);
        t[u] = __ldg(reinterpret_cast<const uchar4 *>(tmask) + v + u * stride);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { one(a[u].x, t[u].x); one(a[u].y, t[u].y); one(a[u].z, t[u].z); one(a[u].w, t[u].w); }
    }
    for (; v < n4; v += stride) {
      const float4 a = __ldg(reinterpret_cast<const float4 *>(w) + v);
      const uchar4 t = __ldg(reinterpret_cast<const uchar4 *>(tmask) + v);
      one(a.x, t.x); one(a.y, t.y); one(a.z, t.z); one(a.w, t.w);
    }
    tail = n4 << 2;
  }
  for (long long i = tail + i0; i < n; i += stride) one(w[i], tmask[i]);
  // block totals of the counters
  for (int o = 16; o > 0; o >>= 1) {
    pool += __shfl_xor_sync(0xffffffffu, pool, o);
    below += __shfl_xor_sync(0xffffffffu, below, o);
    eqlo += __shfl_xor_sync(0xffffffffu, eqlo, o);
    inside += __shfl_xor_sync(0xffffffffu, inside, o);
    nans += __shfl_xor_sync(0xffffffffu, nans, o);
  }
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    red[0][wid] = pool; red[1][wid] = below; red[2][wid] = inside; red[3][wid] = eqlo;
    if (nans) atomicAdd(&stt->nans, (unsigned long long)nans);
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    unsigned long long v = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) v += red[threadIdx.x][k];
    unsigned long long *dst = threadIdx.x == 0 ? &stt->pool : threadIdx.x == 1 ? &stt->below
                              : threadIdx.x == 2 ? &stt->inside : &stt->eqlo;
    if (v) atomicAdd(dst, v);
  }
  for (int i = threadIdx.x; i < RS_BINS; i += blockDim.x) {
    const unsigned c = sh[i];
    if (c) atomicAdd(&stt->hist[i], c);
  }
}

// one block of 1024 threads per layer: k, then the bin of the k-th pool element by a block-wide prefix sum
__global__ void __launch_bounds__(1024)
sel_locate_kernel(SelState *__restrict__ states, double ratio, long long *__restrict__ info) {
  __shared__ unsigned long long wsum[32];
  SelState *stt = states + blockIdx.x;
  long long *inf = info + 4 * blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned long long pool = stt->pool;
  const double kd = rint(ratio * (double)pool);          // python round(): IEEE double product, half to even
  const long long kk = (long long)kd;
  if (threadIdx.x == 0) { stt->k = kk > 0 ? (unsigned long long)kk : 0ull; inf[1] = (long long)pool; inf[2] = kk; inf[3] = 0; }
  if (kk < 1 || (unsigned long long)kk > pool) {          // kthvalue raises -> sys.exit(2) (utils/prune.py:38-42)
    if (threadIdx.x == 0) { stt->status = 2; inf[0] = 2; }
    return;
  }
  const unsigned long long k = (unsigned long long)kk, below = stt->below;
  if (k > pool - stt->nans) {
    // the k-th smallest magnitude is a NaN (torch.kthvalue sorts them last): `abs(w) <= NaN` is false everywhere,
    // nothing is pruned (utils/prune.py:44-47)
    // (abs.f32 returns the canonical NaN 0x7fffffff for every NaN input: that is the key the radix select reports too)
    if (threadIdx.x == 0) { stt->cut_bits = 0x7fffffffu; inf[0] = 0; inf[3] = 0x7fffffffll; }
    return;
  }
  const unsigned long long eqlo = stt->eqlo;
  if (k <= below || k > below + eqlo + stt->inside) {     // the sample's bracket missed the k-th element
    if (threadIdx.x == 0) { stt->status = 3; inf[0] = 3; }
    return;
  }
  if (k <= below + eqlo) {                                 // the k-th element is one of the keys equal to lo
    if (threadIdx.x == 0) {
      stt->exact = 1; stt->cut_bits = stt->lo; stt->located = 1;
      inf[0] = 0; inf[3] = (long long)stt->lo;
    }
    return;
  }
  const unsigned long long want = k - below - eqlo;
  const unsigned long long b0 = stt->hist[2 * threadIdx.x], b1 = stt->hist[2 * threadIdx.x + 1];
  unsigned long long incl = b0 + b1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const unsigned long long v = wsum[lane];
    unsigned long long iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xffffffffu, iv, o);
      if (lane >= o) iv += up;
    }
    wsum[lane] = iv - v;
  }
  __syncthreads();
  incl += wsum[wid];
  const unsigned long long before = incl - b0 - b1;
  if (before < want && want <= incl) {                     // exactly one thread
    const bool first_bin = want <= before + b0;
    const unsigned bin = 2 * threadIdx.x + (first_bin ? 0 : 1);
    const unsigned first = stt->lo + 1u + (bin << stt->shift);
    stt->first = first;
    stt->kremain = want - (first_bin ? before : before + b0);
    stt->located = 1;
    inf[0] = 0;
    if (stt->shift == 0) {                                 // a bin is one exact key: the cut is known
      stt->exact = 1;
      stt->cut_bits = first;
      inf[3] = (long long)first;
    }
  }
}

__global__ void __launch_bounds__(256)
sel_update_kernel(const __grid_constant__ PruneBatch pb, const __grid_constant__ SelGrid sg, int cur,
                  SelState *__restrict__ states, unsigned int *__restrict__ cand_key, unsigned int *__restrict__ cand_idx) {
  const int layer = sel_layer_of(sg, blockIdx.x);
  SelState *stt = states + layer;
  if (stt->status != 0 || !stt->located) return;
  const float *__restrict__ w = pb.w[layer];
  uint8_t *__restrict__ tmask = pb.t[layer];
  const long long n = pb.n[layer];
  const long long gx = sg.blk_start[layer + 1] - sg.blk_start[layer];
  const long long bx = blockIdx.x - sg.blk_start[layer];
  // exact: everything <= cut goes.  Otherwise keys below `first` are pruned for certain and the keys in
  // [first, last] -- the bin of the k-th element -- become candidates.
  const bool exact = stt->exact != 0;
  const unsigned first = exact ? stt->cut_bits : stt->first;
  const unsigned long long last64 = (unsigned long long)first + ((1ull << stt->shift) - 1ull);
  const unsigned last = last64 > 0xffffffffull ? 0xffffffffu : (unsigned)last64;
  unsigned int *ck = cand_key + (long long)layer * SEL_CAP, *ci = cand_idx + (long long)layer * SEL_CAP;
  // returns the new mask byte; records a candidate when the key falls into the bin of the k-th element
  auto one = [&](float wv, unsigned t, long long idx) -> unsigned {
    if (!in_pool(t, cur)) return t;
    const unsigned key = key_of(wv);
    if (key < first || (exact && key == first)) return 0u;          // T == cur -> 0, T == 0 stays 0
    if (!exact && key <= last) {
      const unsigned slot = atomicAdd(&stt->ncand, 1u);             // about |bracket| / 2048 elements per layer
      if (slot < SEL_CAP) { ck[slot] = key; ci[slot] = (unsigned)idx; }
    }
    return t;
  };
  const long long stride = gx * blockDim.x;
  const long long i0 = bx * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(w) & 15) == 0) && ((reinterpret_cast<uintptr_t>(tmask) & 3) == 0);
  long long tail = 0;
  if (vec) {
    const long long n4 = n >> 2;
    long long v = i0;
    for (; v + 3 * stride < n4; v += 4 * stride) {
      float4 a[4]; uchar4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = __ldg(reinterpret_cast<const float4 *>(w) + v + u * stride);
        t[u] = *(reinterpret_cast<const uchar4 *>(tmask) + v + u * stride);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long e = (v + u * stride) << 2;
        uchar4 o;
        o.x = (unsigned char)one(a[u].x, t[u].x, e); o.y = (unsigned char)one(a[u].y, t[u].y, e + 1);
        o.z = (unsigned char)one(a[u].z, t[u].z, e + 2); o.w = (unsigned char)one(a[u].w, t[u].w, e + 3);
        if (o.x != t[u].x || o.y != t[u].y || o.z != t[u].z || o.w != t[u].w)
          *(reinterpret_cast<uchar4 *>(tmask) + v + u * stride) = o;
      }
    }
    for (; v < n4; v += stride) {
      const float4 a = __ldg(reinterpret_cast<const float4 *>(w) + v);
      const uchar4 t = *(reinterpret_cast<const uchar4 *>(tmask) + v);
      const long long e = v << 2;
      uchar4 o;
      o.x = (unsigned char)one(a.x, t.x, e); o.y = (unsigned char)one(a.y, t.y, e + 1);
      o.z = (unsigned char)one(a.z, t.z, e + 2); o.w = (unsigned char)one(a.w, t.w, e + 3);
      if (o.x != t.x || o.y != t.y || o.z != t.z || o.w != t.w) *(reinterpret_cast<uchar4 *>(tmask) + v) = o;
    }
    tail = n4 << 2;
  }
  for (long long i = tail + i0; i < n; i += stride) {
    const unsigned t = tmask[i], o = one(w[i], t, i);
    if (o != t) tmask[i] = (uint8_t)o;
  }
}

// one block per layer: exact kremain-th smallest key among the candidates (all share the key bits above `shift`:
// a radix select over the low bits, 8 bits per round), then prune the candidates <= cut.
__global__ void __launch_bounds__(1024)
sel_finish_kernel(const __grid_constant__ PruneBatch pb, int cur, SelState *__restrict__ states,
                  const unsigned int *__restrict__ cand_key, const unsigned int *__restrict__ cand_idx,
                  long long *__restrict__ info) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_remain;
  const int layer = blockIdx.x;
  SelState *stt = states + layer;
  long long *inf = info + 4 * layer;
  if (stt->status != 0 || !stt->located || stt->exact) return;
  const unsigned nc = stt->ncand;
  if (nc > SEL_CAP) {                                       // one coarse bin held too many keys: exact path instead
    if (threadIdx.x == 0) { stt->status = 3; inf[0] = 3; }
    return;
  }
  const unsigned int *ck = cand_key + (long long)layer * SEL_CAP, *ci = cand_idx + (long long)layer * SEL_CAP;
  uint8_t *__restrict__ tmask = pb.t[layer];
  const unsigned shift = stt->shift;
  const unsigned base_key = stt->first;
  if (threadIdx.x == 0) { s_prefix = 0; s_remain = (unsigned)stt->kremain; }
  // digits of (key - base_key), most significant first
  const int rounds = (int)((shift + 7) / 8);
  for (int rd = rounds - 1; rd >= 0; --rd) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix;
    for (unsigned i = threadIdx.x; i < nc; i += blockDim.x) {
      const unsigned d = ck[i] - base_key;
      if (rd == rounds - 1 || (d >> (8 * (rd + 1))) == prefix) atomicAdd(&hist[(d >> (8 * rd)) & 0xffu], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {                                 // warp 0: 8 bins per lane, shuffle scan over the lanes
      const unsigned remain = s_remain;
      unsigned c[8], tot = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = hist[threadIdx.x * 8 + j]; tot += c[j]; }
      unsigned incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)threadIdx.x >= o) incl += up;
      }
      unsigned acc = incl - tot;
      __syncwarp();                                         // every lane has read s_remain / s_prefix
      if (acc < remain && remain <= incl) {                 // exactly one lane
        int j = 0;
        for (; j < 7; ++j) { if (acc + c[j] >= remain) break; acc += c[j]; }
        s_prefix = (prefix << 8) | (threadIdx.x * 8 + j);
        s_remain = remain - acc;
      }
    }
    __syncthreads();
  }
  const unsigned cut = base_key + s_prefix;
  for (unsigned i = threadIdx.x; i < nc; i += blockDim.x)
    if (ck[i] <= cut) { const unsigned idx = ci[i]; if (tmask[idx] == (uint8_t)cur) tmask[idx] = 0; }
  if (threadIdx.x == 0) { stt->cut_bits = cut; inf[3] = (long long)cut; }
}

}  // namespace cpgb

using namespace cpgb;

extern "C" {

size_t cpgb_prune_workspace_bytes(void) { return sizeof(PruneState); }

int cpgb_prune_select(const float *w, uint8_t *tmask, int64_t n, int32_t cur, double ratio, int64_t *info,
                      void *ws, size_t ws_bytes, void *stream) {
  if (n < 0 || !info || !ws || (n > 0 && (!w || !tmask))) { set_error("cpgb_prune_select: null pointer"); return CPGB_EINVAL; }
  if (ws_bytes < sizeof(PruneState)) { set_error("cpgb_prune_select: workspace %zu < %zu", ws_bytes, sizeof(PruneState)); return CPGB_EWORKSPACE; }
  if (cur < 0 || cur > 255) { set_error("cpgb_prune_select: cur out of range"); return CPGB_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  PruneState *stt = reinterpret_cast<PruneState *>(ws);
  CPGB_CUDA_OK(cudaMemsetAsync(stt, 0, sizeof(PruneState), st));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long want = (n / 4 + 255) / 256;
  int grid = (int)(want < 1 ? 1 : (want > (long long)sms * 8 ? (long long)sms * 8 : want));
  long long *inf = reinterpret_cast<long long *>(info);
  prune_hist_kernel<0><<<grid, 256, 0, st>>>(w, tmask, n, cur, stt);
  prune_scan_kernel<0><<<1, 32, 0, st>>>(stt, ratio, inf);
  prune_hist_kernel<1><<<grid, 256, 0, st>>>(w, tmask, n, cur, stt);
  prune_scan_kernel<1><<<1, 32, 0, st>>>(stt, ratio, inf);
  prune_hist_kernel<2><<<grid, 256, 0, st>>>(w, tmask, n, cur, stt);
  prune_scan_kernel<2><<<1, 32, 0, st>>>(stt, ratio, inf);
  prune_update_kernel<<<grid, 256, 0, st>>>(w, tmask, n, cur, stt);
  CPGB_LAUNCH_OK_N("cpgb_prune_select", 7);
  return CPGB_OK;
}

size_t cpgb_prune_batched_workspace_bytes(int32_t nlayers) {
  return nlayers > 0 ? (size_t)nlayers * sizeof(PruneState) : 0;
}

int cpgb_prune_select_batched(int32_t nlayers, const float *const *w, uint8_t *const *tmask, const int64_t *n,
                              int32_t cur, double ratio, int64_t *info, void *ws, size_t ws_bytes, void *stream) {
  if (nlayers < 0 || nlayers > PRUNE_MAX_LAYERS) { set_error("cpgb_prune_select_batched: 0..%d layers", PRUNE_MAX_LAYERS); return CPGB_EINVAL; }
  if (nlayers == 0) return CPGB_OK;
  if (!w || !tmask || !n || !info || !ws) { set_error("cpgb_prune_select_batched: null pointer"); return CPGB_EINVAL; }
  if (ws_bytes < (size_t)nlayers * sizeof(PruneState)) { set_error("cpgb_prune_select_batched: workspace too small"); return CPGB_EWORKSPACE; }
  if (cur < 0 || cur > 255) { set_error("cpgb_prune_select_batched: cur out of range"); return CPGB_EINVAL; }
  PruneBatch pb;
  long long nmax = 0;
  for (int i = 0; i < nlayers; ++i) {
    if (n[i] < 0 || (n[i] > 0 && (!w[i] || !tmask[i]))) { set_error("cpgb_prune_select_batched: bad layer %d", i); return CPGB_EINVAL; }
    pb.w[i] = w[i]; pb.t[i] = tmask[i]; pb.n[i] = n[i];
    if (n[i] > nmax) nmax = n[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  PruneState *stt = reinterpret_cast<PruneState *>(ws);
  CPGB_CUDA_OK(cudaMemsetAsync(stt, 0, (size_t)nlayers * sizeof(PruneState), st));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long want = (nmax / 4 + 255) / 256;
  int gx = (int)(want < 1 ? 1 : (want > (long long)sms * 4 ? (long long)sms * 4 : want));
  dim3 grid(gx, nlayers);
  long long *inf = reinterpret_cast<long long *>(info);
  prune_hist_batched_kernel<0><<<grid, 256, 0, st>>>(pb, cur, stt);
  prune_scan_batched_kernel<0><<<nlayers, 32, 0, st>>>(stt, ratio, inf);
  prune_hist_batched_kernel<1><<<grid, 256, 0, st>>>(pb, cur, stt);
  prune_scan_batched_kernel<1><<<nlayers, 32, 0, st>>>(stt, ratio, inf);
  prune_hist_batched_kernel<2><<<grid, 256, 0, st>>>(pb, cur, stt);
  prune_scan_batched_kernel<2><<<nlayers, 32, 0, st>>>(stt, ratio, inf);
  prune_update_batched_kernel<<<grid, 256, 0, st>>>(pb, cur, stt);
  CPGB_LAUNCH_OK_N("cpgb_prune_select_batched", 7);
  return CPGB_OK;
}

size_t cpgb_prune_sampled_workspace_bytes(int32_t nlayers) {
  if (nlayers <= 0) return 0;
  return (size_t)nlayers * (sizeof(SelState) + (size_t)SEL_SAMPLE * 4 + (size_t)SEL_CAP * 8) + 256;
}

int cpgb_prune_select_sampled(int32_t nlayers, const float *const *w, uint8_t *const *tmask, const int64_t *n,
                              int32_t cur, double ratio, int64_t *info, void *ws, size_t ws_bytes, void *stream) {
  if (nlayers < 0 || nlayers > PRUNE_MAX_LAYERS) { set_error("cpgb_prune_select_sampled: 0..%d layers", PRUNE_MAX_LAYERS); return CPGB_EINVAL; }
  if (nlayers == 0) return CPGB_OK;
  if (!w || !tmask || !n || !info || !ws) { set_error("cpgb_prune_select_sampled: null pointer"); return CPGB_EINVAL; }
  if (ws_bytes < cpgb_prune_sampled_workspace_bytes(nlayers) || (reinterpret_cast<uintptr_t>(ws) & 7)) {
    set_error("cpgb_prune_select_sampled: workspace too small or misaligned"); return CPGB_EWORKSPACE;
  }
  if (cur < 0 || cur > 255) { set_error("cpgb_prune_select_sampled: cur out of range"); return CPGB_EINVAL; }
  PruneBatch pb;
  long long nmax = 0;
  for (int i = 0; i < nlayers; ++i) {
    if (n[i] < 0 || n[i] > 0xffffffffll || (n[i] > 0 && (!w[i] || !tmask[i]))) {
      set_error("cpgb_prune_select_sampled: bad layer %d", i); return CPGB_EINVAL;
    }
    pb.w[i] = w[i]; pb.t[i] = tmask[i]; pb.n[i] = n[i];
    if (n[i] > nmax) nmax = n[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  SelState *stt = reinterpret_cast<SelState *>(ws);
  unsigned int *samples = reinterpret_cast<unsigned int *>(stt + nlayers);
  unsigned int *cand_key = samples + (size_t)nlayers * SEL_SAMPLE;
  unsigned int *cand_idx = cand_key + (size_t)nlayers * SEL_CAP;
  CPGB_CUDA_OK(cudaMemsetAsync(stt, 0, (size_t)nlayers * sizeof(SelState), st));
  static PerDeviceOnce attr_once;
  if (attr_once.need())
    CPGB_CUDA_OK(cudaFuncSetAttribute(sel_bracket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEL_SAMPLE * 4));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // flat grid of the two streaming passes: about 8 blocks per SM in total, shared out by layer size (each block
  // walks >= 16 elements per thread)
  long long ntot = 0;
  for (int i = 0; i < nlayers; ++i) ntot += n[i];
  const long long budget = (long long)sms * 8;
  SelGrid sg;
  sg.nlayers = nlayers;
  int acc = 0;
  for (int i = 0; i < nlayers; ++i) {
    sg.blk_start[i] = acc;
    long long want = (n[i] / 16 + 255) / 256;                       // blocks that still have 16 elements per thread
    long long share = ntot > 0 ? (n[i] * budget + ntot - 1) / ntot : 1;
    long long b = want < share ? want : share;
    if (b < 1) b = 1;
    acc += (int)b;
  }
  sg.blk_start[nlayers] = acc;
  long long *inf = reinterpret_cast<long long *>(info);
  sel_sample_kernel<<<dim3(SEL_SAMPLE / 256, nlayers), 256, 0, st>>>(pb, cur, samples);
  sel_bracket_kernel<<<nlayers, 1024, SEL_SAMPLE * 4, st>>>(samples, ratio, stt);
  sel_count_kernel<<<acc, 256, 0, st>>>(pb, sg, cur, stt);
  sel_locate_kernel<<<nlayers, 1024, 0, st>>>(stt, ratio, inf);
  sel_update_kernel<<<acc, 256, 0, st>>>(pb, sg, cur, stt, cand_key, cand_idx);
  sel_finish_kernel<<<nlayers, 1024, 0, st>>>(pb, cur, stt, cand_key, cand_idx, inf);
  CPGB_LAUNCH_OK_N("cpgb_prune_select_sampled", 6);
  return CPGB_OK;
}

}  // extern "C"
