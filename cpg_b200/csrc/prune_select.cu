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
  long long nblk = (n + blockDim.x - 1) / blockDim.x;
  const long long gx = nblk < (long long)gridDim.x ? nblk : (long long)gridDim.x;
  if ((long long)blockIdx.x >= gx) return;
  const long long stride = gx * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (tmask[i] == (uint8_t)cur && key_of(w[i]) <= cut) tmask[i] = 0;
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

}  // extern "C"
