// tcgen05 (5th-gen tensor core) implicit-GEMM kernels for the masked convolution / linear path.
//
//   fprop : Y[pix, k]  = sum_{tap, c}  X[pix + tap, c]        * Wt[k, tap, c]      (models/layers.py:108)
//   dgrad : dX[pix, c] = sum_{tap, k}  dY[pix - tap, k]       * Wt[k, tap, c]      (autograd of the above)
//   wgrad : G[k, tap, c] = sum_{pix}   dY[pix, k]             * X[pix + tap, c]    (autograd of the above)
//
// Wt is the *staged* operand: (piggymask > thr ? 1 : 0) * W, rounded to TF32 (rna) and reordered
// from the module's [K][C][R][S] to [K][R*S][Cp] (Cp = C rounded up to 32, zero padded) by
// stage_weights_kernel -- one pass per layer per step that lives in L2 between its producer and
// the GEMMs that read it.  Activations are read in place (NHWC / torch.channels_last) by TMA:
// the im2col gather of a stride-1 convolution is a plain 4-D box load whose (w, h) start
// coordinate is shifted by the filter tap; out-of-bounds rows/columns (the zero padding) are
// filled by the TMA unit.  MMA: tcgen05.mma kind::tf32, M = 128, FP32 accumulators in TMEM.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> registers -> global).  smem ring of NSTAGE stages, each
// guarded by a full (TMA -> MMA) and an empty (tcgen05.commit -> TMA) mbarrier.
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace cpgb {

using namespace ptx;

// ------------------------------------------------------------------------------------------
// driver entry point for cuTensorMapEncodeTiled (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

static EncodeTiledFn get_encode() {
  std::call_once(g_encode_once, [] {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  return g_encode;
}

// MN-major operand description (see ptx.cuh); overridable through cpgb_debug_set_mn for bring-up
struct MnDesc { int layout, lbo, sbo, kadv, tma_swizzle; };
static MnDesc g_mn = {1, 4096, 512, 1024, (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B};
void debug_set_mn(int layout, int lbo, int sbo, int kadv, int tma_swizzle) {
  g_mn = MnDesc{layout, lbo, sbo, kadv, tma_swizzle};
}

// rank-`rank` fp32 tensor map with 128-byte swizzle; dims/strides innermost first; strides[0] is
// implied (4 bytes) -- `strides_bytes[i]` is the stride of dim i+1.
static int make_map(CUtensorMap *m, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                    const uint32_t *box, bool mn_major = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return CPGB_ECUDA; }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? (CUtensorMapSwizzle)g_mn.tma_swizzle : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]", (int)r,
              rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
              (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0),
              (unsigned long long)(rank > 4 ? gd[4] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
              rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
    return CPGB_ECUDA;
  }
  return CPGB_OK;
}

static inline int ilog2_ceil(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
static inline int cdiv_i(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// A tile of T (power of two) output pixels as a (q, p, n) box.
struct PixBox { int lq, lp, ln; int tq, tp, tn; };
static PixBox make_pixbox(int T, int Q, int P, int N) {
  PixBox b;
  int lt = ilog2_ceil(T);
  b.lq = ilog2_ceil(Q); if (b.lq > lt) b.lq = lt;
  b.lp = ilog2_ceil(P); if (b.lp > lt - b.lq) b.lp = lt - b.lq;
  b.ln = lt - b.lq - b.lp;
  b.tq = cdiv_i(Q, 1 << b.lq); b.tp = cdiv_i(P, 1 << b.lp); b.tn = cdiv_i(N, 1 << b.ln);
  return b;
}

// ------------------------------------------------------------------------------------------
// weight staging:  Wt[k][t][c] = tf32_rna(binarize(P[k][c][t]) * W[k][c][t]),  c padded to Cp
// ------------------------------------------------------------------------------------------
constexpr int STAGE_CC = 128;  // channels per block
__global__ void __launch_bounds__(256)
stage_weights_kernel(const float *__restrict__ w, const float *__restrict__ piggy, float *__restrict__ wt, int C,
                     int Cp, int RS, float thr) {
  extern __shared__ float sh[];  // [STAGE_CC][RS]  (+1 padding per row when RS is even)
  const int k = blockIdx.x, c0 = blockIdx.y * STAGE_CC;
  const int cc = min(STAGE_CC, Cp - c0);       // channels this block writes (incl. zero padding)
  const int cv = max(0, min(STAGE_CC, C - c0)); // channels that exist in W
  const int ld = RS | 1;
  const long long base = ((long long)k * C + c0) * RS;
  for (int i = threadIdx.x; i < cv * RS; i += blockDim.x) {
    float v = masked_weight(__ldg(w + base + i), piggy, base + i, thr);
    sh[(i / RS) * ld + (i % RS)] = to_tf32_rna(v);
  }
  __syncthreads();
  float *dst = wt + (long long)k * RS * Cp + c0;
  for (int i = threadIdx.x; i < cc * RS; i += blockDim.x) {
    int t = i / cc, c = i - t * cc;
    dst[(long long)t * Cp + c] = c < cv ? sh[c * ld + t] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// fprop / dgrad kernel
// ------------------------------------------------------------------------------------------
struct ConvGemmParams {
  int tq, tp, tn;          // pixel-tile counts
  int lq, lp;              // log2 box extents along q and p (n extent = 128 >> (lq + lp))
  int Qo, Po, No;          // output pixel extents
  int S, taps;             // filter width, R*S
  int off_h, off_w;        // input coord = output coord + off + tap_index * step
  int step_h, step_w;
  int kblocks;             // 32-wide reduction blocks per tap
  int ncols;               // valid output channels
  int mn_layout, mn_lbo, mn_sbo, mn_kadv;   // MN-major operand descriptor fields
  long long o_sn, o_sh, o_sw;
  float *out;
  const float *bias;
};

constexpr int A_TILE_BYTES = 128 * 128;  // 128 pixels x 32 fp32

template <int BN, bool B_MN>
struct ConvGemmCfg {
  static constexpr int B_TILE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  // two CTAs per SM when the tile is narrow (hides prologue/epilogue), one otherwise
  static constexpr int NSTAGE = BN <= 128 ? 3 : 4;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int BN, bool B_MN>
__global__ void __launch_bounds__(192)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvGemmParams p) {
  using Cfg = ConvGemmCfg<BN, B_MN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + Cfg::NSTAGE * Cfg::STAGE_BYTES);
  uint64_t *empty = full + Cfg::NSTAGE;
  uint64_t *acc_full = empty + Cfg::NSTAGE;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile coordinates
  int tile = blockIdx.x;
  const int tqi = tile % p.tq; tile /= p.tq;
  const int tpi = tile % p.tp; tile /= p.tp;
  const int tni = tile;
  const int q0 = tqi << p.lq, p0 = tpi << p.lp, n0 = tni << (7 - p.lq - p.lp);
  const int col0 = blockIdx.y * BN;
  const int iters = p.taps * p.kblocks;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    for (int s = 0; s < Cfg::NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        const int t = it / p.kblocks, kb = it - t * p.kblocks;
        const int r = t / p.S, s = t - r * p.S;
        mbar_wait(empty + stage, phase ^ 1);
        uint8_t *sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t *sb = sa + A_TILE_BYTES;
        mbar_arrive_expect_tx(full + stage, Cfg::STAGE_BYTES);
        tma_load_4d(sa, &tmA, full + stage, kb * 32, q0 + p.off_w + s * p.step_w, p0 + p.off_h + r * p.step_h, n0);
        if (!B_MN) tma_load_3d(sb, &tmB, full + stage, kb * 32, t, col0);               // [BN k][32 c]
        else       tma_load_4d(sb, &tmB, full + stage, 0, kb * 32, t, col0 >> 5);       // [BN/32][32 k][32 c]
        if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, BN, false, B_MN);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(full + stage, phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = make_smem_desc(sa + ks * 32, 16, 1024);
          const uint64_t bd = B_MN ? make_smem_desc(sb + ks * p.mn_kadv, p.mn_lbo, p.mn_sbo, p.mn_layout)
                                   : make_smem_desc(sb + ks * 32, 16, 1024);
          mma_tf32_ss(tmem_base, ad, bd, idesc, (it | ks) != 0);
        }
        mma_commit(empty + stage);
        if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
      }
      mma_commit(acc_full);
    }
  } else {
    // epilogue: warp w reads TMEM lanes [32*(w%4), +32)
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const int qi = m & ((1 << p.lq) - 1);
    const int pi = (m >> p.lq) & ((1 << p.lp) - 1);
    const int ni = m >> (p.lq + p.lp);
    const int q = q0 + qi, pp = p0 + pi, n = n0 + ni;
    const bool valid = q < p.Qo && pp < p.Po && n < p.No;
    float *orow = p.out + n * p.o_sn + pp * p.o_sh + q * p.o_sw + col0;
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int col = col0 + c + j;
          if (col < p.ncols) {   // ncols % 4 == 0
            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (p.bias) {
              const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + col));
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            *reinterpret_cast<float4 *>(orow + c + j) = o;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// wgrad kernel: G[split][k][tap][c] partial sums over a range of 32-pixel chunks
// ------------------------------------------------------------------------------------------
struct WgradParams {
  int cq, cp, cn;          // chunk counts along q, p, n
  int lq, lp;              // log2 chunk extents (n extent = 32 >> (lq + lp))
  int S, RS;               // filter width, taps
  int pad_h, pad_w, dil_h, dil_w;
  int chunks, chunks_per_split;
  int ctiles;              // number of BN-wide channel tiles
  int K, C;
  int mn_layout, mn_lbo, mn_sbo, mn_kadv;
  float *gpart;            // [splits][K][RS][C]
};

template <int BN, int TG>
struct WgradCfg {
  static constexpr int A_BYTES = 128 * 128;            // [4 blocks][32 pixels][32 k]
  static constexpr int B_BYTES = BN * 128;             // [BN/32 blocks][32 pixels][32 c]
  static constexpr int STAGE_BYTES = A_BYTES + TG * B_BYTES;
  static constexpr int NSTAGE = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 + 256;
  static constexpr int ACC_COLS = TG * BN;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128
                                   : ACC_COLS <= 256 ? 256 : 512;
};

// grid: x = ktiles * ctiles, y = filter rows R, z = splits.  One CTA accumulates the TG = S taps of
// filter row r for a 128(k) x BN(c) tile over its range of pixel chunks.
template <int BN, int TG>
__global__ void __launch_bounds__(192)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                  const WgradParams p) {
  using Cfg = WgradCfg<BN, TG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + Cfg::NSTAGE * Cfg::STAGE_BYTES);
  uint64_t *empty = full + Cfg::NSTAGE;
  uint64_t *acc_full = empty + Cfg::NSTAGE;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x / p.ctiles, ct = blockIdx.x - kt * p.ctiles;
  const int k0 = kt * 128, c0 = ct * BN;
  const int r = blockIdx.y;
  const int split = blockIdx.z;
  const int ch_beg = split * p.chunks_per_split;
  const int ch_end = min(p.chunks, ch_beg + p.chunks_per_split);
  const int iters = ch_end - ch_beg;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmDY);
    prefetch_tensormap(&tmX);
    for (int s = 0; s < Cfg::NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        int ch = ch_beg + it;
        const int cqi = ch % p.cq; ch /= p.cq;
        const int cpi = ch % p.cp; ch /= p.cp;
        const int q0 = cqi << p.lq, p0 = cpi << p.lp, n0 = ch << (5 - p.lq - p.lp);
        mbar_wait(empty + stage, phase ^ 1);
        uint8_t *sa = smem + stage * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(full + stage, Cfg::STAGE_BYTES);
        tma_load_5d(sa, &tmDY, full + stage, 0, q0, p0, n0, k0 >> 5);
#pragma unroll
        for (int s = 0; s < TG; ++s)
          tma_load_5d(sa + Cfg::A_BYTES + s * Cfg::B_BYTES, &tmX, full + stage, 0, q0 - p.pad_w + s * p.dil_w,
                      p0 - p.pad_h + r * p.dil_h, n0, c0 >> 5);
        if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, BN, true, true);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(full + stage, phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
#pragma unroll
        for (int s = 0; s < TG; ++s) {
          const uint32_t sb = sa + Cfg::A_BYTES + s * Cfg::B_BYTES;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = make_smem_desc(sa + ks * p.mn_kadv, p.mn_lbo, p.mn_sbo, p.mn_layout);
            const uint64_t bd = make_smem_desc(sb + ks * p.mn_kadv, p.mn_lbo, p.mn_sbo, p.mn_layout);
            mma_tf32_ss(tmem_base + s * BN, ad, bd, idesc, (it | ks) != 0);
          }
        }
        mma_commit(empty + stage);
        if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
      }
      mma_commit(acc_full);
    }
  } else {
    const int quad = warp & 3;
    const int k = k0 + quad * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int s = 0; s < TG; ++s) {
      float *grow = p.gpart + (((long long)split * p.K + k) * p.RS + (r * p.S + s)) * p.C + c0;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + s * BN + c, v);
        tmem_ld_wait();
        if (k < p.K) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (c0 + c + j < p.C)
              *reinterpret_cast<float4 *>(grow + c + j) = iters > 0 ? make_float4(v[j], v[j + 1], v[j + 2], v[j + 3])
                                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// fused wgrad epilogue over the [split][K][RS][C] partial sums:
//   g = sum_s part[s];  dW = (g*b + wd*W)[T==cur] ...  written in the module's [K][C][R][S] order
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void epi_one_tc(float g, float w, float pv, bool has_p, unsigned t, int cur, float wd,
                                           int mode, float thr, float &dw, float &dp) {
  float gb = has_p ? g * binarize_val(pv, thr) : g;
  if (mode == CPGB_GRAD_RAW) { dw = gb; dp = g * w; return; }
  dw = (t == (unsigned)cur) ? fmaf(wd, w, gb) : 0.f;
  dp = (mode == CPGB_GRAD_FINETUNE && t != 0u && t < (unsigned)cur) ? g * w : 0.f;
}

constexpr int EPI_CC = 64;
__global__ void __launch_bounds__(256)
wgrad_epilogue_krsc_kernel(const float *__restrict__ gpart, int splits, int K, int C, int RS,
                           const float *__restrict__ w, const float *__restrict__ piggy,
                           const uint8_t *__restrict__ tmask, int cur, float wd, int mode, float thr,
                           float *__restrict__ dW, float *__restrict__ dP) {
  extern __shared__ float sh[];  // [RS][EPI_CC + 1]
  const int k = blockIdx.x, c0 = blockIdx.y * EPI_CC;
  const int cc = min(EPI_CC, C - c0);
  const long long split_stride = (long long)K * RS * C;
  for (int i = threadIdx.x; i < RS * cc; i += blockDim.x) {
    const int t = i / cc, c = i - t * cc;
    const float *src = gpart + ((long long)k * RS + t) * C + c0 + c;
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += __ldg(src + sp * split_stride);
    sh[t * (EPI_CC + 1) + c] = s;
  }
  __syncthreads();
  const long long base = ((long long)k * C + c0) * RS;
  const bool has_p = piggy != nullptr;
  for (int i = threadIdx.x; i < cc * RS; i += blockDim.x) {
    const int c = i / RS, t = i - c * RS;
    const float g = sh[t * (EPI_CC + 1) + c];
    const long long idx = base + i;
    float ow, op;
    epi_one_tc(g, __ldg(w + idx), has_p ? __ldg(piggy + idx) : 0.f, has_p, tmask ? tmask[idx] : 0u, cur, wd, mode,
               thr, ow, op);
    dW[idx] = ow;
    if (dP) dP[idx] = op;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static bool aligned16p(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Strides (n, c, h, w) with the don't-care strides of size-1 dims replaced by dense NHWC ones
// (torch reports arbitrary strides there; TMA wants multiples of 16 bytes everywhere).
struct Str4 { int64_t s[4]; };
static Str4 norm_strides(const int64_t in[4], int C, int H, int W, int N) {
  Str4 o;
  o.s[1] = in[1];
  o.s[3] = W == 1 ? C : in[3];
  o.s[2] = H == 1 ? o.s[3] * W : in[2];
  o.s[0] = N == 1 ? o.s[2] * H : in[0];
  if (C == 1) o.s[1] = 1;
  return o;
}
static bool nhwc_ok(const Str4 &st, int C) {
  // channel stride 1, every other stride a multiple of 4 elements (16 bytes)
  const int64_t *s = st.s;
  return s[1] == 1 && s[0] % 4 == 0 && s[2] % 4 == 0 && s[3] % 4 == 0 && s[3] >= C;
}
static Str4 x_strides(const cpgb_conv_desc &d) { return norm_strides(d.xs, d.C, d.H, d.W, d.N); }
static Str4 y_strides(const cpgb_conv_desc &d) { return norm_strides(d.ys, d.K, d.P, d.Q, d.N); }

bool tc_eligible(const cpgb_conv_desc &d, int op) {
  if (d.groups != 1 || d.stride_h != 1 || d.stride_w != 1) return false;
  if (d.C % 4 || d.K % 4) return false;
  if (d.R * d.S > 49 || d.N < 1) return false;
  if (!nhwc_ok(x_strides(d), d.C) || !nhwc_ok(y_strides(d), d.K)) return false;
  if (d.W > 4096 || d.H > 4096) return false;
  if (op == 2) {
    if (d.C % 32 || d.K % 32) return false;
    if (d.S != 1 && d.S != 3) return false;
  }
  return true;
}

static inline int cp_of(const cpgb_conv_desc &d) { return (d.C + 31) / 32 * 32; }

size_t tc_staged_bytes(const cpgb_conv_desc &d) {
  return align_up((size_t)d.K * d.R * d.S * cp_of(d) * sizeof(float) + 256, 256);
}

struct WgradPlan { int BN, ctiles, ktiles, splits, chunks, cps; PixBox box; };
static WgradPlan plan_wgrad(const cpgb_conv_desc &d) {
  WgradPlan pl;
  pl.BN = d.C >= 128 ? 128 : 64;
  pl.ctiles = cdiv_i(d.C, pl.BN);
  pl.ktiles = cdiv_i(d.K, 128);
  pl.box = make_pixbox(32, d.Q, d.P, d.N);
  pl.chunks = pl.box.tq * pl.box.tp * pl.box.tn;
  const int base = pl.ctiles * pl.ktiles * d.R;
  int splits = cdiv_i(2 * 148, base);
  if (splits > pl.chunks) splits = pl.chunks;
  if (splits < 1) splits = 1;
  pl.cps = cdiv_i(pl.chunks, splits);
  pl.splits = cdiv_i(pl.chunks, pl.cps);
  return pl;
}

size_t tc_workspace_bytes(const cpgb_conv_desc &d) {
  if (d.groups <= 0) return 0;
  size_t b = 0;
  if (tc_eligible(d, 0) || tc_eligible(d, 1)) b = tc_staged_bytes(d);
  if (tc_eligible(d, 2)) {
    WgradPlan pl = plan_wgrad(d);
    size_t g = (size_t)pl.splits * d.K * d.R * d.S * d.C * sizeof(float);
    if (g > b) b = g;
  }
  return align_up(b, 256);
}

int tc_stage_weights(const cpgb_conv_desc &d, const float *w, const float *piggy, float thr, void *staged,
                     size_t bytes, cudaStream_t st) {
  if (bytes < tc_staged_bytes(d)) { set_error("staged-weight buffer %zu < %zu", bytes, tc_staged_bytes(d)); return CPGB_EWORKSPACE; }
  const int RS = d.R * d.S, Cp = cp_of(d);
  dim3 grid(d.K, cdiv_i(Cp, STAGE_CC));
  size_t sh = (size_t)STAGE_CC * (RS | 1) * sizeof(float);
  stage_weights_kernel<<<grid, 256, sh, st>>>(w, piggy, reinterpret_cast<float *>(staged), d.C, Cp, RS, thr);
  CPGB_LAUNCH_OK("stage_weights");
  return CPGB_OK;
}

template <int BN, bool B_MN>
static int launch_conv_gemm(const CUtensorMap &ta, const CUtensorMap &tb, const ConvGemmParams &p, int ntiles_n,
                            cudaStream_t st) {
  using Cfg = ConvGemmCfg<BN, B_MN>;
  static bool attr_done = false;
  if (!attr_done) {
    CPGB_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<BN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid(p.tq * p.tp * p.tn, ntiles_n);
  conv_gemm_kernel<BN, B_MN><<<grid, 192, Cfg::SMEM_BYTES, st>>>(ta, tb, p);
  CPGB_LAUNCH_OK("conv_gemm_kernel");
  return CPGB_OK;
}

static int pick_bn(int ncols, long long mtiles) {
  // widest tile that still gives ~one wave of CTAs
  if (ncols > 128 && mtiles * cdiv_i(ncols, 256) >= 120) return 256;
  if (ncols > 64 && mtiles * cdiv_i(ncols, 128) >= 100) return 128;
  if (ncols > 128 && mtiles * cdiv_i(ncols, 64) < 64) return 128;
  return 64;
}

// map of an NHWC activation tensor: dims (C, W, H, N)
static int make_act_map(CUtensorMap *m, const float *base, int C, int W, int H, int N, const Str4 &sv,
                        const PixBox &b) {
  const int64_t *s = sv.s;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t str[3] = {(uint64_t)s[3] * 4, (uint64_t)s[2] * 4, (uint64_t)s[0] * 4};
  uint32_t box[4] = {32, 1u << b.lq, 1u << b.lp, 1u << b.ln};
  return make_map(m, base, 4, dims, str, box);
}

int tc_fprop(const cpgb_conv_desc &d, const float *x, const float *staged, const float *bias, float *y,
             cudaStream_t st) {
  if (!aligned16p(x) || !aligned16p(y) || !aligned16p(staged) || (bias && !aligned16p(bias))) {
    set_error("tcgen05 path needs 16-byte aligned tensors"); return CPGB_EINVAL;
  }
  const int RS = d.R * d.S, Cp = cp_of(d);
  PixBox b = make_pixbox(128, d.Q, d.P, d.N);
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_act_map(&ta, x, d.C, d.W, d.H, d.N, x_strides(d), b))) return rc;
  const long long mtiles = (long long)b.tq * b.tp * b.tn;
  const int BN = pick_bn(d.K, mtiles);
  {
    uint64_t dims[3] = {(uint64_t)Cp, (uint64_t)RS, (uint64_t)d.K};
    uint64_t str[2] = {(uint64_t)Cp * 4, (uint64_t)RS * Cp * 4};
    uint32_t box[3] = {32, 1, (uint32_t)BN};
    if ((rc = make_map(&tb, staged, 3, dims, str, box))) return rc;
  }
  ConvGemmParams p;
  p.tq = b.tq; p.tp = b.tp; p.tn = b.tn; p.lq = b.lq; p.lp = b.lp;
  p.Qo = d.Q; p.Po = d.P; p.No = d.N; p.S = d.S; p.taps = RS;
  p.off_h = -d.pad_h; p.off_w = -d.pad_w; p.step_h = d.dil_h; p.step_w = d.dil_w;
  p.kblocks = Cp / 32; p.ncols = d.K;
  p.mn_layout = g_mn.layout; p.mn_lbo = g_mn.lbo; p.mn_sbo = g_mn.sbo; p.mn_kadv = g_mn.kadv;
  { Str4 ys = y_strides(d); p.o_sn = ys.s[0]; p.o_sh = ys.s[2]; p.o_sw = ys.s[3]; }
  p.out = y; p.bias = bias;
  const int nt = cdiv_i(d.K, BN);
  if (BN == 256) return launch_conv_gemm<256, false>(ta, tb, p, nt, st);
  if (BN == 128) return launch_conv_gemm<128, false>(ta, tb, p, nt, st);
  return launch_conv_gemm<64, false>(ta, tb, p, nt, st);
}

int tc_dgrad(const cpgb_conv_desc &d, const float *dy, const float *staged, float *dx, cudaStream_t st) {
  if (!aligned16p(dy) || !aligned16p(dx) || !aligned16p(staged)) {
    set_error("tcgen05 path needs 16-byte aligned tensors"); return CPGB_EINVAL;
  }
  const int RS = d.R * d.S, Cp = cp_of(d);
  // output pixels = input positions (h, w); A = dy read at (h + pad - r*dil, w + pad - s*dil)
  PixBox b = make_pixbox(128, d.W, d.H, d.N);
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_act_map(&ta, dy, d.K, d.Q, d.P, d.N, y_strides(d), b))) return rc;
  const long long mtiles = (long long)b.tq * b.tp * b.tn;
  const int BN = pick_bn(d.C, mtiles);
  {
    // Wt[k][t][c] as (c_in_block 32, k, t, c_block): B tile = [BN/32][32 k rows][32 c]
    uint64_t dims[4] = {32, (uint64_t)d.K, (uint64_t)RS, (uint64_t)(Cp / 32)};
    uint64_t str[3] = {(uint64_t)RS * Cp * 4, (uint64_t)Cp * 4, 128};
    uint32_t box[4] = {32, 32, 1, (uint32_t)(BN / 32)};
    if ((rc = make_map(&tb, staged, 4, dims, str, box, true))) return rc;
  }
  ConvGemmParams p;
  p.tq = b.tq; p.tp = b.tp; p.tn = b.tn; p.lq = b.lq; p.lp = b.lp;
  p.Qo = d.W; p.Po = d.H; p.No = d.N; p.S = d.S; p.taps = RS;
  p.off_h = d.pad_h; p.off_w = d.pad_w; p.step_h = -d.dil_h; p.step_w = -d.dil_w;
  p.kblocks = cdiv_i(d.K, 32); p.ncols = d.C;
  p.mn_layout = g_mn.layout; p.mn_lbo = g_mn.lbo; p.mn_sbo = g_mn.sbo; p.mn_kadv = g_mn.kadv;
  { Str4 xs = x_strides(d); p.o_sn = xs.s[0]; p.o_sh = xs.s[2]; p.o_sw = xs.s[3]; }
  p.out = dx; p.bias = nullptr;
  const int nt = cdiv_i(d.C, BN);
  if (BN == 256) return launch_conv_gemm<256, true>(ta, tb, p, nt, st);
  if (BN == 128) return launch_conv_gemm<128, true>(ta, tb, p, nt, st);
  return launch_conv_gemm<64, true>(ta, tb, p, nt, st);
}

template <int BN, int TG>
static int launch_wgrad(const CUtensorMap &tdy, const CUtensorMap &tx, const WgradParams &p, dim3 grid,
                        cudaStream_t st) {
  using Cfg = WgradCfg<BN, TG>;
  static bool attr_done = false;
  if (!attr_done) {
    CPGB_CUDA_OK(cudaFuncSetAttribute(wgrad_gemm_kernel<BN, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM_BYTES));
    attr_done = true;
  }
  wgrad_gemm_kernel<BN, TG><<<grid, 192, Cfg::SMEM_BYTES, st>>>(tdy, tx, p);
  CPGB_LAUNCH_OK("wgrad_gemm_kernel");
  return CPGB_OK;
}

// 5-D map (32 channels of a block, W, H, N, channel block) of an NHWC tensor with C % 32 == 0
static int make_act_map5(CUtensorMap *m, const float *base, int C, int W, int H, int N, const Str4 &sv,
                         const PixBox &b, int nblk_box) {
  const int64_t *s = sv.s;
  uint64_t dims[5] = {32, (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)(C / 32)};
  uint64_t str[4] = {(uint64_t)s[3] * 4, (uint64_t)s[2] * 4, (uint64_t)s[0] * 4, 128};
  uint32_t box[5] = {32, 1u << b.lq, 1u << b.lp, 1u << b.ln, (uint32_t)nblk_box};
  return make_map(m, base, 5, dims, str, box, true);
}

int tc_wgrad_fused(const cpgb_conv_desc &d, const float *x, const float *dy, const float *w, const float *piggy,
                   const uint8_t *tmask, int cur, float wd, int mode, float thr, float *dW, float *dP, void *ws,
                   size_t ws_bytes, cudaStream_t st) {
  if (!aligned16p(x) || !aligned16p(dy) || !aligned16p(ws)) {
    set_error("tcgen05 path needs 16-byte aligned tensors"); return CPGB_EINVAL;
  }
  WgradPlan pl = plan_wgrad(d);
  const int RS = d.R * d.S;
  const size_t need = (size_t)pl.splits * d.K * RS * d.C * sizeof(float);
  if (ws_bytes < need) { set_error("workspace %zu < %zu", ws_bytes, need); return CPGB_EWORKSPACE; }
  CUtensorMap tdy, tx;
  int rc;
  if ((rc = make_act_map5(&tdy, dy, d.K, d.Q, d.P, d.N, y_strides(d), pl.box, 4))) return rc;
  if ((rc = make_act_map5(&tx, x, d.C, d.W, d.H, d.N, x_strides(d), pl.box, pl.BN / 32))) return rc;
  WgradParams p;
  p.cq = pl.box.tq; p.cp = pl.box.tp; p.cn = pl.box.tn; p.lq = pl.box.lq; p.lp = pl.box.lp;
  p.S = d.S; p.RS = RS; p.pad_h = d.pad_h; p.pad_w = d.pad_w; p.dil_h = d.dil_h; p.dil_w = d.dil_w;
  p.chunks = pl.chunks; p.chunks_per_split = pl.cps; p.ctiles = pl.ctiles; p.K = d.K; p.C = d.C;
  p.mn_layout = g_mn.layout; p.mn_lbo = g_mn.lbo; p.mn_sbo = g_mn.sbo; p.mn_kadv = g_mn.kadv;
  p.gpart = reinterpret_cast<float *>(ws);
  dim3 grid(pl.ktiles * pl.ctiles, d.R, pl.splits);
  if (d.S == 3) {
    rc = pl.BN == 128 ? launch_wgrad<128, 3>(tdy, tx, p, grid, st) : launch_wgrad<64, 3>(tdy, tx, p, grid, st);
  } else {
    rc = pl.BN == 128 ? launch_wgrad<128, 1>(tdy, tx, p, grid, st) : launch_wgrad<64, 1>(tdy, tx, p, grid, st);
  }
  if (rc) return rc;
  dim3 egrid(d.K, cdiv_i(d.C, EPI_CC));
  size_t sh = (size_t)RS * (EPI_CC + 1) * sizeof(float);
  wgrad_epilogue_krsc_kernel<<<egrid, 256, sh, st>>>(p.gpart, pl.splits, d.K, d.C, RS, w, piggy, tmask, cur, wd, mode,
                                                     thr, dW, dP);
  CPGB_LAUNCH_OK("wgrad_epilogue_krsc");
  return CPGB_OK;
}

}  // namespace cpgb
