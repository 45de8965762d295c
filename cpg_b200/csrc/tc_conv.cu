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
#include <algorithm>
#include <cstdlib>
#include <cstring>
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
static int g_halo_base_mode = 0;   // descriptor base_offset for row-shifted operands: 0 none, 1 (addr>>7)&7, 2 (addr>>7)&3
static int g_halo_enable = 1;
void debug_set_mn(int layout, int lbo, int sbo, int kadv, int tma_swizzle) {
  if (layout == -1) { g_halo_base_mode = lbo; g_halo_enable = sbo; return; }   // (-1, base_mode, enable, *, *)
  g_mn = MnDesc{layout, lbo, sbo, kadv, tma_swizzle};
}

// Launch with programmatic dependent launch enabled: the kernel's prologue (barrier init, TMEM
// allocation, descriptor prefetch) overlaps the tail of the previous kernel on the stream; every
// kernel launched through here calls griddep_wait() before its first global access.
static bool g_pdl = getenv("CPGB_NO_PDL") == nullptr;
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// the same with the grid's z extent grouped into thread-block clusters of (1, 1, cluster_z)
template <class... KArgs, class... Args>
static cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                      int cluster_z, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)cluster_z;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
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

// RS == 1, C % 32 != 0 (the grown FC layers: 627, 5016 input features): rows copied into the zero-padded [K][Cp]
// operand, 16 bytes of output per thread, a grid-stride loop over float4 index range [beg4, end4)
__device__ __forceinline__ void stage_rows_range(const float *__restrict__ w, const float *__restrict__ piggy,
                                                 float *__restrict__ wt, int C, int Cp, float thr, long long beg4,
                                                 long long end4, long long first, long long step) {
  const int cp4 = Cp >> 2;
  const bool vec = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                   (!piggy || (reinterpret_cast<uintptr_t>(piggy) & 15) == 0);
  for (long long i4 = beg4 + first; i4 < end4; i4 += step) {
    const long long k = i4 / cp4;
    const int c = (int)(i4 - k * cp4) << 2;
    const long long src = k * C + c;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && c + 4 <= C) {
      float4 v = __ldg(reinterpret_cast<const float4 *>(w + src));
      if (piggy) {
        const float4 pv = __ldg(reinterpret_cast<const float4 *>(piggy + src));
        v.x *= binarize_val(pv.x, thr); v.y *= binarize_val(pv.y, thr);
        v.z *= binarize_val(pv.z, thr); v.w *= binarize_val(pv.w, thr);
      }
      o = make_float4(to_tf32_rna(v.x), to_tf32_rna(v.y), to_tf32_rna(v.z), to_tf32_rna(v.w));
    } else {
      if (c + 0 < C) o.x = to_tf32_rna(masked_weight(__ldg(w + src + 0), piggy, src + 0, thr));
      if (c + 1 < C) o.y = to_tf32_rna(masked_weight(__ldg(w + src + 1), piggy, src + 1, thr));
      if (c + 2 < C) o.z = to_tf32_rna(masked_weight(__ldg(w + src + 2), piggy, src + 2, thr));
      if (c + 3 < C) o.w = to_tf32_rna(masked_weight(__ldg(w + src + 3), piggy, src + 3, thr));
    }
    reinterpret_cast<float4 *>(wt)[i4] = o;
  }
}
__global__ void __launch_bounds__(256)
stage_weights_rows_kernel(const float *__restrict__ w, const float *__restrict__ piggy, float *__restrict__ wt, int K,
                          int C, int Cp, float thr) {
  stage_rows_range(w, piggy, wt, C, Cp, thr, 0, (long long)K * (Cp >> 2), (long long)blockIdx.x * blockDim.x + threadIdx.x,
                   (long long)gridDim.x * blockDim.x);
}

// RS == 1 and C == Cp: same element order, 16 bytes per thread
__global__ void __launch_bounds__(256)
stage_weights_flat_kernel(const float4 *__restrict__ w, const float4 *__restrict__ piggy, float4 *__restrict__ wt,
                          long long n4, float thr) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = __ldg(w + i);
    if (piggy) {
      const float4 pv = __ldg(piggy + i);
      v.x *= binarize_val(pv.x, thr); v.y *= binarize_val(pv.y, thr);
      v.z *= binarize_val(pv.z, thr); v.w *= binarize_val(pv.w, thr);
    }
    wt[i] = make_float4(to_tf32_rna(v.x), to_tf32_rna(v.y), to_tf32_rna(v.z), to_tf32_rna(v.w));
  }
}

// ------------------------------------------------------------------------------------------
// TF32 numerics.  tcgen05.mma kind::tf32 reads fp32 bit patterns from shared memory and ignores the low
// 13 mantissa bits (truncation).  Every operand is therefore made TF32-exact BEFORE it reaches the tensor
// core, with round-to-nearest (cvt.rna.tf32.f32): the staged weights by the staging kernels, activations
// and activation gradients by their producers (cpgb_bn_relu_* with tf32_out, the im2col kernel) or, when the
// caller does not promise that (CPGB_FLAG_X_TF32 / CPGB_FLAG_DY_TF32), by a rounding pass into workspace.
// The MMA then multiplies exactly the rounded values; no statistical correction is applied anywhere.
// ------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------
// shared pieces of the GEMM kernels
// ------------------------------------------------------------------------------------------
// tail of the dynamic smem, after the pipeline stages
struct SmemTail {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t ready[8];        // in-tile weight masking: transform warps -> MMA issuer
  uint64_t acc_full;
  uint64_t aux_full;        // fused wgrad epilogue: W / P / T tiles have landed
  uint32_t tmem_slot;
  uint32_t pad;
  float *rowptr[4][32];     // per epilogue warp: destination row pointers (nullptr = skip)
};
constexpr int TAIL_BYTES = 2048;
static_assert(sizeof(SmemTail) <= TAIL_BYTES, "tail");

// Epilogue helper.  The warp owns TMEM lanes [32*quad, +32) (one accumulator row per thread).
// NC columns starting at TMEM column `col` are moved to registers, transposed through the warp's
// smem scratch `ts` ([32][NC + 4] floats, conflict-free for 16-byte accesses) and handed to
// f(row, c4, value) with the lanes of a warp covering consecutive 16-byte groups of ONE row, so
// that global accesses made by f are fully coalesced.
template <int NC, class F>
__device__ __forceinline__ void epilogue_rows(uint32_t tmem_base, int quad, int col, float *ts, int lane, F f) {
  constexpr int LD = NC + 4;
#pragma unroll 1
  for (int c = 0; c < NC; c += 32) {
    float v[32];
    tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + col + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4 *>(ts + lane * LD + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
  __syncwarp();
  constexpr int LPR = NC / 4;        // lanes per row
  constexpr int RPP = 32 / LPR;      // rows per pass
  const int rsub = lane / LPR, c4 = (lane % LPR) * 4;
#pragma unroll 4
  for (int r0 = 0; r0 < 32; r0 += RPP) {
    const int rr = r0 + rsub;
    f(rr, c4, *reinterpret_cast<const float4 *>(ts + rr * LD + c4));
  }
  __syncwarp();
}

// bias[col .. col+3], zero beyond the last real channel (ragged channel counts are stored padded to 4)
__device__ __forceinline__ float4 load_bias4(const float *__restrict__ bias, int col, int nvalid) {
  if (col + 4 <= nvalid && (reinterpret_cast<uintptr_t>(bias + col) & 15) == 0)
    return __ldg(reinterpret_cast<const float4 *>(bias + col));
  float4 b;
  b.x = col + 0 < nvalid ? __ldg(bias + col + 0) : 0.f;
  b.y = col + 1 < nvalid ? __ldg(bias + col + 1) : 0.f;
  b.z = col + 2 < nvalid ? __ldg(bias + col + 2) : 0.f;
  b.w = col + 3 < nvalid ? __ldg(bias + col + 3) : 0.f;
  return b;
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
  int ncols;               // output channels stored per pixel (the real count rounded up to 4: pad lanes get zeros)
  int nvalid;              // real output channels (bias is only defined for these)
  int iters_per_split;     // split-K over the (tap, k-block) loop; blockIdx.z = split
  // In-tile weight masking (CPGB_FLAG_W_INTILE, linear / 1x1 layers with C % 32 == 0): the B operand is the module's
  // raw fp32 weight tensor; the four epilogue warps mask each landed tile with the packed Binarizer bits and round
  // it to TF32 in shared memory before the MMAs read it -- the masked copy never exists in global memory.
  int xform;               // 0 off, 1 round only (no piggymask), 2 mask + round
  int bits_ld;             // 64-bit mask words per weight row (= C / 32)
  int w_rows;              // rows of the weight matrix (K)
  const unsigned long long *bits;
  int cluster;             // > 1: the splits of one output tile form a thread-block cluster (1, 1, cluster) and are
                           // summed through distributed shared memory -- no partial sums in global memory
  int nstage;              // depth of the smem ring (3: two CTAs share an SM; more: one CTA, deeper prefetch)
  int mn_layout, mn_lbo, mn_sbo, mn_kadv;   // MN-major operand descriptor fields
  long long o_sn, o_sh, o_sw;
  long long split_stride;  // elements between the partial outputs of consecutive splits
  float *out;              // y / dx, or the partial-sum buffer when split
  const float *bias;       // only when not split
  // fprop feeding a training-mode batch-norm (cpgb_conv2d_fprop_stats): per pixel tile, the column sums and sums of
  // squares of the values this CTA stores, as [tile][cs_ld][2] -- the layout of the batch-norm kernels' partial sums,
  // so that bn_finalize can consume them and the bn_stats pass over y is not needed.  Unsplit tiles, BN <= 128 only.
  float *colstats;
  int cs_ld;
};

constexpr int A_TILE_BYTES = 128 * 128;  // 128 pixels x 32 fp32

template <int BN, bool B_MN>
struct ConvGemmCfg {
  static constexpr int B_TILE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  // narrow tiles with >= 2 waves of CTAs: two CTAs per SM (one CTA's epilogue overlaps the other's
  // main loop); otherwise one CTA per SM with as deep a ring as fits (the loop is latency-bound)
  static constexpr int NSTAGE_PAIR = BN == 256 ? 4 : BN == 128 ? 3 : 4;
  static constexpr int NSTAGE_SOLO = BN == 256 ? 4 : BN == 128 ? 6 : 8;
  static constexpr int NSTAGE = NSTAGE_PAIR;   // minimum, for the scratch static_assert
  static constexpr int smem_bytes(int nstage) { return nstage * STAGE_BYTES + 1024 /*align slack*/ + TAIL_BYTES; }
  // cluster split-K: the whole fp32 accumulator tile, [128][BN + 4], overlays the (then idle) stage ring
  static constexpr int RED_LD = BN + 4;
  static constexpr int RED_BYTES = 128 * RED_LD * 4;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int NC = BN < 128 ? BN : 128;       // columns per epilogue pass
  static_assert(4 * 32 * (NC + 4) * 4 <= NSTAGE * STAGE_BYTES, "epilogue scratch must fit in the stage ring");
  static_assert(RED_BYTES <= NSTAGE * STAGE_BYTES, "cluster reduction tile must fit in the stage ring");
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One 16-byte chunk of a landed weight tile: apply the Binarizer bits (b = 1 keeps, 0 multiplies by zero -- a true
// multiply, so Inf / NaN weights behave as in models/layers.py:103) and round to the nearest TF32 value.
__device__ __forceinline__ void xform_chunk(float4 *ptr, unsigned bits4, bool mask) {
  float4 v = *ptr;
  if (mask) {
    v.x *= (bits4 & 1u) ? 1.f : 0.f; v.y *= (bits4 & 2u) ? 1.f : 0.f;
    v.z *= (bits4 & 4u) ? 1.f : 0.f; v.w *= (bits4 & 8u) ? 1.f : 0.f;
  }
  *ptr = make_float4(to_tf32_rna(v.x), to_tf32_rna(v.y), to_tf32_rna(v.z), to_tf32_rna(v.w));
}

template <int BN, bool B_MN, bool XF = false>
__global__ void __launch_bounds__(192)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvGemmParams p) {
  using Cfg = ConvGemmCfg<BN, B_MN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NSTAGE = p.nstage;
  SmemTail *tail = reinterpret_cast<SmemTail *>(smem + NSTAGE * Cfg::STAGE_BYTES);
  uint64_t *full = tail->full, *empty = tail->empty, *acc_full = &tail->acc_full;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile coordinates
  int tile = blockIdx.x;
  const int tqi = tile % p.tq; tile /= p.tq;
  const int tpi = tile % p.tp; tile /= p.tp;
  const int tni = tile;
  const int q0 = tqi << p.lq, p0 = tpi << p.lp, n0 = tni << (7 - p.lq - p.lp);
  const int col0 = blockIdx.y * BN;
  const int it_beg = blockIdx.z * p.iters_per_split;
  const int it_end = min(p.taps * p.kblocks, it_beg + p.iters_per_split);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); mbar_init(tail->ready + s, 4); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tail->tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_slot;
  griddep_launch_dependents();     // the next kernel may start its own prologue
  griddep_wait();                  // ours ends here: inputs are produced by the previous kernel
  const bool xf = XF && p.xform != 0;

  if (XF && warp >= 2 && xf) {
    // ---- in-tile weight masking: the epilogue warps are idle during the main loop ----
    // fprop (K-major B, [BN rows = output channels][32 c], 128-byte swizzle: the 16-byte chunk at physical position jp
    // of row r holds input channels 4*(jp ^ (r & 7)) ..): warp tw owns rows [tw*BN/4, +BN/4).
    // dgrad (MN-major B, [BN/32 blocks of 32 c][32 rows = k][32 c], 32-byte-atom swizzle: the 32-byte unit at physical
    // position up of row r holds channels 8*(up ^ (r & 3)) ..): warp tw owns block tw.
    const int tw = warp - 2;
    const bool mask = p.xform == 2;
    constexpr int ROWS_W = B_MN ? 32 : BN / 4;            // rows this warp transforms per stage
    const bool active = B_MN ? (tw < BN / 32) : true;
    const int r_sub = lane >> 3, ch = lane & 7;
    int stage = 0; uint32_t phase = 0;
    // the mask word of "my" row for k-block kb: lane l <-> row l of the warp's share.  Fetched one iteration ahead
    // (an L2 round trip per stage would otherwise sit between the tile landing and the MMAs).
    auto load_word = [&](int it) -> unsigned {
      if (!(mask && active && lane < ROWS_W) || it >= it_end) return 0xffffffffu;
      const int kb = it % p.kblocks;                       // one tap (R*S == 1)
      long long wrow, wcol;
      if (!B_MN) { wrow = col0 + tw * ROWS_W + lane; wcol = kb; }              // row = output channel, word = k-block
      else       { wrow = kb * 32 + lane;            wcol = (col0 >> 5) + tw; }  // row = reduction index k, word = c-block
      return (wrow < p.w_rows && wcol < p.bits_ld) ? (unsigned)__ldg(p.bits + wrow * p.bits_ld + wcol) : 0u;
    };
    unsigned next_word = load_word(it_beg);
    for (int it = it_beg; it < it_end; ++it) {
      const unsigned word = next_word;
      next_word = load_word(it + 1);
      mbar_wait(full + stage, phase);
      if (active) {
        uint8_t *tile = smem + stage * Cfg::STAGE_BYTES + A_TILE_BYTES + (B_MN ? tw * 4096 : tw * ROWS_W * 128);
#pragma unroll
        for (int i = 0; i < ROWS_W / 4; ++i) {
          const int r = 4 * i + r_sub;
          const unsigned wr = __shfl_sync(0xffffffffu, word, r);
          int cbase;
          if (!B_MN) cbase = 4 * (ch ^ ((tw * ROWS_W + r) & 7));
          else       cbase = 8 * ((ch >> 1) ^ (r & 3)) + 4 * (ch & 1);
          xform_chunk(reinterpret_cast<float4 *>(tile + r * 128 + ch * 16), (wr >> cbase) & 0xFu, mask);
        }
      }
      fence_proxy_async_smem();       // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(tail->ready + stage);
      if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
    }
  }

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = it_beg; it < it_end; ++it) {
        const int t = it / p.kblocks, kb = it - t * p.kblocks;
        const int r = t / p.S, s = t - r * p.S;
        mbar_wait(empty + stage, phase ^ 1);
        uint8_t *sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t *sb = sa + A_TILE_BYTES;
        mbar_arrive_expect_tx(full + stage, Cfg::STAGE_BYTES);
        tma_load_4d(sa, &tmA, full + stage, kb * 32, q0 + p.off_w + s * p.step_w, p0 + p.off_h + r * p.step_h, n0);
        if (!B_MN) tma_load_3d(sb, &tmB, full + stage, kb * 32, t, col0);               // [BN k][32 c]
        else       tma_load_4d(sb, &tmB, full + stage, 0, kb * 32, t, col0 >> 5);       // [BN/32][32 k][32 c]
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, BN, false, B_MN);
      // descriptor = (constant upper half) | (start address >> 4): everything but the stage base is
      // hoisted out of the loop -- this single thread issues every MMA of the CTA
      const uint64_t a_tmpl = make_smem_desc(0, 16, 1024);
      const uint64_t b_tmpl = B_MN ? make_smem_desc(0, p.mn_lbo, p.mn_sbo, p.mn_layout) : make_smem_desc(0, 16, 1024);
      const uint32_t b_kstep = (B_MN ? p.mn_kadv : 32) >> 4;
      int stage = 0; uint32_t phase = 0;
      for (int it = it_beg; it < it_end; ++it) {
        mbar_wait(full + stage, phase);
        if (xf) mbar_wait(tail->ready + stage, phase);    // ... and the weight tile has been masked and rounded
        tc_fence_after();
        const uint32_t sa = desc_addr(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + (A_TILE_BYTES >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = a_tmpl | (uint64_t)(sa + ks * 2);
          const uint64_t bd = b_tmpl | (uint64_t)(sb + ks * b_kstep);
          mma_tf32_ss(tmem_base, ad, bd, idesc, (it > it_beg) | (ks != 0));
        }
        mma_commit(empty + stage);
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      mma_commit(acc_full);
    }
  } else {
    // epilogue: warp w reads TMEM lanes [32*(w%4), +32); row m of the tile = one output pixel
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const int qi = m & ((1 << p.lq) - 1);
    const int pi = (m >> p.lq) & ((1 << p.lp) - 1);
    const int ni = m >> (p.lq + p.lp);
    const int q = q0 + qi, pp = p0 + pi, n = n0 + ni;
    const bool valid = q < p.Qo && pp < p.Po && n < p.No;
    tail->rowptr[quad][lane] =
        valid ? p.out + blockIdx.z * p.split_stride + n * p.o_sn + pp * p.o_sh + q * p.o_sw + col0 : nullptr;
    __syncwarp();
    mbar_wait(acc_full, 0);
    tc_fence_after();
    // the stage ring is idle now (every TMA load landed and every MMA retired): reuse it as scratch
    float *ts = reinterpret_cast<float *>(smem) + quad * 32 * (Cfg::NC + 4);
    float *const *rows = tail->rowptr[quad];
    const float *bias = p.bias;
    if (p.cluster > 1) {
      // park this CTA's partial tile in shared memory for the cluster-wide sum below: thread = accumulator row.
      // The tile overlays the stage ring, which these four warps may have written themselves (in-tile weight
      // masking): acc_full already orders every such write before this point through the mbarrier chain; the named
      // barrier states the same thing in terms compute-sanitizer's racecheck understands.
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float *red = reinterpret_cast<float *>(smem) + m * Cfg::RED_LD;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4 *>(red + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    } else {
      float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), cq = cs;    // column sums / sums of squares of this lane's columns
      const bool stats = p.colstats != nullptr;
#pragma unroll 1
      for (int h = 0; h < BN; h += Cfg::NC) {
        epilogue_rows<Cfg::NC>(tmem_base, quad, h, ts, lane, [&](int rr, int c4, float4 v) {
          float *rp = rows[rr];
          const int col = col0 + h + c4;
          if (rp != nullptr && col < p.ncols) {   // ncols % 4 == 0
            if (bias) {
              const float4 b = load_bias4(bias, col, p.nvalid);
              v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            *reinterpret_cast<float4 *>(rp + h + c4) = v;
            if (stats) {
              cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
              cq.x = fmaf(v.x, v.x, cq.x); cq.y = fmaf(v.y, v.y, cq.y);
              cq.z = fmaf(v.z, v.z, cq.z); cq.w = fmaf(v.w, v.w, cq.w);
            }
          }
        });
      }
      if (stats) {
        // (host side: BN <= 128, so NC == BN and a lane kept ONE column group throughout)
        constexpr int NCs = Cfg::NC, LPR = NCs / 4;
        if (LPR < 32) {                          // BN = 64: lanes l and l + 16 hold the same columns (even / odd rows)
#pragma unroll
          for (int o = LPR; o < 32; o <<= 1) {
            cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
            cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
            cq.x += __shfl_xor_sync(0xffffffffu, cq.x, o); cq.y += __shfl_xor_sync(0xffffffffu, cq.y, o);
            cq.z += __shfl_xor_sync(0xffffffffu, cq.z, o); cq.w += __shfl_xor_sync(0xffffffffu, cq.w, o);
          }
        }
        const int c4 = (lane % LPR) * 4;
        // the warp's scratch is free again (epilogue_rows ends with __syncwarp): park the warp's sums there, ...
        if (lane < LPR) {
          *reinterpret_cast<float4 *>(ts + c4) = cs;
          *reinterpret_cast<float4 *>(ts + NCs + c4) = cq;
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");      // ... the four epilogue warps meet, ...
        if (quad == 0 && lane < LPR && col0 + c4 < p.ncols) {
          // ... and one warp adds the four in warp order and writes the tile's partial pairs
          const float *base = reinterpret_cast<const float *>(smem);
          constexpr int WS = 32 * (NCs + 4);     // floats between the scratch regions of consecutive warps
          float4 s4 = *reinterpret_cast<const float4 *>(base + c4), q4 = *reinterpret_cast<const float4 *>(base + NCs + c4);
#pragma unroll
          for (int w = 1; w < 4; ++w) {
            const float4 a = *reinterpret_cast<const float4 *>(base + w * WS + c4);
            const float4 b = *reinterpret_cast<const float4 *>(base + w * WS + NCs + c4);
            s4.x += a.x; s4.y += a.y; s4.z += a.z; s4.w += a.w;
            q4.x += b.x; q4.y += b.y; q4.z += b.z; q4.w += b.w;
          }
          float *dst = p.colstats + ((long long)blockIdx.x * p.cs_ld + col0 + c4) * 2;
          reinterpret_cast<float4 *>(dst)[0] = make_float4(s4.x, q4.x, s4.y, q4.y);
          reinterpret_cast<float4 *>(dst)[1] = make_float4(s4.z, q4.z, s4.w, q4.w);
        }
      }
    }
    tc_fence_before();
  }
  if (p.cluster > 1) {
    // Sum the `cluster` partial tiles through distributed shared memory: CTA r of the cluster owns the output rows
    // r, r + cluster, ...; a warp takes whole rows (coalesced 16-byte global stores) and adds the peers' copies in
    // rank order, so the result does not depend on scheduling.
    __syncwarp();
    cluster_sync_all();
    if (warp >= 2) {
      constexpr int LPR = BN / 4;                   // lanes per row
      constexpr int RPP = LPR >= 32 ? 1 : 32 / LPR; // rows per warp pass
      constexpr int CPL = LPR > 32 ? LPR / 32 : 1;  // float4 per lane and row (BN = 256: 2)
      const uint32_t rank = cluster_ctarank();
      const int nranks = p.cluster;
      const int ew = warp - 2;                      // 0..3
      const int rsub = LPR >= 32 ? 0 : lane / LPR;
      const int lc = LPR >= 32 ? lane : lane % LPR;
      const uint32_t red0 = smem_u32(smem);
      const float *bias = p.bias;
      for (int i = ew * RPP + rsub; ; i += 4 * RPP) {
        const int row = (int)rank + i * nranks;
        if (row >= 128) break;
        float *rp = tail->rowptr[row >> 5][row & 31];
#pragma unroll
        for (int cc = 0; cc < CPL; ++cc) {
          const int c4 = (lc + cc * 32) * 4;
          const uint32_t off = red0 + (uint32_t)(row * Cfg::RED_LD + c4) * 4u;
          float4 acc = ld_dsmem_f4(mapa_shared(off, 0));
          for (int q = 1; q < nranks; ++q) {
            const float4 b = ld_dsmem_f4(mapa_shared(off, (uint32_t)q));
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
          }
          const int col = col0 + c4;
          if (rp != nullptr && col < p.ncols) {
            if (bias) {
              const float4 b = load_bias4(bias, col, p.nvalid);
              acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
            }
            *reinterpret_cast<float4 *>(rp + c4) = acc;
          }
        }
      }
    }
    __syncwarp();
    cluster_sync_all();                              // peers may still be reading this CTA's tile
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// out[i] = sum_s part[s][i] (+ bias[i % ncols])
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float4 *__restrict__ part, int splits, long long n4, long long split_stride4, int ncols4,
                     int nvalid, const float *__restrict__ bias, float4 *__restrict__ out) {
  griddep_launch_dependents();
  griddep_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = __ldg(part + i);
    int s = 1;
    for (; s + 4 <= splits; s += 4) {             // four loads in flight; the sum order stays 0,1,2,...
      const float4 b0 = __ldg(part + (s + 0) * split_stride4 + i);
      const float4 b1 = __ldg(part + (s + 1) * split_stride4 + i);
      const float4 b2 = __ldg(part + (s + 2) * split_stride4 + i);
      const float4 b3 = __ldg(part + (s + 3) * split_stride4 + i);
      a.x += b0.x; a.y += b0.y; a.z += b0.z; a.w += b0.w;
      a.x += b1.x; a.y += b1.y; a.z += b1.z; a.w += b1.w;
      a.x += b2.x; a.y += b2.y; a.z += b2.z; a.w += b2.w;
      a.x += b3.x; a.y += b3.y; a.z += b3.z; a.w += b3.w;
    }
    for (; s < splits; ++s) {
      const float4 b = __ldg(part + s * split_stride4 + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (bias) {
      const float4 b = load_bias4(bias, (int)(i % ncols4) * 4, nvalid);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    out[i] = a;
  }
}

// ------------------------------------------------------------------------------------------
// wgrad kernel: G[k][tap][c] = sum over 32-pixel chunks of dY^T X, both operands MN-major
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void epi_one_tc(float g, float w, float p, bool has_p, unsigned t, int cur, float wd,
                                        int mode, float thr, float &dw, float &dp) {
  grad_epilogue_elem(g, w, p, has_p, t, cur, wd, mode, thr, dw, dp);
}

struct WgradParams {
  int cq, cp, cn;          // chunk counts along q, p, n
  int lq, lp;              // log2 chunk extents (n extent = 32 >> (lq + lp))
  int S, RS;               // filter width, taps
  int pad_h, pad_w, dil_h, dil_w;
  int chunks, chunks_per_split;
  int ragged;              // operands come as one bounded 4-D box per 32-channel block (ragged channel counts)
  int ctiles;              // number of BN-wide channel tiles
  int K, C, Cg;            // Cg = C rounded up to 4: row stride of gpart
  int mn_layout, mn_lbo, mn_sbo, mn_kadv;
  float *gpart;            // [splits][K][RS][Cg]
  // halo mode (HALO = true): one X tile per stage serves the S taps of a filter row
  int h_pitch;             // bytes between the 32-channel blocks of the X tile (multiple of 512)
  int h_box_bytes;         // bytes one X-block TMA box delivers
  int h_stage_bytes, h_nstage;
  int h_krow[4];           // first X-tile row of each 8-pixel K step (tap 0)
  int h_base_mode;
  // fused epilogue (only RS == 1, one split): SURVEY K6-K8 applied straight from TMEM
  int f_nstage, f_region_bytes;   // fused: ring depth, size of the [stage ring | epilogue scratch] region
  int fused, cur, mode;
  float wd, thr;
  const float *w, *piggy;
  const uint8_t *tmask;
  float *dW, *dP;
};

template <int BN, int TG>
struct WgradCfg {
  static constexpr int A_BYTES = 128 * 128;            // [4 blocks][32 pixels][32 k]
  static constexpr int B_BYTES = BN * 128;             // [BN/32 blocks][32 pixels][32 c]
  static constexpr int STAGE_BYTES = A_BYTES + TG * B_BYTES;
  // TG == 1: three stages so that two CTAs share an SM (epilogue of one overlaps the other)
  static constexpr int NSTAGE = TG == 1 ? 3 : ((200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES);
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 + TAIL_BYTES;
  static constexpr int ACC_COLS = TG * BN;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128
                                   : ACC_COLS <= 256 ? 256 : 512;
  static_assert(4 * 32 * (BN + 4) * 4 <= NSTAGE * STAGE_BYTES, "epilogue scratch must fit in the stage ring");
};

// grid: x = ktiles * ctiles, y = tap groups (R groups of TG = S taps, or R*S groups of one tap),
// z = splits.  One CTA accumulates TG taps for a 128(k) x BN(c) tile over its pixel chunks.
template <int BN, int TG, bool HALO>
__global__ void __launch_bounds__(192)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmP,
                  const __grid_constant__ CUtensorMap tmT, const WgradParams p) {
  using Cfg = WgradCfg<BN, TG>;
  constexpr bool CAN_FUSE = TG == 1 && BN == 128 && !HALO;
  const bool fused = CAN_FUSE && p.fused;
  const int NSTAGE = HALO ? p.h_nstage : fused ? p.f_nstage : Cfg::NSTAGE;
  const int STAGE_BYTES = HALO ? p.h_stage_bytes : Cfg::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // fused: [ring / scratch region][W tile 64 KB][P tile 64 KB if piggy][T tile 16 KB][tail]
  const bool f_has_p = fused && p.piggy != nullptr;
  uint8_t *sW = smem + p.f_region_bytes;
  uint8_t *sP = sW + 128 * 128 * 4;
  uint8_t *sT = sP + (f_has_p ? 128 * 128 * 4 : 0);
  SmemTail *tail = reinterpret_cast<SmemTail *>(fused ? sT + 128 * 128 : smem + NSTAGE * STAGE_BYTES);
  uint64_t *full = tail->full, *empty = tail->empty, *acc_full = &tail->acc_full;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x / p.ctiles, ct = blockIdx.x - kt * p.ctiles;
  const int k0 = kt * 128, c0 = ct * BN;
  const int tap0 = blockIdx.y * TG;               // first tap of this CTA
  const int r = tap0 / p.S, s0 = tap0 - r * p.S;
  const int split = blockIdx.z;
  const int ch_beg = split * p.chunks_per_split;
  const int ch_end = min(p.chunks, ch_beg + p.chunks_per_split);
  const int iters = ch_end - ch_beg;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmDY);
    prefetch_tensormap(&tmX);
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(acc_full, 1);
    mbar_init(&tail->aux_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tail->tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_slot;
  griddep_launch_dependents();
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      if (fused) {
        // the epilogue's operands: in flight during the whole main loop
        mbar_arrive_expect_tx(&tail->aux_full, 128 * 128 * 4 * (f_has_p ? 2 : 1) + (p.tmask ? 128 * 128 : 0));
        tma_load_2d(sW, &tmW, &tail->aux_full, c0, k0);
        if (f_has_p) tma_load_2d(sP, &tmP, &tail->aux_full, c0, k0);
        if (p.tmask) tma_load_2d(sT, &tmT, &tail->aux_full, c0, k0);
      }
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        int ch = ch_beg + it;
        const int cqi = ch % p.cq; ch /= p.cq;
        const int cpi = ch % p.cp; ch /= p.cp;
        const int q0 = cqi << p.lq, p0 = cpi << p.lp, n0 = ch << (5 - p.lq - p.lp);
        mbar_wait(empty + stage, phase ^ 1);
        uint8_t *sa = smem + stage * STAGE_BYTES;
        if (HALO) {
          // one X tile with (S-1)*dil extra columns per block; the taps are row offsets into it
          mbar_arrive_expect_tx(full + stage, Cfg::A_BYTES + (BN / 32) * p.h_box_bytes);
          if (p.ragged) {
#pragma unroll
            for (int b = 0; b < 4; ++b) tma_load_4d(sa + b * 4096, &tmDY, full + stage, k0 + 32 * b, q0, p0, n0);
#pragma unroll
            for (int b = 0; b < BN / 32; ++b)
              tma_load_4d(sa + Cfg::A_BYTES + b * p.h_pitch, &tmX, full + stage, c0 + 32 * b, q0 - p.pad_w,
                          p0 - p.pad_h + r * p.dil_h, n0);
          } else {
            tma_load_5d(sa, &tmDY, full + stage, 0, q0, p0, n0, k0 >> 5);
#pragma unroll
            for (int b = 0; b < BN / 32; ++b)
              tma_load_5d(sa + Cfg::A_BYTES + b * p.h_pitch, &tmX, full + stage, 0, q0 - p.pad_w,
                          p0 - p.pad_h + r * p.dil_h, n0, (c0 >> 5) + b);
          }
        } else {
          mbar_arrive_expect_tx(full + stage, Cfg::STAGE_BYTES);
          if (p.ragged) {
            // out-of-range channels (and whole out-of-range blocks) are zero filled by the TMA unit
#pragma unroll
            for (int b = 0; b < 4; ++b) tma_load_4d(sa + b * 4096, &tmDY, full + stage, k0 + 32 * b, q0, p0, n0);
#pragma unroll
            for (int s = 0; s < TG; ++s)
#pragma unroll
              for (int b = 0; b < BN / 32; ++b)
                tma_load_4d(sa + Cfg::A_BYTES + s * Cfg::B_BYTES + b * 4096, &tmX, full + stage, c0 + 32 * b,
                            q0 - p.pad_w + (s0 + s) * p.dil_w, p0 - p.pad_h + r * p.dil_h, n0);
          } else {
            tma_load_5d(sa, &tmDY, full + stage, 0, q0, p0, n0, k0 >> 5);
#pragma unroll
            for (int s = 0; s < TG; ++s)
              tma_load_5d(sa + Cfg::A_BYTES + s * Cfg::B_BYTES, &tmX, full + stage, 0,
                          q0 - p.pad_w + (s0 + s) * p.dil_w, p0 - p.pad_h + r * p.dil_h, n0, c0 >> 5);
          }
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, BN, true, true);
      // all descriptor arithmetic hoisted: per MMA only "template | (stage base + constant offset)"
      const uint64_t a_tmpl = make_smem_desc(0, p.mn_lbo, p.mn_sbo, p.mn_layout);
      const uint64_t b_tmpl = make_smem_desc(0, HALO ? p.h_pitch : p.mn_lbo, p.mn_sbo, p.mn_layout);
      uint32_t a_off[4], b_off[TG][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        a_off[ks] = (ks * p.mn_kadv) >> 4;
#pragma unroll
        for (int s = 0; s < TG; ++s)
          b_off[s][ks] = (HALO ? Cfg::A_BYTES + (p.h_krow[ks] + s * p.dil_w) * 128
                               : Cfg::A_BYTES + s * Cfg::B_BYTES + ks * p.mn_kadv) >> 4;
      }
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(full + stage, phase);
        tc_fence_after();
        const uint32_t sa = desc_addr(smem + stage * STAGE_BYTES);
#pragma unroll
        for (int s = 0; s < TG; ++s) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = a_tmpl | (uint64_t)(sa + a_off[ks]);
            const uint64_t bd = b_tmpl | (uint64_t)(sa + b_off[s][ks]);
            mma_tf32_ss(tmem_base + s * BN, ad, bd, idesc, (it | ks) != 0);
          }
        }
        mma_commit(empty + stage);
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      mma_commit(acc_full);
    }
  } else {
    const int quad = warp & 3;
    const int kbase = k0 + quad * 32;               // accumulator row rr of this warp = out channel kbase + rr
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float *ts = reinterpret_cast<float *>(smem) + quad * 32 * (BN + 4);
    if (fused) {
      // linear / 1x1 layer, single split: G has the module's [K][C] layout -- finish the gradient here
      // from the prefetched W / P / T tiles; the epilogue only stores to global memory.
      const bool has_p = f_has_p;
      mbar_wait(&tail->aux_full, 0);
      const float *w_s = reinterpret_cast<const float *>(sW);
      const float *p_s = reinterpret_cast<const float *>(sP);
      epilogue_rows<BN>(tmem_base, quad, 0, ts, lane, [&](int rr, int c4, float4 g) {
        const int row = quad * 32 + rr, k = kbase + rr, c = c0 + c4;
        if (k < p.K && c < p.C) {
          const float4 wv = *reinterpret_cast<const float4 *>(w_s + row * 128 + c4);
          const float4 pv = has_p ? *reinterpret_cast<const float4 *>(p_s + row * 128 + c4) : make_float4(0, 0, 0, 0);
          const uchar4 tv = p.tmask ? *reinterpret_cast<const uchar4 *>(sT + row * 128 + c4) : make_uchar4(0, 0, 0, 0);
          float4 ow, op;
          epi_one_tc(g.x, wv.x, pv.x, has_p, tv.x, p.cur, p.wd, p.mode, p.thr, ow.x, op.x);
          epi_one_tc(g.y, wv.y, pv.y, has_p, tv.y, p.cur, p.wd, p.mode, p.thr, ow.y, op.y);
          epi_one_tc(g.z, wv.z, pv.z, has_p, tv.z, p.cur, p.wd, p.mode, p.thr, ow.z, op.z);
          epi_one_tc(g.w, wv.w, pv.w, has_p, tv.w, p.cur, p.wd, p.mode, p.thr, ow.w, op.w);
          const long long idx = (long long)k * p.C + c;
          *reinterpret_cast<float4 *>(p.dW + idx) = ow;
          if (p.dP) *reinterpret_cast<float4 *>(p.dP + idx) = op;
        }
      });
    } else {
#pragma unroll 1
      for (int s = 0; s < TG; ++s) {
        float *gbase = p.gpart + (((long long)split * p.K + kbase) * p.RS + (tap0 + s)) * p.Cg + c0;
        const long long row_stride = (long long)p.RS * p.Cg;
        epilogue_rows<BN>(tmem_base, quad, s * BN, ts, lane, [&](int rr, int c4, float4 g) {
          if (kbase + rr < p.K && c0 + c4 < p.Cg) *reinterpret_cast<float4 *>(gbase + rr * row_stride + c4) = g;
        });
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
// One block handles `kb` consecutive output channels x `cc` input channels (blockIdx.y = channel
// chunk).  Phase 1: sum the splits of G[k][t][c0..c0+cc) (16-byte loads) into smem [k][t][c]; phase 2:
// walk the module's [k][c][t] order -- cc*RS contiguous elements per k -- with 16-byte accesses to
// W / P / T / dW / dP.  RS_T > 0: compile-time tap count.
// SCALAR: channel counts that are not a multiple of 4 (78 / 313 / 627): the module-order pass uses 4-byte accesses
// (still coalesced: consecutive threads, consecutive elements); the partial sums are read 16 bytes at a time either way
// (their rows are padded to Cg).
template <int RS_T, bool SCALAR = false>
__global__ void __launch_bounds__(256)
wgrad_epilogue_krsc_kernel(const float *__restrict__ gpart, int splits, int K, int C, int Cg, int RS_rt, int kb, int cc,
                           int sgroups, const float *__restrict__ w, const float *__restrict__ piggy,
                           const uint8_t *__restrict__ tmask, int cur, float wd, int mode, float thr,
                           float *__restrict__ dW, float *__restrict__ dP) {
  extern __shared__ float sh[];  // [kb][RS][cc + 1]
  griddep_launch_dependents();
  griddep_wait();
  const int RS = RS_T > 0 ? RS_T : RS_rt;
  const int k0 = blockIdx.x * kb, c0 = blockIdx.y * cc;
  const int nk = min(kb, K - k0);
  const int ncc = min(cc, C - c0);               // multiple of 4 unless SCALAR
  const int ld = cc + 1;
  const long long split_stride = (long long)K * RS * Cg;
  const int row4 = (ncc + 3) >> 2;               // float4 per (k, t) row segment (gpart rows are padded to Cg)
  const int items = nk * RS * row4;
  // sum of splits [s0, s1) of one float4 of G, four loads in flight, ascending order
  auto sum_range = [&](const float *src, int s0, int s1) {
    float4 a = __ldg(reinterpret_cast<const float4 *>(src + s0 * split_stride));
    int sp = s0 + 1;
    for (; sp + 4 <= s1; sp += 4) {
      const float4 b0 = __ldg(reinterpret_cast<const float4 *>(src + (sp + 0) * split_stride));
      const float4 b1 = __ldg(reinterpret_cast<const float4 *>(src + (sp + 1) * split_stride));
      const float4 b2 = __ldg(reinterpret_cast<const float4 *>(src + (sp + 2) * split_stride));
      const float4 b3 = __ldg(reinterpret_cast<const float4 *>(src + (sp + 3) * split_stride));
      a.x += b0.x; a.y += b0.y; a.z += b0.z; a.w += b0.w;
      a.x += b1.x; a.y += b1.y; a.z += b1.z; a.w += b1.w;
      a.x += b2.x; a.y += b2.y; a.z += b2.z; a.w += b2.w;
      a.x += b3.x; a.y += b3.y; a.z += b3.z; a.w += b3.w;
    }
    for (; sp < s1; ++sp) {
      const float4 b = __ldg(reinterpret_cast<const float4 *>(src + sp * split_stride));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    return a;
  };
  if (sgroups > 1) {
    // few float4 per block and many splits (the 32x32 / 16x16 layers: 72 items x 49 splits): `sgroups`
    // thread groups each sum a contiguous range of splits, the ranges are then added in order --
    // the serial chain of L2 round trips per thread shrinks by the group count, the result stays
    // deterministic
    float4 *gsum = reinterpret_cast<float4 *>(sh + kb * RS * ld + 4) ;   // [sgroups][items], 16-byte aligned below
    gsum = reinterpret_cast<float4 *>((reinterpret_cast<uintptr_t>(gsum) + 15) & ~uintptr_t(15));
    const int grp = threadIdx.x / items, i = threadIdx.x - grp * items;
    const int per = (splits + sgroups - 1) / sgroups;
    if (grp < sgroups) {
      const int s0 = grp * per, s1 = min(splits, s0 + per);
      const int row = i / row4, c = (i - row * row4) * 4;
      const float *src = gpart + ((long long)k0 * RS + row) * Cg + c0 + c;
      gsum[grp * items + i] = s0 < s1 ? sum_range(src, s0, s1) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (threadIdx.x < items) {
      const int ii = threadIdx.x;
      float4 a = gsum[ii];
      for (int gq = 1; gq < sgroups; ++gq) {
        const float4 b = gsum[gq * items + ii];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      const int row = ii / row4, c = (ii - row * row4) * 4;
      float *d = sh + row * ld + c;
      d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
    }
  } else {
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
      const int row = i / row4, c = (i - row * row4) * 4;     // row = kk * RS + t
      const float *src = gpart + ((long long)k0 * RS + row) * Cg + c0 + c;
      const float4 a = sum_range(src, 0, splits);
      float *d = sh + row * ld + c;
      d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
    }
  }
  __syncthreads();
  const bool has_p = piggy != nullptr;
  const int per_k = ncc * RS;                    // contiguous elements of one k in this chunk (multiple of 4)
  if (SCALAR) {
    for (int i = threadIdx.x; i < nk * per_k; i += blockDim.x) {
      const int kk = i / per_k, rem = i - kk * per_k;
      const int c = rem / RS, t = rem - c * RS;
      const long long idx = ((long long)(k0 + kk) * C + c0) * RS + rem;
      float ow, op;
      epi_one_tc(sh[(kk * RS + t) * ld + c], __ldg(w + idx), has_p ? __ldg(piggy + idx) : 0.f, has_p,
                 tmask ? (unsigned)__ldg(tmask + idx) : 0u, cur, wd, mode, thr, ow, op);
      dW[idx] = ow;
      if (dP) dP[idx] = op;
    }
    return;
  }
  for (int i4 = threadIdx.x; i4 < (nk * per_k) >> 2; i4 += blockDim.x) {
    const int i = i4 << 2;
    const int kk = i / per_k;
    int rem = i - kk * per_k;
    const long long idx = ((long long)(k0 + kk) * C + c0) * RS + rem;
    const float4 wv = __ldg(reinterpret_cast<const float4 *>(w + idx));
    const float4 pv = has_p ? __ldg(reinterpret_cast<const float4 *>(piggy + idx)) : make_float4(0, 0, 0, 0);
    const uchar4 tv = tmask ? __ldg(reinterpret_cast<const uchar4 *>(tmask + idx)) : make_uchar4(0, 0, 0, 0);
    int c = rem / RS, t = rem - c * RS;
    const float *srow = sh + kk * RS * ld;
    float g[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      g[j] = srow[t * ld + c];
      if (++t == RS) { t = 0; ++c; }
    }
    float4 ow, op;
    epi_one_tc(g[0], wv.x, pv.x, has_p, tv.x, cur, wd, mode, thr, ow.x, op.x);
    epi_one_tc(g[1], wv.y, pv.y, has_p, tv.y, cur, wd, mode, thr, ow.y, op.y);
    epi_one_tc(g[2], wv.z, pv.z, has_p, tv.z, cur, wd, mode, thr, ow.z, op.z);
    epi_one_tc(g[3], wv.w, pv.w, has_p, tv.w, cur, wd, mode, thr, ow.w, op.w);
    *reinterpret_cast<float4 *>(dW + idx) = ow;
    if (dP) *reinterpret_cast<float4 *>(dP + idx) = op;
  }
}

// scalar variant for channel counts that are not a multiple of 4 (the 3-channel stem)
__global__ void __launch_bounds__(256)
wgrad_epilogue_krsc_scalar_kernel(const float *__restrict__ gpart, int splits, int K, int C, int Cg, int RS,
                                  const float *__restrict__ w, const float *__restrict__ piggy,
                                  const uint8_t *__restrict__ tmask, int cur, float wd, int mode, float thr,
                                  float *__restrict__ dW, float *__restrict__ dP) {
  const long long n = (long long)K * C * RS;
  const long long split_stride = (long long)K * RS * Cg;
  const bool has_p = piggy != nullptr;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    const int t = (int)(idx % RS);
    const long long kc = idx / RS;
    const int c = (int)(kc % C);
    const long long k = kc / C;
    const float *src = gpart + (k * RS + t) * Cg + c;
    float g = 0.f;
    for (int sp = 0; sp < splits; ++sp) g += __ldg(src + sp * split_stride);
    float ow, op;
    epi_one_tc(g, __ldg(w + idx), has_p ? __ldg(piggy + idx) : 0.f, has_p, tmask ? tmask[idx] : 0u, cur, wd, mode,
               thr, ow, op);
    dW[idx] = ow;
    if (dP) dP[idx] = op;
  }
}

// RS == 1: [K][C] order already; sum the splits, 16 bytes per thread
__global__ void __launch_bounds__(256)
wgrad_epilogue_flat_kernel(const float4 *__restrict__ gpart, int splits, long long n4,
                           const float4 *__restrict__ w, const float4 *__restrict__ piggy,
                           const uchar4 *__restrict__ tmask, int cur, float wd, int mode, float thr,
                           float4 *__restrict__ dW, float4 *__restrict__ dP) {
  griddep_launch_dependents();
  griddep_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool has_p = piggy != nullptr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 g = __ldg(gpart + i);
    for (int s = 1; s < splits; ++s) {
      const float4 b = __ldg(gpart + s * n4 + i);
      g.x += b.x; g.y += b.y; g.z += b.z; g.w += b.w;
    }
    const float4 ww = __ldg(w + i);
    const float4 pv = has_p ? __ldg(piggy + i) : make_float4(0, 0, 0, 0);
    const uchar4 t = tmask ? __ldg(tmask + i) : make_uchar4(0, 0, 0, 0);
    float4 ow, op;
    epi_one_tc(g.x, ww.x, pv.x, has_p, t.x, cur, wd, mode, thr, ow.x, op.x);
    epi_one_tc(g.y, ww.y, pv.y, has_p, t.y, cur, wd, mode, thr, ow.y, op.y);
    epi_one_tc(g.z, ww.z, pv.z, has_p, t.z, cur, wd, mode, thr, ow.z, op.z);
    epi_one_tc(g.w, ww.w, pv.w, has_p, t.w, cur, wd, mode, thr, ow.w, op.w);
    dW[i] = ow;
    if (dP) dP[i] = op;
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
  // pixel stride of a one-pixel-wide tensor: whatever the outer dimensions say (a padded layout keeps its padding)
  o.s[3] = W == 1 ? (H > 1 ? in[2] : N > 1 ? in[0] : (int64_t)((C + 3) & ~3)) : in[3];
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

// Shapes the tensor-core kernels take (everything else runs on the CUDA-core kernels):
// stride 1, groups 1, NHWC activations whose pixel stride is a multiple of 16 bytes (the channel
// count itself may be odd, e.g. the 3-channel stem stored with a pixel stride of 4), K % 4 == 0
// for vector stores; dgrad additionally C % 4 == 0; wgrad K % 32 == 0, C % 32 == 0 or C < 32,
// filter width 1 or 3.
// Ragged channel counts (the grown networks: sqrt(1.5) * {64, 128, 256, 512, 4096} = 78 / 156 / 313 / 627 / 5016,
// CPG_cifar100_main_normal.py:115) are "padded on pack": the staged weight operand is zero padded to whole
// 32-channel blocks, TMA bounds every activation load at the real channel count (out-of-bounds lanes read as
// zero), and an activation tensor whose channel count is not a multiple of 4 is STORED with its pixel stride
// rounded up to 4 floats -- the kernels then write whole 16-byte groups, zeros in the pad lanes.
static inline int up4(int v) { return (v + 3) & ~3; }
static bool implicit_eligible(const cpgb_conv_desc &d, int op) {
  if (d.groups != 1 || d.stride_h != 1 || d.stride_w != 1) return false;
  if (d.R * d.S > 49 || d.N < 1) return false;
  const Str4 xs = x_strides(d), ys = y_strides(d);
  if (!nhwc_ok(xs, d.C) || !nhwc_ok(ys, d.K)) return false;
  if (d.W > 4096 || d.H > 4096) return false;
  if (op == 0 && ys.s[3] < up4(d.K)) return false;     // y is written in 16-byte groups
  if (op == 1 && xs.s[3] < up4(d.C)) return false;     // dx likewise
  if (op == 2 && d.S != 1 && d.S != 3) return false;
  return true;
}
// wgrad operand loads: whole 32-channel blocks through one 5-D box per operand, or (ragged) one bounded 4-D box
// per 32-channel block
static inline bool wgrad_ragged(const cpgb_conv_desc &d) { return (d.C % 32 && d.C > 32) || d.K % 32; }
static inline int cg_of(const cpgb_conv_desc &d) { return (d.C + 3) & ~3; }   // row stride of the wgrad partials

static inline int cp_of(const cpgb_conv_desc &d) { return (d.C + 31) / 32 * 32; }

// ---- TF32 rounding pre-pass of an activation operand the caller did not declare exact ------------------
// elements spanned by a strided 4-D tensor (its lowest address is the base pointer: strides are positive)
static long long span_elems(int n0, int n1, int n2, int n3, const Str4 &st) {
  return 1 + (long long)(n0 - 1) * st.s[0] + (long long)(n1 - 1) * st.s[1] + (long long)(n2 - 1) * st.s[2] +
         (long long)(n3 - 1) * st.s[3];
}
static long long span_x(const cpgb_conv_desc &d) { return span_elems(d.N, d.C, d.H, d.W, x_strides(d)); }
static long long span_dy(const cpgb_conv_desc &d) { return span_elems(d.N, d.K, d.P, d.Q, y_strides(d)); }
static size_t round_bytes_x(const cpgb_conv_desc &d) {
  return (d.flags & CPGB_FLAG_X_TF32) ? 0 : align_up((size_t)span_x(d) * 4, 256);
}
static size_t round_bytes_dy(const cpgb_conv_desc &d) {
  return (d.flags & CPGB_FLAG_DY_TF32) ? 0 : align_up((size_t)span_dy(d) * 4, 256);
}
// Replace `src` by a rounded copy (same strides) carved from the front of the scratch region.
static int round_operand(const float *&src, size_t bytes, long long span, void *&ws, size_t &ws_bytes, cudaStream_t st) {
  if (bytes == 0) return CPGB_OK;
  if (!ws || ws_bytes < bytes) { set_error("workspace %zu < %zu (TF32 rounding copy)", ws_bytes, bytes); return CPGB_EWORKSPACE; }
  float *dst = reinterpret_cast<float *>(ws);
  int rc = round_tf32(src, dst, span, st);
  if (rc) return rc;
  src = dst;
  ws = reinterpret_cast<char *>(ws) + bytes;
  ws_bytes -= bytes;
  return CPGB_OK;
}

static size_t implicit_staged_bytes(const cpgb_conv_desc &d) {
  return align_up((size_t)d.K * d.R * d.S * cp_of(d) * sizeof(float) + 256, 256);
}

// SM count the grids are planned for.  CPGB_SM_MARGIN=m plans for (SMs - m): under data parallelism the NCCL
// all-reduce kernels overlap the backward pass and hold a few SMs, which would turn the "exactly one wave"
// grids of the split-K / split-pixel plans into two waves.
static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    const char *m = getenv("CPGB_SM_MARGIN");
    if (m) { int mv = atoi(m); if (mv > 0 && mv < v / 2) v -= mv; }
    n = v;
  }
  return n;
}

// ---- fprop / dgrad plan: tile width and split-K factor -------------------------------------
struct GemmPlan { PixBox box; int BN, ntiles, iters, splits, ips; bool dense, cluster; long long out_elems; };
// split-K partial tiles summed inside a thread-block cluster (distributed shared memory) instead of through a
// partial-sum buffer + splitk_reduce_kernel; CPGB_SPLITK_CLUSTER=0 restores the round trip through global memory
static const bool g_splitk_cluster = !(getenv("CPGB_SPLITK_CLUSTER") && atoi(getenv("CPGB_SPLITK_CLUSTER")) == 0);
constexpr int MAX_CLUSTER = 16;  // 8 is the portable cluster size; 9..16 need cudaFuncAttributeNonPortableClusterSizeAllowed
static const int g_cluster_max_splits = std::max(1, std::min(MAX_CLUSTER, getenv("CPGB_CLUSTER_SPLITS") ? atoi(getenv("CPGB_CLUSTER_SPLITS")) : 8));
static GemmPlan plan_gemm(int Qo, int Po, int No, int ncols, int iters, const Str4 &os, bool conv) {
  GemmPlan g;
  g.box = make_pixbox(128, Qo, Po, No);
  const long long mtiles = (long long)g.box.tq * g.box.tp * g.box.tn;
  const int sms = num_sms();
  if (ncols <= 64) g.BN = 64;
  else if (ncols <= 128) g.BN = 128;
  else g.BN = (mtiles * cdiv_i(ncols, 256) >= sms) ? 256 : 128;
  // Experiment, off by default (CPGB_NARROW_FEW_TILES=1): half-width tiles for layers with very few pixel tiles (the 2x2
  // maps of VGG16: 4), so that a cluster of <= 8 splits fills the machine and the split-K sum stays in shared memory
  // instead of partial sums + a reduction kernel.  No gain in the step (1.474 vs 1.473-1.484 ms), and the half-width
  // plans it creates at other batch sizes (non-cluster split-K at BN = 64 with several channel tiles) gave gradients
  // that were 3 % off in the 16-vs-32 sample shard check (tools/shard_check.py): not a plan the tests cover, so not on.
  static const bool narrow = getenv("CPGB_NARROW_FEW_TILES") && atoi(getenv("CPGB_NARROW_FEW_TILES")) == 1;
  if (narrow && conv && g_splitk_cluster && g.BN == 128 && ncols >= 128 && ncols % 64 == 0 &&
      mtiles * cdiv_i(ncols, 128) * g_cluster_max_splits < sms)
    g.BN = 64;
  // experiment knobs (read per call): CPGB_GEMM_BN forces the tile width where the layer is at least that
  // wide, CPGB_GEMM_SPLITS the split-K factor
  const char *ebn = getenv("CPGB_GEMM_BN"), *esp = getenv("CPGB_GEMM_SPLITS");
  if (ebn) { const int v = atoi(ebn); if ((v == 64 || v == 128 || v == 256) && ncols >= v) g.BN = v; }
  g.ntiles = cdiv_i(ncols, g.BN);
  g.iters = iters;
  // partial sums are addressed like the output, so splitting needs a dense NHWC output
  g.dense = os.s[3] == ncols && os.s[2] == (int64_t)Qo * os.s[3] && os.s[0] == (int64_t)Po * os.s[2];
  g.out_elems = (long long)No * Po * Qo * ncols;
  const long long ctas = mtiles * g.ntiles;
  int splits = 1;
  const bool splittable = g.dense || g_splitk_cluster;
  if (splittable && ctas * 5 < sms * 4) {
    // two CTAs share an SM: fill the 2 * sms slots as evenly as the cluster size allows (a 1.5-wave grid leaves half
    // of the SMs with twice the work of the others)
    splits = cdiv_i(sms * 3 / 2, ctas);
    // up to MAX_CLUSTER splits: a cluster, filled towards the 2 * sms resident slots (7 -> 8 for the FC layers: a
    // 1.5-wave grid leaves half of the SMs with twice the work of the others).  Layers that want more splits than
    // a portable cluster holds (the 2x2 maps: 14) keep the partial-sum buffer: measured, 8 splits of 18 stages are
    // slower there (30.7 us) than 14 splits of 10 stages plus the reduction kernel (21.5 us) -- these CTAs are
    // latency-bound, their number matters more than the round trip.
    if (g_splitk_cluster && splits <= g_cluster_max_splits && g_cluster_max_splits >= 2)
      splits = std::max(splits, std::min(g_cluster_max_splits, (int)(2LL * sms / ctas)));
    if (splits > iters / 4) splits = iters / 4;
    if (splits < 1) splits = 1;
  }
  // One CTA per SM and a long reduction loop (conv 256->256 @ 8x8: 128 CTAs x 72 stages): halving the loop puts
  // two CTAs on every SM, so one CTA's prologue / epilogue hides under the other's main loop (measured
  // 36.9 -> 31.7 us; shorter loops lose more to the extra reduction pass than they gain)
  if (splittable && splits == 1 && g.BN <= 128 && ctas * 2 > sms && ctas <= sms && iters >= 64) splits = 2;
  if (esp && splittable) { const int v = atoi(esp); if (v >= 1 && v <= iters) splits = v; }
  // cluster reduction: at most MAX_CLUSTER splits per tile, and no dense-output requirement (nothing is addressed
  // "like the output" any more) -- a padded or strided output can be split as well
  g.ips = cdiv_i(iters, splits);
  g.splits = cdiv_i(iters, g.ips);
  g.cluster = g_splitk_cluster && g.splits > 1 && g.splits <= g_cluster_max_splits;
  if (g.splits > 1 && !g.cluster && !g.dense) {      // the partial-sum buffer is addressed like a dense output
    g.splits = 1; g.ips = iters;
  }
  return g;
}
static GemmPlan plan_fprop(const cpgb_conv_desc &d) {
  return plan_gemm(d.Q, d.P, d.N, up4(d.K), d.R * d.S * (cp_of(d) / 32), y_strides(d), d.R * d.S > 1);
}
static GemmPlan plan_dgrad(const cpgb_conv_desc &d) {
  return plan_gemm(d.W, d.H, d.N, up4(d.C), d.R * d.S * cdiv_i(d.K, 32), x_strides(d), d.R * d.S > 1);
}
static size_t plan_partial_bytes(const GemmPlan &g) {
  return (g.splits > 1 && !g.cluster) ? (size_t)g.splits * g.out_elems * sizeof(float) : 0;
}

// ---- wgrad plan ------------------------------------------------------------------------------
struct WgradPlan { int BN, TG, groups, ctiles, ktiles, splits, chunks, cps; PixBox box; bool halo; };
static WgradPlan plan_wgrad(const cpgb_conv_desc &d) {
  WgradPlan pl;
  pl.BN = d.C >= 128 ? 128 : 64;
  pl.ctiles = cdiv_i(d.C, pl.BN);
  pl.ktiles = cdiv_i(d.K, 128);
  pl.box = make_pixbox(32, d.Q, d.P, d.N);
  pl.chunks = pl.box.tq * pl.box.tp * pl.box.tn;
  const int RS = d.R * d.S, sms = num_sms();
  // Few pixel chunks: one CTA per tap (no split-K traffic).  Many: the S taps of a filter row share
  // the dY tile of every chunk (TG = S accumulators) and the chunk range is split.
  // Halo mode: chunks made of whole 8-pixel runs of an image row, filter width 3 -> the three taps of
  // a filter row are row offsets into ONE X tile (loaded with (S-1)*dil extra columns).
  pl.halo = g_halo_enable && d.S == 3 && (1 << pl.box.lq) >= 8 && (1 << pl.box.lq) + 2 * d.dil_w <= 256;
  pl.TG = pl.halo ? 3 : (d.S == 3 && pl.ctiles * pl.ktiles * RS * 2 > sms && pl.chunks <= 256) ? 1 : d.S;
  pl.groups = RS / pl.TG;
  const int base = pl.ctiles * pl.ktiles * pl.groups;
  // fill exactly one wave: TG = 3 CTAs own 384 TMEM columns and ~200 KB of smem (one per SM), TG = 1
  // CTAs fit two per SM
  const int target = pl.TG == 1 ? 2 * sms : sms;
  int splits = target / base;
  if (splits > pl.chunks / 4) splits = pl.chunks / 4;
  if (splits < 1) splits = 1;
  pl.cps = cdiv_i(pl.chunks, splits);
  pl.splits = cdiv_i(pl.chunks, pl.cps);
  return pl;
}
static bool wgrad_fusable(const cpgb_conv_desc &d, const WgradPlan &pl) {
  // Linear / 1x1 layers with a single split CAN finish the gradient inside the GEMM (CPGB_WGRAD_FUSE=1): the
  // CTA's W / P / T tiles are prefetched into shared memory by TMA during the main loop and the epilogue only
  // stores.  Measured on B200 (tools/time_ops.py, FC 4096x4096 @ batch 128) it loses to "write G (stays in
  // L2) + streaming epilogue kernel": 93 us fused with TMA prefetch (one 183 KB CTA per SM, 7 rounds), 133 us
  // fused with the epilogue warps loading W / T themselves, 79 us unfused (two CTAs per SM).  Off by default.
  static const bool fuse = getenv("CPGB_WGRAD_FUSE") != nullptr;
  return fuse && d.R * d.S == 1 && pl.splits == 1 && pl.BN == 128 && pl.TG == 1 && !pl.halo && d.C % 16 == 0;
}

static size_t implicit_workspace_bytes(const cpgb_conv_desc &d) {
  if (d.groups <= 0) return 0;
  size_t staged = 0, b = 0;
  if (implicit_eligible(d, 0) || implicit_eligible(d, 1)) staged = implicit_staged_bytes(d);
  if (implicit_eligible(d, 0)) b = std::max(b, plan_partial_bytes(plan_fprop(d)));
  if (implicit_eligible(d, 1)) b = std::max(b, plan_partial_bytes(plan_dgrad(d)));
  if (implicit_eligible(d, 2)) {
    WgradPlan pl = plan_wgrad(d);
    if (!wgrad_fusable(d, pl)) b = std::max(b, (size_t)pl.splits * d.K * d.R * d.S * cg_of(d) * sizeof(float));
  }
  // layout of ws: [staged operand (when the caller passes none)] [rounded x] [rounded dy] [partial sums]
  return staged + round_bytes_x(d) + round_bytes_dy(d) + align_up(b, 256) + 256;
}

static int implicit_stage_weights(const cpgb_conv_desc &d, const float *w, const float *piggy, float thr, void *staged,
                                  size_t bytes, cudaStream_t st) {
  if (bytes < implicit_staged_bytes(d)) { set_error("staged-weight buffer %zu < %zu", bytes, implicit_staged_bytes(d)); return CPGB_EWORKSPACE; }
  const int RS = d.R * d.S, Cp = cp_of(d);
  if (RS == 1 && Cp == d.C && aligned16p(w) && aligned16p(staged) && (!piggy || aligned16p(piggy))) {
    const long long n4 = (long long)d.K * d.C / 4;
    int grid = (int)std::min<long long>((n4 + 255) / 256, (long long)num_sms() * 8);
    stage_weights_flat_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(w),
                                                    reinterpret_cast<const float4 *>(piggy),
                                                    reinterpret_cast<float4 *>(staged), n4, thr);
    CPGB_LAUNCH_OK("stage_weights_flat");
    return CPGB_OK;
  }
  if (RS == 1 && aligned16p(staged)) {
    const long long n4 = (long long)d.K * (Cp / 4);
    int grid = (int)std::min<long long>((n4 + 255) / 256, (long long)num_sms() * 8);
    stage_weights_rows_kernel<<<grid, 256, 0, st>>>(w, piggy, reinterpret_cast<float *>(staged), d.K, d.C, Cp, thr);
    CPGB_LAUNCH_OK("stage_weights_rows");
    return CPGB_OK;
  }
  dim3 grid(d.K, cdiv_i(Cp, STAGE_CC));
  size_t sh = (size_t)STAGE_CC * (RS | 1) * sizeof(float);
  stage_weights_kernel<<<grid, 256, sh, st>>>(w, piggy, reinterpret_cast<float *>(staged), d.C, Cp, RS, thr);
  CPGB_LAUNCH_OK("stage_weights");
  return CPGB_OK;
}

template <int BN, bool B_MN, bool XF = false>
static int launch_conv_gemm(const CUtensorMap &ta, const CUtensorMap &tb, ConvGemmParams &p, int ntiles_n,
                            int splits, cudaStream_t st, bool cluster = false) {
  using Cfg = ConvGemmCfg<BN, B_MN>;
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    CPGB_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<BN, B_MN, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::smem_bytes(Cfg::NSTAGE_SOLO)));
    if (g_cluster_max_splits > 8)
      CPGB_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<BN, B_MN, XF>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  }
  dim3 grid(p.tq * p.tp * p.tn, ntiles_n, splits);
  const long long ctas = (long long)grid.x * grid.y * grid.z;
  p.nstage = ctas > num_sms() ? Cfg::NSTAGE_PAIR : Cfg::NSTAGE_SOLO;   // solo only when no SM gets two CTAs anyway
  p.cluster = cluster ? splits : 1;
  if (cluster)
    CPGB_CUDA_OK(launch_pdl_cluster(conv_gemm_kernel<BN, B_MN, XF>, grid, dim3(192), Cfg::smem_bytes(p.nstage), st, splits,
                                    ta, tb, p));
  else
  CPGB_CUDA_OK(launch_pdl(conv_gemm_kernel<BN, B_MN, XF>, grid, dim3(192), Cfg::smem_bytes(p.nstage), st, ta, tb, p));
  CPGB_LAUNCH_OK("conv_gemm_kernel");
  return CPGB_OK;
}

// map of an NHWC activation tensor: dims (C, W, H, N)
static int make_act_map(CUtensorMap *m, const float *base, int C, int W, int H, int N, const Str4 &sv,
                        const PixBox &b, bool mn_major = false, int box_w = 0) {
  const int64_t *s = sv.s;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t str[3] = {(uint64_t)s[3] * 4, (uint64_t)s[2] * 4, (uint64_t)s[0] * 4};
  uint32_t box[4] = {32, box_w ? (uint32_t)box_w : 1u << b.lq, 1u << b.lp, 1u << b.ln};
  return make_map(m, base, 4, dims, str, box, mn_major);
}

// common tail of fprop / dgrad: launch the GEMM (split or not) and, when split, the reduction
template <bool B_MN>
static int run_gemm(const GemmPlan &g, const CUtensorMap &ta, const CUtensorMap &tb, ConvGemmParams &p, float *out,
                    const float *bias, void *part, size_t part_bytes, cudaStream_t st) {
  p.tq = g.box.tq; p.tp = g.box.tp; p.tn = g.box.tn; p.lq = g.box.lq; p.lp = g.box.lp;
  p.iters_per_split = g.ips;
  p.mn_layout = g_mn.layout; p.mn_lbo = g_mn.lbo; p.mn_sbo = g_mn.sbo; p.mn_kadv = g_mn.kadv;
  if (g.splits > 1 && !g.cluster) {
    if (!part || part_bytes < plan_partial_bytes(g)) {
      set_error("workspace %zu < %zu (split-K partial sums)", part_bytes, plan_partial_bytes(g));
      return CPGB_EWORKSPACE;
    }
    p.out = reinterpret_cast<float *>(part); p.bias = nullptr; p.split_stride = g.out_elems;
  } else {
    p.out = out; p.bias = bias; p.split_stride = 0;
  }
  int rc;
  if (p.xform) {      // in-tile weight masking: 128- or 64-wide tiles (plan_gemm never picks 256 for these layers)
    if (g.BN == 128) rc = launch_conv_gemm<128, B_MN, true>(ta, tb, p, g.ntiles, g.splits, st, g.cluster);
    else if (g.BN == 64) rc = launch_conv_gemm<64, B_MN, true>(ta, tb, p, g.ntiles, g.splits, st, g.cluster);
    else { set_error("in-tile weight masking needs a 64- or 128-wide tile"); rc = CPGB_EINVAL; }
  } else if (g.BN == 256) rc = launch_conv_gemm<256, B_MN>(ta, tb, p, g.ntiles, g.splits, st, g.cluster);
  else if (g.BN == 128) rc = launch_conv_gemm<128, B_MN>(ta, tb, p, g.ntiles, g.splits, st, g.cluster);
  else rc = launch_conv_gemm<64, B_MN>(ta, tb, p, g.ntiles, g.splits, st, g.cluster);
  if (rc || g.splits == 1 || g.cluster) return rc;
  const long long n4 = g.out_elems / 4;
  int grid = (int)std::min<long long>((n4 + 255) / 256, (long long)num_sms() * 8);
  CPGB_CUDA_OK(launch_pdl(splitk_reduce_kernel, dim3(grid), dim3(256), 0, st, reinterpret_cast<const float4 *>(part),
                          g.splits, n4, n4, p.ncols / 4, p.nvalid, bias, reinterpret_cast<float4 *>(out)));
  CPGB_LAUNCH_OK("splitk_reduce");
  return CPGB_OK;
}

// in-tile weight masking (CPGB_FLAG_W_INTILE): which layers, and the operand description handed to the kernels
static bool intile_shape(int K, int C, int R, int S, int stride_h, int stride_w, int groups) {
  return groups == 1 && R * S == 1 && stride_h == 1 && stride_w == 1 && C % 32 == 0 && C >= 32 && K >= 1;
}
// Worth it where a weight tile is consumed by very few CTAs (the FC layers at batch <= 256: ONE pixel tile), so the
// in-shared-memory work is not repeated; layers with many pixel tiles keep the staged operand.
bool tc_intile_eligible(const cpgb_conv_desc &d) {
  static const bool off = getenv("CPGB_NO_INTILE") != nullptr;
  if (off || !intile_shape(d.K, d.C, d.R, d.S, d.stride_h, d.stride_w, d.groups)) return false;
  if (!implicit_eligible(d, 0)) return false;
  return (long long)d.N * d.P * d.Q <= 256 && d.K > 64 && d.C > 64;
}
struct IntileArgs { int xform; const unsigned long long *bits; };
static void set_intile(ConvGemmParams &p, const cpgb_conv_desc &d, const IntileArgs *it) {
  p.xform = it ? it->xform : 0;
  p.bits = it ? it->bits : nullptr;
  p.bits_ld = d.C / 32;
  p.w_rows = d.K;
}

// Can the fprop of d hand per-tile column statistics to a batch-norm that follows (ConvGemmParams::colstats)?  Only the
// plain epilogue does it: an unsplit plan with tiles of at most 128 channels.  Returns the number of pixel tiles.
static int implicit_colstats_parts(const cpgb_conv_desc &d) {
  if (!implicit_eligible(d, 0)) return 0;
  const GemmPlan g = plan_fprop(d);
  if (g.splits != 1 || g.BN > 128) return 0;
  return g.box.tq * g.box.tp * g.box.tn;
}

static int implicit_fprop(const cpgb_conv_desc &d, const float *x, const float *staged, const float *bias, float *y,
                          void *part, size_t part_bytes, cudaStream_t st, bool raw = false,
                          const IntileArgs *intile = nullptr, float *colstats = nullptr) {
  if (!aligned16p(x) || !aligned16p(y) || !aligned16p(staged) || (bias && !aligned16p(bias)) || !aligned16p(part)) {
    set_error("tcgen05 path needs 16-byte aligned tensors"); return CPGB_EINVAL;
  }
  const int RS = d.R * d.S, Cp = cp_of(d);
  GemmPlan g = plan_fprop(d);
  CUtensorMap ta, tb;
  int rc;
  if ((rc = round_operand(x, round_bytes_x(d), span_x(d), part, part_bytes, st))) return rc;
  if ((rc = make_act_map(&ta, x, d.C, d.W, d.H, d.N, x_strides(d), g.box))) return rc;
  {
    uint64_t dims[3] = {(uint64_t)Cp, (uint64_t)RS, (uint64_t)d.K};
    uint64_t str[2] = {(uint64_t)Cp * 4, (uint64_t)RS * Cp * 4};
    uint32_t box[3] = {32, 1, (uint32_t)g.BN};
    if ((rc = make_map(&tb, staged, 3, dims, str, box))) return rc;
  }
  ConvGemmParams p;
  p.Qo = d.Q; p.Po = d.P; p.No = d.N; p.S = d.S; p.taps = RS;
  p.off_h = -d.pad_h; p.off_w = -d.pad_w; p.step_h = d.dil_h; p.step_w = d.dil_w;
  p.kblocks = Cp / 32; p.ncols = up4(d.K); p.nvalid = d.K;
  set_intile(p, d, intile);
  p.colstats = (colstats && !intile && g.splits == 1 && g.BN <= 128) ? colstats : nullptr;
  p.cs_ld = up4(d.K);
  { Str4 ys = y_strides(d); p.o_sn = ys.s[0]; p.o_sh = ys.s[2]; p.o_sw = ys.s[3]; }
  return run_gemm<false>(g, ta, tb, p, y, bias, part, part_bytes, st);
}

static int implicit_dgrad(const cpgb_conv_desc &d, const float *dy, const float *staged, float *dx, void *part,
                          size_t part_bytes, cudaStream_t st, bool raw = false, const IntileArgs *intile = nullptr) {
  if (!aligned16p(dy) || !aligned16p(dx) || !aligned16p(staged) || !aligned16p(part)) {
    set_error("tcgen05 path needs 16-byte aligned tensors"); return CPGB_EINVAL;
  }
  const int RS = d.R * d.S, Cp = cp_of(d);
  // output pixels = input positions (h, w); A = dy read at (h + pad - r*dil, w + pad - s*dil)
  GemmPlan g = plan_dgrad(d);
  CUtensorMap ta, tb;
  int rc;
  if ((rc = round_operand(dy, round_bytes_dy(d), span_dy(d), part, part_bytes, st))) return rc;
  if ((rc = make_act_map(&ta, dy, d.K, d.Q, d.P, d.N, y_strides(d), g.box))) return rc;
  {
    // Wt[k][t][c] as (c_in_block 32, k, t, c_block): B tile = [BN/32][32 k rows][32 c]
    uint64_t dims[4] = {32, (uint64_t)d.K, (uint64_t)RS, (uint64_t)(Cp / 32)};
    uint64_t str[3] = {(uint64_t)RS * Cp * 4, (uint64_t)Cp * 4, 128};
    uint32_t box[4] = {32, 32, 1, (uint32_t)(g.BN / 32)};
    if ((rc = make_map(&tb, staged, 4, dims, str, box, true))) return rc;
  }
  ConvGemmParams p;
  p.Qo = d.W; p.Po = d.H; p.No = d.N; p.S = d.S; p.taps = RS;
  p.off_h = d.pad_h; p.off_w = d.pad_w; p.step_h = -d.dil_h; p.step_w = -d.dil_w;
  p.kblocks = cdiv_i(d.K, 32); p.ncols = up4(d.C); p.nvalid = d.C;
  p.colstats = nullptr; p.cs_ld = 0;
  set_intile(p, d, intile);
  { Str4 xs = x_strides(d); p.o_sn = xs.s[0]; p.o_sh = xs.s[2]; p.o_sw = xs.s[3]; }
  return run_gemm<true>(g, ta, tb, p, dx, nullptr, part, part_bytes, st);
}

struct FusedMaps { CUtensorMap w, p, t; };
template <int BN, int TG, bool HALO>
static int launch_wgrad(const CUtensorMap &tdy, const CUtensorMap &tx, const FusedMaps &fm, const WgradParams &p,
                        dim3 grid, cudaStream_t st) {
  using Cfg = WgradCfg<BN, TG>;
  int smem_bytes = HALO ? p.h_nstage * p.h_stage_bytes + 1024 + TAIL_BYTES : Cfg::SMEM_BYTES;
  if (p.fused) smem_bytes = p.f_region_bytes + 128 * 128 * 4 * (p.piggy ? 2 : 1) + 128 * 128 + 1024 + TAIL_BYTES;
  static PerDeviceOnce attr_once;
  if (attr_once.need())
    CPGB_CUDA_OK(cudaFuncSetAttribute(wgrad_gemm_kernel<BN, TG, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (HALO || (TG == 1 && BN == 128)) ? 227 * 1024 : Cfg::SMEM_BYTES));
  CPGB_CUDA_OK(launch_pdl(wgrad_gemm_kernel<BN, TG, HALO>, grid, dim3(192), (size_t)smem_bytes, st, tdy, tx, fm.w, fm.p,
                          fm.t, p));
  CPGB_LAUNCH_OK("wgrad_gemm_kernel");
  return CPGB_OK;
}

// 5-D map (32 channels of a block, W, H, N, channel block) of an NHWC tensor with C % 32 == 0 or C < 32
static int make_act_map5(CUtensorMap *m, const float *base, int C, int W, int H, int N, const Str4 &sv,
                         const PixBox &b, int nblk_box) {
  const int64_t *s = sv.s;
  uint64_t dims[5] = {(uint64_t)(C < 32 ? C : 32), (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)((C + 31) / 32)};
  uint64_t str[4] = {(uint64_t)s[3] * 4, (uint64_t)s[2] * 4, (uint64_t)s[0] * 4, 128};
  uint32_t box[5] = {32, 1u << b.lq, 1u << b.lp, 1u << b.ln, (uint32_t)nblk_box};
  return make_map(m, base, 5, dims, str, box, true);
}

// Fork of the epilogue stream: everything queued on `est` from here on runs after what `st` holds now.  One event per
// device, re-recorded per call (a wait keeps the state the event had when the wait was queued).
static int fork_epilogue_stream(cudaStream_t st, cudaStream_t est) {
  static cudaEvent_t ev[64] = {};
  int dev = 0;
  CPGB_CUDA_OK(cudaGetDevice(&dev));
  dev &= 63;
  if (!ev[dev]) CPGB_CUDA_OK(cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming));
  CPGB_CUDA_OK(cudaEventRecord(ev[dev], st));
  CPGB_CUDA_OK(cudaStreamWaitEvent(est, ev[dev], 0));
  return CPGB_OK;
}

// est: stream of the epilogue kernels (== st: everything in order on one stream).  A different stream lets the
// latency-bound epilogue of layer l run under the GEMM of layer l - 1 that follows on `st`; the caller joins `est`
// before anything consumes dW / dP.
static int implicit_wgrad_fused(const cpgb_conv_desc &d, const float *x, const float *dy, const float *w,
                                const float *piggy, const uint8_t *tmask, int cur, float wd, int mode, float thr,
                                float *dW, float *dP, void *ws, size_t ws_bytes, cudaStream_t st,
                                cudaStream_t est_in = nullptr, bool est_given = false) {
  cudaStream_t est = est_given ? est_in : st;
  if (!aligned16p(x) || !aligned16p(dy) || !aligned16p(ws)) {
    set_error("tcgen05 path needs 16-byte aligned tensors"); return CPGB_EINVAL;
  }
  WgradPlan pl = plan_wgrad(d);
  const int RS = d.R * d.S;
  const bool vec_ok = aligned16p(w) && aligned16p(dW) && (!piggy || aligned16p(piggy)) && (!dP || aligned16p(dP)) &&
                      (!tmask || (reinterpret_cast<uintptr_t>(tmask) & 3) == 0);
  const bool fused = wgrad_fusable(d, pl) && vec_ok && (!tmask || aligned16p(tmask));
  int rc;
  if ((rc = round_operand(x, round_bytes_x(d), span_x(d), ws, ws_bytes, st))) return rc;
  if ((rc = round_operand(dy, round_bytes_dy(d), span_dy(d), ws, ws_bytes, st))) return rc;
  const size_t need = fused ? 0 : (size_t)pl.splits * d.K * RS * cg_of(d) * sizeof(float);
  if (ws_bytes < need) { set_error("workspace %zu < %zu", ws_bytes, need); return CPGB_EWORKSPACE; }
  CUtensorMap tdy, tx;
  const bool ragged = wgrad_ragged(d);
  if (ragged) rc = make_act_map(&tdy, dy, d.K, d.Q, d.P, d.N, y_strides(d), pl.box, true);
  else rc = make_act_map5(&tdy, dy, d.K, d.Q, d.P, d.N, y_strides(d), pl.box, 4);
  if (rc) return rc;
  WgradParams p;
  p.ragged = ragged ? 1 : 0;
  if (pl.halo) {
    const int bq = 1 << pl.box.lq, bp = 1 << pl.box.lp, bn = 1 << pl.box.ln;
    const int wq = bq + (d.S - 1) * d.dil_w;           // X-tile columns per image row
    const int rows = bn * bp * wq;
    const int rows_p = (rows + 3) & ~3;                // block pitch: whole 4-row swizzle atoms
    p.h_pitch = rows_p * 128; p.h_box_bytes = rows * 128;
    p.h_stage_bytes = (int)align_up((size_t)16384 + (pl.BN / 32) * p.h_pitch, 1024);
    p.h_nstage = std::min(8, (224 * 1024 - 1024 - TAIL_BYTES) / p.h_stage_bytes);
    for (int ks = 0; ks < 4; ++ks) {
      const int pix = 8 * ks, qo = pix % bq, pi = (pix / bq) % bp, ni = pix / (bq * bp);
      p.h_krow[ks] = (ni * bp + pi) * wq + qo;
    }
    p.h_base_mode = g_halo_base_mode;
    const Str4 xs = x_strides(d);
    uint64_t dims[5] = {(uint64_t)(d.C < 32 ? d.C : 32), (uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.N,
                        (uint64_t)((d.C + 31) / 32)};
    uint64_t str[4] = {(uint64_t)xs.s[3] * 4, (uint64_t)xs.s[2] * 4, (uint64_t)xs.s[0] * 4, 128};
    uint32_t box[5] = {32, (uint32_t)wq, (uint32_t)bp, (uint32_t)bn, 1};
    if (ragged) rc = make_act_map(&tx, x, d.C, d.W, d.H, d.N, xs, pl.box, true, wq);
    else rc = make_map(&tx, x, 5, dims, str, box, true);
    if (rc) return rc;
  } else {
    p.h_pitch = p.h_box_bytes = p.h_stage_bytes = p.h_nstage = p.h_base_mode = 0;
    for (int ks = 0; ks < 4; ++ks) p.h_krow[ks] = 0;
    if (ragged) rc = make_act_map(&tx, x, d.C, d.W, d.H, d.N, x_strides(d), pl.box, true);
    else rc = make_act_map5(&tx, x, d.C, d.W, d.H, d.N, x_strides(d), pl.box, pl.BN / 32);
    if (rc) return rc;
  }
  p.cq = pl.box.tq; p.cp = pl.box.tp; p.cn = pl.box.tn; p.lq = pl.box.lq; p.lp = pl.box.lp;
  p.S = d.S; p.RS = RS; p.pad_h = d.pad_h; p.pad_w = d.pad_w; p.dil_h = d.dil_h; p.dil_w = d.dil_w;
  p.chunks = pl.chunks; p.chunks_per_split = pl.cps; p.ctiles = pl.ctiles; p.K = d.K; p.C = d.C; p.Cg = cg_of(d);
  p.mn_layout = g_mn.layout; p.mn_lbo = g_mn.lbo; p.mn_sbo = g_mn.sbo; p.mn_kadv = g_mn.kadv;
  p.gpart = reinterpret_cast<float *>(ws);
  p.fused = fused ? 1 : 0; p.cur = cur; p.mode = mode; p.wd = wd; p.thr = thr;
  p.w = w; p.piggy = piggy; p.tmask = tmask; p.dW = dW; p.dP = dP;
  dim3 grid(pl.ktiles * pl.ctiles, pl.groups, pl.splits);
  FusedMaps fm;
  fm.w = fm.p = fm.t = tdy;        // placeholders unless fused
  p.f_nstage = 0; p.f_region_bytes = 0;
  if (fused) {
    // W / P tiles: fp32 [K][C] boxes of 128 x 128, T tile: uint8; no swizzle (read row-wise by the epilogue)
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return CPGB_ECUDA; }
    cuuint64_t gd[2] = {(cuuint64_t)d.C, (cuuint64_t)d.K};
    cuuint32_t bx[2] = {128, 128}, es[2] = {1, 1};
    cuuint64_t gs4[1] = {(cuuint64_t)d.C * 4}, gs1[1] = {(cuuint64_t)d.C};
    CUresult r1 = enc(&fm.w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(w), gd, gs4, bx, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = piggy ? enc(&fm.p, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(piggy), gd, gs4, bx, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                        : CUDA_SUCCESS;
    CUresult r3 = tmask ? enc(&fm.t, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(tmask), gd, gs1, bx, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                        : CUDA_SUCCESS;
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS || r3 != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled failed for the fused wgrad epilogue tiles (%d %d %d)", (int)r1, (int)r2, (int)r3);
      return CPGB_ECUDA;
    }
    p.f_nstage = piggy ? 2 : 3;
    const int scratch = 4 * 32 * (128 + 4) * 4;
    p.f_region_bytes = (int)align_up((size_t)std::max(p.f_nstage * WgradCfg<128, 1>::STAGE_BYTES, scratch), 1024);
  }
  if (pl.halo) {
    rc = pl.BN == 128 ? launch_wgrad<128, 3, true>(tdy, tx, fm, p, grid, st) : launch_wgrad<64, 3, true>(tdy, tx, fm, p, grid, st);
  } else if (pl.TG == 3) {
    rc = pl.BN == 128 ? launch_wgrad<128, 3, false>(tdy, tx, fm, p, grid, st) : launch_wgrad<64, 3, false>(tdy, tx, fm, p, grid, st);
  } else {
    rc = pl.BN == 128 ? launch_wgrad<128, 1, false>(tdy, tx, fm, p, grid, st) : launch_wgrad<64, 1, false>(tdy, tx, fm, p, grid, st);
  }
  if (rc || fused) return rc;
  if (est != st && (rc = fork_epilogue_stream(st, est))) return rc;
  if (RS == 1 && vec_ok && d.C % 4 == 0) {
    const long long n4 = (long long)d.K * d.C / 4;
    int egrid = (int)std::min<long long>((n4 + 255) / 256, (long long)num_sms() * 8);
    CPGB_CUDA_OK(launch_pdl(wgrad_epilogue_flat_kernel, dim3(egrid), dim3(256), 0, est,
                            reinterpret_cast<const float4 *>(p.gpart), pl.splits, n4, reinterpret_cast<const float4 *>(w),
                            reinterpret_cast<const float4 *>(piggy), reinterpret_cast<const uchar4 *>(tmask), cur, wd,
                            mode, thr, reinterpret_cast<float4 *>(dW), reinterpret_cast<float4 *>(dP)));
    CPGB_LAUNCH_OK("wgrad_epilogue_flat");
    return CPGB_OK;
  }
  if (d.C % 4 == 0 && vec_ok) {
    // ~1152 elements of (k, channel chunk) per block, >= 2 blocks per SM: channel chunks of 128, or 32
    // when the layer has few output channels
    int cc = std::min(d.C, 128), kb = 1;
    while (cc > 32 && (long long)d.K * cdiv_i(d.C, cc) < 2LL * num_sms()) cc >>= 1;
    cc = (cc + 3) & ~3;
    while (kb < 8 && (long long)(d.K / (2 * kb)) * cdiv_i(d.C, cc) >= 2LL * num_sms() && 2 * kb * cc * RS <= 4608) kb *= 2;
    size_t sh = (size_t)kb * RS * (cc + 1) * sizeof(float);
    // split groups of the partial-sum phase (see the kernel): only when every block is tile-complete
    int sgroups = 1;
    const int items = kb * RS * (cc / 4);
    if (items <= 128 && d.K % kb == 0 && d.C % cc == 0 && pl.splits >= 8)
      sgroups = std::min(std::min(256 / items, pl.splits / 4), 8);
    if (sgroups > 1) sh += (size_t)sgroups * items * 16 + 48;
    if (sh <= 96 * 1024 && (cc * RS) % 4 == 0) {
      static PerDeviceOnce attr_once;
      if (attr_once.need()) {
        CPGB_CUDA_OK(cudaFuncSetAttribute(wgrad_epilogue_krsc_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        CPGB_CUDA_OK(cudaFuncSetAttribute(wgrad_epilogue_krsc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      }
      const dim3 egrid(cdiv_i(d.K, kb), cdiv_i(d.C, cc));
      if (RS == 9)
        CPGB_CUDA_OK(launch_pdl(wgrad_epilogue_krsc_kernel<9>, egrid, dim3(256), sh, est, (const float *)p.gpart,
                                pl.splits, d.K, d.C, p.Cg, RS, kb, cc, sgroups, w, piggy, tmask, cur, wd, mode, thr, dW, dP));
      else
        CPGB_CUDA_OK(launch_pdl(wgrad_epilogue_krsc_kernel<0>, egrid, dim3(256), sh, est, (const float *)p.gpart,
                                pl.splits, d.K, d.C, p.Cg, RS, kb, cc, sgroups, w, piggy, tmask, cur, wd, mode, thr, dW, dP));
      CPGB_LAUNCH_OK("wgrad_epilogue_krsc");
      return CPGB_OK;
    }
  }
  if (d.C > 32 && aligned16p(p.gpart)) {
    // ragged channel count: same two-phase kernel, module-order pass with 4-byte accesses
    int cc = std::min((d.C + 3) & ~3, 128), kb = 1;
    while (cc > 32 && (long long)d.K * cdiv_i(d.C, cc) < 2LL * num_sms()) cc >>= 1;
    cc = (cc + 3) & ~3;
    while (kb < 8 && (long long)(d.K / (2 * kb)) * cdiv_i(d.C, cc) >= 2LL * num_sms() && 2 * kb * cc * RS <= 4608) kb *= 2;
    const size_t sh = (size_t)kb * RS * (cc + 1) * sizeof(float);
    if (sh <= 96 * 1024) {
      static PerDeviceOnce attr_once2;
      if (attr_once2.need()) {
        CPGB_CUDA_OK(cudaFuncSetAttribute(wgrad_epilogue_krsc_kernel<9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        CPGB_CUDA_OK(cudaFuncSetAttribute(wgrad_epilogue_krsc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      }
      const dim3 egrid(cdiv_i(d.K, kb), cdiv_i(d.C, cc));
      if (RS == 9)
        CPGB_CUDA_OK(launch_pdl(wgrad_epilogue_krsc_kernel<9, true>, egrid, dim3(256), sh, est, (const float *)p.gpart,
                                pl.splits, d.K, d.C, p.Cg, RS, kb, cc, 1, w, piggy, tmask, cur, wd, mode, thr, dW, dP));
      else
        CPGB_CUDA_OK(launch_pdl(wgrad_epilogue_krsc_kernel<0, true>, egrid, dim3(256), sh, est, (const float *)p.gpart,
                                pl.splits, d.K, d.C, p.Cg, RS, kb, cc, 1, w, piggy, tmask, cur, wd, mode, thr, dW, dP));
      CPGB_LAUNCH_OK("wgrad_epilogue_krsc_ragged");
      return CPGB_OK;
    }
  }
  const long long n = (long long)d.K * d.C * RS;
  int egrid = (int)std::min<long long>((n + 255) / 256, (long long)num_sms() * 8);
  wgrad_epilogue_krsc_scalar_kernel<<<egrid, 256, 0, est>>>(p.gpart, pl.splits, d.K, d.C, p.Cg, RS, w, piggy, tmask, cur,
                                                          wd, mode, thr, dW, dP);
  CPGB_LAUNCH_OK("wgrad_epilogue_krsc_scalar");
  return CPGB_OK;
}

}  // namespace cpgb


// ------------------------------------------------------------------------------------------
// Explicit-im2col tier ("xcol"): any stride / padding / dilation, groups = 1, any channel count.
// The convolution becomes a plain GEMM over X_col[pixels][R*S*C] (tight (tap, channel) packing,
// padded to 32), materialised in workspace by im2col_kernel; dgrad is the GEMM into dX_col followed
// by a gathering col2im.  Used where the in-place TMA gather does not apply: stride-2 layers
// (ResNet / SphereNet down-sampling) and the 3-channel stems, whose implicit-GEMM K blocks would be
// 29/32 padding.  The GEMMs are the same tcgen05 kernels in "linear" geometry.
// ------------------------------------------------------------------------------------------
namespace cpgb {

static inline int kc_of(const cpgb_conv_desc &d) { return d.R * d.S * d.C; }
static inline int kcp_of(const cpgb_conv_desc &d) { return (kc_of(d) + 31) / 32 * 32; }
static inline long long pixels_of(const cpgb_conv_desc &d) { return (long long)d.N * d.P * d.Q; }

// y / dy as a [pixels][K] matrix: NHWC, pixels in (n, p, q) order, the pixel stride a multiple of 4 floats that
// holds at least K rounded up to 4 (dense when K % 4 == 0, padded otherwise)
static bool y_pixel_matrix(const cpgb_conv_desc &d) {
  const Str4 ys = y_strides(d);
  const int64_t ps = ys.s[3];
  return ys.s[1] == 1 && ps % 4 == 0 && ps >= ((d.K + 3) & ~3) && ys.s[2] == (int64_t)d.Q * ps &&
         ys.s[0] == (int64_t)d.P * d.Q * ps;
}

static bool xcol_eligible(const cpgb_conv_desc &d, int op) {
  (void)op;
  if (d.groups != 1 || d.N < 1) return false;
  if (!y_pixel_matrix(d)) return false;
  if (pixels_of(d) >= (1ll << 31) - 256) return false;
  if ((size_t)pixels_of(d) * kcp_of(d) * sizeof(float) > ((size_t)8 << 30)) return false;   // X_col <= 8 GiB
  return true;
}

// ONE staging layout per descriptor (fprop and dgrad share the staged operand): xcol for strided
// layers and for channel counts so small that the implicit K blocks would be mostly padding.
static bool prefer_xcol(const cpgb_conv_desc &d) {
  return d.stride_h != 1 || d.stride_w != 1 || d.C < 16;
}
enum TcMode { TC_NONE = 0, TC_IMPLICIT = 1, TC_XCOL = 2 };
static TcMode tc_mode(const cpgb_conv_desc &d, int op) {
  if (prefer_xcol(d)) return xcol_eligible(d, op) ? TC_XCOL : TC_NONE;
  return implicit_eligible(d, op) ? TC_IMPLICIT : TC_NONE;
}

// derived "linear" descriptor: M pixels x KCp features -> K outputs
static cpgb_conv_desc xcol_desc(const cpgb_conv_desc &d) {
  cpgb_conv_desc l;
  memset(&l, 0, sizeof(l));
  const int kcp = kcp_of(d);
  l.N = (int)pixels_of(d); l.C = kcp; l.H = l.W = 1; l.K = d.K; l.R = l.S = 1; l.P = l.Q = 1;
  l.stride_h = l.stride_w = l.dil_h = l.dil_w = 1; l.groups = 1;
  const int64_t yps = y_strides(d).s[3];            // pixel stride of y / dy (K rounded up to 4 when padded)
  l.xs[0] = kcp; l.xs[1] = 1; l.xs[2] = kcp; l.xs[3] = kcp;
  l.ys[0] = yps; l.ys[1] = 1; l.ys[2] = yps; l.ys[3] = yps;
  // X_col is rounded to TF32 by im2col_kernel (dX_col is an output); dy is the caller's tensor
  l.flags = CPGB_FLAG_X_TF32 | (d.flags & CPGB_FLAG_DY_TF32);
  return l;
}

// X_col[pix][t*C + c] = x[n, p*sh - ph + r*dh, q*sw - pw + s*dw, c]   (0 outside, 0 for padding columns)
__global__ void __launch_bounds__(256)
im2col_kernel(Geom g, const float *__restrict__ x, float *__restrict__ xcol, int KC, int KCp, long long total4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int row4 = KCp >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const long long pix = i / row4;
    const int kc0 = (int)(i - pix * row4) << 2;
    const int n = (int)(pix / g.PQ), pq = (int)(pix - (long long)n * g.PQ), p = pq / g.Q, q = pq - p * g.Q;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kc = kc0 + j;
      float val = 0.f;
      if (kc < KC) {
        const int t = kc / g.C, c = kc - t * g.C, r = t / g.S, s = t - r * g.S;
        const int h = p * g.sh - g.ph + r * g.dh, ww = q * g.sw - g.pw + s * g.dw;
        if ((unsigned)h < (unsigned)g.H && (unsigned)ww < (unsigned)g.W)
          val = to_tf32_rna(__ldg(x + n * g.xs0 + c * g.xs1 + h * g.xs2 + ww * g.xs3));
      }
      v[j] = val;
    }
    reinterpret_cast<float4 *>(xcol)[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// dx[n,h,w,c] = sum over taps (r,s) and output pixels (p,q) with p*sh - ph + r*dh == h, q*sw - pw + s*dw == w
//               of dXcol[pix(n,p,q)][(r*S+s)*C + c]
__global__ void __launch_bounds__(256)
col2im_kernel(Geom g, const float *__restrict__ dxcol, float *__restrict__ dx, int KCp, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % g.C);
    long long rest = i / g.C;
    const int ww = (int)(rest % g.W); rest /= g.W;
    const int h = (int)(rest % g.H);
    const int n = (int)(rest / g.H);
    float acc = 0.f;
    for (int r = 0; r < g.R; ++r) {
      const int hp = h + g.ph - r * g.dh;
      if (hp < 0) continue;
      const int p = hp / g.sh;
      if (p * g.sh != hp || p >= g.P) continue;
      for (int s = 0; s < g.S; ++s) {
        const int wq = ww + g.pw - s * g.dw;
        if (wq < 0) continue;
        const int q = wq / g.sw;
        if (q * g.sw != wq || q >= g.Q) continue;
        const long long pix = ((long long)n * g.P + p) * g.Q + q;
        acc += __ldg(dxcol + pix * KCp + (r * g.S + s) * g.C + c);
      }
    }
    dx[n * g.xs0 + c * g.xs1 + h * g.xs2 + ww * g.xs3] = acc;
  }
}

// staged[k][t*C + c] = tf32_rna(binarize(P[k][c][t]) * W[k][c][t]), zero padded to KCp
__global__ void __launch_bounds__(256)
stage_weights_tight_kernel(const float *__restrict__ w, const float *__restrict__ piggy, float *__restrict__ wt, int C,
                           int RS, int KCp, long long total, float thr) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int KC = C * RS;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long k = i / KCp;
    const int kc = (int)(i - k * KCp);
    float v = 0.f;
    if (kc < KC) {
      const int t = kc / C, c = kc - t * C;
      const long long idx = (k * C + c) * RS + t;
      v = to_tf32_rna(masked_weight(__ldg(w + idx), piggy, idx, thr));
    }
    wt[i] = v;
  }
}

// (dW, dP) in the module's [K][C][R][S] order from G2[K][KCp] ((tap, channel) columns)
__global__ void __launch_bounds__(256)
wgrad_epilogue_xcol_kernel(const float *__restrict__ g2, int K, int C, int RS, int KCp, const float *__restrict__ w,
                           const float *__restrict__ piggy, const uint8_t *__restrict__ tmask, int cur, float wd,
                           int mode, float thr, float *__restrict__ dW, float *__restrict__ dP) {
  const long long n = (long long)K * C * RS;
  const bool has_p = piggy != nullptr;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    const int t = (int)(idx % RS);
    const long long kc_ = idx / RS;
    const int c = (int)(kc_ % C);
    const long long k = kc_ / C;
    const float g = __ldg(g2 + k * KCp + t * C + c);
    float ow, op;
    epi_one_tc(g, __ldg(w + idx), has_p ? __ldg(piggy + idx) : 0.f, has_p, tmask ? tmask[idx] : 0u, cur, wd, mode,
               thr, ow, op);
    dW[idx] = ow;
    if (dP) dP[idx] = op;
  }
}

static size_t xcol_staged_bytes(const cpgb_conv_desc &d) {
  return align_up((size_t)d.K * kcp_of(d) * sizeof(float) + 256, 256);
}
static size_t xcol_buffer_bytes(const cpgb_conv_desc &d) {
  return align_up((size_t)pixels_of(d) * kcp_of(d) * sizeof(float), 256);
}
static size_t xcol_workspace_bytes(const cpgb_conv_desc &d) {
  // [staged (when the caller passes none)] [X_col / dX_col] [G2] [workspace of the inner GEMM]
  const cpgb_conv_desc l = xcol_desc(d);
  return xcol_staged_bytes(d) + xcol_buffer_bytes(d) + align_up((size_t)d.K * kcp_of(d) * sizeof(float), 256) +
         implicit_workspace_bytes(l) + 256;
}

static int xcol_stage_weights(const cpgb_conv_desc &d, const float *w, const float *piggy, float thr, void *staged,
                              size_t bytes, cudaStream_t st) {
  if (bytes < xcol_staged_bytes(d)) { set_error("staged-weight buffer %zu < %zu", bytes, xcol_staged_bytes(d)); return CPGB_EWORKSPACE; }
  const long long total = (long long)d.K * kcp_of(d);
  int grid = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 8);
  stage_weights_tight_kernel<<<grid, 256, 0, st>>>(w, piggy, reinterpret_cast<float *>(staged), d.C, d.R * d.S,
                                                   kcp_of(d), total, thr);
  CPGB_LAUNCH_OK("stage_weights_tight");
  return CPGB_OK;
}

static int run_im2col(const cpgb_conv_desc &d, const float *x, float *xcol, cudaStream_t st) {
  const long long total4 = pixels_of(d) * (kcp_of(d) / 4);
  int grid = (int)std::min<long long>((total4 + 255) / 256, (long long)num_sms() * 16);
  im2col_kernel<<<grid, 256, 0, st>>>(make_geom(d), x, xcol, kc_of(d), kcp_of(d), total4);
  CPGB_LAUNCH_OK("im2col");
  return CPGB_OK;
}

static int xcol_fprop(const cpgb_conv_desc &d, const float *x, const float *staged, const float *bias, float *y,
                      void *ws, size_t ws_bytes, cudaStream_t st) {
  const size_t need = xcol_buffer_bytes(d);
  if (!ws || ws_bytes < need) { set_error("workspace %zu < %zu (im2col buffer)", ws_bytes, need); return CPGB_EWORKSPACE; }
  float *xcol = reinterpret_cast<float *>(ws);
  int rc;
  if ((rc = run_im2col(d, x, xcol, st))) return rc;
  const cpgb_conv_desc l = xcol_desc(d);
  return implicit_fprop(l, xcol, staged, bias, y, reinterpret_cast<char *>(ws) + need, ws_bytes - need, st);
}

static int xcol_dgrad(const cpgb_conv_desc &d, const float *dy, const float *staged, float *dx, void *ws,
                      size_t ws_bytes, cudaStream_t st) {
  const size_t need = xcol_buffer_bytes(d);
  if (!ws || ws_bytes < need) { set_error("workspace %zu < %zu (col2im buffer)", ws_bytes, need); return CPGB_EWORKSPACE; }
  float *dxcol = reinterpret_cast<float *>(ws);
  const cpgb_conv_desc l = xcol_desc(d);
  int rc;
  if ((rc = implicit_dgrad(l, dy, staged, dxcol, reinterpret_cast<char *>(ws) + need, ws_bytes - need, st))) return rc;
  const long long total = (long long)d.N * d.H * d.W * d.C;
  int grid = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16);
  col2im_kernel<<<grid, 256, 0, st>>>(make_geom(d), dxcol, dx, kcp_of(d), total);
  CPGB_LAUNCH_OK("col2im");
  return CPGB_OK;
}

static int xcol_wgrad_fused(const cpgb_conv_desc &d, const float *x, const float *dy, const float *w,
                            const float *piggy, const uint8_t *tmask, int cur, float wd, int mode, float thr,
                            float *dW, float *dP, void *ws, size_t ws_bytes, cudaStream_t st) {
  const size_t nb = xcol_buffer_bytes(d), ng = align_up((size_t)d.K * kcp_of(d) * sizeof(float), 256);
  if (!ws || ws_bytes < nb + ng) { set_error("workspace %zu < %zu (im2col buffer)", ws_bytes, nb + ng); return CPGB_EWORKSPACE; }
  float *xcol = reinterpret_cast<float *>(ws);
  float *g2 = reinterpret_cast<float *>(reinterpret_cast<char *>(ws) + nb);
  int rc;
  if ((rc = run_im2col(d, x, xcol, st))) return rc;
  const cpgb_conv_desc l = xcol_desc(d);
  // raw G2 = dY^T X_col through the linear wgrad kernel (the "weight" pointer is only read to form a dP that
  // is not requested: any K*KCp-float buffer will do, G2 itself is one)
  if ((rc = implicit_wgrad_fused(l, xcol, dy, g2, nullptr, nullptr, 0, 0.f, CPGB_GRAD_RAW, thr, g2, nullptr,
                                 reinterpret_cast<char *>(ws) + nb + ng, ws_bytes - nb - ng, st)))
    return rc;
  const long long n = (long long)d.K * d.C * d.R * d.S;
  int grid = (int)std::min<long long>((n + 255) / 256, (long long)num_sms() * 8);
  wgrad_epilogue_xcol_kernel<<<grid, 256, 0, st>>>(g2, d.K, d.C, d.R * d.S, kcp_of(d), w, piggy, tmask, cur, wd, mode, thr,
                                                   dW, dP);
  CPGB_LAUNCH_OK("wgrad_epilogue_xcol");
  return CPGB_OK;
}

// ---- batched staging: every sharable layer of a model in one launch ---------------------------
constexpr int STAGE_MAX_ITEMS = 56;
struct StageBatch {
  const float *w[STAGE_MAX_ITEMS];
  const float *piggy[STAGE_MAX_ITEMS];
  float *out[STAGE_MAX_ITEMS];
  int K[STAGE_MAX_ITEMS], C[STAGE_MAX_ITEMS], RS[STAGE_MAX_ITEMS];
  int kind[STAGE_MAX_ITEMS];       // 0: [K][RS][Cp] through smem transpose, 1: flat copy, 2: tight [K][KCp], 3: padded rows
  int blk0[STAGE_MAX_ITEMS + 1];   // first block of each item
  float thr[STAGE_MAX_ITEMS];
  int n;
};
static_assert(sizeof(StageBatch) <= 4000, "kernel parameter space");

constexpr int STAGE_FLAT_PER_BLOCK = 256 * 8;   // float4 per block of the flat / tight kinds
__global__ void __launch_bounds__(256) stage_weights_batched_kernel(const __grid_constant__ StageBatch sb) {
  extern __shared__ float sh[];
  // which item does this block belong to (<= 56 items: linear search over the prefix sums)
  int it = 0;
  while (it + 1 < sb.n && (int)blockIdx.x >= sb.blk0[it + 1]) ++it;
  const int b = blockIdx.x - sb.blk0[it];
  const float *__restrict__ w = sb.w[it];
  const float *__restrict__ piggy = sb.piggy[it];
  float *__restrict__ wt = sb.out[it];
  const int K = sb.K[it], C = sb.C[it], RS = sb.RS[it];
  const float thr = sb.thr[it];
  if (sb.kind[it] == 0) {
    const int Cp = (C + 31) / 32 * 32, cchunks = (Cp + STAGE_CC - 1) / STAGE_CC;
    const int k = b / cchunks, c0 = (b - k * cchunks) * STAGE_CC;
    const int cc = min(STAGE_CC, Cp - c0), cv = max(0, min(STAGE_CC, C - c0)), ld = RS | 1;
    const long long base = ((long long)k * C + c0) * RS;
    float *dst = wt + (long long)k * RS * Cp + c0;
    const bool vec = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                     (!piggy || (reinterpret_cast<uintptr_t>(piggy) & 15) == 0);
    if (vec) {
      // 16-byte loads of the contiguous [cv][RS] slab, 16-byte stores of the [RS][cc] rows (Cp, c0 % 4 == 0)
      for (int i4 = threadIdx.x; i4 < (cv * RS) >> 2; i4 += blockDim.x) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(w + base) + i4);
        if (piggy) {
          const float4 pv = __ldg(reinterpret_cast<const float4 *>(piggy + base) + i4);
          v.x *= binarize_val(pv.x, thr); v.y *= binarize_val(pv.y, thr);
          v.z *= binarize_val(pv.z, thr); v.w *= binarize_val(pv.w, thr);
        }
        const float e[4] = {v.x, v.y, v.z, v.w};
        int c = (i4 * 4) / RS, t = i4 * 4 - c * RS;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sh[c * ld + t] = to_tf32_rna(e[j]);
          if (++t == RS) { t = 0; ++c; }
        }
      }
      __syncthreads();
      const int cc4 = cc >> 2;
      for (int i4 = threadIdx.x; i4 < cc4 * RS; i4 += blockDim.x) {
        const int t = i4 / cc4, c = (i4 - t * cc4) * 4;
        float4 o;
        o.x = c + 0 < cv ? sh[(c + 0) * ld + t] : 0.f;
        o.y = c + 1 < cv ? sh[(c + 1) * ld + t] : 0.f;
        o.z = c + 2 < cv ? sh[(c + 2) * ld + t] : 0.f;
        o.w = c + 3 < cv ? sh[(c + 3) * ld + t] : 0.f;
        *reinterpret_cast<float4 *>(dst + (long long)t * Cp + c) = o;
      }
    } else {
      for (int i = threadIdx.x; i < cv * RS; i += blockDim.x)
        sh[(i / RS) * ld + (i % RS)] = to_tf32_rna(masked_weight(__ldg(w + base + i), piggy, base + i, thr));
      __syncthreads();
      for (int i = threadIdx.x; i < cc * RS; i += blockDim.x) {
        const int t = i / cc, c = i - t * cc;
        dst[(long long)t * Cp + c] = c < cv ? sh[c * ld + t] : 0.f;
      }
    }
  } else if (sb.kind[it] == 3) {
    const int Cp = (C + 31) / 32 * 32;
    const long long n4 = (long long)K * (Cp >> 2);
    const long long beg = (long long)b * STAGE_FLAT_PER_BLOCK, end = min(n4, beg + STAGE_FLAT_PER_BLOCK);
    stage_rows_range(w, piggy, wt, C, Cp, thr, beg, end, threadIdx.x, blockDim.x);
  } else if (sb.kind[it] == 1) {
    const long long n4 = (long long)K * C * RS / 4;
    const long long beg = (long long)b * STAGE_FLAT_PER_BLOCK, end = min(n4, beg + STAGE_FLAT_PER_BLOCK);
    for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
      float4 v = __ldg(reinterpret_cast<const float4 *>(w) + i);
      if (piggy) {
        const float4 pv = __ldg(reinterpret_cast<const float4 *>(piggy) + i);
        v.x *= binarize_val(pv.x, thr); v.y *= binarize_val(pv.y, thr);
        v.z *= binarize_val(pv.z, thr); v.w *= binarize_val(pv.w, thr);
      }
      reinterpret_cast<float4 *>(wt)[i] = make_float4(to_tf32_rna(v.x), to_tf32_rna(v.y), to_tf32_rna(v.z), to_tf32_rna(v.w));
    }
  } else {
    const int KC = C * RS, KCp = (KC + 31) / 32 * 32;
    const long long total = (long long)K * KCp;
    const long long beg = (long long)b * STAGE_FLAT_PER_BLOCK * 4, end = min(total, beg + STAGE_FLAT_PER_BLOCK * 4);
    for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const long long k = i / KCp;
      const int kc = (int)(i - k * KCp);
      float v = 0.f;
      if (kc < KC) {
        const int t = kc / C, c = kc - t * C;
        const long long idx = (k * C + c) * RS + t;
        v = to_tf32_rna(masked_weight(__ldg(w + idx), piggy, idx, thr));
      }
      wt[i] = v;
    }
  }
}

// weight-only view of the staging decision (the input shape is not known when a whole model is staged)
static cpgb_conv_desc weight_desc(int K, int C, int R, int S, int stride_h, int stride_w, int groups) {
  cpgb_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.K = K; d.C = C; d.R = R; d.S = S; d.stride_h = stride_h; d.stride_w = stride_w; d.groups = groups;
  d.N = 1; d.H = d.W = d.P = d.Q = 1; d.dil_h = d.dil_w = 1;
  return d;
}
size_t tc_staged_bytes_for_weight(int K, int C, int R, int S, int stride_h, int stride_w, int groups) {
  if (groups != 1 || K <= 0 || C <= 0 || R * S > 49) return 0;
  const cpgb_conv_desc d = weight_desc(K, C, R, S, stride_h, stride_w, groups);
  return tc_staged_bytes(d);
}

int tc_stage_weights_batched(int n, const float *const *w, const float *const *piggy, void *const *staged,
                             const int *K, const int *C, const int *R, const int *S, const int *stride_h,
                             const int *stride_w, const float *thr, cudaStream_t st) {
  for (int base = 0; base < n; base += STAGE_MAX_ITEMS) {
    StageBatch sb;
    sb.n = std::min(STAGE_MAX_ITEMS, n - base);
    int blocks = 0, rs_max = 1;
    for (int j = 0; j < sb.n; ++j) {
      const int i = base + j;
      const cpgb_conv_desc d = weight_desc(K[i], C[i], R[i], S[i], stride_h[i], stride_w[i], 1);
      sb.w[j] = w[i]; sb.piggy[j] = piggy[i]; sb.out[j] = reinterpret_cast<float *>(staged[i]);
      sb.K[j] = K[i]; sb.C[j] = C[i]; sb.RS[j] = R[i] * S[i]; sb.thr[j] = thr[i];
      sb.blk0[j] = blocks;
      if (!w[i] || !staged[i] || !aligned16p(staged[i])) { set_error("stage_weights_batched: bad item %d", i); return CPGB_EINVAL; }
      if (prefer_xcol(d)) {
        sb.kind[j] = 2;
        blocks += cdiv_i((long long)K[i] * kcp_of(d), STAGE_FLAT_PER_BLOCK * 4);
      } else if (sb.RS[j] == 1 && C[i] % 32 == 0 && aligned16p(w[i]) && (!piggy[i] || aligned16p(piggy[i]))) {
        sb.kind[j] = 1;
        blocks += cdiv_i((long long)K[i] * C[i] / 4, STAGE_FLAT_PER_BLOCK);
      } else if (sb.RS[j] == 1) {
        sb.kind[j] = 3;
        blocks += cdiv_i((long long)K[i] * (cp_of(d) / 4), STAGE_FLAT_PER_BLOCK);
      } else {
        sb.kind[j] = 0;
        blocks += K[i] * cdiv_i(cp_of(d), STAGE_CC);
        rs_max = std::max(rs_max, sb.RS[j]);
      }
    }
    sb.blk0[sb.n] = blocks;
    if (blocks == 0) continue;
    size_t sh = (size_t)STAGE_CC * (rs_max | 1) * sizeof(float);
    stage_weights_batched_kernel<<<blocks, 256, sh, st>>>(sb);
    CPGB_LAUNCH_OK("stage_weights_batched");
  }
  return CPGB_OK;
}

// ---- public dispatchers (common.cuh) -------------------------------------------------------
bool tc_eligible(const cpgb_conv_desc &d, int op) { return tc_mode(d, op) != TC_NONE; }

size_t tc_staged_bytes(const cpgb_conv_desc &d) { return prefer_xcol(d) ? xcol_staged_bytes(d) : implicit_staged_bytes(d); }

size_t tc_workspace_bytes(const cpgb_conv_desc &d) {
  if (d.groups <= 0) return 0;
  if (prefer_xcol(d)) return (xcol_eligible(d, 0) || xcol_eligible(d, 1) || xcol_eligible(d, 2)) ? xcol_workspace_bytes(d) : 0;
  return implicit_workspace_bytes(d);
}

int tc_stage_weights(const cpgb_conv_desc &d, const float *w, const float *piggy, float thr, void *staged, size_t bytes,
                     cudaStream_t st) {
  return prefer_xcol(d) ? xcol_stage_weights(d, w, piggy, thr, staged, bytes, st)
                        : implicit_stage_weights(d, w, piggy, thr, staged, bytes, st);
}

// Round 1 fed the raw fp32 weight tensor of linear / 1x1 layers without a piggymask straight to the tensor
// core (same [K][C] layout) and compensated the truncation statistically.  Operands are now always rounded to
// nearest, so the weight is never consumed raw; the entry point stays for ABI stability.
bool tc_weights_usable_raw(const cpgb_conv_desc &) { return false; }

int tc_fprop(const cpgb_conv_desc &d, const float *x, const float *staged, const float *bias, float *y, void *part,
             size_t part_bytes, cudaStream_t st, bool raw, float *colstats) {
  return tc_mode(d, 0) == TC_XCOL ? xcol_fprop(d, x, staged, bias, y, part, part_bytes, st)
                                  : implicit_fprop(d, x, staged, bias, y, part, part_bytes, st, raw, nullptr, colstats);
}
int tc_fprop_colstats_parts(const cpgb_conv_desc &d) {
  return tc_mode(d, 0) == TC_IMPLICIT ? implicit_colstats_parts(d) : 0;
}

int tc_dgrad(const cpgb_conv_desc &d, const float *dy, const float *staged, float *dx, void *part, size_t part_bytes,
             cudaStream_t st, bool raw) {
  return tc_mode(d, 1) == TC_XCOL ? xcol_dgrad(d, dy, staged, dx, part, part_bytes, st)
                                  : implicit_dgrad(d, dy, staged, dx, part, part_bytes, st, raw);
}

// CPGB_FLAG_W_INTILE: the B operand is `w` itself; `bits` (cpgb_pack_mask words, NULL = no piggymask) masks it in
// shared memory
int tc_fprop_intile(const cpgb_conv_desc &d, const float *x, const float *w, const void *bits, const float *bias, float *y,
                    void *part, size_t part_bytes, cudaStream_t st) {
  IntileArgs it{bits ? 2 : 1, reinterpret_cast<const unsigned long long *>(bits)};
  return implicit_fprop(d, x, w, bias, y, part, part_bytes, st, false, &it);
}
int tc_dgrad_intile(const cpgb_conv_desc &d, const float *dy, const float *w, const void *bits, float *dx, void *part,
                    size_t part_bytes, cudaStream_t st) {
  IntileArgs it{bits ? 2 : 1, reinterpret_cast<const unsigned long long *>(bits)};
  return implicit_dgrad(d, dy, w, dx, part, part_bytes, st, false, &it);
}
bool tc_intile_weight_shape(int K, int C, int R, int S, int stride_h, int stride_w, int groups) {
  static const bool off = getenv("CPGB_NO_INTILE") != nullptr;
  return !off && intile_shape(K, C, R, S, stride_h, stride_w, groups) && K > 64 && C > 64;
}

int tc_wgrad_fused(const cpgb_conv_desc &d, const float *x, const float *dy, const float *w, const float *piggy,
                   const uint8_t *tmask, int cur, float wd, int mode, float thr, float *dW, float *dP, void *ws,
                   size_t ws_bytes, cudaStream_t st, cudaStream_t est) {
  return tc_mode(d, 2) == TC_XCOL
             ? xcol_wgrad_fused(d, x, dy, w, piggy, tmask, cur, wd, mode, thr, dW, dP, ws, ws_bytes, st)
             : implicit_wgrad_fused(d, x, dy, w, piggy, tmask, cur, wd, mode, thr, dW, dP, ws, ws_bytes, st, est, true);
}

}  // namespace cpgb

// ------------------------------------------------------------------------------------------
// bring-up microbenchmark: issue rate of tcgen05.mma kind::tf32 M=128 for the operand layouts the
// kernels use (no loads: descriptors point at whatever is in shared memory).
// ------------------------------------------------------------------------------------------
namespace cpgb {
template <int BN>
__global__ void __launch_bounds__(128) mma_rate_kernel(int a_mn, int b_mn, int iters, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + BN * 128) / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, BN < 32 ? 32 : BN); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && lane == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    const uint32_t idesc = make_idesc_tf32(128, BN, a_mn != 0, b_mn != 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = a_mn ? make_smem_desc(sa + ks * 1024, 4096, 512, 1) : make_smem_desc(sa + ks * 32, 16, 1024);
        const uint64_t bd = b_mn ? make_smem_desc(sb + ks * 1024, 4096, 512, 1) : make_smem_desc(sb + ks * 32, 16, 1024);
        mma_tf32_ss(tmem, ad, bd, idesc, (it | ks) != 0);
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, BN < 32 ? 32 : BN); }
}
}  // namespace cpgb

extern "C" int cpgb_debug_mma_rate(int bn, int a_mn, int b_mn, int iters, int grid, long long *out_dev, void *stream) {
  using namespace cpgb;
  const int smem = 16384 + bn * 128 + 1024;
  cudaStream_t st = (cudaStream_t)stream;
  if (bn == 64) { cudaFuncSetAttribute(mma_rate_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                  mma_rate_kernel<64><<<grid, 128, smem, st>>>(a_mn, b_mn, iters, out_dev); }
  else if (bn == 128) { cudaFuncSetAttribute(mma_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                  mma_rate_kernel<128><<<grid, 128, smem, st>>>(a_mn, b_mn, iters, out_dev); }
  else { cudaFuncSetAttribute(mma_rate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                  mma_rate_kernel<256><<<grid, 128, smem, st>>>(a_mn, b_mn, iters, out_dev); }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -3;
}

// bring-up probe: is a (1, 1, z) cluster launched through launch_pdl_cluster really a cluster, and what do shared
// memory addresses look like inside one?  out[block][0..5] = {%cluster_ctarank, %cluster_nctarank, token read from
// the next rank's shared memory, cvta.to.shared address of a static variable, of the dynamic window, mapa(own rank)}
namespace cpgb {
__global__ void cluster_probe_kernel(int *out) {
  extern __shared__ uint8_t dyn[];
  __shared__ int token;
  uint32_t rank = ptx::cluster_ctarank(), n;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(n));
  if (threadIdx.x == 0) token = 1000 + (int)blockIdx.z;
  __syncthreads();
  ptx::cluster_sync_all();
  int peer = -1;
  if (threadIdx.x == 0) {
    const uint32_t a = ptx::mapa_shared(ptx::smem_u32(&token), (rank + 1) % n);
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(peer) : "r"(a) : "memory");
    int *o = out + blockIdx.z * 6;
    o[0] = (int)rank; o[1] = (int)n; o[2] = peer;
    o[3] = (int)ptx::smem_u32(&token); o[4] = (int)ptx::smem_u32(dyn); o[5] = (int)ptx::mapa_shared(ptx::smem_u32(dyn), rank);
  }
  ptx::cluster_sync_all();
}
}  // namespace cpgb

extern "C" int cpgb_debug_cluster_probe(int z, int use_pdl, int smem, int *out_dev, void *stream) {
  using namespace cpgb;
  const bool saved = g_pdl;
  g_pdl = use_pdl != 0;
  cudaFuncSetAttribute(cluster_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaError_t e = launch_pdl_cluster(cluster_probe_kernel, dim3(1, 1, z), dim3(64), (size_t)smem, (cudaStream_t)stream, z,
                                     out_dev);
  g_pdl = saved;
  if (e != cudaSuccess) return cuda_fail(e, "cluster probe launch");
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "cluster probe");
}
