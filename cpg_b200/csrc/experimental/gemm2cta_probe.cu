// EXPERIMENTAL -- not part of libcpgb200.so, never run on hardware yet (round 1 ended without GPU budget
// for it).  Bring-up probe for the `cta_group::2` version of the conv GEMMs:
// a plain TF32 GEMM  C[M][N] = A[M][K] * B[N][K]^T  (both operands K-major fp32, the layout of the fprop
// kernel's A / B tiles) on 256 x BN tiles computed by a CTA PAIR:
//
//   cluster (2,1,1); CTA r of the pair owns rows [128 r, 128 r + 128) of the tile.
//   TMA      : each CTA loads ITS A tile (128 x 32) and ITS half of B (BN/2 x 32) into its own shared
//              memory with the .cta_group::2 form; both complete on the LEADER's full[s] barrier (address
//              from mapa), which the leader arms for 2 x STAGE_BYTES.
//   MMA      : the leader's one thread issues tcgen05.mma.cta_group::2 (M = 256, N = BN, K = 8); the
//              descriptors name the same shared-memory offsets in both CTAs.
//   release  : tcgen05.commit.cta_group::2 ... multicast::cluster to empty[s] of BOTH CTAs (each producer
//              waits on its own), and at the end to acc_full of both.
//   epilogue : every CTA drains its own 128 TMEM lanes (its 128 rows of C).
//   TMEM     : tcgen05.alloc / dealloc .cta_group::2 by one warp of each CTA, cluster barrier before dealloc.
//
// Build (syntax check on the build host, run on a B200):
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared \
//        -o cpg_b200/libcpgb_probe.so cpg_b200/csrc/experimental/gemm2cta_probe.cu -lcudart
// Driver: tools/gemm2cta_probe.py (compares against torch.matmul, times 1-CTA vs 2-CTA).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../ptx.cuh"

using namespace cpgb::ptx;

namespace {

constexpr int BM = 128;                 // rows per CTA
constexpr int BK = 32;                  // fp32 per stage row = 128 bytes (one 128B-swizzle row)
constexpr int A_BYTES = BM * BK * 4;    // 16 KB

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_tf32_ss_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit_2cta_mc(uint64_t *bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
// 2-D tile load of a CTA pair: data lands in the EXECUTING CTA's shared memory, the transaction bytes are
// signalled on `bar_cluster_addr` (a shared::cluster address, here the leader's barrier)
__device__ __forceinline__ void tma_load_2d_2cta(void *dst, const CUtensorMap *m, uint32_t bar_cluster_addr, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

struct Tail {
  uint64_t full[8], empty[8], acc_full;
  uint32_t tmem_slot, pad;
};

// PAIR = true: 256 x BN tile by two CTAs; PAIR = false: the 128 x BN single-CTA kernel with the same
// structure, as the baseline of the comparison.
template <int BN, bool PAIR>
__global__ void __launch_bounds__(192)
gemm_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float *__restrict__ C,
                  int M, int N, int K, int nstage) {
  constexpr int B_ROWS = PAIR ? BN / 2 : BN;           // rows of B held by one CTA
  constexpr int B_BYTES = B_ROWS * BK * 4;
  constexpr int STAGE = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Tail *tail = reinterpret_cast<Tail *>(smem + nstage * STAGE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  // tile coordinates: blockIdx.x counts CTAs along M (pairs are adjacent), blockIdx.y along N
  const int row0 = blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;
  const int kiters = K / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    for (int s = 0; s < nstage; ++s) { mbar_init(tail->full + s, 1); mbar_init(tail->empty + s, 1); }
    mbar_init(&tail->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_2cta(&tail->tmem_slot, BN); tmem_relinquish_2cta(); }
    else      { tmem_alloc(&tail->tmem_slot, BN); tmem_relinquish(); }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // peers' barriers are initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_slot;

  if (warp == 0 && lane == 0) {
    // ---- producer (both CTAs)
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < kiters; ++it) {
      mbar_wait(tail->empty + stage, phase ^ 1);
      uint8_t *sa = smem + stage * STAGE, *sb = sa + A_BYTES;
      if (PAIR) {
        const uint32_t ldr_full = mapa_u32(smem_u32(tail->full + stage), 0);
        if (leader) mbar_arrive_expect_tx(tail->full + stage, 2 * STAGE);      // both CTAs' bytes
        tma_load_2d_2cta(sa, &tmA, ldr_full, it * BK, row0);
        tma_load_2d_2cta(sb, &tmB, ldr_full, it * BK, col0 + (int)rank * B_ROWS);
      } else {
        mbar_arrive_expect_tx(tail->full + stage, STAGE);
        tma_load_2d(sa, &tmA, tail->full + stage, it * BK, row0);
        tma_load_2d(sb, &tmB, tail->full + stage, it * BK, col0);
      }
      if (++stage == nstage) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ---- MMA issuer (leader CTA only)
    constexpr uint32_t idesc = make_idesc_tf32(PAIR ? 256 : 128, BN, false, false);
    const uint64_t tmpl = make_smem_desc(0, 16, 1024);
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < kiters; ++it) {
      mbar_wait(tail->full + stage, phase);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + stage * STAGE) >> 4, sb = sa + (A_BYTES >> 4);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = tmpl | (uint64_t)(sa + ks * 2), bd = tmpl | (uint64_t)(sb + ks * 2);
        if (PAIR) mma_tf32_ss_2cta(tmem_base, ad, bd, idesc, (it > 0) | (ks != 0));
        else      mma_tf32_ss(tmem_base, ad, bd, idesc, (it > 0) | (ks != 0));
      }
      if (PAIR) mma_commit_2cta_mc(tail->empty + stage, 0b11); else mma_commit(tail->empty + stage);
      if (++stage == nstage) { stage = 0; phase ^= 1; }
    }
    if (PAIR) mma_commit_2cta_mc(&tail->acc_full, 0b11); else mma_commit(&tail->acc_full);
  } else if (warp >= 2) {
    // ---- epilogue: warp w drains TMEM lanes [32 (w % 4), +32) = rows of this CTA's half of the tile
    const int quad = warp & 3;
    const int row = row0 + quad * 32 + lane;
    mbar_wait(&tail->acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c, v);
      tmem_ld_wait();
      if (row < M) {
        float *dst = C + (long long)row * N + col0 + c;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (col0 + c + j < N) *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    tc_fence_before();
  }
  // nobody leaves while the peer may still signal its barriers or read its shared memory
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2cta(tmem_base, BN); else tmem_dealloc(tmem_base, BN);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map2d(CUtensorMap *m, const float *base, int rows, int cols, int box_rows) {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -1;
  cuuint64_t gd[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, gs[1] = {(cuuint64_t)cols * 4};
  cuuint32_t bx[2] = {BK, (cuuint32_t)box_rows}, es[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(fn)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gd, gs,
                                                   bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

template <int BN, bool PAIR>
int launch(const float *A, const float *B, float *C, int M, int N, int K, int nstage, cudaStream_t st) {
  constexpr int B_ROWS = PAIR ? BN / 2 : BN;
  constexpr int STAGE = A_BYTES + B_ROWS * BK * 4;
  CUtensorMap ta, tb;
  if (make_map2d(&ta, A, M, K, BM) || make_map2d(&tb, B, N, K, B_ROWS)) return -2;
  const size_t smem = (size_t)nstage * STAGE + 1024 + sizeof(Tail);
  auto kern = gemm_probe_kernel<BN, PAIR>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
  cudaLaunchConfig_t cfg = {};
  int mt = (M + BM - 1) / BM;
  if (PAIR) mt = (mt + 1) & ~1;                        // whole pairs; the odd CTA works on out-of-range rows
  cfg.gridDim = dim3(mt, (N + BN - 1) / BN, 1);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, ta, tb, C, M, N, K, nstage) == cudaSuccess ? 0 : -4;
}

}  // namespace

// C[M][N] = A[M][K] * B[N][K]^T, TF32.  K % 32 == 0, N % 4 == 0, bn in {128, 256}, pair in {0, 1}.
extern "C" int cpgb_probe_gemm(const float *A, const float *B, float *C, int M, int N, int K, int bn, int pair,
                               int nstage, void *stream) {
  if (K % BK || N % 4 || nstage < 2 || nstage > 8) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  if (bn == 128) return pair ? launch<128, true>(A, B, C, M, N, K, nstage, st) : launch<128, false>(A, B, C, M, N, K, nstage, st);
  if (bn == 256) return pair ? launch<256, true>(A, B, C, M, N, K, nstage, st) : launch<256, false>(A, B, C, M, N, K, nstage, st);
  return -1;
}
