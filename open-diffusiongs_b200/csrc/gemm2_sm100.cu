// gemm2_sm100.cu -- CTA-pair (cta_group::2) variant of the DiT GEMM:  256 x 256 output tile per 2-CTA cluster.
//
// The DEFAULT GEMM of the DiT (dispatch in gemm_sm100.cu).  Why: at 128 x 256 tiles per CTA the single-CTA kernel stages
// 48 KB per 128x256x64 MAC block (43 MAC/B) and a cta_group::1 MMA reads all of it from one SM's shared memory; its
// mainloop runs at ~770 clk per k-block against 512 ideal.  With a CTA pair each CTA stages only
// its 128 rows of A and HALF of the B tile (128 of the 256 weight rows); one tcgen05.mma.cta_group::2 of shape
// 256 x 256 x 16, issued by the leader CTA, reads both CTAs' shared memory and writes 128 x 256 fp32 accumulators into
// EACH CTA's TMEM.  Operand traffic drops to 32 KB per CTA per k-block (64 MAC/B).
//
// Protocol (leader = cluster rank 0):
//   full[s]    lives in the LEADER: count 1 = the leader producer's arrive.expect_tx(2 x 32 KB); both CTAs' TMA loads
//              complete_tx on it (cp.async.bulk.tensor ... cta_group::2, leader address).  The peer producer does NOT
//              arrive: its loads for phase n+1 can only be issued after its empty[s] completed phase n, i.e. after the
//              leader's full[s] phase n was consumed, so a transiently negative tx-count is the worst that can happen.
//              (A remote mbarrier.arrive.release.cluster per k-block in the peer's producer loop cost ~1400 clk per
//              iteration and was THE reason this kernel ran at half the single-CTA rate: profiles/r1_gemm_probe_v1.txt.)
//   empty[s]   one per CTA, released by the leader's tcgen05.commit ... multicast::cluster (mask 0b11)
//   tfull[a]   one per CTA, same multicast commit when an accumulator is complete
//   tempty[a]  lives in the leader: count 2 = one elected epilogue thread per CTA (after a 128-thread named barrier);
//              the peer's arrives remotely, once per tile
// Same fused epilogues as gemm_sm100.cu (each CTA drains its own 128 rows).
#include <cstdlib>

#include "dgs_internal.h"
#include "dit_kernels.h"
#include "gemm_epilogue.cuh"
#include "sm100_ptx.cuh"

namespace dgs {

using namespace ptx;

namespace g2 {

constexpr int BM_CTA = 128, BN = 256, BN_CTA = 128, BK = 64, UMMA_K = 16, THREADS = 192;
constexpr int A_BYTES = BM_CTA * BK * 2, B_BYTES = BN_CTA * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int TMEM_COLS = 2 * BN;
// TMAEPI (gemm_epilogue.cuh: outputs leave through shared-memory staging + TMA store / reduce-add): 5 operand stages +
// 32 KB of staging; else 6 stages.  (The probe measured no difference between 4 and 7 stages on the DiT shapes.)
// TMAEPI = 2: additionally the bf16 aux (pre-activation) store of the training-mode fc1 epilogue: 4 stages + 64 KB.
template <int TMAEPI>
struct Cfg2 {
  static constexpr int STAGES = TMAEPI == 2 ? 4 : TMAEPI == 1 ? 5 : 6;
  static constexpr int STG_BYTES = TMAEPI * EPI_TMA_STAGING_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024 + 256 + 2 * 2 * BN * 4;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar_cluster_addr,
                                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once the issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}


// MN = false: C = A[M,K] x W[N,K]^T (K-major operands).  MN = true: C = A^T x W for A [K,M], W [K,N] row-major (both
// operands MN-major: the weight-gradient GEMM, see gemm_sm100.cu); a CTA's stage then holds 2 + 2 swizzle atoms of
// [64 K rows x 64 M/N elements].
// splits > 1 (split-K, fp32 TMA reduce-add epilogue only): work unit u = (tile u % num_tiles, K range u / num_tiles);
// every unit adds its partial product into the (pre-zeroed) output.  Used by the weight-gradient GEMMs, whose K = tokens is
// long and whose 256 x 256 output tiles are too few to fill 74 clusters (proj: 16 tiles).
// UNI (default; DGS_GEMM_UNI=0 selects the old path): the TMA-producer and MMA-issuer warps run converged and issue under elect.sync instead of
// `lane == 0` (elect_one_sync in sm100_ptx.cuh): without it ptxas wraps every TMA / tcgen05 instruction of those roles in an
// ELECT / BRA.U.ANY serialisation loop with R2UR operand moves (99 such loops in this file's kernels).
template <int EPI, bool MN, int TMAEPI, bool UNI = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmX, GemmEpilogue ep,
                      int M, int N, int K, int splits) {
  constexpr int STAGES = Cfg2<TMAEPI>::STAGES, STG_BYTES = Cfg2<TMAEPI>::STG_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* stg_base = smem + STAGES * STAGE_BYTES;  // TMAEPI: 4 warps x 2 x 4 KB, 1024-byte aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_m = (M + 2 * BM_CTA - 1) / (2 * BM_CTA), num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n, num_k = (K + BK - 1) / BK;
  const int num_units = num_tiles * splits, kpb = (num_k + splits - 1) / splits;  // k-blocks per unit

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (TMAEPI) prefetch_tmap(&tmO);
    if (TMAEPI == 2) prefetch_tmap(&tmX);
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; s++) { mbar_init(tfull_bar + s, 1); mbar_init(tempty_bar + s, 2); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // both CTAs' barriers are initialised before any remote arrive / multicast commit / 2SM TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();  // PDL: see sm100_ptx.cuh
  griddep_wait();

  if (warp == 0) {
    // ===================== TMA producer (one per CTA) =====================
    if (UNI || lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
        const int tile = unit % num_tiles, kb0 = (unit / num_tiles) * kpb, kb1 = min(num_k, kb0 + kpb);
        const int m0 = (tile / num_n) * (2 * BM_CTA) + (int)rank * BM_CTA;
        const int n0 = (tile % num_n) * BN + (int)rank * BN_CTA;
        for (int kb = kb0; kb < kb1; kb++) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          const uint32_t leader_full = mapa(smem_u32(full_bar + stage), 0);
          if (!UNI || elect_one_sync()) {
            if (leader) mbar_arrive_expect_tx(full_bar + stage, 2 * STAGE_BYTES);
            if (!MN) {
              tma_load_2d_2sm(sA + stage * A_BYTES, &tmA, leader_full, kb * BK, m0);
              tma_load_2d_2sm(sB + stage * B_BYTES, &tmB, leader_full, kb * BK, n0);
            } else {
#pragma unroll
              for (int a = 0; a < BM_CTA / 64; a++)
                tma_load_2d_2sm(sA + stage * A_BYTES + a * 8192, &tmA, leader_full, m0 + a * 64, kb * BK);
#pragma unroll
              for (int a = 0; a < BN_CTA / 64; a++)
                tma_load_2d_2sm(sB + stage * B_BYTES + a * 8192, &tmB, leader_full, n0 + a * 64, kb * BK);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one thread of the LEADER CTA =====================
    if (leader && (UNI || lane == 0)) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM_CTA, BN, MN, MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
        const int kb0 = (unit / num_tiles) * kpb, kb1 = min(num_k, kb0 + kpb);
        mbar_wait(tempty_bar + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; kb++) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + stage * A_BYTES), MN ? 8192 : 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + stage * B_BYTES), MN ? 8192 : 16, 1024);
          if (!UNI || elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; k++) {
              const uint64_t adv = MN ? (uint64_t)(128 * k) : (uint64_t)(2 * k);  // 16 K rows = 2048 B  |  16 bf16 = 32 B
              umma_bf16_2sm(d_tmem, adesc + adv, bdesc + adv, idesc, (kb > kb0 || k) ? 1u : 0u);
            }
            umma_commit_2sm(empty_bar + stage);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (!UNI || elect_one_sync()) umma_commit_2sm(tfull_bar + acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps 2..5: each CTA drains its own 128 rows (gemm_epilogue.cuh) ==========
    const int quad = warp & 3;
    const int et = (warp - 2) * 32 + lane;
    float* s_vec_all = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + STG_BYTES + 256);
    const uint32_t leader_tempty0 = mapa(smem_u32(tempty_bar), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
      const int tile = unit % num_tiles;
      const int m0 = (tile / num_n) * (2 * BM_CTA) + (int)rank * BM_CTA, n0 = (tile % num_n) * BN;
      const int row = m0 + quad * 32 + lane;
      float* s_vec = s_vec_all + acc * 2 * BN;
      bool uniform_gate;
      epilogue_stage_vectors<EPI, BN>(ep, s_vec, et, m0, n0, M, N, &uniform_gate);  // under the main loop
      mbar_wait(tfull_bar + acc, acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (TMAEPI != 0)
        epilogue_drain_row_tma<EPI, BN, TMAEPI == 2>(ep, s_vec, uniform_gate, t_row, row, m0 + quad * 32, n0, M, N,
                                                     stg_base + (warp - 2) * (TMAEPI * 8192), lane, &tmO, &tmX, splits > 1);
      else
        epilogue_drain_row<EPI, BN>(ep, s_vec, uniform_gate, t_row, row, n0, M, N);
      tc_fence_before();
      epi_bar_sync<128>();
      if (et == 0) {
        if (leader) mbar_arrive(tempty_bar + acc);
        else mbar_arrive_remote(leader_tempty0 + (uint32_t)(acc * 8));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (TMAEPI && lane == 0) epi_bulk_wait_all();  // the staging buffers must outlive the bulk stores reading them
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // neither CTA leaves (or frees TMEM) while its peer may still touch its smem / barriers / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

}  // namespace g2

template <int EPI, bool MN = false, int TMAEPI = 0>
static int launch_2cta(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const GemmEpilogue& ep, int M,
                       int N, int K, int num_sms, cudaStream_t st, const CUtensorMap* tmX = nullptr, int splits = 1) {
  constexpr int SMEM = g2::Cfg2<TMAEPI>::SMEM_BYTES;
  static bool configured = false;
  static int uni = 0;
  if (!configured) {
    DGS_CUDA_OK(cudaFuncSetAttribute(g2::gemm_bf16_2cta_kernel<EPI, MN, TMAEPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    DGS_CUDA_OK(cudaFuncSetAttribute(g2::gemm_bf16_2cta_kernel<EPI, MN, TMAEPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    const char* eu = getenv("DGS_GEMM_UNI");
    uni = (eu && eu[0] == '0') ? 0 : 1;  // default since round 2 (measured: r2 first GPU call)
    configured = true;
  }
  auto kern = uni ? g2::gemm_bf16_2cta_kernel<EPI, MN, TMAEPI, true> : g2::gemm_bf16_2cta_kernel<EPI, MN, TMAEPI, false>;
  const int tiles = ceil_div(M, 2 * g2::BM_CTA) * ceil_div(N, g2::BN) * splits;
  int clusters = num_sms / 2;
  if (tiles < clusters) clusters = tiles;
  // cluster dims come from the kernel's __cluster_dims__ attribute; PDL as for the single-CTA kernel
  DGS_CUDA_OK(launch_pdl(kern, dim3(2 * clusters), dim3(g2::THREADS), SMEM, st, tmA, tmB, tmO, tmX ? *tmX : tmO, ep, M, N, K, splits));
  DGS_POST_LAUNCH();
  return DGS_OK;
}

// same contract as gemm_bf16(); requires N % 256 == 0
int gemm_bf16_2cta(const void* A, const void* W, int M, int N, int K, int epi, const GemmEpilogue& ep, cudaStream_t st) {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    DGS_CUDA_OK(cudaGetDevice(&dev));
    DGS_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[1] = {(uint64_t)(ep.lda ? ep.lda : K) * 2};
    uint32_t box[2] = {g2::BK, g2::BM_CTA};
    int rc = make_tmap_bf16(&tmA, A, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[1] = {(uint64_t)(ep.ldb ? ep.ldb : K) * 2};
    uint32_t box[2] = {g2::BK, g2::BN_CTA};
    int rc = make_tmap_bf16(&tmB, W, 2, dims, str, box);
    if (rc) return rc;
  }
  // TMA epilogue (gemm_epilogue.cuh): the inference-mode forms of the three block epilogues -- no aux store, residual
  // updated in place -- with a 16-byte aligned output whose row stride is a multiple of 16 bytes.  DGS_GEMM_TMA_EPI=0 disables.
  static int tma_epi = -1;
  if (tma_epi < 0) {
    const char* e = getenv("DGS_GEMM_TMA_EPI");
    tma_epi = (e && e[0] == '0') ? 0 : 1;
  }
  const bool out_f32 = epi == EPI_GATE_RESID_F32;
  const bool gelu_aux = epi == EPI_BIAS_GELU_BF16 && ep.aux;  // training-mode fc1: also stores the pre-activation
  const bool can_tma = tma_epi && (!ep.aux || gelu_aux) && !ep.resid &&
                       (epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16 || out_f32) && ((uintptr_t)ep.out % 16) == 0 &&
                       ((uintptr_t)ep.aux % 16) == 0 && ((size_t)ep.ldc * (out_f32 ? 4 : 2)) % 16 == 0;
  CUtensorMap tmO = tmA;  // placeholder when the TMA epilogue is not used (never dereferenced)
  if (can_tma) {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, str[1] = {(uint64_t)ep.ldc * (out_f32 ? 4 : 2)};
    uint32_t box[2] = {out_f32 ? 32u : 64u, 32u};  // 128-byte rows x 32 rows
    int rc = out_f32 ? make_tmap_f32(&tmO, ep.out, 2, dims, str, box) : make_tmap_bf16(&tmO, ep.out, 2, dims, str, box);
    if (rc) return rc;
    if (gelu_aux) {
      CUtensorMap tmX;
      rc = make_tmap_bf16(&tmX, ep.aux, 2, dims, str, box);
      if (rc) return rc;
      return launch_2cta<EPI_BIAS_GELU_BF16, false, 2>(tmA, tmB, tmO, ep, M, N, K, num_sms, st, &tmX);
    }
    switch (epi) {
      case EPI_BIAS_BF16: return launch_2cta<EPI_BIAS_BF16, false, 1>(tmA, tmB, tmO, ep, M, N, K, num_sms, st);
      case EPI_BIAS_GELU_BF16: return launch_2cta<EPI_BIAS_GELU_BF16, false, 1>(tmA, tmB, tmO, ep, M, N, K, num_sms, st);
      default: return launch_2cta<EPI_GATE_RESID_F32, false, 1>(tmA, tmB, tmO, ep, M, N, K, num_sms, st);
    }
  }
  switch (epi) {
    case EPI_BIAS_BF16: return launch_2cta<EPI_BIAS_BF16>(tmA, tmB, tmO, ep, M, N, K, num_sms, st);
    case EPI_BIAS_GELU_BF16: return launch_2cta<EPI_BIAS_GELU_BF16>(tmA, tmB, tmO, ep, M, N, K, num_sms, st);
    case EPI_GATE_RESID_F32: return launch_2cta<EPI_GATE_RESID_F32>(tmA, tmB, tmO, ep, M, N, K, num_sms, st);
    case EPI_F32: return launch_2cta<EPI_F32>(tmA, tmB, tmO, ep, M, N, K, num_sms, st);
    case EPI_DGELU_BF16: return launch_2cta<EPI_DGELU_BF16>(tmA, tmB, tmO, ep, M, N, K, num_sms, st);
    default: set_error("gemm: unknown epilogue %d", epi); return DGS_ERR_INVALID_ARGUMENT;
  }
}

// CTA-pair variant of gemm_bf16_tn (C[M,N] fp32 = A^T W, MN-major operands); requires N % 256 == 0
int gemm_bf16_tn_2cta(const void* A, const void* W, int M, int N, int K, const GemmEpilogue& ep, cudaStream_t st) {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    DGS_CUDA_OK(cudaGetDevice(&dev));
    DGS_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int lda = ep.lda ? ep.lda : M, ldb = ep.ldb ? ep.ldb : N;
  CUtensorMap tmA, tmB;
  uint32_t box[2] = {64, g2::BK};  // [64 contiguous M/N elements (128 B) x 64 K rows]
  {
    uint64_t dims[2] = {(uint64_t)M, (uint64_t)K}, str[1] = {(uint64_t)lda * 2};
    int rc = make_tmap_bf16(&tmA, A, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)K}, str[1] = {(uint64_t)ldb * 2};
    int rc = make_tmap_bf16(&tmB, W, 2, dims, str, box);
    if (rc) return rc;
  }
  // Split-K: the output has ceil(M/256) * (N/256) tiles (16 ... 64 for the DiT weight gradients) but K = tokens is long
  // (257 k-blocks at 4 samples): cut K so that the work units fill ~2 waves of the 74 clusters; every unit adds its
  // partial sum into the zeroed output with the TMA (fp32 reduce-add), rows >= M clipped by the tensor map.
  const int tiles = ceil_div(M, 2 * g2::BM_CTA) * (N / g2::BN), num_k = ceil_div(K, g2::BK), clusters = num_sms / 2;
  static int splitk = -1;
  if (splitk < 0) {
    const char* e = getenv("DGS_GEMM_SPLITK");
    splitk = (e && e[0] == '0') ? 0 : 1;
  }
  int splits = 1;
  if (splitk && !ep.bias && tiles < 2 * clusters && num_k >= 32 && ((uintptr_t)ep.out % 16) == 0 &&
      ((size_t)(ep.ldc ? ep.ldc : N) * 4) % 16 == 0) {
    splits = (2 * clusters + tiles - 1) / tiles;            // ~2 waves of work units
    if (splits > num_k / 8) splits = num_k / 8;             // at least 8 k-blocks per unit
    const int kpb = ceil_div(num_k, splits);
    splits = ceil_div(num_k, kpb);                          // no empty unit
  }
  if (splits <= 1) return launch_2cta<EPI_F32, true>(tmA, tmB, tmA, ep, M, N, K, num_sms, st);
  const int ldc = ep.ldc ? ep.ldc : N;
  CUtensorMap tmO;
  {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, str[1] = {(uint64_t)ldc * 4};
    uint32_t obox[2] = {32u, 32u};
    int rc = make_tmap_f32(&tmO, ep.out, 2, dims, str, obox);
    if (rc) return rc;
  }
  DGS_CUDA_OK(cudaMemset2DAsync(ep.out, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st));
  return launch_2cta<EPI_F32, true, 1>(tmA, tmB, tmO, ep, M, N, K, num_sms, st, nullptr, splits);
}

}  // namespace dgs
