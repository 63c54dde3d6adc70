// gemm3_sm100.cu -- 256 x 256 tile, single CTA: the operand-traffic variant of the DiT GEMM.
//
// EXPERIMENT, opt-in (DGS_GEMM_M256=1), slower than both gemm_sm100.cu and the default CTA-pair kernel gemm2_sm100.cu.
// Premise at the time: the 128 x 256 kernel moves 48 KB of A/B per 128x256x64 MAC block (43 MAC/B) and looked bound by
// operand delivery.  (The CTA-pair kernel later showed what the single-CTA mainloop really loses: see DESIGN.md section 4.)
// Here one CTA computes TWO 128-row halves against the SAME 256-column B block: 64 KB per 256x256x64 block = 64 MAC/B,
// 1.5x less traffic per FLOP.  Each k-step issues two tcgen05.mma (top / bottom half) that share the B descriptor.
// The two 128 x 256 fp32 accumulators fill all 512 TMEM columns, so the accumulator is NOT double-buffered; to keep the
// exposed epilogue short it is drained by 8 epilogue warps (two per TMEM lane quadrant: one per half) side by side.
// 320 threads: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), warps 2..9 epilogue.  3-stage 64 KB ring.
#include "dgs_internal.h"
#include "dit_kernels.h"
#include "gemm_epilogue.cuh"
#include "sm100_ptx.cuh"

namespace dgs {

using namespace ptx;

namespace g3 {

constexpr int BM = 256, BN = 256, BK = 64, UMMA_K = 16, STAGES = 3, THREADS = 320;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;  // 32 KB + 32 KB
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + 2 * BN * 4;

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_m256_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmEpilogue ep,
                      int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);
  float* s_vec = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n, num_k = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 256);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * BN;
        for (int kb = 0; kb < num_k; kb++) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          mbar_arrive_expect_tx(full_bar + stage, STAGE_BYTES);
          tma_load_2d(sA + stage * A_BYTES, &tmA, full_bar + stage, kb * BK, m0);  // 256 rows: top half, bottom half
          tma_load_2d(sB + stage * B_BYTES, &tmB, full_bar + stage, kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, false, false);
      int stage = 0;
      uint32_t phase = 0, tile_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar, tile_phase ^ 1);  // both accumulators drained
        tc_fence_after();
        for (int kb = 0; kb < num_k; kb++) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint64_t adesc_top = make_smem_desc_sw128(smem_u32(sA + stage * A_BYTES), 16, 1024);
          const uint64_t adesc_bot = make_smem_desc_sw128(smem_u32(sA + stage * A_BYTES + 128 * 128), 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + stage * B_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++) {
            umma_bf16(tmem_base, adesc_top + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            umma_bf16(tmem_base + BN, adesc_bot + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(empty_bar + stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar);
        tile_phase ^= 1;
      }
    }
  } else {
    // ===================== 8 epilogue warps: (quadrant, half) =====================
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int et = (warp - 2) * 32 + lane;
    uint32_t tile_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * BN;
      const int row = m0 + half * 128 + quad * 32 + lane;
      bool uniform_gate;
      epilogue_stage_vectors<EPI, BN, 256, 256>(ep, s_vec, et, m0, n0, M, N, &uniform_gate);
      mbar_wait(tfull_bar, tile_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * BN);
      epilogue_drain_row<EPI, BN>(ep, s_vec, uniform_gate, t_row, row, n0, M, N);
      tc_fence_before();
      mbar_arrive(tempty_bar);
      // s_vec is single-buffered: nobody may restage it for the next tile before everyone has finished this one
      epi_bar_sync<256>();
      tile_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace g3

template <int EPI>
static int launch_m256(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpilogue& ep, int M, int N, int K,
                       int num_sms, cudaStream_t st) {
  auto kern = g3::gemm_bf16_m256_kernel<EPI>;
  static bool configured = false;
  if (!configured) {
    DGS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g3::SMEM_BYTES));
    configured = true;
  }
  const int tiles = ceil_div(M, g3::BM) * ceil_div(N, g3::BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, g3::THREADS, g3::SMEM_BYTES, st>>>(tmA, tmB, ep, M, N, K);
  DGS_POST_LAUNCH();
  return DGS_OK;
}

// same contract as gemm_bf16(); requires N % 256 == 0
int gemm_bf16_m256(const void* A, const void* W, int M, int N, int K, int epi, const GemmEpilogue& ep, cudaStream_t st) {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    DGS_CUDA_OK(cudaGetDevice(&dev));
    DGS_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {g3::BK, g3::BM};
    int rc = make_tmap_bf16(&tmA, A, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {g3::BK, g3::BN};
    int rc = make_tmap_bf16(&tmB, W, 2, dims, str, box);
    if (rc) return rc;
  }
  switch (epi) {
    case EPI_BIAS_BF16: return launch_m256<EPI_BIAS_BF16>(tmA, tmB, ep, M, N, K, num_sms, st);
    case EPI_BIAS_GELU_BF16: return launch_m256<EPI_BIAS_GELU_BF16>(tmA, tmB, ep, M, N, K, num_sms, st);
    case EPI_GATE_RESID_F32: return launch_m256<EPI_GATE_RESID_F32>(tmA, tmB, ep, M, N, K, num_sms, st);
    case EPI_F32: return launch_m256<EPI_F32>(tmA, tmB, ep, M, N, K, num_sms, st);
    default: set_error("gemm: unknown epilogue %d", epi); return DGS_ERR_INVALID_ARGUMENT;
  }
}

}  // namespace dgs
