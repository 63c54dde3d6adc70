// gemm_sm100.cu -- persistent, warp-specialised tcgen05 GEMM for the DiT linears (sm_100a).
//
//   C[M,N] = epilogue( A[M,K] (bf16, row-major) x W[N,K]^T (bf16, row-major = nn.Linear weight) )
//
// One CTA per SM, 192 threads:  warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma
// issuer, warps 2..5 = epilogue (TMEM -> registers -> fused epilogue -> global).  Operands are staged by TMA
// into a 128-byte-swizzled shared-memory ring (BK = 64 bf16 = one swizzle row); fp32 accumulators live in
// TMEM, double-buffered so the epilogue of tile i overlaps the main loop of tile i+1.
// Tile = 128 x BN (BN = 256 or 128), UMMA 128 x BN x 16, cta_group::1.
//
// Fused epilogues = the elementwise tails of the reference DiT block
// (diffusionGS/models/transformers/utils_transformer.py:270-290, timm Attention/Mlp):
//   EPI_BIAS_BF16       y = acc + b                         (qkv)
//   EPI_BIAS_GELU_BF16  y = gelu_tanh(acc + b)              (mlp.fc1 + act)
//   EPI_GATE_RESID_F32  x += gate[sample] * (acc + b)       (attn.proj / mlp.fc2 + gate + residual, fp32 stream)
//   EPI_F32             y = acc (+ b), fp32                 (tokenizer, decoder head, weight gradients)
//   EPI_DGELU_BF16      y = acc * gelu'(u)                  (backward of mlp.fc2 -> act: u = saved pre-activation)
// Training mode (GemmEpilogue::aux / resid): fc1 also stores its pre-activation, the gate epilogues also store the
// pre-gate branch output and may read the residual from a different buffer than they write.
#include <cstdlib>
#include <cstring>

#include "dgs_internal.h"
#include "dit_kernels.h"
#include "gemm_epilogue.cuh"
#include "sm100_ptx.cuh"

namespace dgs {

using namespace ptx;

// ---------------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_tmap_typed(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box);
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  return make_tmap_typed(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  return make_tmap_typed(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}
static int make_tmap_typed(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)"); return DGS_ERR_CUDA; }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; i++) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; i++) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return DGS_ERR_CUDA; }
  return DGS_OK;
}

// ---------------------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------------------
constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator buffers (512 or 256: powers of two)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * 2 * BN * 4 /*bias+gate x2*/;
};

// MN = false:  C = A[M,K] x W[N,K]^T, both operands K-major (rows of 64 K elements = one 128-byte swizzle row).
// MN = true :  C = A^T x W   for A [K, M], W [K, N] row-major, i.e. both operands MN-major (the weight-gradient GEMM
//              dW = dY^T X with K = tokens: neither operand has to be transposed in memory).  A stage then holds
//              BM/64 (resp. BN/64) swizzle atoms of [64 K rows x 64 M/N elements]; the UMMA descriptors use
//              LBO = 8192 B (atom to atom along M/N), SBO = 1024 B (8 K rows), and step 16 K rows = 2048 B per MMA.
template <int BN, int EPI, bool MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmEpilogue ep,
                 int M, int N, int K) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n, num_k = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < Cfg::STAGES; s++) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; s++) { mbar_init(tfull_bar + s, 1); mbar_init(tempty_bar + s, 128); }
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
  griddep_launch_dependents();  // PDL: see sm100_ptx.cuh
  griddep_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * BN;
        for (int kb = 0; kb < num_k; kb++) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          mbar_arrive_expect_tx(full_bar + stage, Cfg::STAGE_BYTES);
          if (!MN) {
            tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, full_bar + stage, kb * BK, m0);
            tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, full_bar + stage, kb * BK, n0);
          } else {  // one [64 K x 64 MN] box per swizzle atom
#pragma unroll
            for (int a = 0; a < BM / 64; a++)
              tma_load_2d(sA + stage * Cfg::A_BYTES + a * 8192, &tmA, full_bar + stage, m0 + a * 64, kb * BK);
#pragma unroll
            for (int a = 0; a < BN / 64; a++)
              tma_load_2d(sB + stage * Cfg::B_BYTES + a * 8192, &tmB, full_bar + stage, n0 + a * 64, kb * BK);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, MN, MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_k; kb++) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES), MN ? 8192 : 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + stage * Cfg::B_BYTES), MN ? 8192 : 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++) {
            // K-major: advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (>>4) address field;
            // MN-major: advance 16 K rows = 2048 bytes: +128
            const uint64_t adv = MN ? (uint64_t)(128 * k) : (uint64_t)(2 * k);
            umma_bf16(d_tmem, adesc + adv, bdesc + adv, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(empty_bar + stage);  // frees the smem slot once these MMAs have read it
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar + acc);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps 2..5 (gemm_epilogue.cuh) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int et = (warp - 2) * 32 + lane;
    float* s_vec_all = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * BN;
      const int row = m0 + quad * 32 + lane;
      float* s_vec = s_vec_all + acc * 2 * BN;
      bool uniform_gate;
      epilogue_stage_vectors<EPI, BN>(ep, s_vec, et, m0, n0, M, N, &uniform_gate);  // under the main loop
      mbar_wait(tfull_bar + acc, acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      epilogue_drain_row<EPI, BN>(ep, s_vec, uniform_gate, t_row, row, n0, M, N);
      tc_fence_before();
      mbar_arrive(tempty_bar + acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
static int g_num_sms = 0;

template <int BN, int EPI, bool MN = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpilogue& ep, int M, int N, int K,
                       cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_kernel<BN, EPI, MN>;
  static bool configured = false;
  if (!configured) {
    DGS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  if (!g_num_sms) {
    int dev = 0;
    DGS_CUDA_OK(cudaGetDevice(&dev));
    DGS_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int tiles = ceil_div(M, BM) * ceil_div(N, BN);
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  DGS_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, st, tmA, tmB, ep, M, N, K));
  DGS_POST_LAUNCH();
  return DGS_OK;
}

int gemm_bf16(const void* A, const void* W, int M, int N, int K, int epi, const GemmEpilogue& ep, cudaStream_t st) {
  DGS_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape %dx%dx%d", M, N, K);
  DGS_REQUIRE(N % 32 == 0, "gemm: need N %% 32 == 0 (got N=%d)", N);
  DGS_REQUIRE((ep.lda ? ep.lda : K) % 8 == 0 && (ep.ldb ? ep.ldb : K) % 8 == 0,
              "gemm: operand row strides must be multiples of 8 elements (K=%d lda=%d ldb=%d)", K, ep.lda, ep.ldb);
  DGS_REQUIRE(epi != EPI_DGELU_BF16 || ep.aux, "gemm: EPI_DGELU_BF16 needs aux = saved pre-activation");
  DGS_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "gemm: operands must be 16-byte aligned");
  // CTA-pair kernel (gemm2_sm100.cu: 256 x 256 tiles per 2-CTA cluster, tcgen05.mma.cta_group::2, 2/3 of the operand
  // traffic per CTA, TMA-store epilogues): the DEFAULT wherever a GEMM has enough 256 x 256 tiles.  Measured on B200
  // (profiles/r1_gemm_probe_v3.txt, r1_bench_v9.json): 1.70 vs 1.55 PFLOP/s at 8192 x 4096 x 4096 and 19-33 % less time
  // on the four DiT linears at N = 4098 than the single-CTA kernel below.  DGS_GEMM_2CTA=0 falls back to the latter.
  static int use_2cta = -1;
  if (use_2cta < 0) {
    const char* e = getenv("DGS_GEMM_2CTA");
    use_2cta = (e && e[0] == '0') ? 0 : 1;
  }
  if (use_2cta && N % 256 == 0 && ceil_div(M, 256) * (N / 256) >= 48) return gemm_bf16_2cta(A, W, M, N, K, epi, ep, st);
  // wide tiles when they still fill the machine, else 128-wide tiles for more CTAs
  static int wide_min = -1;
  if (wide_min < 0) {
    const char* e = getenv("DGS_GEMM_WIDE_MIN");
    wide_min = e ? atoi(e) : 120;
  }
  const bool wide = (N % 256 == 0) && (ceil_div(M, BM) * (N / 256) >= wide_min);
  const int BN = wide ? 256 : 128;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[1] = {(uint64_t)(ep.lda ? ep.lda : K) * 2};
    uint32_t box[2] = {BK, BM};
    int rc = make_tmap_bf16(&tmA, A, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[1] = {(uint64_t)(ep.ldb ? ep.ldb : K) * 2};
    uint32_t box[2] = {BK, (uint32_t)BN};
    int rc = make_tmap_bf16(&tmB, W, 2, dims, str, box);
    if (rc) return rc;
  }
#define DGS_GEMM_CASE(E)                                                                  \
  case E:                                                                                 \
    return wide ? launch_gemm<256, E>(tmA, tmB, ep, M, N, K, st) : launch_gemm<128, E>(tmA, tmB, ep, M, N, K, st);
  switch (epi) {
    DGS_GEMM_CASE(EPI_BIAS_BF16)
    DGS_GEMM_CASE(EPI_BIAS_GELU_BF16)
    DGS_GEMM_CASE(EPI_GATE_RESID_F32)
    DGS_GEMM_CASE(EPI_F32)
    DGS_GEMM_CASE(EPI_DGELU_BF16)
    default:
      set_error("gemm: unknown epilogue %d", epi);
      return DGS_ERR_INVALID_ARGUMENT;
  }
#undef DGS_GEMM_CASE
}

// C[M, N] (fp32) = A^T W for A [K, M], W [K, N] bf16 row-major (lda / ldb = row strides in elements, 0 = M / N):
// the weight-gradient GEMM  dW[n_out, n_in] = dY[tokens, n_out]^T  X[tokens, n_in]  without transposed copies.
int gemm_bf16_tn(const void* A, const void* W, int M, int N, int K, const GemmEpilogue& ep, cudaStream_t st) {
  DGS_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_tn: bad shape %dx%dx%d", M, N, K);
  const int lda = ep.lda ? ep.lda : M, ldb = ep.ldb ? ep.ldb : N;
  DGS_REQUIRE(N % 32 == 0 && lda % 8 == 0 && ldb % 8 == 0, "gemm_tn: need N %% 32 == 0 and row strides %% 8 == 0");
  DGS_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "gemm_tn: operands must be 16-byte aligned");
  {
    // CTA-pair kernel by default (see gemm_bf16); with its split-K path (long K = tokens) even a handful of 256 x 256
    // output tiles fill the machine, so the tile-count condition only applies to short K
    const char* e = getenv("DGS_GEMM_2CTA");
    const int tiles2 = ceil_div(M, 256) * (N / 256);
    if (!(e && e[0] == '0') && N % 256 == 0 && M >= 128 && (tiles2 >= 48 || K >= 2048))
      return gemm_bf16_tn_2cta(A, W, M, N, K, ep, st);
  }
  const bool wide = (N % 256 == 0) && (ceil_div(M, BM) * (N / 256) >= 120);
  CUtensorMap tmA, tmB;
  uint32_t box[2] = {64, BK};  // [64 contiguous M/N elements (128 B) x 64 K rows]
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
  return wide ? launch_gemm<256, EPI_F32, true>(tmA, tmB, ep, M, N, K, st) : launch_gemm<128, EPI_F32, true>(tmA, tmB, ep, M, N, K, st);
}

}  // namespace dgs
