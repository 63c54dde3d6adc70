// attention2_sm100.cu -- EXPERIMENT (opt-in, DGS_ATT_TPR2=1): the forward attention kernel of attention_sm100.cu with TWO
// softmax threads per query row (8 softmax warps per CTA, 16 per SM instead of 8).
//
// Why: the default kernel keeps the MUFU pipe -- its binding unit -- 57 % busy: with one softmax warp per scheduler and CTA
// (two per scheduler and SM) the pipe idles whenever both warps are between their exponential phases (TMEM load, row
// max, store of P, barrier).  Four warps per scheduler should overlap those phases better.  Each thread takes 32 of a
// block's 64 key columns; the row max is completed through shared memory under a 64-thread named barrier of the two warps
// that share a TMEM lane quadrant.  TMA producer, MMA issuer, TMEM layout, barriers and numerics are those of the default
// kernel.  An earlier two-threads-per-row version (r1 "v4", P through shared memory) was slower because the MMAs' shared-memory
// operand fetch was the limit then; P now lives in tensor memory.
#include <cstdlib>

#include "dgs_internal.h"
#include "dit_kernels.h"
#include "sm100_ptx.cuh"

namespace dgs {

using namespace ptx;

namespace {
constexpr int ATT_BM = 128, ATT_BN = 64, ATT_HD = 64, ATT_KV_STAGES = 4, ATT2_THREADS = 320;
constexpr int ATT_Q_BYTES = ATT_BM * ATT_HD * 2;
constexpr int ATT_KV_BYTES = ATT_BN * ATT_HD * 2;
constexpr int ATT_ONES_BYTES = 16 * 128;
constexpr int ATT2_XMAX_BYTES = 2 * 2 * ATT_BM * 4;
constexpr int ATT2_SMEM_BYTES = ATT_Q_BYTES + 2 * ATT_KV_STAGES * ATT_KV_BYTES + ATT_ONES_BYTES + 1024 + 256 + ATT2_XMAX_BYTES;
constexpr uint32_t TMEM_S = 0, TMEM_O = 2 * ATT_BN, TMEM_L = TMEM_O + ATT_HD, ATT_TMEM_COLS = 256;
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(ATT2_THREADS, 2)
attention_fwd_tpr2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     __nv_bfloat16* __restrict__ out, float* __restrict__ lse2, int Np, int N, int H) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_Q_BYTES;
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_KV_BYTES;
  uint8_t* sOnes = sV + ATT_KV_STAGES * ATT_KV_BYTES;  // 2 KB, 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + ATT_ONES_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + ATT_KV_STAGES;
  uint64_t* kv_empty = v_full + ATT_KV_STAGES;
  uint64_t* s_full = kv_empty + ATT_KV_STAGES;
  uint64_t* p_full = s_full + 2;
  uint64_t* pv_full = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + 2);
  float* s_xmax = reinterpret_cast<float*>(bars) + 64;  // [2 block parities][2 halves][128 rows], after the 256-byte barrier area

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM, h = blockIdx.y, b = blockIdx.z;
  const int n_blocks = (N + ATT_BN - 1) / ATT_BN;
  const int D = H * ATT_HD;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_q);
    prefetch_tmap(&tm_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT_KV_STAGES; s++) { mbar_init(k_full + s, 1); mbar_init(v_full + s, 1); mbar_init(kv_empty + s, 1); }
    for (int s = 0; s < 2; s++) { mbar_init(s_full + s, 1); mbar_init(p_full + s, 256); mbar_init(pv_full + s, 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < ATT_ONES_BYTES / 4; i += ATT2_THREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
  fence_proxy_async();  // the ones tile is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();  // programmatic dependent launch: the next kernel may start its prologue now ...
  griddep_wait();               // ... and this one touches global memory only after its predecessor has completed

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, ATT_Q_BYTES);
      tma_load_3d(sQ, &tm_q, q_full, h * ATT_HD, q0, b);
      for (int j = 0; j < n_blocks; j++) {
        const int s = j % ATT_KV_STAGES;
        const uint32_t use = (uint32_t)(j / ATT_KV_STAGES);
        mbar_wait(kv_empty + s, (use & 1) ^ 1);
        mbar_arrive_expect_tx(k_full + s, ATT_KV_BYTES);
        tma_load_3d(sK + s * ATT_KV_BYTES, &tm_kv, k_full + s, D + h * ATT_HD, j * ATT_BN, b);
        mbar_arrive_expect_tx(v_full + s, ATT_KV_BYTES);
        tma_load_3d(sV + s * ATT_KV_BYTES, &tm_kv, v_full + s, 2 * D + h * ATT_HD, j * ATT_BN, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, false, false);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BM, ATT_HD, false, true);   // P (K-major) x V (MN-major)
      constexpr uint32_t idesc_l = make_idesc_bf16(ATT_BM, 16, false, false);      // P (K-major) x ONES (K-major)
      const uint64_t odesc = make_smem_desc_sw128(smem_u32(sOnes), 16, 1024);
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      auto issue_s = [&](int j) {
        const int s = j % ATT_KV_STAGES;
        mbar_wait(k_full + s, (uint32_t)(j / ATT_KV_STAGES) & 1);
        tc_fence_after();
        const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + s * ATT_KV_BYTES), 16, 1024);
        const uint32_t d = tmem_base + TMEM_S + (uint32_t)((j & 1) * ATT_BN);
#pragma unroll
        for (int k = 0; k < ATT_HD / 16; k++) umma_bf16(d, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k ? 1u : 0u);
        umma_commit(s_full + (j & 1));
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_blocks; j++) {
        if (j + 1 < n_blocks) issue_s(j + 1);
        const int s = j % ATT_KV_STAGES;
        mbar_wait(p_full + (j & 1), (uint32_t)(j >> 1) & 1);
        mbar_wait(v_full + s, (uint32_t)(j / ATT_KV_STAGES) & 1);
        tc_fence_after();
        const uint32_t p_tmem = tmem_base + TMEM_S + (uint32_t)((j & 1) * ATT_BN);  // P_j: packed bf16 over S_j
        const uint32_t vbase = smem_u32(sV + s * ATT_KV_BYTES);
        const uint32_t d = tmem_base + TMEM_O;
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; k++) {
          // A = P from TMEM: 16 keys = 8 packed columns;  B = V: MN-major ([key][64 dims] rows of 128 bytes),
          // 16 keys = 2 groups of 8 rows = 2048 bytes
          const uint64_t vdesc = make_smem_desc_sw128(vbase + (uint32_t)(k * 2048), ATT_KV_BYTES, 1024);
          umma_bf16_ts(d, p_tmem + (uint32_t)(k * 8), vdesc, idesc_pv, (j | k) ? 1u : 0u);  // O += P_j V_j
          umma_bf16_ts(tmem_base + TMEM_L, p_tmem + (uint32_t)(k * 8), odesc + (uint64_t)(2 * k), idesc_l, (j | k) ? 1u : 0u);
        }
        umma_commit(pv_full + (j & 1));
        umma_commit(kv_empty + s);
      }
    }
  } else {
    // ===================== softmax / output (warps 2..9): TWO threads per query row =====================
    // warps w and w + 4 share a TMEM lane quadrant; thread `half` of a row owns key columns [32 half, 32 half + 32) of every
    // S block, probability words [16 half, 16 half + 16) and output dims [32 half, 32 half + 32).
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint32_t t_o = t_lane + TMEM_O + (uint32_t)(half * 32), t_l = t_lane + TMEM_L;
    float m_run = -INFINITY;

    for (int j = 0; j < n_blocks; j++) {
      const int buf = j & 1;
      mbar_wait(s_full + buf, (uint32_t)(j >> 1) & 1);
      tc_fence_after();
      const uint32_t t_s = t_lane + TMEM_S + (uint32_t)(buf * ATT_BN);
      const int kv_valid = N - j * ATT_BN - half * 32;  // valid keys among this thread's 32 columns (may be <= 0)
      uint32_t r[32];
      tmem_ld_32x32(t_s + (uint32_t)(half * 32), r);
      tmem_ld_wait();
      if (kv_valid < 32) {  // warp-uniform: mask the zero-filled tail keys
#pragma unroll
        for (int i = 0; i < 32; i++)
          if (i >= kv_valid) r[i] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(r[i])); mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(r[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
      }
      // row max = max over both halves: exchanged through shared memory, double-buffered by block parity.  The named
      // barrier of the two warps of this lane quadrant also orders "both threads have loaded their S columns" before
      // either overwrites the head of S with its probabilities (P words 16..31 lie over the OTHER thread's S columns).
      s_xmax[(buf * 2 + half) * ATT_BM + row] = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      const float m_blk = fmaxf(s_xmax[(buf * 2) * ATT_BM + row], s_xmax[(buf * 2 + 1) * ATT_BM + row]);
      // lazy rescale: keep the stale max unless it is exceeded by more than the threshold (both threads of a row see the
      // same m_blk and m_run, hence take the same decision)
      float alpha = 1.0f;
      const bool grow = (m_blk - m_run) * sl2 > ATT_RESCALE_THRESHOLD;  // true on the first block (m_run = -inf)
      if (grow) {
        alpha = ex2_approx((m_run - m_blk) * sl2);  // 0 on the first block
        m_run = m_blk;
      }
      if (j >= 1 && __any_sync(0xffffffffu, grow)) {  // O must hold every earlier block
        mbar_wait(pv_full + (buf ^ 1), (uint32_t)((j - 1) >> 1) & 1);
        tc_fence_after();
        uint32_t q[32];
        tmem_ld_32x32(t_o, q);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) q[i] = __float_as_uint(__uint_as_float(q[i]) * alpha);
        tmem_st_32x32(t_o, q);
        if (half == 0) {
          const uint32_t lsum = tmem_ld_32x1(t_l);
          tmem_ld_wait();
          tmem_st_32x1(t_l, __float_as_uint(__uint_as_float(lsum) * alpha));
        }
        tmem_st_wait();
      }
      const float moff = m_run * sl2;
      uint32_t pk[16];  // 32 probabilities, two bf16 per word
      const uint64_t sl2_2 = pack_f32x2(sl2, sl2), moff_2 = pack_f32x2(-moff, -moff);  // x = s * sl2 - moff as FFMA2
#pragma unroll
      for (int i = 0; i < 16; i++) {
        float x0, x1;
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sl2_2, moff_2), x0, x1);
        pk[i] = pack2_bf16(ex2_approx(x0), ex2_approx(x1));
      }
      tmem_st_32x16(t_s + (uint32_t)(half * 16), pk);
      tmem_st_wait();
      tc_fence_before();  // our tcgen05.ld of S_j / O and the store of P_j are complete before the issuer proceeds
      mbar_arrive(p_full + buf);
    }
    {  // all blocks accumulated -> normalise and store this thread's 32 output dims
      const int last = n_blocks - 1;
      mbar_wait(pv_full + (last & 1), (uint32_t)(last >> 1) & 1);
      tc_fence_after();
      uint32_t q[32];
      tmem_ld_32x32(t_o, q);
      const uint32_t lsum = tmem_ld_32x1(t_l);
      tmem_ld_wait();
      if (q0 + row < N) {
        if (lse2 && half == 0) lse2[((size_t)b * H + h) * Np + q0 + row] = fmaf(m_run, sl2, log2f(__uint_as_float(lsum)));
        const float inv = 1.0f / __uint_as_float(lsum);
        __nv_bfloat16* dst = out + ((size_t)b * N + q0 + row) * D + h * ATT_HD + half * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v;
          v.x = pack2_bf16(__uint_as_float(q[i]) * inv, __uint_as_float(q[i + 1]) * inv);
          v.y = pack2_bf16(__uint_as_float(q[i + 2]) * inv, __uint_as_float(q[i + 3]) * inv);
          v.z = pack2_bf16(__uint_as_float(q[i + 4]) * inv, __uint_as_float(q[i + 5]) * inv);
          v.w = pack2_bf16(__uint_as_float(q[i + 6]) * inv, __uint_as_float(q[i + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + i) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

}  // namespace

int attention_fwd_tpr2(const void* qkv, void* out, float* lse2, int B, int N, int H, cudaStream_t st) {
  DGS_REQUIRE(B > 0 && N > 0 && H > 0, "attention: bad shape B=%d N=%d H=%d", B, N, H);
  const int D = H * ATT_HD;
  const int Np = attention_lse_stride(N);
  CUtensorMap tm_q, tm_kv;
  uint64_t dims[3] = {(uint64_t)(3 * D), (uint64_t)N, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)(3 * D) * 2, (uint64_t)N * 3 * D * 2};
  uint32_t box_q[3] = {ATT_HD, ATT_BM, 1}, box_kv[3] = {ATT_HD, ATT_BN, 1};
  int rc = make_tmap_bf16(&tm_q, qkv, 3, dims, str, box_q);
  if (rc) return rc;
  rc = make_tmap_bf16(&tm_kv, qkv, 3, dims, str, box_kv);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    DGS_CUDA_OK(cudaFuncSetAttribute(attention_fwd_tpr2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM_BYTES));
    configured = true;
  }
  dim3 grid(ceil_div(N, ATT_BM), H, B);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  DGS_CUDA_OK(launch_pdl(attention_fwd_tpr2_kernel, grid, dim3(ATT2_THREADS), ATT2_SMEM_BYTES, st, tm_q, tm_kv, o, lse2, Np, N, H));
  DGS_POST_LAUNCH();
  return DGS_OK;
}

}  // namespace dgs
