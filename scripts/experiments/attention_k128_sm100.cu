// attention_k128_sm100.cu -- forward attention, second generation (round 2): 128-key blocks.
//
// Same contract as attention_sm100.cu (out = softmax(q k^T / 8) v straight from the fused qkv tensor, head_dim 64, any N,
// optional log-sum-exp for the backward); what changed is the work per synchronisation:
//   * S_j = Q K_j^T covers 128 keys with FOUR tcgen05.mma of N = 128 (the 64-key kernel issued 4 per 64 keys),
//     P_j V_j is 8 MMAs per 128 keys, and the row sums are accumulated in registers with packed add.f32x2 -- 12 tensor
//     instructions per 128 keys instead of 24, one S / P hand-over per 128 keys instead of two;
//   * S (128 fp32 columns) is single-buffered and P_j (128 packed bf16 = 64 columns) has its OWN tensor-memory region, so
//     S_{j+1} is issued as soon as the softmax has READ S_j (s_free), before P_j V_j: TMEM = S 128 + P 64 + O 64 = 256
//     columns, two CTAs per SM.  While one CTA waits for its next S the co-resident CTA owns the MUFU pipe -- the
//     two-tile ping-pong of FlashAttention-4 with the tiles in two CTAs;
//   * the softmax makes two passes over S in tensor memory (row max, then exp / pack chunk by chunk), so a thread never
//     holds more than 32 scores: <= 168 registers with 128-key blocks;
//   * the last key block is issued at its own width (N = 4098 = 32 x 128 + 2: a 16-key S MMA and ONE P V MMA instead of
//     a full block of zero padding).
// Roles (192 threads): warp 0 TMA producer (Q once; K and V through separate 2-stage rings -- K_j is released after S_j,
// V_j after P_j V_j), warp 1 TMEM allocator + MMA issuer (converged warp, elect.sync), warps 2..5 softmax, one query row
// per thread.  O accumulates in TMEM with the lazy (2^8) rescale of the first-generation kernel.
#include <cstdlib>

#include "dgs_internal.h"
#include "dit_kernels.h"
#include "sm100_ptx.cuh"

namespace dgs {

using namespace ptx;

namespace k128 {

constexpr int BM = 128, BN = 128, HD = 64, STAGES = 2, THREADS = 192;
constexpr int Q_BYTES = BM * HD * 2;   // 16 KB
constexpr int KV_BYTES = BN * HD * 2;  // 16 KB per K or V block
constexpr int SMEM_BYTES = Q_BYTES + 2 * STAGES * KV_BYTES + 1024 + 256;
constexpr uint32_t TM_S = 0, TM_P = 128, TM_O = 192, TM_COLS = 256;
constexpr float RESCALE_THRESHOLD = 8.0f;

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int POLY_OF_8>
__global__ void __launch_bounds__(THREADS, 2)
attention_fwd_k128_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                          __nv_bfloat16* __restrict__ out, float* __restrict__ lse2, int Np, int N, int H) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + STAGES * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;             // [STAGES]
  uint64_t* v_full = k_full + STAGES;      // [STAGES]
  uint64_t* k_empty = v_full + STAGES;     // [STAGES]  commit after S_j
  uint64_t* v_empty = k_empty + STAGES;    // [STAGES]  commit after P_j V_j
  uint64_t* s_full = v_empty + STAGES;     // S_j complete            (commit)
  uint64_t* s_free = s_full + 1;           // S_j read by all rows    (128 arrivals)
  uint64_t* p_full = s_free + 1;           // P_j stored              (128 arrivals)
  uint64_t* p_free = p_full + 1;           // P_j V_j complete        (commit): P region reusable, O up to date
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
  const int n_blocks = (N + BN - 1) / BN;
  const int n_tail = N - (n_blocks - 1) * BN;        // valid keys of the last block, 1..128
  const int tail16 = (n_tail + 15) & ~15;            // ... as the MMA sees it (zero-filled by TMA up to here)
  const int D = H * HD;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_q);
    prefetch_tmap(&tm_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; s++) { mbar_init(k_full + s, 1); mbar_init(v_full + s, 1); mbar_init(k_empty + s, 1); mbar_init(v_empty + s, 1); }
    mbar_init(s_full, 1); mbar_init(s_free, 128); mbar_init(p_full, 128); mbar_init(p_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();
  griddep_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_3d(sQ, &tm_q, q_full, h * HD, q0, b);
      for (int j = 0; j < n_blocks; j++) {
        const int s = j % STAGES;
        const uint32_t use = (uint32_t)(j / STAGES);
        mbar_wait(k_empty + s, (use & 1) ^ 1);
        mbar_arrive_expect_tx(k_full + s, KV_BYTES);
        tma_load_3d(sK + s * KV_BYTES, &tm_kv, k_full + s, D + h * HD, j * BN, b);
        mbar_wait(v_empty + s, (use & 1) ^ 1);
        mbar_arrive_expect_tx(v_full + s, KV_BYTES);
        tma_load_3d(sV + s * KV_BYTES, &tm_kv, v_full + s, 2 * D + h * HD, j * BN, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: converged warp, instructions under elect.sync =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN, false, false);  // Q (K-major) x K_j (K-major), N = 128 keys
    constexpr uint32_t idesc_pv = make_idesc_bf16(BM, HD, false, true);  // P (TMEM)    x V_j (MN-major), N = 64 dims
    const uint32_t idesc_s_tail = make_idesc_bf16(BM, tail16, false, false);
    const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
    const uint32_t t_s = tmem_base + TM_S, t_p = tmem_base + TM_P, t_o = tmem_base + TM_O;
    auto issue_pv = [&](int j) {  // O (+)= P_j V_j, then release P / V_j
      const int s = j % STAGES;
      mbar_wait(p_full, (uint32_t)j & 1);
      mbar_wait(v_full + s, (uint32_t)(j / STAGES) & 1);
      tc_fence_after();
      const uint32_t vbase = smem_u32(sV + s * KV_BYTES);
      const int ksteps = (j == n_blocks - 1) ? tail16 / 16 : BN / 16;
      if (elect_one_sync()) {
        for (int k = 0; k < ksteps; k++) {
          // A = P from TMEM: 16 keys = 8 packed columns;  B = V MN-major: 16 keys = 2 groups of 8 rows = 2048 bytes
          const uint64_t vdesc = make_smem_desc_sw128(vbase + (uint32_t)(k * 2048), KV_BYTES, 1024);
          umma_bf16_ts(t_o, t_p + (uint32_t)(k * 8), vdesc, idesc_pv, (j | k) ? 1u : 0u);
        }
        umma_commit(p_free);
        umma_commit(v_empty + s);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    for (int j = 0; j < n_blocks; j++) {
      const int s = j % STAGES;
      mbar_wait(k_full + s, (uint32_t)(j / STAGES) & 1);
      if (j > 0) mbar_wait(s_free, (uint32_t)(j - 1) & 1);  // every row has read S_{j-1}
      tc_fence_after();
      const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + s * KV_BYTES), 16, 1024);
      const uint32_t idesc = (j == n_blocks - 1) ? idesc_s_tail : idesc_s;
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < HD / 16; k++) umma_bf16(t_s, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
        umma_commit(s_full);
        umma_commit(k_empty + s);
      }
      __syncwarp();
      if (j > 0) issue_pv(j - 1);  // after S_j so that the rows can start on block j while P_{j-1} V_{j-1} runs
    }
    issue_pv(n_blocks - 1);
  } else {
    // ===================== softmax / output (warps 2..5): one query row per thread =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t t_s = t_lane + TM_S, t_p = t_lane + TM_P, t_o = t_lane + TM_O;
    const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    float m_run = -INFINITY;
    uint64_t l2 = pack_f32x2(0.f, 0.f);  // two partial row sums (packed add)

    for (int j = 0; j < n_blocks; j++) {
      const bool last = (j == n_blocks - 1);
      const int kv_valid = last ? n_tail : BN;           // warp-uniform
      const int chunks = (kv_valid + 31) >> 5;           // 32-key chunks that hold at least one valid key
      mbar_wait(s_full, (uint32_t)j & 1);
      tc_fence_after();
      // ---- pass 1: row max over the block (S stays in tensor memory); two 32-column loads in flight per wait ----
      float mx0 = -INFINITY, mx1 = -INFINITY;
      for (int c = 0; c < chunks; c += 2) {
        uint32_t ra[32], rb[32];
        const bool two = c + 1 < chunks;  // warp-uniform
        tmem_ld_32x32(t_s + (uint32_t)(c * 32), ra);
        if (two) tmem_ld_32x32(t_s + (uint32_t)(c * 32 + 32), rb);
        tmem_ld_wait();
        if (last) {
#pragma unroll
          for (int i = 0; i < 32; i++) {
            if (c * 32 + i >= kv_valid) ra[i] = 0xff800000u;  // -inf: zero-filled / stale tail columns
            if (c * 32 + 32 + i >= kv_valid) rb[i] = 0xff800000u;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(ra[i]), __uint_as_float(ra[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(ra[i + 2]), __uint_as_float(ra[i + 3])));
        }
        if (two) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(rb[i]), __uint_as_float(rb[i + 1])));
            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(rb[i + 2]), __uint_as_float(rb[i + 3])));
          }
        }
      }
      const float m_blk = fmaxf(mx0, mx1);
      float alpha = 1.0f;
      const bool grow = (m_blk - m_run) * sl2 > RESCALE_THRESHOLD;  // true on the first block (m_run = -inf)
      if (grow) {
        alpha = ex2_approx((m_run - m_blk) * sl2);  // 0 on the first block
        m_run = m_blk;
      }
      const bool any_grow = __any_sync(0xffffffffu, grow);
      // P_{j-1} V_{j-1} must be complete before P_j overwrites the P region and before O is rescaled
      if (j >= 1) {
        mbar_wait(p_free, (uint32_t)(j - 1) & 1);
        tc_fence_after();
      }
      if (j >= 1 && any_grow) {
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(t_o, o0);
        tmem_ld_32x32(t_o + 32u, o1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) {
          o0[i] = __float_as_uint(__uint_as_float(o0[i]) * alpha);
          o1[i] = __float_as_uint(__uint_as_float(o1[i]) * alpha);
        }
        tmem_st_32x32(t_o, o0);
        tmem_st_32x32(t_o + 32u, o1);
        float la, lb;
        unpack_f32x2(l2, la, lb);
        l2 = pack_f32x2(la * alpha, lb * alpha);
      }
      // ---- pass 2: p = 2^(s * sl2 - m), row sum, bf16 pack; the load of chunk c+1 is in flight while chunk c is
      //      exponentiated (two register buffers) ----
      const float moff = m_run * sl2;
      const uint64_t sl2_2 = pack_f32x2(sl2, sl2), moff_2 = pack_f32x2(-moff, -moff);
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(t_s, ra);
      auto do_chunk = [&](int c, uint32_t (&r)[32]) {
        uint32_t pk[16];
        if (last) {
#pragma unroll
          for (int i = 0; i < 32; i++)
            if (c * 32 + i >= kv_valid) r[i] = 0xff800000u;
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
          float x0, x1, p0, p1;
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sl2_2, moff_2), x0, x1);
          if ((i & 7) < POLY_OF_8) ex2_poly3_x2(x0, x1, p0, p1);
          else { p0 = ex2_approx(x0); p1 = ex2_approx(x1); }
          l2 = add_f32x2(l2, pack_f32x2(p0, p1));
          pk[i] = pack2_bf16(p0, p1);
        }
        tmem_st_32x16(t_p + (uint32_t)(c * 16), pk);  // 32 keys = 16 packed columns
      };
      for (int c = 0; c < chunks; c += 2) {
        tmem_ld_wait();                                              // chunk c is in ra
        if (c + 1 < chunks) tmem_ld_32x32(t_s + (uint32_t)(c * 32 + 32), rb);
        else { tc_fence_before(); mbar_arrive(s_free); }             // every score of S_j has left tensor memory
        do_chunk(c, ra);
        if (c + 1 < chunks) {
          tmem_ld_wait();                                            // chunk c+1 is in rb
          if (c + 2 < chunks) tmem_ld_32x32(t_s + (uint32_t)(c * 32 + 64), ra);
          else { tc_fence_before(); mbar_arrive(s_free); }
          do_chunk(c + 1, rb);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    {  // all blocks accumulated -> normalise and store
      mbar_wait(p_free, (uint32_t)(n_blocks - 1) & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(t_o, o0);
      tmem_ld_32x32(t_o + 32u, o1);
      tmem_ld_wait();
      float la, lb;
      unpack_f32x2(l2, la, lb);
      const float lsum = la + lb;
      if (q0 + row < N) {
        if (lse2) lse2[((size_t)b * H + h) * Np + q0 + row] = fmaf(m_run, sl2, log2f(lsum));
        const float inv = 1.0f / lsum;
        __nv_bfloat16* dst = out + ((size_t)b * N + q0 + row) * D + h * HD;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v;
          v.x = pack2_bf16(__uint_as_float(o0[i]) * inv, __uint_as_float(o0[i + 1]) * inv);
          v.y = pack2_bf16(__uint_as_float(o0[i + 2]) * inv, __uint_as_float(o0[i + 3]) * inv);
          v.z = pack2_bf16(__uint_as_float(o0[i + 4]) * inv, __uint_as_float(o0[i + 5]) * inv);
          v.w = pack2_bf16(__uint_as_float(o0[i + 6]) * inv, __uint_as_float(o0[i + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + i) = v;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v;
          v.x = pack2_bf16(__uint_as_float(o1[i]) * inv, __uint_as_float(o1[i + 1]) * inv);
          v.y = pack2_bf16(__uint_as_float(o1[i + 2]) * inv, __uint_as_float(o1[i + 3]) * inv);
          v.z = pack2_bf16(__uint_as_float(o1[i + 4]) * inv, __uint_as_float(o1[i + 5]) * inv);
          v.w = pack2_bf16(__uint_as_float(o1[i + 6]) * inv, __uint_as_float(o1[i + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 32 + i) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

}  // namespace k128

int attention_fwd_k128(const void* qkv, void* out, float* lse2, int B, int N, int H, int poly, cudaStream_t st) {
  using namespace k128;
  const int D = H * HD;
  const int Np = attention_lse_stride(N);
  CUtensorMap tm_q, tm_kv;
  uint64_t dims[3] = {(uint64_t)(3 * D), (uint64_t)N, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)(3 * D) * 2, (uint64_t)N * 3 * D * 2};
  uint32_t box_q[3] = {HD, BM, 1}, box_kv[3] = {HD, BN, 1};
  int rc = make_tmap_bf16(&tm_q, qkv, 3, dims, str, box_q);
  if (rc) return rc;
  rc = make_tmap_bf16(&tm_kv, qkv, 3, dims, str, box_kv);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    DGS_CUDA_OK(cudaFuncSetAttribute(attention_fwd_k128_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    DGS_CUDA_OK(cudaFuncSetAttribute(attention_fwd_k128_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    DGS_CUDA_OK(cudaFuncSetAttribute(attention_fwd_k128_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    DGS_CUDA_OK(cudaFuncSetAttribute(attention_fwd_k128_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    DGS_CUDA_OK(cudaFuncSetAttribute(attention_fwd_k128_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  dim3 grid(ceil_div(N, BM), H, B);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  auto kern = poly == 1 ? attention_fwd_k128_kernel<1> : poly == 2 ? attention_fwd_k128_kernel<2> : poly == 3 ? attention_fwd_k128_kernel<3>
              : poly >= 4 ? attention_fwd_k128_kernel<4> : attention_fwd_k128_kernel<0>;
  DGS_CUDA_OK(launch_pdl(kern, grid, dim3(THREADS), SMEM_BYTES, st, tm_q, tm_kv, o, lse2, Np, N, H));
  DGS_POST_LAUNCH();
  return DGS_OK;
}

}  // namespace dgs
