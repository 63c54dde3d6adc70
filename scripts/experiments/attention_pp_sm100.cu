// attention_pp_sm100.cu -- forward attention, round-2 generation: two query tiles per CTA in explicit ping-pong.
//
// Same contract as attention_sm100.cu (out = softmax(q k^T / 8) v straight from the fused qkv tensor, head_dim 64, any N,
// optional log-sum-exp for the backward).  Why a new structure (measured, profiles/r2_att_probe_uni.txt, r2_umma_rate.txt):
// in the first-generation kernel the two CTAs of an SM each own one 128-query tile and share the MUFU pipe.  Their
// softmax warps lock IN PHASE -- whoever lags gets the whole pipe while the leader waits for its next S tile, so the lag
// closes -- and then both exponentiate together at half rate (1160 clk per 64-key block) and both wait together with the
// pipe idle (250 clk): MUFU busy 73 %, and taking half of the exponentials off the MUFU pipe gains only 8 %.  A start-up
// stagger does not survive (measured).  The phase has to be ENFORCED:
//   * ONE CTA per SM owns two 128-query tiles of a head; softmax warpgroup t (4 warps, one query row per thread) serves
//     tile t.  The exponentiation phase is handed back and forth with two named barriers: while warpgroup 0 turns S_0 into
//     P_0 alone on the MUFU pipe, warpgroup 1 is in its non-MUFU part (wait for S_1, row max, lazy rescale) and the
//     tensor core runs P_1 V and the next Q_1 K^T -- and vice versa (the FlashAttention-3/4 schedule);
//   * 128-key blocks: S_t = Q_t K_j^T is four tcgen05.mma of N = 128 (70 clk each; the 64-key form costs 53 clk for half the
//     work: the operand fetch from shared memory binds it), P_t V_j eight MMAs with P read from tensor memory, the row sums
//     stay in registers (packed add.f32x2) -- 24 tensor instructions per 128 keys and TWO tiles, K_j / V_j fetched once
//     for both;
//   * tensor memory (all 512 columns): S_0 S_1 (128 fp32 columns each), P_0 P_1 (128 packed bf16 = 64 columns each, their
//     own regions so that S_t of the next block can be issued before P_t V has consumed P_t), O_0 O_1 (64 each);
//   * the softmax makes two passes over S in tensor memory (row max, then exp / pack 32 keys at a time with the next
//     chunk's load in flight), so the register file never holds a whole 128-key row;
//   * a fraction of the exponentials runs as a packed degree-3 polynomial on the FMA pipe (ex2_poly3_x2);
//   * the last key block is issued at its own width (N = 4098 = 32 x 128 + 2).
// Roles (320 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer (converged warp, elect.sync),
// warps 2..5 softmax warpgroup 0, warps 6..9 softmax warpgroup 1.
#include <cstdlib>

#include "dgs_internal.h"
#include "dit_kernels.h"
#include "sm100_ptx.cuh"

namespace dgs {

using namespace ptx;

namespace pp {

// Timeline probe (scripts/att_pp_probe.cu defines DGS_ATT_PROBE and includes this file): per-role cycle accounting; with
// the macro undefined (the product build) these lines do not exist.
#ifdef DGS_ATT_PROBE
__device__ unsigned long long* g_pp_dbg = nullptr;  // [CTAs][32]
#define PP_T0() const long long pp_t0 = clock64()
#define PP_ACC(i) pp_acc[i] += (unsigned long long)(clock64() - pp_t0)
#define PP_DECL(n) unsigned long long pp_acc[n] = {}
#define PP_OUT(slot, i)                                                                                               \
  do {                                                                                                                \
    if (g_pp_dbg) g_pp_dbg[(size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32 + (slot)] = pp_acc[i]; \
  } while (0)
#else
#define PP_T0()
#define PP_ACC(i)
#define PP_DECL(n)
#define PP_OUT(slot, i)
#endif

constexpr int BM = 128, BN = 128, HD = 64, STAGES = 3, THREADS = 320;
constexpr int Q_BYTES = BM * HD * 2;   // 16 KB per query tile
constexpr int KV_BYTES = BN * HD * 2;  // 16 KB per K or V block
constexpr int SMEM_BYTES = 2 * Q_BYTES + 2 * STAGES * KV_BYTES + 1024 + 512;
constexpr uint32_t TM_S = 0, TM_P = 256, TM_O = 384, TM_COLS = 512;  // + tile * {128, 64, 64}
constexpr float RESCALE_THRESHOLD = 8.0f;

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// named barriers 1 / 2 = "warpgroup 0 / 1 may exponentiate" (immediate ids: a register id makes ptxas reserve all 16)
__device__ __forceinline__ void turn_wait(int t) {
  if (t == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
  else asm volatile("bar.sync 2, 256;" ::: "memory");
}
__device__ __forceinline__ void turn_give(int t_other) {
  if (t_other == 0) asm volatile("bar.arrive 1, 256;" ::: "memory");
  else asm volatile("bar.arrive 2, 256;" ::: "memory");
}

template <int POLY_OF_8>
__global__ void __launch_bounds__(THREADS, 1)
attention_fwd_pp_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                        __nv_bfloat16* __restrict__ out, float* __restrict__ lse2, int Np, int N, int H) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;                      // two tiles
  uint8_t* sK = sQ + 2 * Q_BYTES;
  uint8_t* sV = sK + STAGES * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;             // [STAGES]
  uint64_t* v_full = k_full + STAGES;      // [STAGES]
  uint64_t* k_empty = v_full + STAGES;     // [STAGES]  commit after S_1 of the block (both tiles have used K_j)
  uint64_t* v_empty = k_empty + STAGES;    // [STAGES]  commit after P_1 V of the block
  uint64_t* s_full = v_empty + STAGES;     // [2] S_t complete                  (commit)
  uint64_t* s_free = s_full + 2;           // [2] S_t read by all 128 rows      (128 arrivals)
  uint64_t* p_full = s_free + 2;           // [2] P_t stored                    (128 arrivals)
  uint64_t* p_free = p_full + 2;           // [2] P_t V complete                (commit): P_t reusable, O_t up to date
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * BM, h = blockIdx.y, b = blockIdx.z;
  const int n_blocks = (N + BN - 1) / BN;
  const int n_tail = N - (n_blocks - 1) * BN;        // valid keys of the last block, 1..128
  const int tail16 = (n_tail + 15) & ~15;            // ... as the MMA sees it (TMA zero-fills up to here)
  const int D = H * HD;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_q);
    prefetch_tmap(&tm_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; s++) { mbar_init(k_full + s, 1); mbar_init(v_full + s, 1); mbar_init(k_empty + s, 1); mbar_init(v_empty + s, 1); }
    for (int t = 0; t < 2; t++) { mbar_init(s_full + t, 1); mbar_init(s_free + t, 128); mbar_init(p_full + t, 128); mbar_init(p_free + t, 1); }
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
      mbar_arrive_expect_tx(q_full, 2 * Q_BYTES);
      tma_load_3d(sQ, &tm_q, q_full, h * HD, q0, b);
      tma_load_3d(sQ + Q_BYTES, &tm_q, q_full, h * HD, q0 + BM, b);  // rows >= N are zero-filled
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
    // per key block j:  S_0(j)  P_0 V(j-1)  S_1(j)  P_1 V(j-1)  -- each S_t as soon as its rows have read S_t(j-1), each P_t V
    // right behind it, so that tile t's tensor work runs while the OTHER warpgroup owns the MUFU pipe
    constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN, false, false);  // Q_t (K-major) x K_j (K-major), N = 128 keys
    constexpr uint32_t idesc_pv = make_idesc_bf16(BM, HD, false, true);  // P_t (TMEM)    x V_j (MN-major), N = 64 dims
    const uint32_t idesc_s_tail = make_idesc_bf16(BM, tail16, false, false);
    PP_DECL(8);
#ifdef DGS_ATT_PROBE
    const long long pp_begin = clock64();
#endif
    auto issue_pv = [&](int t, int j) {
      const int s = j % STAGES;
      { PP_T0(); mbar_wait(p_full + t, (uint32_t)j & 1); PP_ACC(2 + t); }
      { PP_T0(); mbar_wait(v_full + s, (uint32_t)(j / STAGES) & 1); PP_ACC(4); }
      tc_fence_after();
      const uint32_t vbase = smem_u32(sV + s * KV_BYTES);
      const uint32_t t_p = tmem_base + TM_P + (uint32_t)(t * 64), t_o = tmem_base + TM_O + (uint32_t)(t * 64);
      const int ksteps = (j == n_blocks - 1) ? tail16 / 16 : BN / 16;
      if (elect_one_sync()) {
        for (int k = 0; k < ksteps; k++) {
          // A = P from TMEM: 16 keys = 8 packed columns;  B = V MN-major: 16 keys = 2 groups of 8 rows = 2048 bytes
          const uint64_t vdesc = make_smem_desc_sw128(vbase + (uint32_t)(k * 2048), KV_BYTES, 1024);
          umma_bf16_ts(t_o, t_p + (uint32_t)(k * 8), vdesc, idesc_pv, (j | k) ? 1u : 0u);
        }
        umma_commit(p_free + t);
        if (t == 1) umma_commit(v_empty + s);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    for (int j = 0; j < n_blocks; j++) {
      const int s = j % STAGES;
      { PP_T0(); mbar_wait(k_full + s, (uint32_t)(j / STAGES) & 1); PP_ACC(5); }
      const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + s * KV_BYTES), 16, 1024);
      const uint32_t idesc = (j == n_blocks - 1) ? idesc_s_tail : idesc_s;
#pragma unroll
      for (int t = 0; t < 2; t++) {
        if (j > 0) { PP_T0(); mbar_wait(s_free + t, (uint32_t)(j - 1) & 1); PP_ACC(t); }  // every row of tile t has read S_t(j-1)
        tc_fence_after();
        const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ + t * Q_BYTES), 16, 1024);
        const uint32_t t_s = tmem_base + TM_S + (uint32_t)(t * 128);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < HD / 16; k++) umma_bf16(t_s, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
          umma_commit(s_full + t);
          if (t == 1) umma_commit(k_empty + s);
        }
        __syncwarp();
        if (j > 0) issue_pv(t, j - 1);
      }
    }
    issue_pv(0, n_blocks - 1);
    issue_pv(1, n_blocks - 1);
#ifdef DGS_ATT_PROBE
    pp_acc[6] = (unsigned long long)(clock64() - pp_begin);
    if (lane == 0) { PP_OUT(16, 0); PP_OUT(17, 1); PP_OUT(18, 2); PP_OUT(19, 3); PP_OUT(20, 4); PP_OUT(21, 5); PP_OUT(22, 6); }
#endif
  } else {
    // ===================== softmax warpgroup t = (warp - 2) / 4, one query row per thread =====================
    const int t = (warp - 2) >> 2;
    const int quad = warp & 3;                       // the TMEM lane quarter this warp may access
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t t_s = t_lane + TM_S + (uint32_t)(t * 128), t_p = t_lane + TM_P + (uint32_t)(t * 64);
    const uint32_t t_o = t_lane + TM_O + (uint32_t)(t * 64);
    uint64_t* const my_s_full = s_full + t;
    uint64_t* const my_s_free = s_free + t;
    uint64_t* const my_p_full = p_full + t;
    uint64_t* const my_p_free = p_free + t;
    const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    float m_run = -INFINITY;
    uint64_t l2 = pack_f32x2(0.f, 0.f);  // two partial row sums (packed add)
    if (t == 1) turn_give(0);  // warpgroup 0 goes first
    PP_DECL(8);
#ifdef DGS_ATT_PROBE
    const long long pp_begin = clock64();
#endif

    for (int j = 0; j < n_blocks; j++) {
      const bool last = (j == n_blocks - 1);
      const int kv_valid = last ? n_tail : BN;           // warp-uniform
      const int chunks = (kv_valid + 31) >> 5;           // 32-key chunks that hold at least one valid key
      { PP_T0(); mbar_wait(my_s_full, (uint32_t)j & 1); PP_ACC(0); }
      tc_fence_after();
#ifdef DGS_ATT_PROBE
      const long long pp_p1 = clock64();
#endif
      // ---- pass 1 (no MUFU): row max over the block; S stays in tensor memory; two 32-column loads per wait ----
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
#ifdef DGS_ATT_PROBE
      pp_acc[1] += (unsigned long long)(clock64() - pp_p1);
#endif
      const float m_blk = fmaxf(mx0, mx1);
      float alpha = 1.0f;
      const bool grow = (m_blk - m_run) * sl2 > RESCALE_THRESHOLD;  // true on the first block (m_run = -inf)
      if (grow) {
        alpha = ex2_approx((m_run - m_blk) * sl2);  // 0 on the first block
        m_run = m_blk;
      }
      const bool any_grow = __any_sync(0xffffffffu, grow);
      // P_t V(j-1) must be complete before P_t(j) overwrites the P region and before O_t is rescaled
      if (j >= 1) {
        PP_T0();
        mbar_wait(my_p_free, (uint32_t)(j - 1) & 1);
        PP_ACC(2);
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
      // ---- pass 2 (the exclusive phase): p = 2^(s * sl2 - m), row sum, bf16 pack; the load of chunk c+1 is in flight
      //      while chunk c is exponentiated ----
      const float moff = m_run * sl2;
      const uint64_t sl2_2 = pack_f32x2(sl2, sl2), moff_2 = pack_f32x2(-moff, -moff);
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(t_s, ra);          // chunk 0 travels while we wait for our turn
      { PP_T0(); turn_wait(t); PP_ACC(3); }   // the other warpgroup has finished its exponentiation phase
#ifdef DGS_ATT_PROBE
      const long long pp_p2 = clock64();
#endif
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
        else { tc_fence_before(); mbar_arrive(my_s_free); }          // every score of S_t(j) has left tensor memory
        do_chunk(c, ra);
        if (c + 1 < chunks) {
          tmem_ld_wait();                                            // chunk c+1 is in rb
          if (c + 2 < chunks) tmem_ld_32x32(t_s + (uint32_t)(c * 32 + 64), ra);
          else { tc_fence_before(); mbar_arrive(my_s_free); }
          do_chunk(c + 1, rb);
        }
      }
#ifdef DGS_ATT_PROBE
      pp_acc[4] += (unsigned long long)(clock64() - pp_p2);
#endif
      if (!(t == 1 && last)) turn_give(t ^ 1);                       // hand the MUFU pipe over (arrivals balance the syncs)
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(my_p_full);
    }
#ifdef DGS_ATT_PROBE
    pp_acc[5] = (unsigned long long)(clock64() - pp_begin);
    if (quad == 0 && lane == 0) { for (int i = 0; i < 6; i++) PP_OUT(t * 8 + i, i); }
#endif
    {  // all blocks accumulated -> normalise and store
      mbar_wait(my_p_free, (uint32_t)(n_blocks - 1) & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(t_o, o0);
      tmem_ld_32x32(t_o + 32u, o1);
      tmem_ld_wait();
      float la, lb;
      unpack_f32x2(l2, la, lb);
      const float lsum = la + lb;
      const int qrow = q0 + t * BM + row;
      if (qrow < N) {
        if (lse2) lse2[((size_t)b * H + h) * Np + qrow] = fmaf(m_run, sl2, log2f(lsum));
        const float inv = 1.0f / lsum;
        __nv_bfloat16* dst = out + ((size_t)b * N + qrow) * D + h * HD;
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

}  // namespace pp

int attention_fwd_pp(const void* qkv, void* out, float* lse2, int B, int N, int H, int poly, cudaStream_t st) {
  using namespace pp;
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
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, __nv_bfloat16*, float*, int, int, int);
  static Kern table[5] = {attention_fwd_pp_kernel<0>, attention_fwd_pp_kernel<1>, attention_fwd_pp_kernel<2>,
                          attention_fwd_pp_kernel<3>, attention_fwd_pp_kernel<4>};
  static bool configured = false;
  if (!configured) {
    for (int q = 0; q < 5; q++) DGS_CUDA_OK(cudaFuncSetAttribute(table[q], cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  poly = poly < 0 ? 0 : poly > 4 ? 4 : poly;
  dim3 grid(ceil_div(N, 2 * BM), H, B);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  DGS_CUDA_OK(launch_pdl(table[poly], grid, dim3(THREADS), SMEM_BYTES, st, tm_q, tm_kv, o, lse2, Np, N, H));
  DGS_POST_LAUNCH();
  return DGS_OK;
}

}  // namespace dgs
