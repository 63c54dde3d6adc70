// attention_sk_sm100.cu -- forward attention, round-2 generation: four softmax warpgroups per SM.
//
// Same contract as attention_sm100.cu (out = softmax(q k^T / 8) v straight from the fused qkv tensor, head_dim 64, any N,
// optional log-sum-exp for the backward).  What the round-2 measurements say about the first-generation kernel
// (profiles/r2_att_probe_uni.txt, r2_att_pp_probe.txt, r2_umma_rate.txt):
//   * ONE warp cannot drive the MUFU pipe of its SM sub-partition: alone it retires one ex2 per 16 clk (a 128-key block took
//     2086 clk in a kernel that gave each warp the pipe for itself), the pipe takes one per 8 clk -- at least two warps per
//     sub-partition must be exponentiating at any time;
//   * with exactly two (one 128-query tile per CTA, two CTAs per SM) both are needed ALL the time, but each also spends ~250 clk
//     per 64-key block outside its exp phase (S hand-over, TMEM load / store, bookkeeping), and the pair locks in phase, so
//     the pipe idles 27 % of the time; neither a start-up stagger nor moving half of the exponentials to the FMA pipe helps
//     much (-8 %).
// Hence FOUR softmax warps per sub-partition: one CTA per SM owns TWO 128-query tiles, and every tile has TWO softmax
// warpgroups that split the KEY blocks by parity -- warpgroup (t, p) takes blocks j = p, p+2, ... of tile t with its own
// running max / row sum and its own output accumulator O_{t,p}; the two partial softmaxes of a tile are merged once at the
// end (flash-decoding style: O = (a_0 O_0 + a_1 O_1) / (a_0 l_0 + a_1 l_1), a_p = 2^(m_p - m)).  While one warpgroup waits
// for its P V and its next S, the other three keep the MUFU pipe busy.
//   tensor memory (512 columns): S_{t,p} 4 x 64 fp32 columns (P_{t,p}: packed bf16 over the first 32 columns of its S),
//                                O_{t,p} 4 x 64;  row sums in registers (packed add.f32x2)
//   shared memory: Q_0 Q_1 (32 KB), K / V 64-key blocks through a 4-stage ring, fetched ONCE for both tiles
//   tensor core per key block and tile: S = 4 MMAs (N = 64), P V = 4 MMAs (A = P from tensor memory); issue order
//     P_t V(j) then S_t(j+2) into the buffer P_t(j) just left (in-order pipe, no extra hand-over)
//   softmax: two passes over the 64 scores in tensor memory, 32 at a time (<= 96 registers with 18 warps per SM);
//     POLY_OF_8 of every 8 element pairs take the packed polynomial exp2 on the FMA pipe
// Roles (576 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer (converged warp, elect.sync),
// warps 2..17 softmax: warpgroup g = (warp - 2) / 4 = 2 t + p.
#include <cstdlib>

#include "dgs_internal.h"
#include "dit_kernels.h"
#include "sm100_ptx.cuh"

namespace dgs {

using namespace ptx;

namespace sk {

constexpr int BM = 128, BN = 64, HD = 64, STAGES = 4, THREADS = 576;
constexpr int Q_BYTES = BM * HD * 2;   // 16 KB per query tile
constexpr int KV_BYTES = BN * HD * 2;  // 8 KB per K or V block
constexpr int MERGE_BYTES = 4 * 128 * 8;  // (m, l) of every row of the four warpgroups
constexpr int SMEM_BYTES = 2 * Q_BYTES + 2 * STAGES * KV_BYTES + MERGE_BYTES + 1024 + 512;
constexpr uint32_t TM_S = 0, TM_O = 256, TM_COLS = 512;  // + g * 64, g = 2 t + p
constexpr float RESCALE_THRESHOLD = 8.0f;

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int POLY_OF_8>
__global__ void __launch_bounds__(THREADS, 1)
attention_fwd_sk_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                        __nv_bfloat16* __restrict__ out, float* __restrict__ lse2, int Np, int N, int H) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;                      // two tiles
  uint8_t* sK = sQ + 2 * Q_BYTES;
  uint8_t* sV = sK + STAGES * KV_BYTES;
  float2* s_merge = reinterpret_cast<float2*>(sV + STAGES * KV_BYTES);  // [4][128] (m, l)
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_merge) + MERGE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;             // [STAGES]
  uint64_t* v_full = k_full + STAGES;      // [STAGES]
  uint64_t* kv_empty = v_full + STAGES;    // [STAGES]  commit after P_1 V of the block (both tiles have used K_j and V_j)
  uint64_t* s_full = kv_empty + STAGES;    // [4] S_g complete                       (commit)
  uint64_t* p_full = s_full + 4;           // [4] P_g stored                         (128 arrivals)
  uint64_t* pv_done = p_full + 4;          // [4] P_g V complete                     (commit): O_g up to date
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * BM, h = blockIdx.y, b = blockIdx.z;
  const int n_blocks = (N + BN - 1) / BN;
  const int n_tail = N - (n_blocks - 1) * BN;        // valid keys of the last block, 1..64
  const int tail16 = (n_tail + 15) & ~15;            // ... as the MMA sees it (TMA zero-fills up to here)
  const int D = H * HD;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_q);
    prefetch_tmap(&tm_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; s++) { mbar_init(k_full + s, 1); mbar_init(v_full + s, 1); mbar_init(kv_empty + s, 1); }
    for (int g = 0; g < 4; g++) { mbar_init(s_full + g, 1); mbar_init(p_full + g, 128); mbar_init(pv_done + g, 1); }
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
        mbar_wait(kv_empty + s, ((uint32_t)(j / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(k_full + s, KV_BYTES);
        tma_load_3d(sK + s * KV_BYTES, &tm_kv, k_full + s, D + h * HD, j * BN, b);
        mbar_arrive_expect_tx(v_full + s, KV_BYTES);
        tma_load_3d(sV + s * KV_BYTES, &tm_kv, v_full + s, 2 * D + h * HD, j * BN, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: converged warp, instructions under elect.sync =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN, false, false);  // Q_t (K-major) x K_j (K-major), N = 64 keys
    constexpr uint32_t idesc_pv = make_idesc_bf16(BM, HD, false, true);  // P_g (TMEM)    x V_j (MN-major), N = 64 dims
    const uint32_t idesc_s_tail = make_idesc_bf16(BM, tail16, false, false);
    auto issue_s = [&](int t, int j) {  // S_{t, j&1} = Q_t K_j^T
      const int s = j % STAGES, g = 2 * t + (j & 1);
      mbar_wait(k_full + s, (uint32_t)(j / STAGES) & 1);
      tc_fence_after();
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ + t * Q_BYTES), 16, 1024);
      const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + s * KV_BYTES), 16, 1024);
      const uint32_t t_s = tmem_base + TM_S + (uint32_t)(g * 64);
      const uint32_t idesc = (j == n_blocks - 1) ? idesc_s_tail : idesc_s;
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < HD / 16; k++) umma_bf16(t_s, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
        umma_commit(s_full + g);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    for (int j = 0; j < 2 && j < n_blocks; j++) { issue_s(0, j); issue_s(1, j); }
    for (int j = 0; j < n_blocks; j++) {
      const int s = j % STAGES, p = j & 1;
      const uint32_t ph = (uint32_t)(j >> 1) & 1;  // phase of this parity's barriers
#pragma unroll
      for (int t = 0; t < 2; t++) {
        const int g = 2 * t + p;
        mbar_wait(p_full + g, ph);
        mbar_wait(v_full + s, (uint32_t)(j / STAGES) & 1);
        tc_fence_after();
        const uint32_t vbase = smem_u32(sV + s * KV_BYTES);
        const uint32_t t_p = tmem_base + TM_S + (uint32_t)(g * 64), t_o = tmem_base + TM_O + (uint32_t)(g * 64);
        const int ksteps = (j == n_blocks - 1) ? tail16 / 16 : BN / 16;
        if (elect_one_sync()) {
          for (int k = 0; k < ksteps; k++) {
            // A = P from TMEM: 16 keys = 8 packed columns;  B = V MN-major: 16 keys = 2 groups of 8 rows = 2048 bytes
            const uint64_t vdesc = make_smem_desc_sw128(vbase + (uint32_t)(k * 2048), KV_BYTES, 1024);
            umma_bf16_ts(t_o, t_p + (uint32_t)(k * 8), vdesc, idesc_pv, (j >= 2 || k) ? 1u : 0u);  // first block of this parity overwrites
          }
          umma_commit(pv_done + g);
          if (t == 1) umma_commit(kv_empty + s);
        }
        __syncwarp();
        if (j + 2 < n_blocks) issue_s(t, j + 2);  // into the buffer P_g(j) just left (the pipe executes in issue order)
      }
    }
  } else {
    // ===================== softmax warpgroup g = 2 t + p, one query row per thread =====================
    const int g = (warp - 2) >> 2, t = g >> 1, p = g & 1;
    const int quad = warp & 3;                       // the TMEM lane quarter this warp may access
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t t_s = t_lane + TM_S + (uint32_t)(g * 64), t_o = t_lane + TM_O + (uint32_t)(g * 64);
    uint64_t* const my_s_full = s_full + g;
    uint64_t* const my_p_full = p_full + g;
    uint64_t* const my_pv_done = pv_done + g;
    const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    float m_run = -INFINITY;
    uint64_t l2 = pack_f32x2(0.f, 0.f);  // two partial row sums (packed add)
    int n_mine = 0;

    for (int j = p; j < n_blocks; j += 2, n_mine++) {
      const uint32_t ph = (uint32_t)(j >> 1) & 1;
      const bool last = (j == n_blocks - 1);
      const int kv_valid = last ? n_tail : BN;           // warp-uniform
      mbar_wait(my_s_full, ph);
      tc_fence_after();
      // ---- pass 1: row max (S stays in tensor memory) ----
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; c++) {
        if (c * 32 < kv_valid) {
          uint32_t r[32];
          tmem_ld_32x32(t_s + (uint32_t)(c * 32), r);
          tmem_ld_wait();
          if (last) {
#pragma unroll
            for (int i = 0; i < 32; i++)
              if (c * 32 + i >= kv_valid) r[i] = 0xff800000u;  // -inf: zero-filled / stale tail columns
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
          }
        }
      }
      const float m_blk = fmaxf(mx0, mx1);
      float alpha = 1.0f;
      const bool grow = (m_blk - m_run) * sl2 > RESCALE_THRESHOLD;  // true on this warpgroup's first block (m_run = -inf)
      if (grow) {
        alpha = ex2_approx((m_run - m_blk) * sl2);  // 0 on the first block
        m_run = m_blk;
      }
      if (n_mine >= 1 && __any_sync(0xffffffffu, grow)) {
        // O_g must hold every earlier block of this parity: wait for P_g V of block j-2
        mbar_wait(my_pv_done, (uint32_t)((j - 2) >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; c++) {
          uint32_t o[32];
          tmem_ld_32x32(t_o + (uint32_t)(c * 32), o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32(t_o + (uint32_t)(c * 32), o);
        }
        float la, lb;
        unpack_f32x2(l2, la, lb);
        l2 = pack_f32x2(la * alpha, lb * alpha);
      }
      // ---- pass 2: p = 2^(s * sl2 - m), row sum, bf16 pack; P_g overwrites the head of S_g (32 keys -> 16 columns, so
      //      chunk 1's scores in columns 32..63 are still intact when chunk 0's probabilities land in columns 0..15) ----
      const float moff = m_run * sl2;
      const uint64_t sl2_2 = pack_f32x2(sl2, sl2), moff_2 = pack_f32x2(-moff, -moff);
#pragma unroll
      for (int c = 0; c < 2; c++) {
        if (c * 32 < kv_valid) {
          uint32_t r[32], pk[16];
          tmem_ld_32x32(t_s + (uint32_t)(c * 32), r);
          tmem_ld_wait();
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
          tmem_st_32x16(t_s + (uint32_t)(c * 16), pk);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(my_p_full);
    }
    // ---- merge the two key-parity halves of tile t and store: warpgroup (t, p) writes output dims [32 p, 32 p + 32) ----
    float la, lb;
    unpack_f32x2(l2, la, lb);
    const float l_mine = la + lb;  // 0 when this warpgroup had no block (N <= 64 and p == 1)
    s_merge[g * 128 + row] = make_float2(m_run, l_mine);
    if (n_mine >= 1) {
      const int jl = p + 2 * (n_mine - 1);  // this warpgroup's last block
      mbar_wait(my_pv_done, (uint32_t)(jl >> 1) & 1);
      tc_fence_after();
    }
    tc_fence_before();
    // both warpgroups of the tile: (m, l) published, own accumulator final  (named barrier 1 + t, 256 threads)
    if (t == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
    else asm volatile("bar.sync 2, 256;" ::: "memory");
    tc_fence_after();
    const float2 other = s_merge[(g ^ 1) * 128 + row];
    const float m_all = fmaxf(m_run, other.x);
    const float a_mine = (l_mine > 0.f) ? ex2_approx((m_run - m_all) * sl2) : 0.f;
    const float a_oth = (other.y > 0.f) ? ex2_approx((other.x - m_all) * sl2) : 0.f;
    const float l_all = a_mine * l_mine + a_oth * other.y;
    const int qrow = q0 + t * BM + row;
    uint32_t om[32], oo[32];
    const uint32_t col = (uint32_t)(p * 32);
    // an accumulator that never received a block holds stale tensor memory: it is read (addresses are valid) but not used
    tmem_ld_32x32(t_o + col, om);
    tmem_ld_32x32(t_lane + TM_O + (uint32_t)((g ^ 1) * 64) + col, oo);
    tmem_ld_wait();
    if (qrow < N) {
      if (lse2 && p == 0) lse2[((size_t)b * H + h) * Np + qrow] = fmaf(m_all, sl2, log2f(l_all));
      const float inv = 1.0f / l_all;
      const float wm = a_mine * inv, wo = a_oth * inv;
      __nv_bfloat16* dst = out + ((size_t)b * N + qrow) * D + h * HD + col;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const float x = (a_mine > 0.f) ? __uint_as_float(om[i + e]) * wm : 0.f;
          const float y = (a_oth > 0.f) ? __uint_as_float(oo[i + e]) * wo : 0.f;
          v[e] = x + y;
        }
        uint4 w;
        w.x = pack2_bf16(v[0], v[1]);
        w.y = pack2_bf16(v[2], v[3]);
        w.z = pack2_bf16(v[4], v[5]);
        w.w = pack2_bf16(v[6], v[7]);
        *reinterpret_cast<uint4*>(dst + i) = w;
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

}  // namespace sk

int attention_fwd_sk(const void* qkv, void* out, float* lse2, int B, int N, int H, int poly, cudaStream_t st) {
  using namespace sk;
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
  static Kern table[5] = {attention_fwd_sk_kernel<0>, attention_fwd_sk_kernel<1>, attention_fwd_sk_kernel<2>,
                          attention_fwd_sk_kernel<3>, attention_fwd_sk_kernel<4>};
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
