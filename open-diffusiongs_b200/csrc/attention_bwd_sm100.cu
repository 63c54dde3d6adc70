// attention_bwd_sm100.cu -- backward of the non-causal multi-head attention on tcgen05 (sm_100a), head_dim 64.
//
// Backward of  O = softmax(Q K^T / 8) V  (timm Attention.forward -> F.scaled_dot_product_attention, instantiated at
// utils_transformer.py:254-256; the reference differentiates it through torch autograd / the SDPA backward).
// With P = exp2(S * c - lse2) (c = log2(e)/8, lse2 saved by the forward), Dsum = rowsum(dO * O):
//     dV = P^T dO          dP = dO V^T          dS = P * (dP - Dsum) / 8          dQ = dS K          dK = dS^T Q
// Two kernels, no atomics, every operand read straight from the [B, N, 3, H, 64] qkv tensor / the [B, N, H*64] dO
// tensor through 3-D TMA maps and written straight into the [B, N, 3, H, 64] dqkv tensor:
//   attn_bwd_dq_kernel   one CTA per (128 queries, head, sample), loops over 64-key blocks:
//                        S = Q K_j^T, dP = dO V_j^T (TMEM) -> dS_j (bf16, written back into TENSOR MEMORY over dP_j) ->
//                        dQ += dS_j K_j (A operand from TMEM, TMEM accumulator; K_j consumed MN-major from the same
//                        smem tile that fed S)
//   attn_bwd_dkv_kernel  one CTA per (128 keys, head, sample), loops over 64-query blocks, works on the TRANSPOSED
//                        scores so that the key is the TMEM lane / the thread:  S^T = K Q_i^T, dP^T = V dO_i^T ->
//                        P^T, dS^T (bf16, written back into TENSOR MEMORY over dP^T: the dV / dK MMAs read their A operands
//                        from TMEM) -> dV += P^T dO_i, dK += dS^T Q_i (dO_i, Q_i consumed MN-major from the tiles that
//                        fed S^T / dP^T)
// Both: 320 threads (TMA warp, MMA warp, 8 softmax warps), 2 CTAs per SM, 256 TMEM columns.  No row reductions are
// needed in the backward (lse2 and Dsum are inputs), so a row is split between two threads (32 of the 64 columns each)
// at no cost: 16 softmax warps per SM hide the TMEM-load / MUFU latencies that bound the 4-warp version
// (profiles/r1_ncu_attn_bwd_v1.txt: XU 27-35 %, issue 31-38 %, 12 warps resident).  Per element: exp2 + 3 FMA-pipe ops
// + half a pack; the tail-key mask is applied only in the last key block.
#include "dgs_internal.h"
#include "dit_kernels.h"
#include "sm100_ptx.cuh"

namespace dgs {

using namespace ptx;

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

namespace {

constexpr int AB_THREADS = 320, AB_HD = 64;  // TMA warp + MMA warp + 8 softmax warps
constexpr int AB_SOFT = 256;                // softmax threads: two per row / key, 32 columns each
constexpr int AB_T128 = 128 * AB_HD * 2;  // [128 x 64] bf16 tile, 16 KB
constexpr int AB_T64 = 64 * AB_HD * 2;    // [64 x 64] bf16 tile, 8 KB
constexpr int AB_STAGES = 2;
constexpr float AB_SL2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
constexpr float AB_SCALE = 0.125f;

__device__ __forceinline__ uint32_t ab_pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Dsum[b, h, n] = sum_d O[b, n, h, d] * dO[b, n, h, d];  pads n in [N, Np): lse2 = +inf (=> P = 0), Dsum = 0
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ O, const __nv_bfloat16* __restrict__ dO,
                                     float* __restrict__ lse2, float* __restrict__ dsum, int N, int Np, int H) {
  const int n = blockIdx.x, b = blockIdx.y, t = threadIdx.x;  // blockDim.x = H * 16
  const int D = H * AB_HD;
  griddep_launch_dependents();  // PDL (sm100_ptx.cuh)
  griddep_wait();
  if (n >= N) {
    if (t < H) {
      lse2[((size_t)b * H + t) * Np + n] = INFINITY;
      dsum[((size_t)b * H + t) * Np + n] = 0.f;
    }
    return;
  }
  float s = 0.f;
  if (t < H * 16) {  // (blockDim.x is padded to a full warp when H == 1)
    const size_t off = ((size_t)b * N + n) * D + (size_t)t * 4;
    const uint2 a = *reinterpret_cast<const uint2*>(O + off);
    const uint2 g = *reinterpret_cast<const uint2*>(dO + off);
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const float2 x = __bfloat1622float2(a2[i]), y = __bfloat1622float2(g2[i]);
      s = fmaf(x.x, y.x, s);
      s = fmaf(x.y, y.y, s);
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((t & 15) == 0 && t < H * 16) dsum[((size_t)b * H + (t >> 4)) * Np + n] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// dQ
// ---------------------------------------------------------------------------------------------------------------
constexpr int DQ_STAGES = 3;
constexpr int DQ_SMEM = 2 * AB_T128 /*Q, dO*/ + 2 * DQ_STAGES * AB_T64 /*K, V*/ + 1024 + 256;
constexpr uint32_t DQ_TM_S = 0 /* 2 buffers x 64 */, DQ_TM_DP = 128, DQ_TM_DQ = 192, DQ_TMEM_COLS = 256;

// UNI (default; DGS_ATT_UNI=0 selects the old path): MMA issue by the converged warp under elect.sync (see attention_sm100.cu / sm100_ptx.cuh)
constexpr int ATTB_POLY_DEFAULT = 1;  // index into {0, 2, 4} of 8 pairs; measured r2 (B = 1 / 4): 295 / 1064 -> 284 / 1033 us with 2 of 8, 289 / 1047 with 4 of 8

template <bool UNI, int POLY>
__global__ void __launch_bounds__(AB_THREADS, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tm_q128, const __grid_constant__ CUtensorMap tm_kv64,
                   const __grid_constant__ CUtensorMap tm_do128, const float* __restrict__ lse2,
                   const float* __restrict__ dsum, __nv_bfloat16* __restrict__ dqkv, int N, int Np, int H) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + AB_T128;
  uint8_t* sK = sdO + AB_T128;
  uint8_t* sV = sK + DQ_STAGES * AB_T64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + DQ_STAGES * AB_T64);
  uint64_t* q_full = bars;                    // Q + dO landed
  uint64_t* kv_full = bars + 1;               // [DQ_STAGES]
  uint64_t* kv_empty = kv_full + DQ_STAGES;   // [DQ_STAGES]
  uint64_t* s_full = kv_empty + DQ_STAGES;    // [2] S of block j in TMEM (double-buffered: issued one block ahead)
  uint64_t* dp_full = s_full + 2;             // dP of block j in TMEM (single buffer)
  uint64_t* ds_full = dp_full + 1;            // [2] dS_j written over dP_j in TMEM (and S_j read out)
  uint64_t* dq_done = ds_full + 2;            // [2] dQ MMA of block j retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int n_blocks = (N + 63) / 64;
  const int D = H * AB_HD;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_q128);
    prefetch_tmap(&tm_kv64);
    prefetch_tmap(&tm_do128);
    mbar_init(q_full, 1);
    for (int s = 0; s < DQ_STAGES; s++) { mbar_init(kv_full + s, 1); mbar_init(kv_empty + s, 1); }
    mbar_init(dp_full, 1);
    for (int s = 0; s < 2; s++) { mbar_init(s_full + s, 1); mbar_init(ds_full + s, AB_SOFT); mbar_init(dq_done + s, 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, DQ_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();  // PDL (sm100_ptx.cuh)
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * AB_T128);
      tma_load_3d(sQ, &tm_q128, q_full, h * AB_HD, q0, b);
      tma_load_3d(sdO, &tm_do128, q_full, h * AB_HD, q0, b);
      for (int j = 0; j < n_blocks; j++) {
        const int s = j % DQ_STAGES;
        mbar_wait(kv_empty + s, ((uint32_t)(j / DQ_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(kv_full + s, 2 * AB_T64);
        tma_load_3d(sK + s * AB_T64, &tm_kv64, kv_full + s, D + h * AB_HD, j * 64, b);
        tma_load_3d(sV + s * AB_T64, &tm_kv64, kv_full + s, 2 * D + h * AB_HD, j * 64, b);
      }
    }
  } else if (warp == 1) {
    if (UNI || lane == 0) {
      constexpr uint32_t idesc_kk = make_idesc_bf16(128, 64, false, false);  // A K-major x B K-major
      constexpr uint32_t idesc_kmn = make_idesc_bf16(128, 64, false, true);  // A K-major x B MN-major
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t dodesc = make_smem_desc_sw128(smem_u32(sdO), 16, 1024);
      // Issue order (the softmax needs S first and dP only for its second half, so S runs one block ahead):
      //   S(0) dP(0) | S(1) .. wait dS(0) .. dQ(0) dP(1) | S(2) .. wait dS(1) .. dQ(1) dP(2) | ...
      auto issue_s = [&](int j) {
        const int s = j % DQ_STAGES;
        mbar_wait(kv_full + s, (uint32_t)(j / DQ_STAGES) & 1);
        tc_fence_after();
        const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + s * AB_T64), 16, 1024);
        const uint32_t d = tmem_base + DQ_TM_S + (uint32_t)((j & 1) * 64);
        if (!UNI || elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; k++) umma_bf16(d, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u);
          umma_commit(s_full + (j & 1));
        }
      };
      auto issue_dp = [&](int j) {  // kv_full(j) already observed by issue_s(j)
        const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV + (j % DQ_STAGES) * AB_T64), 16, 1024);
        if (!UNI || elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; k++)
            umma_bf16(tmem_base + DQ_TM_DP, dodesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u);
          umma_commit(dp_full);
        }
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      issue_dp(0);
      for (int j = 0; j < n_blocks; j++) {
        const int s = j % DQ_STAGES;
        // S buffer (j+1)&1 was last read by the softmax of block j-1, whose ds_full was observed one iteration ago
        if (j + 1 < n_blocks) issue_s(j + 1);
        mbar_wait(ds_full + (j & 1), (uint32_t)(j >> 1) & 1);
        tc_fence_after();
        const uint32_t kbase = smem_u32(sK + s * AB_T64);
        if (!UNI || elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          // A = dS_j from TENSOR MEMORY: packed bf16 pairs written by the softmax threads over their own dP columns --
          // keys 0..31 in columns [0,16), keys 32..63 in columns [32,48) of the dP buffer; 16 keys = 8 columns.
          // B = K_j: MN-major ([key][64 dims] rows of 128 B), 16 keys = 2 groups of 8 rows = 2048 B
          const uint32_t a_tmem = tmem_base + DQ_TM_DP + (uint32_t)((k >> 1) * 32 + (k & 1) * 8);
          const uint64_t bdesc = make_smem_desc_sw128(kbase + (uint32_t)(k * 2048), AB_T64, 1024);
          umma_bf16_ts(tmem_base + DQ_TM_DQ, a_tmem, bdesc, idesc_kmn, (j | k) ? 1u : 0u);
        }
        umma_commit(dq_done + (j & 1));
        umma_commit(kv_empty + s);
        }
        if (j + 1 < n_blocks) issue_dp(j + 1);  // dP(j) has been read (ds_full(j))
      }
    }
  } else {
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int ch = (warp - 2) >> 2;       // which 32 of the 64 key columns this thread handles
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const size_t stat = ((size_t)b * H + h) * Np + q0 + row;  // q0 + row < Np always
    const float lse = lse2[stat];                             // +inf on pad rows -> P = 0
    const float dsc = dsum[stat] * AB_SCALE;
    const uint64_t sl2_2 = pack_f32x2(AB_SL2, AB_SL2), nlse_2 = pack_f32x2(-lse, -lse);
    const uint64_t scale_2 = pack_f32x2(AB_SCALE, AB_SCALE), ndsc_2 = pack_f32x2(-dsc, -dsc);
    for (int j = 0; j < n_blocks; j++) {
      const int buf = j & 1;
      const int kv_valid = N - j * 64;
      // ---- phase 1: P from S (runs while the tensor core still produces dQ(j-1) and dP(j)) ----
      mbar_wait(s_full + buf, (uint32_t)(j >> 1) & 1);
      tc_fence_after();
      float p[32];
#pragma unroll
      for (int sub = 0; sub < 2; sub++) {
        uint32_t rs[16];
        tmem_ld_32x16(t_lane + DQ_TM_S + (uint32_t)(buf * 64 + ch * 32 + sub * 16), rs);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 2) {  // packed fp32 pairs (FFMA2): x = s * c - lse
          float x0, x1;
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(rs[i]), __uint_as_float(rs[i + 1])), sl2_2, nlse_2), x0, x1);
          // POLY of the 8 pairs of each 16-column load: packed polynomial exp2 on the FMA pipe (ex2_poly3_x2)
          if ((i >> 1) < POLY) ex2_poly3_x2(x0, x1, p[sub * 16 + i], p[sub * 16 + i + 1]);
          else { p[sub * 16 + i] = ex2_approx(x0); p[sub * 16 + i + 1] = ex2_approx(x1); }
        }
      }
      if (kv_valid < 64) {  // warp-uniform, last key block only: zero-filled tail keys contribute nothing
#pragma unroll
        for (int i = 0; i < 32; i++)
          if (ch * 32 + i >= kv_valid) p[i] = 0.f;
      }
      // ---- phase 2: dS = P * (dP - Dsum) / 8 ----
      mbar_wait(dp_full, (uint32_t)j & 1);
      tc_fence_after();
      // (dP_j is complete => every earlier MMA, including dQ(j-1) which read dS_{j-1} from these columns, is too)
      uint32_t dsw[16];  // this thread's 32 dS values, two bf16 per word
#pragma unroll
      for (int sub = 0; sub < 2; sub++) {
        uint32_t rp[16];
        tmem_ld_32x16(t_lane + DQ_TM_DP + (uint32_t)(ch * 32 + sub * 16), rp);
        tmem_ld_wait();
        float ds[16];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {  // dS = P * (dP / 8 - Dsum / 8): FFMA2 + FMUL2
          const uint64_t t = fma_f32x2(pack_f32x2(__uint_as_float(rp[i]), __uint_as_float(rp[i + 1])), scale_2, ndsc_2);
          unpack_f32x2(mul_f32x2(pack_f32x2(p[sub * 16 + i], p[sub * 16 + i + 1]), t), ds[i], ds[i + 1]);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) dsw[sub * 8 + i] = ab_pack2(ds[2 * i], ds[2 * i + 1]);
      }
      tmem_st_32x16(t_lane + DQ_TM_DP + (uint32_t)(ch * 32), dsw);  // over the head of this thread's own dP columns
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ds_full + buf);
    }
    const int last = n_blocks - 1;
    mbar_wait(dq_done + (last & 1), (uint32_t)(last >> 1) & 1);
    tc_fence_after();
    uint32_t r0[32];
    tmem_ld_32x32(t_lane + DQ_TM_DQ + (uint32_t)(ch * 32), r0);
    tmem_ld_wait();
    if (q0 + row < N) {
      __nv_bfloat16* dst = dqkv + ((size_t)b * N + q0 + row) * 3 * D + h * AB_HD + ch * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 pk;
        pk.x = ab_pack2(__uint_as_float(r0[i]), __uint_as_float(r0[i + 1]));
        pk.y = ab_pack2(__uint_as_float(r0[i + 2]), __uint_as_float(r0[i + 3]));
        pk.z = ab_pack2(__uint_as_float(r0[i + 4]), __uint_as_float(r0[i + 5]));
        pk.w = ab_pack2(__uint_as_float(r0[i + 6]), __uint_as_float(r0[i + 7]));
        *reinterpret_cast<uint4*>(dst + i) = pk;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, DQ_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dK, dV
// ---------------------------------------------------------------------------------------------------------------
constexpr int DKV_STAGES = 3;
constexpr int DKV_SMEM = 2 * AB_T128 /*K, V*/ + 2 * DKV_STAGES * AB_T64 /*Q, dO*/ +
                         2 * 128 * 4 /*lse2 | Dsum of a query block, x2*/ + 1024 + 256;
constexpr uint32_t DKV_TM_ST = 0, DKV_TM_DPT = 64, DKV_TM_DV = 128, DKV_TM_DK = 192, DKV_TMEM_COLS = 256;

template <bool UNI, int POLY>
__global__ void __launch_bounds__(AB_THREADS, 2)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tm_kv128, const __grid_constant__ CUtensorMap tm_q64,
                    const __grid_constant__ CUtensorMap tm_do64, const float* __restrict__ lse2,
                    const float* __restrict__ dsum, __nv_bfloat16* __restrict__ dqkv, int N, int Np, int H) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sK = smem;
  uint8_t* sV = sK + AB_T128;
  uint8_t* sQ = sV + AB_T128;
  uint8_t* sdO = sQ + DKV_STAGES * AB_T64;
  float* s_stat = reinterpret_cast<float*>(sdO + DKV_STAGES * AB_T64);  // [2][128]: lse2[64] | Dsum/8 [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + 2 * 128);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;              // [DKV_STAGES]
  uint64_t* q_empty = q_full + DKV_STAGES;   // [DKV_STAGES]
  uint64_t* st_full = q_empty + DKV_STAGES;  // S^T of block i in TMEM
  uint64_t* dpt_full = st_full + 1;         // dP^T of block i in TMEM
  uint64_t* pt_full = dpt_full + 1;         // P^T, dS^T written (and S^T/dP^T read out of TMEM)
  uint64_t* acc_done = pt_full + 1;         // dV/dK MMAs of block i retired (P^T/dS^T buffers free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int n_blocks = (N + 63) / 64;  // query blocks
  const int D = H * AB_HD;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_kv128);
    prefetch_tmap(&tm_q64);
    prefetch_tmap(&tm_do64);
    mbar_init(kv_full, 1);
    for (int s = 0; s < DKV_STAGES; s++) { mbar_init(q_full + s, 1); mbar_init(q_empty + s, 1); }
    mbar_init(st_full, 1);
    mbar_init(dpt_full, 1);
    mbar_init(pt_full, AB_SOFT);
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, DKV_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();  // PDL (sm100_ptx.cuh)
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * AB_T128);
      tma_load_3d(sK, &tm_kv128, kv_full, D + h * AB_HD, k0, b);
      tma_load_3d(sV, &tm_kv128, kv_full, 2 * D + h * AB_HD, k0, b);
      for (int i = 0; i < n_blocks; i++) {
        const int s = i % DKV_STAGES;
        mbar_wait(q_empty + s, ((uint32_t)(i / DKV_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(q_full + s, 2 * AB_T64);
        tma_load_3d(sQ + s * AB_T64, &tm_q64, q_full + s, h * AB_HD, i * 64, b);
        tma_load_3d(sdO + s * AB_T64, &tm_do64, q_full + s, h * AB_HD, i * 64, b);
      }
    }
  } else if (warp == 1) {
    if (UNI || lane == 0) {
      constexpr uint32_t idesc_kk = make_idesc_bf16(128, 64, false, false);
      constexpr uint32_t idesc_kmn = make_idesc_bf16(128, 64, false, true);
      const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
      // Issue order: the softmax of block i+1 only needs S^T to start (exp), so S^T(i+1) goes in FRONT of the
      // dV/dK accumulation of block i, and dP^T(i+1) behind it:
      //   St(0) dPt(0) | wait PT(0) : St(1) dV(0) dK(0) dPt(1) | wait PT(1) : St(2) dV(1) dK(1) dPt(2) | ...
      auto issue_st = [&](int i) {
        const int s = i % DKV_STAGES;
        mbar_wait(q_full + s, (uint32_t)(i / DKV_STAGES) & 1);
        tc_fence_after();
        const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ + s * AB_T64), 16, 1024);
        if (!UNI || elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; k++)  // S^T = K Q_i^T : [128 keys x 64 queries]
            umma_bf16(tmem_base + DKV_TM_ST, kdesc + (uint64_t)(2 * k), qdesc + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u);
          umma_commit(st_full);
        }
      };
      auto issue_dpt = [&](int i) {
        const uint64_t dodesc = make_smem_desc_sw128(smem_u32(sdO + (i % DKV_STAGES) * AB_T64), 16, 1024);
        if (!UNI || elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; k++)  // dP^T = V dO_i^T
            umma_bf16(tmem_base + DKV_TM_DPT, vdesc + (uint64_t)(2 * k), dodesc + (uint64_t)(2 * k), idesc_kk, k ? 1u : 0u);
          umma_commit(dpt_full);
        }
      };
      mbar_wait(kv_full, 0);
      issue_st(0);
      issue_dpt(0);
      for (int i = 0; i < n_blocks; i++) {
        const int s = i % DKV_STAGES;
        mbar_wait(pt_full, (uint32_t)i & 1);  // softmax(i) has read S^T(i), dP^T(i) and written P^T, dS^T
        tc_fence_after();
        if (i + 1 < n_blocks) issue_st(i + 1);
        const uint32_t qbase = smem_u32(sQ + s * AB_T64), dobase = smem_u32(sdO + s * AB_T64);
        if (!UNI || elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; k++) {  // reduction over the 64 queries of the block, 16 per instruction
          // P^T and dS^T come from TENSOR MEMORY: packed bf16 pairs written by the softmax threads over their own
          // dP^T columns -- for queries 0..31: dS^T in columns [0,16), P^T in [16,32); for queries 32..63: [32,48), [48,64)
          const uint32_t ds_tmem = tmem_base + DKV_TM_DPT + (uint32_t)((k >> 1) * 32 + (k & 1) * 8);
          const uint32_t p_tmem = ds_tmem + 16u;
          const uint64_t dob = make_smem_desc_sw128(dobase + (uint32_t)(k * 2048), AB_T64, 1024);  // MN-major
          const uint64_t qb = make_smem_desc_sw128(qbase + (uint32_t)(k * 2048), AB_T64, 1024);    // MN-major
          umma_bf16_ts(tmem_base + DKV_TM_DV, p_tmem, dob, idesc_kmn, (i | k) ? 1u : 0u);   // dV += P^T dO_i
          umma_bf16_ts(tmem_base + DKV_TM_DK, ds_tmem, qb, idesc_kmn, (i | k) ? 1u : 0u);  // dK += dS^T Q_i
        }
        umma_commit(acc_done);
        umma_commit(q_empty + s);
        }
        if (i + 1 < n_blocks) issue_dpt(i + 1);
      }
    }
  } else {
    const int quad = warp & 3;
    const int ch = (warp - 2) >> 2;       // which 32 of the 64 query columns this thread handles
    const int row = quad * 32 + lane;     // key row of this thread
    const bool key_valid = k0 + row < N;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const size_t stat_base = ((size_t)b * H + h) * Np;
    // per query block, softmax threads 0..127 stage -lse2[64] | -Dsum[64]/8 in smem (prefetched one block ahead)
    const int st_t = (warp - 2) * 32 + lane;  // 0..255
    const float* stat_src = (st_t < 64 ? lse2 : dsum) + stat_base + (st_t & 63);
    const float stat_mul = st_t < 64 ? -1.0f : -AB_SCALE;  // staged NEGATED: they are the addends of the packed FMAs
    const uint64_t sl2_2 = pack_f32x2(AB_SL2, AB_SL2), scale_2 = pack_f32x2(AB_SCALE, AB_SCALE);
    float pre = st_t < 128 ? stat_src[0] * stat_mul : 0.f;  // block 0 (Np >= 128: in bounds)
    for (int i = 0; i < n_blocks; i++) {
      float* st = s_stat + (i & 1) * 128;
      if (st_t < 128) st[st_t] = pre;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (st_t < 128 && i + 1 < n_blocks) pre = stat_src[(i + 1) * 64] * stat_mul;  // (i+1)*64 + 63 < Np
      // ---- phase 1: P^T from S^T (overlaps the dV/dK MMAs of block i-1) ----
      mbar_wait(st_full, (uint32_t)i & 1);
      tc_fence_after();
      float p[32];
#pragma unroll
      for (int sub = 0; sub < 2; sub++) {
        const int c0 = ch * 32 + sub * 16;
        uint32_t rs[16];
        tmem_ld_32x16(t_lane + DKV_TM_ST + (uint32_t)c0, rs);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          // smem broadcast of -lse2 for 4 queries (pad queries: -inf -> P = 0), consumed as two packed fp32 pairs
          const ulonglong2 l4 = *reinterpret_cast<const ulonglong2*>(st + c0 + c);
          float x0, x1, x2, x3;
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(rs[c]), __uint_as_float(rs[c + 1])), sl2_2, l4.x), x0, x1);
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(rs[c + 2]), __uint_as_float(rs[c + 3])), sl2_2, l4.y), x2, x3);
          if ((c >> 1) < POLY) ex2_poly3_x2(x0, x1, p[sub * 16 + c], p[sub * 16 + c + 1]);
          else { p[sub * 16 + c] = ex2_approx(x0); p[sub * 16 + c + 1] = ex2_approx(x1); }
          if ((c >> 1) + 1 < POLY) ex2_poly3_x2(x2, x3, p[sub * 16 + c + 2], p[sub * 16 + c + 3]);
          else { p[sub * 16 + c + 2] = ex2_approx(x2); p[sub * 16 + c + 3] = ex2_approx(x3); }
        }
      }
      if (!key_valid) {  // zero-filled tail key rows (last key block only)
#pragma unroll
        for (int c = 0; c < 32; c++) p[c] = 0.f;
      }
      uint32_t pw[16];  // P^T of this thread's 32 queries, two bf16 per word
#pragma unroll
      for (int e = 0; e < 16; e++) pw[e] = ab_pack2(p[2 * e], p[2 * e + 1]);
      // ---- phase 2: dS^T = P^T * (dP^T - Dsum) / 8 ----
      mbar_wait(dpt_full, (uint32_t)i & 1);
      tc_fence_after();
      uint32_t dsw[16];  // this thread's 32 dS^T values, two bf16 per word
#pragma unroll
      for (int sub = 0; sub < 2; sub++) {
        const int c0 = ch * 32 + sub * 16;
        uint32_t rp[16];
        tmem_ld_32x16(t_lane + DKV_TM_DPT + (uint32_t)c0, rp);
        tmem_ld_wait();
        float ds[16];
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          const ulonglong2 d4 = *reinterpret_cast<const ulonglong2*>(st + 64 + c0 + c);  // -Dsum / 8 of 4 queries
          const uint64_t t0 = fma_f32x2(pack_f32x2(__uint_as_float(rp[c]), __uint_as_float(rp[c + 1])), scale_2, d4.x);
          const uint64_t t1 = fma_f32x2(pack_f32x2(__uint_as_float(rp[c + 2]), __uint_as_float(rp[c + 3])), scale_2, d4.y);
          unpack_f32x2(mul_f32x2(pack_f32x2(p[sub * 16 + c], p[sub * 16 + c + 1]), t0), ds[c], ds[c + 1]);
          unpack_f32x2(mul_f32x2(pack_f32x2(p[sub * 16 + c + 2], p[sub * 16 + c + 3]), t1), ds[c + 2], ds[c + 3]);
        }
#pragma unroll
        for (int e = 0; e < 8; e++) dsw[sub * 8 + e] = ab_pack2(ds[2 * e], ds[2 * e + 1]);
      }
      // dP^T(i) complete => dV/dK(i-1), which read the previous P^T / dS^T from these columns, are complete too
      tmem_st_32x16(t_lane + DKV_TM_DPT + (uint32_t)(ch * 32), dsw);
      tmem_st_32x16(t_lane + DKV_TM_DPT + (uint32_t)(ch * 32 + 16), pw);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(pt_full);
    }
    mbar_wait(acc_done, (uint32_t)(n_blocks - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int which = 0; which < 2; which++) {  // 0: dK (qkv section 1), 1: dV (section 2)
      uint32_t r0[32];
      const uint32_t col = (which ? DKV_TM_DV : DKV_TM_DK) + (uint32_t)(ch * 32);
      tmem_ld_32x32(t_lane + col, r0);
      tmem_ld_wait();
      if (key_valid) {
        __nv_bfloat16* dst = dqkv + ((size_t)b * N + k0 + row) * 3 * D + (size_t)(1 + which) * D + h * AB_HD + ch * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk;
          pk.x = ab_pack2(__uint_as_float(r0[i]), __uint_as_float(r0[i + 1]));
          pk.y = ab_pack2(__uint_as_float(r0[i + 2]), __uint_as_float(r0[i + 3]));
          pk.z = ab_pack2(__uint_as_float(r0[i + 4]), __uint_as_float(r0[i + 5]));
          pk.w = ab_pack2(__uint_as_float(r0[i + 6]), __uint_as_float(r0[i + 7]));
          *reinterpret_cast<uint4*>(dst + i) = pk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, DKV_TMEM_COLS);
  }
}

}  // namespace

int attention_bwd(const void* qkv, const void* out, const void* dout, float* lse2, float* dsum, void* dqkv, int B, int N,
                  int H, cudaStream_t st) {
  DGS_REQUIRE(B > 0 && N > 0 && H > 0 && H <= 64, "attention_bwd: bad shape B=%d N=%d H=%d", B, N, H);
  const int D = H * AB_HD, Np = attention_lse_stride(N);
  CUtensorMap tm_qkv128, tm_qkv64, tm_do128, tm_do64;
  {
    uint64_t dims[3] = {(uint64_t)(3 * D), (uint64_t)N, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)(3 * D) * 2, (uint64_t)N * 3 * D * 2};
    uint32_t b128[3] = {AB_HD, 128, 1}, b64[3] = {AB_HD, 64, 1};
    int rc = make_tmap_bf16(&tm_qkv128, qkv, 3, dims, str, b128);
    if (rc) return rc;
    rc = make_tmap_bf16(&tm_qkv64, qkv, 3, dims, str, b64);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)D, (uint64_t)N, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)D * 2, (uint64_t)N * D * 2};
    uint32_t b128[3] = {AB_HD, 128, 1}, b64[3] = {AB_HD, 64, 1};
    int rc = make_tmap_bf16(&tm_do128, dout, 3, dims, str, b128);
    if (rc) return rc;
    rc = make_tmap_bf16(&tm_do64, dout, 3, dims, str, b64);
    if (rc) return rc;
  }
  using KernT = decltype(&attn_bwd_dq_kernel<true, 0>);
  static KernT dq_tab[2][3] = {{attn_bwd_dq_kernel<false, 0>, attn_bwd_dq_kernel<false, 2>, attn_bwd_dq_kernel<false, 4>},
                               {attn_bwd_dq_kernel<true, 0>, attn_bwd_dq_kernel<true, 2>, attn_bwd_dq_kernel<true, 4>}};
  static KernT dkv_tab[2][3] = {{attn_bwd_dkv_kernel<false, 0>, attn_bwd_dkv_kernel<false, 2>, attn_bwd_dkv_kernel<false, 4>},
                                {attn_bwd_dkv_kernel<true, 0>, attn_bwd_dkv_kernel<true, 2>, attn_bwd_dkv_kernel<true, 4>}};
  static bool configured = false;
  static int uni = 0, poly = 0;
  if (!configured) {
    for (int u = 0; u < 2; u++)
      for (int q = 0; q < 3; q++) {
        DGS_CUDA_OK(cudaFuncSetAttribute(dq_tab[u][q], cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
        DGS_CUDA_OK(cudaFuncSetAttribute(dkv_tab[u][q], cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM));
      }
    const char* eu = getenv("DGS_ATT_UNI");
    uni = (eu && eu[0] == '0') ? 0 : 1;  // default since round 2 (measured: r2 first GPU call)
    const char* ep = getenv("DGS_ATTB_POLY");  // 0 / 1 / 2 = none / 2 of 8 / 4 of 8 pairs on the packed polynomial exp2
    poly = ep ? atoi(ep) : ATTB_POLY_DEFAULT;
    poly = poly < 0 ? 0 : poly > 2 ? 2 : poly;
    configured = true;
  }
  DGS_CUDA_OK(launch_pdl(attn_bwd_prep_kernel, dim3(Np, B), dim3(H * 16 < 32 ? 32 : H * 16), 0, st,
                         (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, lse2, dsum, N, Np, H));
  DGS_POST_LAUNCH();
  dim3 grid(ceil_div(N, 128), H, B);
  DGS_CUDA_OK(launch_pdl(dq_tab[uni][poly], grid, dim3(AB_THREADS), DQ_SMEM, st, tm_qkv128,
                         tm_qkv64, tm_do128, (const float*)lse2, (const float*)dsum, (__nv_bfloat16*)dqkv, N, Np, H));
  DGS_POST_LAUNCH();
  DGS_CUDA_OK(launch_pdl(dkv_tab[uni][poly], grid, dim3(AB_THREADS), DKV_SMEM, st,
                         tm_qkv128, tm_qkv64, tm_do64, (const float*)lse2, (const float*)dsum, (__nv_bfloat16*)dqkv, N, Np, H));
  DGS_POST_LAUNCH();
  return DGS_OK;
}

}  // namespace dgs
