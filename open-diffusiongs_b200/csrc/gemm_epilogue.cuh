// gemm_epilogue.cuh -- the fused epilogue shared by the single-CTA and the CTA-pair GEMM kernels.
//
// 128 epilogue threads (warps 2..5) drain one 128 x BN fp32 accumulator tile from TMEM, one output row per thread,
// 32 columns at a time.  ncu (profiles/r1_ncu_gemm2cta_v1.txt) showed the first version of this loop exposed a
// dependent global-load latency per 32-column chunk (bias via __ldg, plus the fp32 residual read of the gate+residual
// epilogue), which made the epilogue LONGER than the tile's MMA main loop and starved the tensor pipe.  Now:
//   * bias / gate rows of the tile are staged once per tile in shared memory (double-buffered by accumulator index)
//     by the epilogue threads themselves, BEFORE they wait for the accumulator -- i.e. under the main loop;
//   * the residual stream values of chunk c+1 are prefetched into registers while chunk c is processed.
#pragma once
#include "dit_kernels.h"
#include "sm100_ptx.cuh"

namespace dgs {

__device__ __forceinline__ float epi_gelu_tanh(float x) {  // nn.GELU(approximate="tanh"), tanh on the MUFU pipe
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * (x + k1 * x * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.0f + t);
}
__device__ __forceinline__ float epi_dgelu_tanh(float x) {  // d/dx of the above
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float x2 = x * x;
  const float u = k0 * (x + k1 * x * x2);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float du = k0 * (1.0f + 3.0f * k1 * x2);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}
__device__ __forceinline__ uint32_t epi_pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <int NT>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// Stage bias[n0 .. n0+BN) and (gate epilogue) gate[sample(row0)][n0 .. n0+BN) into shared memory.
// `s_vec` = this accumulator buffer's [2][BN] floats.  All 128 epilogue threads call it.
// NOTE gate: a 128-row tile may straddle two samples; rows then read their own sample's gate directly from global
// memory (rare: one tile per sample boundary), signalled by *uniform_gate == false.
// NT = number of epilogue threads sharing s_vec (128, or 256 when two 128-row halves are drained side by side: then
// ROWS = 256 and the sample-uniformity check spans both halves).
template <int EPI, int BN, int NT = 128, int ROWS = 128>
__device__ __forceinline__ void epilogue_stage_vectors(const GemmEpilogue& ep, float* s_vec, int et /*0..NT-1*/, int m0,
                                                       int n0, int M, int N, bool* uniform_gate) {
  const int last_row = min(m0 + ROWS - 1, M - 1);
  const bool uni = (EPI != EPI_GATE_RESID_F32) || (m0 / ep.rows_per_sample == last_row / ep.rows_per_sample);
  *uniform_gate = uni;
  for (int i = et; i < BN; i += NT) {
    const int n = n0 + i;
    s_vec[i] = (ep.bias && n < N) ? __ldg(ep.bias + n) : 0.f;
    if (EPI == EPI_GATE_RESID_F32)
      s_vec[BN + i] = (uni && n < N) ? __ldg(ep.gate + (size_t)(m0 / ep.rows_per_sample) * ep.gate_stride + n) : 0.f;
  }
  epi_bar_sync<NT>();
}

// Drain one accumulator row (this thread's) of BN columns.
template <int EPI, int BN>
__device__ __forceinline__ void epilogue_drain_row(const GemmEpilogue& ep, const float* s_vec, bool uniform_gate,
                                                   uint32_t t_row, int row, int n0, int M, int N) {
  using namespace ptx;
  const bool valid = row < M;
  const float* gate_row = nullptr;
  if (EPI == EPI_GATE_RESID_F32 && valid && !uniform_gate)
    gate_row = ep.gate + (size_t)(row / ep.rows_per_sample) * ep.gate_stride;
  float4 xn[8];  // residual prefetch (gate epilogue only)
  float* xrow = (EPI == EPI_GATE_RESID_F32 && valid) ? reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldc : nullptr;
  const float* rrow = (EPI == EPI_GATE_RESID_F32 && valid) ? (ep.resid ? ep.resid + (size_t)row * ep.ldc : xrow) : nullptr;
  __nv_bfloat16* arow = (ep.aux && valid) ? reinterpret_cast<__nv_bfloat16*>(ep.aux) + (size_t)row * ep.ldc : nullptr;
  if (EPI == EPI_GATE_RESID_F32 && valid && n0 < N) {
#pragma unroll
    for (int j = 0; j < 8; j++) xn[j] = *reinterpret_cast<const float4*>(rrow + n0 + 4 * j);
  }
#pragma unroll 1
  for (int c = 0; c < BN / 32; c++) {
    const int n = n0 + c * 32;
    if (n >= N) break;  // warp-uniform (N % 32 == 0)
    uint32_t r[32];
    tmem_ld_32x32(t_row + (uint32_t)(c * 32), r);
    float4 xc[8];
    if (EPI == EPI_GATE_RESID_F32 && valid) {
#pragma unroll
      for (int j = 0; j < 8; j++) xc[j] = xn[j];
      if (n + 32 < N && c + 1 < BN / 32) {
#pragma unroll
        for (int j = 0; j < 8; j++) xn[j] = *reinterpret_cast<const float4*>(rrow + n + 32 + 4 * j);
      }
    }
    uint4 ua[4];  // EPI_DGELU_BF16: the saved pre-activation of this chunk
    if (EPI == EPI_DGELU_BF16 && valid) {
#pragma unroll
      for (int j = 0; j < 4; j++) ua[j] = *reinterpret_cast<const uint4*>(arow + n + 8 * j);
    }
    tmem_ld_wait();
    if (!valid) continue;
    float v[32];
    const float* sb = s_vec + c * 32;
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]) + sb[j];
    if ((EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_GATE_RESID_F32) && arow) {  // training: keep acc + b (bf16)
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 pk;
        pk.x = epi_pack_bf16(v[j], v[j + 1]); pk.y = epi_pack_bf16(v[j + 2], v[j + 3]);
        pk.z = epi_pack_bf16(v[j + 4], v[j + 5]); pk.w = epi_pack_bf16(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(arow + n + j) = pk;
      }
    }
    if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_DGELU_BF16) {
      if (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = epi_gelu_tanh(v[j]);
      }
      if (EPI == EPI_DGELU_BF16) {
        const __nv_bfloat16* ub = reinterpret_cast<const __nv_bfloat16*>(ua);
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] *= epi_dgelu_tanh(__bfloat162float(ub[j]));
      }
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ldc + n;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 pk;
        pk.x = epi_pack_bf16(v[j], v[j + 1]); pk.y = epi_pack_bf16(v[j + 2], v[j + 3]);
        pk.z = epi_pack_bf16(v[j + 4], v[j + 5]); pk.w = epi_pack_bf16(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(o + j) = pk;
      }
    } else if (EPI == EPI_GATE_RESID_F32) {
      const float* sg = s_vec + BN + c * 32;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float4 g4;
        if (uniform_gate) g4 = make_float4(sg[4 * j], sg[4 * j + 1], sg[4 * j + 2], sg[4 * j + 3]);
        else g4 = __ldg(reinterpret_cast<const float4*>(gate_row + n + 4 * j));
        float4 x4 = xc[j];
        x4.x += g4.x * v[4 * j]; x4.y += g4.y * v[4 * j + 1]; x4.z += g4.z * v[4 * j + 2]; x4.w += g4.w * v[4 * j + 3];
        *reinterpret_cast<float4*>(xrow + n + 4 * j) = x4;
      }
    } else {  // EPI_F32
      float* o = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldc + n;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
}

// ---- TMA epilogue (CTA-pair kernel, inference-mode epilogues) -------------------------------------------------------
// The row-per-thread global stores above touch 32 different 128-byte lines per warp instruction (one per row), so the
// drain of a 128 x 256 tile is bound by LSU wavefronts (~4-7k clk, profiles/r1_gemm_probe_v2.txt) -- fully exposed on
// the last tile of every CTA and on the single-wave GEMMs (proj, fc2).  Here each epilogue warp stages its 32-row slab
// in shared memory in the 128-byte-swizzled layout of a tensor map (thread = row, conflict-free 16-byte stores) and ONE
// lane hands a [32 rows x 128 bytes] box to the TMA:
//   bf16 outputs:  cp.async.bulk.tensor store, 64 columns per box;
//   gate+residual: x += gate * (acc + b) as cp.reduce.async.bulk.tensor ... add.f32, 32 columns per box -- the fp32
//                  residual stream is updated in L2 by the TMA and never read by the SM.
// Rows >= M are clipped by the tensor map.  Two staging buffers per warp; a buffer is reused once the bulk group that
// read it has completed its shared-memory reads (cp.async.bulk.wait_group.read).
__device__ __forceinline__ void epi_tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void epi_tma_reduce_add_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void epi_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void epi_bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void epi_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr int EPI_TMA_STAGING_BYTES = 4 * 2 * 4096;  // 4 epilogue warps x 2 buffers x (32 rows x 128 bytes)

// `stg` = this warp's staging buffers (1024-byte aligned): two 4 KB buffers for the output, and with AUX two more (at
// +8 KB) for the bf16 pre-activation the training-mode fc1 epilogue also stores (tm_aux).  slab_row0 = first output row
// of the warp's 32-row slab.  `reduce`: EPI_F32 only -- add into the output instead of storing (split-K partial sums).
template <int EPI, int BN, bool AUX = false>
__device__ __forceinline__ void epilogue_drain_row_tma(const GemmEpilogue& ep, const float* s_vec, bool uniform_gate,
                                                       uint32_t t_row, int row, int slab_row0, int n0, int M, int N,
                                                       uint8_t* stg, int lane, const CUtensorMap* tm_out,
                                                       const CUtensorMap* tm_aux = nullptr, bool reduce = false) {
  using namespace ptx;
  static_assert(EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_GATE_RESID_F32 || EPI == EPI_F32,
                "TMA epilogue: unsupported");
  static_assert(!AUX || EPI == EPI_BIAS_GELU_BF16, "TMA epilogue: the aux store exists for the fc1 epilogue only");
  const int crow = row < M ? row : M - 1;  // rows past the end compute on a valid row's gate; their box rows are clipped
  const float* gate_row = nullptr;
  if (EPI == EPI_GATE_RESID_F32 && !uniform_gate) gate_row = ep.gate + (size_t)(crow / ep.rows_per_sample) * ep.gate_stride;
  const uint32_t sw = (uint32_t)(lane & 7);
  if (EPI == EPI_GATE_RESID_F32 || EPI == EPI_F32) {
#pragma unroll 1
    for (int c = 0; c < BN / 32; c++) {
      const int n = n0 + c * 32;
      if (n >= N) break;
      uint32_t r[32];
      tmem_ld_32x32(t_row + (uint32_t)(c * 32), r);
      uint8_t* buf = stg + (c & 1) * 4096;
      if (lane == 0) epi_bulk_wait_read1();
      __syncwarp();
      tmem_ld_wait();
      const float* sb = s_vec + c * 32;
      const float* sg = s_vec + BN + c * 32;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float4 v4;
        v4.x = __uint_as_float(r[4 * j]) + sb[4 * j];
        v4.y = __uint_as_float(r[4 * j + 1]) + sb[4 * j + 1];
        v4.z = __uint_as_float(r[4 * j + 2]) + sb[4 * j + 2];
        v4.w = __uint_as_float(r[4 * j + 3]) + sb[4 * j + 3];
        if (EPI == EPI_GATE_RESID_F32) {
          float4 g4;
          if (uniform_gate) g4 = make_float4(sg[4 * j], sg[4 * j + 1], sg[4 * j + 2], sg[4 * j + 3]);
          else g4 = __ldg(reinterpret_cast<const float4*>(gate_row + n + 4 * j));
          v4.x *= g4.x; v4.y *= g4.y; v4.z *= g4.z; v4.w *= g4.w;
        }
        *reinterpret_cast<float4*>(buf + lane * 128 + (((uint32_t)j ^ sw) * 16)) = v4;  // SWIZZLE_128B
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (EPI == EPI_GATE_RESID_F32 || reduce) epi_tma_reduce_add_2d(tm_out, buf, n, slab_row0);
        else epi_tma_store_2d(tm_out, buf, n, slab_row0);
        epi_bulk_commit();
      }
    }
  } else {
#pragma unroll 1
    for (int c = 0; c < BN / 64; c++) {
      const int n = n0 + c * 64;
      if (n >= N) break;
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(t_row + (uint32_t)(c * 64), r0);
      tmem_ld_32x32(t_row + (uint32_t)(c * 64 + 32), r1);
      uint8_t* buf = stg + (c & 1) * 4096;
      uint8_t* abuf = stg + 8192 + (c & 1) * 4096;
      if (lane == 0) epi_bulk_wait_read1();
      __syncwarp();
      tmem_ld_wait();
      const float* sb = s_vec + c * 64;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t* r = (j < 4) ? (r0 + 8 * j) : (r1 + 8 * (j - 4));
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]) + sb[8 * j + i];
        if (AUX) {  // training: keep the pre-activation acc + b (bf16)
          uint4 pa;
          pa.x = epi_pack_bf16(v[0], v[1]); pa.y = epi_pack_bf16(v[2], v[3]);
          pa.z = epi_pack_bf16(v[4], v[5]); pa.w = epi_pack_bf16(v[6], v[7]);
          *reinterpret_cast<uint4*>(abuf + lane * 128 + (((uint32_t)j ^ sw) * 16)) = pa;
        }
        if (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
          for (int i = 0; i < 8; i++) v[i] = epi_gelu_tanh(v[i]);
        }
        uint4 pk;
        pk.x = epi_pack_bf16(v[0], v[1]); pk.y = epi_pack_bf16(v[2], v[3]);
        pk.z = epi_pack_bf16(v[4], v[5]); pk.w = epi_pack_bf16(v[6], v[7]);
        *reinterpret_cast<uint4*>(buf + lane * 128 + (((uint32_t)j ^ sw) * 16)) = pk;  // SWIZZLE_128B
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        epi_tma_store_2d(tm_out, buf, n, slab_row0);
        if (AUX) epi_tma_store_2d(tm_aux, abuf, n, slab_row0);
        epi_bulk_commit();  // one group per step: wait_group.read 1 frees both buffers of the step before last
      }
    }
  }
}

}  // namespace dgs
