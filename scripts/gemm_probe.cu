// gemm_probe.cu -- stand-alone diagnosis of the CTA-pair (cta_group::2) tcgen05 GEMM (NOT part of the product).
//
// Question: why does the 256 x 256 CTA-pair kernel (csrc/gemm2_sm100.cu) run 1.6x slower than the
// single-CTA 128 x 256 kernel although it moves 2/3 of the operand bytes?  This binary runs the same pipeline protocol
// with switches that remove one suspect at a time, and lets the producer / MMA / epilogue threads accumulate the cycles
// they spend waiting on each barrier.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I open-diffusiongs_b200/csrc -I include \
//        scripts/gemm_probe.cu -o scripts/bin/gemm_probe -lcuda
//   scripts/bin/gemm_probe            (on the GPU box)
//
// Template switches:  WAIT 0 = product's mbar_wait (try_wait, then try_wait with a 20 us suspend hint)
//                          1 = plain try_wait loop (no hint)      2 = test_wait spin (never suspends)
//                     EPI  0 = drain TMEM -> bf16 -> global       1 = no drain (arrive only)
//                     LOAD 0 = TMA                                1 = no TMA at all (MMA on whatever is in smem)
//                     STAGES, and CL4 = cluster of 4 (two pairs) with the B half multicast across the pairs.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sm100_ptx.cuh"

using namespace dgs::ptx;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
  } while (0)

namespace dgs {
static CUtensorMapDataType g_tmap_dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; i++) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; i++) gstr[i] = strides_bytes[i];
  CUresult r = cuTensorMapEncodeTiled(out, g_tmap_dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim,
                                      gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}
}  // namespace dgs

constexpr int BM_CTA = 128, BN = 256, BN_CTA = 128, BK = 64, UMMA_K = 16, THREADS = 192;
constexpr int A_BYTES = BM_CTA * BK * 2, B_BYTES = BN_CTA * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// multicast variant: the box lands at the same smem offset in every CTA of `mask`; each destination CTA's barrier AT THE SAME
// OFFSET as `bar_cluster_addr`'s receives the complete_tx -- with cta_group::2 the signal goes to the barrier of the
// addressed CTA-pair member (here: the leader of each destination pair).
__device__ __forceinline__ void tma_load_2d_2sm_mc(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                                   uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst, uint32_t n) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(n) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t t, uint32_t n) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(n) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

template <int WAIT>
__device__ __forceinline__ void pwait(uint64_t* bar, uint32_t parity) {
  if (WAIT == 0) { mbar_wait(bar, parity); return; }
  unsigned long long n = 0;
  if (WAIT == 1) {
    while (!mbar_try_wait(bar, parity)) { if (++n > 20000000ull) { printf("probe: try_wait timeout\n"); __trap(); } }
  } else {
    while (!mbar_test_wait(bar, parity)) { if (++n > 200000000ull) { printf("probe: test_wait timeout\n"); __trap(); } }
  }
}

// dbg layout per CTA (8 x u64): 0 producer empty-wait, 1 mma full-wait, 2 mma tempty-wait, 3 epi tfull-wait, 4 epi drain,
//                               5 total cycles of the mma thread, 6 tiles, 7 total cycles of the producer
template <int WAIT, int EPI, int LOAD, int STAGES, int CL4, int NOARR, int TE1>
__global__ void __launch_bounds__(THREADS, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmF, __nv_bfloat16* out, int ldc,
             int M, int N, int K, unsigned long long* dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  constexpr int STG_BYTES = (EPI >= 2) ? 4 * 2 * 4096 : 0;  // per epilogue warp: two 32-row x 128-byte staging buffers
  uint8_t* stg_base = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();      // 0..1 (pair) or 0..3 (two pairs)
  const uint32_t prank = rank & 1;              // rank inside the CTA pair
  const uint32_t pair = rank >> 1;              // which pair of the cluster (CL4)
  const bool leader = prank == 0;
  constexpr int CSZ = CL4 ? 4 : 2;
  const int cluster_id = blockIdx.x / CSZ, num_clusters = gridDim.x / CSZ;
  // a cluster's output block: (CL4 ? 512 : 256) rows x 256 columns; pair p takes rows [p*256, p*256+256)
  constexpr int CM = CL4 ? 512 : 256;
  const int num_m = (M + CM - 1) / CM, num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n, num_k = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    // full: the leader's own arrive.expect_tx + one remote arrive from its pair peer
    // empty (CL4): a stage's B half is written by the OTHER pair's TMA too, so a slot is free only when BOTH pairs' MMAs
    //              have consumed it: 2 commits
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar + s, NOARR ? 1 : 2); mbar_init(empty_bar + s, CL4 ? 2 : 1); }
    for (int s = 0; s < 2; s++) { mbar_init(tfull_bar + s, 1); mbar_init(tempty_bar + s, TE1 ? 2 : 256); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long t_begin = clock64();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      unsigned long long w_empty = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = (tile / num_n) * CM + (int)pair * 256 + (int)prank * BM_CTA;
        // B rows of this CTA: pair layout: prank selects the 128-column half.  CL4: each half is further split in two 64-row
        // quarters, quarter `pair` loaded by this CTA and multicast to the same-prank CTA of the other pair.
        const int n0 = (tile % num_n) * BN + (int)prank * BN_CTA;
        for (int kb = 0; kb < num_k; kb++) {
          const long long t0 = clock64();
          pwait<WAIT>(empty_bar + stage, phase ^ 1);
          w_empty += clock64() - t0;
          const uint32_t leader_full = mapa(smem_u32(full_bar + stage), rank & ~1u);
          if (leader) mbar_arrive_expect_tx(full_bar + stage, LOAD == 0 ? 2 * STAGE_BYTES : 0);
          else if (!NOARR) mbar_arrive_remote(leader_full);
          if (LOAD == 0) {
            tma_load_2d_2sm(sA + stage * A_BYTES, &tmA, leader_full, kb * BK, m0);
            if (!CL4) {
              tma_load_2d_2sm(sB + stage * B_BYTES, &tmB, leader_full, kb * BK, n0);
            } else {
              // 64 of this CTA's 128 B rows, delivered to both CTAs with the same prank (ranks prank and prank + 2)
              const uint16_t mask = (uint16_t)(0x5u << prank);
              tma_load_2d_2sm_mc(sB + stage * B_BYTES + (int)pair * (64 * BK * 2), &tmB, leader_full, kb * BK,
                                 n0 + (int)pair * 64, mask);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      dbg[blockIdx.x * 8 + 0] = w_empty;
      dbg[blockIdx.x * 8 + 7] = clock64() - t_begin;
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM_CTA, BN, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      unsigned long long w_full = 0, w_tempty = 0, tiles = 0;
      const uint16_t commit_mask = CL4 ? 0xF : 0x3;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        long long t0 = clock64();
        pwait<WAIT>(tempty_bar + acc, acc_phase ^ 1);
        w_tempty += clock64() - t0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_k; kb++) {
          t0 = clock64();
          pwait<WAIT>(full_bar + stage, phase);
          w_full += clock64() - t0;
          tc_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + stage * A_BYTES), 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + stage * B_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++)
            umma_bf16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          umma_commit_2sm(empty_bar + stage, commit_mask);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(tfull_bar + acc, (uint16_t)(0x3u << (rank & ~1u)));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        tiles++;
      }
      dbg[blockIdx.x * 8 + 1] = w_full;
      dbg[blockIdx.x * 8 + 2] = w_tempty;
      dbg[blockIdx.x * 8 + 5] = clock64() - t_begin;
      dbg[blockIdx.x * 8 + 6] = tiles;
    }
  } else {
    const int quad = warp & 3;
    const uint32_t leader_tempty0 = mapa(smem_u32(tempty_bar), rank & ~1u);
    int acc = 0;
    uint32_t acc_phase = 0;
    unsigned long long w_tfull = 0, w_drain = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m0 = (tile / num_n) * CM + (int)pair * 256 + (int)prank * BM_CTA, n0 = (tile % num_n) * BN;
      const int row = m0 + quad * 32 + lane;
      long long t0 = clock64();
      pwait<WAIT>(tfull_bar + acc, acc_phase);
      long long t1 = clock64();
      w_tfull += t1 - t0;
      tc_fence_after();
      if (EPI == 0) {
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
        for (int c = 0; c < BN / 32; c++) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + (uint32_t)(c * 32), r);
          tmem_ld_wait();
          if (row < M) {
            __nv_bfloat16* o = out + (size_t)row * ldc + n0 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk;
              __nv_bfloat162 v;
              v = __floats2bfloat162_rn(__uint_as_float(r[j]), __uint_as_float(r[j + 1])); pk.x = *reinterpret_cast<uint32_t*>(&v);
              v = __floats2bfloat162_rn(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])); pk.y = *reinterpret_cast<uint32_t*>(&v);
              v = __floats2bfloat162_rn(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5])); pk.z = *reinterpret_cast<uint32_t*>(&v);
              v = __floats2bfloat162_rn(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7])); pk.w = *reinterpret_cast<uint32_t*>(&v);
              *reinterpret_cast<uint4*>(o + j) = pk;
            }
          }
        }
      }
      if (EPI == 2) {
        // bf16 tile through swizzled smem staging + TMA store: one 32-row x 64-column box (128-byte rows) per step
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
        uint8_t* stg = stg_base + (warp - 2) * 8192;
#pragma unroll 1
        for (int c = 0; c < BN / 64; c++) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(t_row + (uint32_t)(c * 64), r0);
          tmem_ld_32x32(t_row + (uint32_t)(c * 64 + 32), r1);
          uint8_t* buf = stg + (c & 1) * 4096;
          if (lane == 0) bulk_wait_read<1>();  // the store that read this buffer two steps ago is done with it
          __syncwarp();
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const uint32_t* r = (j < 4) ? (r0 + 8 * j) : (r1 + 8 * (j - 4));
            uint4 pk;
            __nv_bfloat162 v;
            v = __floats2bfloat162_rn(__uint_as_float(r[0]), __uint_as_float(r[1])); pk.x = *reinterpret_cast<uint32_t*>(&v);
            v = __floats2bfloat162_rn(__uint_as_float(r[2]), __uint_as_float(r[3])); pk.y = *reinterpret_cast<uint32_t*>(&v);
            v = __floats2bfloat162_rn(__uint_as_float(r[4]), __uint_as_float(r[5])); pk.z = *reinterpret_cast<uint32_t*>(&v);
            v = __floats2bfloat162_rn(__uint_as_float(r[6]), __uint_as_float(r[7])); pk.w = *reinterpret_cast<uint32_t*>(&v);
            *reinterpret_cast<uint4*>(buf + lane * 128 + ((j ^ (lane & 7)) * 16)) = pk;  // SWIZZLE_128B
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmO, buf, n0 + c * 64, m0 + quad * 32);
            bulk_commit();
          }
        }
      }
      if (EPI == 3) {
        // fp32 tile added INTO global memory by the TMA (cp.reduce ... add): x += v without reading x on the SM
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
        uint8_t* stg = stg_base + (warp - 2) * 8192;
#pragma unroll 1
        for (int c = 0; c < BN / 32; c++) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + (uint32_t)(c * 32), r);
          uint8_t* buf = stg + (c & 1) * 4096;
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; j++)
            *reinterpret_cast<uint4*>(buf + lane * 128 + ((j ^ (lane & 7)) * 16)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tmF, buf, n0 + c * 32, m0 + quad * 32);
            bulk_commit();
          }
        }
      }
      w_drain += clock64() - t1;
      tc_fence_before();
      if (TE1) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
          if (leader) mbar_arrive(tempty_bar + acc);
          else mbar_arrive_remote(leader_tempty0 + (uint32_t)(acc * 8));
        }
      } else {
        if (leader) mbar_arrive(tempty_bar + acc);
        else mbar_arrive_remote(leader_tempty0 + (uint32_t)(acc * 8));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (EPI >= 2 && lane == 0) bulk_wait<0>();
    if (threadIdx.x == 64) { dbg[blockIdx.x * 8 + 3] = w_tfull; dbg[blockIdx.x * 8 + 4] = w_drain; }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

struct Bufs {
  __nv_bfloat16 *A, *W, *O;
  float* F;
  unsigned long long* dbg;
};

template <int WAIT, int EPI, int LOAD, int STAGES, int CL4, int NOARR = 0, int TE1 = 0>
static void run(const char* name, const Bufs& b, int M, int N, int K, int num_sms, int check) {
  auto kern = probe_kernel<WAIT, EPI, LOAD, STAGES, CL4, NOARR, TE1>;
  const int smem = STAGES * STAGE_BYTES + 1024 + 256 + (EPI >= 2 ? 32768 : 0);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {BK, BM_CTA};
    if (dgs::make_tmap_bf16(&tmA, b.A, 2, dims, str, box)) { printf("tmap A failed\n"); exit(1); }
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {BK, (uint32_t)(CL4 ? 64 : BN_CTA)};
    if (dgs::make_tmap_bf16(&tmB, b.W, 2, dims, str, box)) { printf("tmap B failed\n"); exit(1); }
  }
  CUtensorMap tmO, tmF;
  {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, str[1] = {(uint64_t)N * 2};
    uint32_t box[2] = {64, 32};
    if (dgs::make_tmap_bf16(&tmO, b.O, 2, dims, str, box)) { printf("tmap O failed\n"); exit(1); }
  }
  {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, str[1] = {(uint64_t)N * 4};
    uint32_t box[2] = {32, 32};
    dgs::g_tmap_dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const int rc = dgs::make_tmap_bf16(&tmF, b.F, 2, dims, str, box);
    dgs::g_tmap_dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (rc) { printf("tmap F failed\n"); exit(1); }
  }
  if (EPI == 3) {
    std::vector<float> ones((size_t)M * N, 1.0f);
    CK(cudaMemcpy(b.F, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice));
  }
  if (EPI == 2) CK(cudaMemset(b.O, 0, (size_t)M * N * 2));
  constexpr int CSZ = CL4 ? 4 : 2;
  constexpr int CM = CL4 ? 512 : 256;
  const int tiles = ((M + CM - 1) / CM) * (N / BN);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_sms / CSZ * CSZ);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CSZ; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  // a persistent kernel must not launch more clusters than can be co-resident (GPC sizes limit clusters of 4)
  int clusters = 0;
  CK(cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg));
  static int printed[2] = {0, 0};
  if (!printed[CL4]) { printf("  [max co-resident clusters of %d: %d]\n", CSZ, clusters); printed[CL4] = 1; }
  if (clusters > num_sms / CSZ) clusters = num_sms / CSZ;
  if (tiles < clusters) clusters = tiles;
  cfg.gridDim = dim3(CSZ * clusters);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaMemset(b.dbg, 0, 8 * 8 * 160));
  float best = 1e30f;
  for (int it = 0; it < 8; it++) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmO, tmF, b.O, N, M, N, K, b.dbg));
    CK(cudaEventRecord(e1));
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { printf("%-28s FAILED: %s\n", name, cudaGetErrorString(e)); exit(1); }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  std::vector<unsigned long long> h(8 * CSZ * clusters);
  CK(cudaMemcpy(h.data(), b.dbg, h.size() * 8, cudaMemcpyDeviceToHost));
  // averages over leader CTAs (mma counters) and over all CTAs (producer / epilogue counters)
  double pe = 0, mf = 0, mt = 0, ef = 0, ed = 0, tot = 0, tl = 0, ptot = 0;
  int nl = 0, nc = CSZ * clusters;
  for (int c = 0; c < nc; c++) {
    pe += h[c * 8 + 0]; ef += h[c * 8 + 3]; ed += h[c * 8 + 4]; ptot += h[c * 8 + 7];
    if ((c & 1) == 0) { mf += h[c * 8 + 1]; mt += h[c * 8 + 2]; tot += h[c * 8 + 5]; tl += h[c * 8 + 6]; nl++; }
  }
  const double kblocks = (tl / nl) * ((K + BK - 1) / BK);
  printf("%-28s M=%d N=%d K=%d  %.1f us  %.0f TFLOP/s | tiles/cluster %.2f  mma: total %.0f clk (%.0f/kblock) wait_full %.0f wait_tempty %.0f"
         " | producer: total %.0f wait_empty %.0f | epi: wait_tfull %.0f drain %.0f (%.0f/tile)\n",
         name, M, N, K, best * 1e3, 2.0 * M * N * K / best / 1e9, tl / nl, tot / nl, tot / nl / kblocks, mf / nl, mt / nl,
         ptot / nc, pe / nc, ef / nc, ed / nc, ed / nc / (tl / nl));
  if (check && EPI == 3 && LOAD == 0) {
    // F started at 1 and received 8 launches' worth of += acc
    std::vector<__nv_bfloat16> hA((size_t)M * K), hW((size_t)N * K);
    std::vector<float> hF((size_t)M * N);
    CK(cudaMemcpy(hA.data(), b.A, hA.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hW.data(), b.W, hW.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hF.data(), b.F, hF.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int s = 0; s < 256; s++) {
      const int r = (s < 4) ? (M - 1 - s) : (int)((s * 2654435761u) % (unsigned)M), c = (int)((s * 40503u + 17u * s * s) % (unsigned)N);
      double acc = 0;
      for (int k = 0; k < K; k++) acc += (double)__bfloat162float(hA[(size_t)r * K + k]) * (double)__bfloat162float(hW[(size_t)c * K + k]);
      const double want = 1.0 + 8.0 * acc, err = fabs(hF[(size_t)r * N + c] - want) / (fabs(want) + 1.0);
      if (err > worst) worst = err;
    }
    printf("    check (fp32 reduce-add, 8 launches): worst rel err %.3e %s\n", worst, worst < 1e-4 ? "OK" : "MISMATCH");
  }
  if (check && (EPI == 0 || EPI == 2) && LOAD == 0) {
    // spot check 64 entries against a host dot product
    std::vector<__nv_bfloat16> hA((size_t)M * K), hW((size_t)N * K), hO((size_t)M * N);
    CK(cudaMemcpy(hA.data(), b.A, hA.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hW.data(), b.W, hW.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hO.data(), b.O, hO.size() * 2, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int s = 0; s < 256; s++) {
      const int r = (s < 4) ? (M - 1 - s) : (int)((s * 2654435761u) % (unsigned)M), c = (int)((s * 40503u + 17u * s * s) % (unsigned)N);
      double acc = 0;
      for (int k = 0; k < K; k++) acc += (double)__bfloat162float(hA[(size_t)r * K + k]) * (double)__bfloat162float(hW[(size_t)c * K + k]);
      const double got = __bfloat162float(hO[(size_t)r * N + c]);
      const double err = fabs(got - acc) / (fabs(acc) + 1.0);
      if (err > worst) worst = err;
    }
    printf("    check: worst rel err over 256 samples %.3e %s\n", worst, worst < 2e-2 ? "OK" : "MISMATCH");
  }
}

extern "C" int dgs_gemm_bf16(const void* A, const void* Wt, const float* bias, const float* gate, void* out, int M, int N, int K,
                             int epi, int ldc, int gate_stride, int rows_per_sample, void* stream);
extern "C" const char* dgs_last_error();
static void run_lib(const Bufs& b, int M, int N, int K) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int it = 0; it < 8; it++) {
    CK(cudaEventRecord(e0));
    if (dgs_gemm_bf16(b.A, b.W, nullptr, nullptr, b.O, M, N, K, 0, N, 0, 1, nullptr)) { printf("lib gemm failed: %s\n", dgs_last_error()); return; }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  printf("%-28s M=%d N=%d K=%d  %.1f us  %.0f TFLOP/s\n", "library 1-CTA 128x256", M, N, K, best * 1e3, 2.0 * M * N * K / best / 1e9);
}

__global__ void fill_kernel(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = (uint32_t)i * 747796405u + seed;
  x = ((x >> ((x >> 28) + 4)) ^ x) * 277803737u;
  x = (x >> 22) ^ x;
  p[i] = __float2bfloat16(((x & 0xFFFF) / 65536.0f - 0.5f) * scale);
}

int main(int argc, char** argv) {
  int dev = 0, num_sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  printf("SMs %d\n", num_sms);
  const int shapes[][3] = {{4098, 3072, 1024}, {4098, 4096, 1024}, {4098, 1024, 4096}, {8192, 4096, 4096}};
  size_t maxA = 0, maxW = 0, maxO = 0;
  for (auto& s : shapes) {
    maxA = std::max(maxA, (size_t)s[0] * s[2]); maxW = std::max(maxW, (size_t)s[1] * s[2]); maxO = std::max(maxO, (size_t)s[0] * s[1]);
  }
  Bufs b;
  CK(cudaMalloc(&b.A, maxA * 2));
  CK(cudaMalloc(&b.W, maxW * 2));
  CK(cudaMalloc(&b.O, maxO * 2));
  CK(cudaMalloc(&b.F, maxO * 4));
  CK(cudaMalloc(&b.dbg, 8 * 8 * 160));
  fill_kernel<<<(unsigned)((maxA + 255) / 256), 256>>>(b.A, maxA, 1u, 2.0f);
  fill_kernel<<<(unsigned)((maxW + 255) / 256), 256>>>(b.W, maxW, 2u, 0.1f);
  CK(cudaDeviceSynchronize());
  for (auto& s : shapes) {
    const int M = s[0], N = s[1], K = s[2];
    const int check = (M == 4098 && N == 3072);
    run_lib(b, M, N, K);
    run<0, 0, 0, 6, 0, 1, 1>("pair (product protocol)", b, M, N, K, num_sms, check);
    run<0, 0, 0, 4, 0, 1, 1>("pair 4 stages", b, M, N, K, num_sms, 0);
    run<0, 2, 0, 4, 0, 1, 1>("pair 4st bf16 TMA-store epi", b, M, N, K, num_sms, 1);
    run<0, 2, 0, 5, 0, 1, 1>("pair 5st bf16 TMA-store epi", b, M, N, K, num_sms, 0);
    run<0, 3, 0, 4, 0, 1, 1>("pair 4st f32 TMA-reduce epi", b, M, N, K, num_sms, 1);
    run<0, 3, 0, 5, 0, 1, 1>("pair 5st f32 TMA-reduce epi", b, M, N, K, num_sms, 0);
    printf("\n");
    fflush(stdout);
  }
  return 0;
}
