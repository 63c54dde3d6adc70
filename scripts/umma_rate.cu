// umma_rate.cu -- stand-alone micro-benchmark (NOT part of the product): how long does ONE thread need to issue a
// tcgen05.mma of a given shape, and at what rate does the tensor pipe retire back-to-back MMAs of that shape, with one and
// with two CTAs per SM?  Decides the attention redesign: is the forward attention kernel, which
// issues 12 small MMAs per 64-key block, bound by the instruction rate of the tensor pipe or by its math rate?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I open-diffusiongs_b200/csrc \
//        -I include scripts/umma_rate.cu -o scripts/bin/umma_rate
//   scripts/bin/umma_rate            (on the GPU box)
#include <cstdio>
#include <vector>

#include "../open-diffusiongs_b200/csrc/sm100_ptx.cuh"

using namespace dgs::ptx;

struct Result { unsigned long long issue, total; };

// mode 0: A and B from shared memory (K-major, SW128); mode 1: A from tensor memory, B from smem (MN-major like V)
template <int N_MMA, int MODE>
__global__ void __launch_bounds__(128) rate_kernel(Result* out, int reps, int per_commit) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;               // 128 rows x 64 bf16 (16 KB)
  uint8_t* sB = smem + 16384;       // up to 256 rows x 64 bf16 (32 KB)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N_MMA, false, MODE == 1);
    const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA), 16, 1024);
    const uint64_t bdesc = MODE == 1 ? make_smem_desc_sw128(smem_u32(sB), 8192, 1024) : make_smem_desc_sw128(smem_u32(sB), 16, 1024);
    unsigned long long t_issue = 0, t0 = clock64();
    uint32_t phase = 0;
    for (int r = 0; r < reps; r += per_commit) {
      const unsigned long long a = clock64();
      if (elect_one_sync()) {
        for (int i = 0; i < per_commit; i++) {
          const int k = i & 3;
          if (MODE == 0) umma_bf16(tmem + 128, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
          else umma_bf16_ts(tmem + 128, tmem + (uint32_t)(k * 8), bdesc + (uint64_t)(k * 128), idesc, 1u);
        }
        umma_commit(bar);
      }
      __syncwarp();
      t_issue += clock64() - a;
      mbar_wait(bar, phase);
      phase ^= 1;
    }
    const unsigned long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x].issue = t_issue; out[blockIdx.x].total = t1 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <int N_MMA, int MODE>
static void run(const char* name, int ctas_per_sm, int per_commit) {
  const int sms = 148, reps = 4096;
  Result* d;
  cudaMalloc(&d, sizeof(Result) * sms * 2);
  // smem sized so that exactly ctas_per_sm CTAs fit (1: > 114 KB; 2: ~50 KB + padding)
  const int smem = ctas_per_sm == 1 ? 120 * 1024 : 60 * 1024;
  cudaFuncSetAttribute(rate_kernel<N_MMA, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int it = 0; it < 2; it++) rate_kernel<N_MMA, MODE><<<sms * ctas_per_sm, 128, smem>>>(d, reps, per_commit);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return; }
  std::vector<Result> h(sms * ctas_per_sm);
  cudaMemcpy(h.data(), d, sizeof(Result) * h.size(), cudaMemcpyDeviceToHost);
  double si = 0, st = 0;
  for (auto& r : h) { si += (double)r.issue; st += (double)r.total; }
  si /= h.size(); st /= h.size();
  const double macs = 128.0 * N_MMA * 16;
  printf("%-34s %d CTA/SM, %2d MMA per commit: issue %6.1f clk/MMA, retire %6.1f clk/MMA per CTA (%6.1f clk/MMA per SM) = %5.0f MAC/clk/SM\n",
         name, ctas_per_sm, per_commit, si / reps, st / reps, st / reps / ctas_per_sm, macs * ctas_per_sm / (st / reps));
  cudaFree(d);
}

int main() {
  for (int c = 1; c <= 2; c++) {
    for (int pc : {4, 16, 64}) {
      run<64, 0>("SS 128x64x16  (S, 64 keys)", c, pc);
      run<128, 0>("SS 128x128x16 (S, 128 keys)", c, pc);
      run<256, 0>("SS 128x256x16", c, pc);
      run<64, 1>("TS 128x64x16  (P V)", c, pc);
      run<16, 1>("TS 128x16x16  (row sum)", c, pc);
      run<128, 1>("TS 128x128x16", c, pc);
    }
  }
  return 0;
}
