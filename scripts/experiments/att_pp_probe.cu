// att_pp_probe.cu -- timeline probe of the ping-pong attention kernel (NOT part of the product): compiles the PRODUCT source
// (csrc/attention_pp_sm100.cu) with DGS_ATT_PROBE; every role accumulates the cycles it waits on each hand-over.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I open-diffusiongs_b200/csrc \
//        -I include scripts/att_pp_probe.cu -o scripts/bin/att_pp_probe -L open-diffusiongs_b200/dgs_b200/lib -ldgs_b200 \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../../open-diffusiongs_b200/dgs_b200/lib'
#define DGS_ATT_PROBE 1
#include "../open-diffusiongs_b200/csrc/attention_pp_sm100.cu"

#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__global__ void fill_qkv(__nv_bfloat16* p, size_t n, uint32_t seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = (uint32_t)i * 747796405u + seed;
  x = ((x >> ((x >> 28) + 4)) ^ x) * 277803737u;
  x = (x >> 22) ^ x;
  const float u = ((x & 0xFFFF) + ((x >> 16) & 0xFFFF)) / 65536.0f - 1.0f;
  p[i] = __float2bfloat16(u * 3.7f);
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 4098, H = argc > 2 ? atoi(argv[2]) : 16, B = argc > 3 ? atoi(argv[3]) : 1;
  const int poly = argc > 4 ? atoi(argv[4]) : 0;
  const int D = H * 64;
  const size_t n_qkv = (size_t)B * N * 3 * D;
  __nv_bfloat16 *qkv, *out;
  CK(cudaMalloc(&qkv, n_qkv * 2));
  CK(cudaMalloc(&out, (size_t)B * N * D * 2));
  fill_qkv<<<(unsigned)((n_qkv + 255) / 256), 256>>>(qkv, n_qkv, 12345u);
  const int ctas = ((N + 255) / 256) * H * B, nb = (N + 127) / 128;
  unsigned long long* dbg;
  CK(cudaMalloc(&dbg, (size_t)ctas * 32 * 8));
  CK(cudaMemset(dbg, 0, (size_t)ctas * 32 * 8));
  CK(cudaMemcpyToSymbol(dgs::pp::g_pp_dbg, &dbg, sizeof(dbg)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int it = 0; it < 6; it++) {
    CK(cudaEventRecord(e0));
    if (dgs::attention_fwd_pp(qkv, out, nullptr, B, N, H, poly, nullptr)) { printf("attention_fwd_pp failed: %s\n", dgs_last_error()); return 1; }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  std::vector<unsigned long long> h((size_t)ctas * 32);
  CK(cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost));
  double s[32] = {};
  for (int c = 0; c < ctas; c++)
    for (int i = 0; i < 32; i++) s[i] += (double)h[(size_t)c * 32 + i] / ctas / nb;
  printf("pp attention N=%d H=%d B=%d poly=%d: %.1f us (with probe overhead), %.0f TFLOP/s; %d CTAs x %d key blocks; cycles PER KEY BLOCK:\n",
         N, H, B, poly, best * 1e3, 4.0 * N * N * D * B / best / 1e9, ctas, nb);
  for (int t = 0; t < 2; t++)
    printf("  softmax WG%d  loop %.0f   wait S %.0f   pass 1 (max) %.0f   wait PV(j-1) %.0f   wait turn %.0f   pass 2 (exp) %.0f\n", t,
           s[t * 8 + 5], s[t * 8 + 0], s[t * 8 + 1], s[t * 8 + 2], s[t * 8 + 3], s[t * 8 + 4]);
  printf("  MMA issuer   loop %.0f   wait s_free0 %.0f  s_free1 %.0f   wait p_full0 %.0f  p_full1 %.0f   wait V %.0f   wait K %.0f\n", s[22], s[16],
         s[17], s[18], s[19], s[20], s[21]);
  return 0;
}
