// att_probe.cu -- stand-alone timeline probe of the forward attention kernel (NOT part of the product).
//
// Compiles the PRODUCT kernel source (csrc/attention_sm100.cu) with DGS_ATT_PROBE defined: every role then accumulates the
// cycles it waits on each barrier (producer: free K/V stage; MMA issuer: K, P, V; softmax thread: S, TMEM load, previous
// P V before a rescale, TMEM store) and the softmax thread its whole loop.  Answers the question the round-1 experiments
// left open (attention row): WHAT does the per-block chain of one CTA wait for at N = 4098?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I open-diffusiongs_b200/csrc \
//        -I include scripts/att_probe.cu -o scripts/bin/att_probe -L open-diffusiongs_b200/dgs_b200/lib -ldgs_b200 \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../../open-diffusiongs_b200/dgs_b200/lib'
//   scripts/bin/att_probe [N] [H] [B]          (on the GPU box; default 4098 16 1)
#define DGS_ATT_PROBE 1
#include "../open-diffusiongs_b200/csrc/attention_sm100.cu"

#include <cstdio>
#include <vector>

#define CK(x)                                                                                                     \
  do {                                                                                                            \
    cudaError_t e_ = (x);                                                                                         \
    if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } \
  } while (0)

__global__ void fill_qkv(__nv_bfloat16* p, size_t n, uint32_t seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = (uint32_t)i * 747796405u + seed;
  x = ((x >> ((x >> 28) + 4)) ^ x) * 277803737u;
  x = (x >> 22) ^ x;
  // sum of two uniforms ~ triangular, scaled to std ~ 1.5 like the parity tests
  const float u = ((x & 0xFFFF) + ((x >> 16) & 0xFFFF)) / 65536.0f - 1.0f;
  p[i] = __float2bfloat16(u * 3.7f);
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 4098, H = argc > 2 ? atoi(argv[2]) : 16, B = argc > 3 ? atoi(argv[3]) : 1;
  const int D = H * 64;
  const size_t n_qkv = (size_t)B * N * 3 * D;
  __nv_bfloat16 *qkv, *out;
  CK(cudaMalloc(&qkv, n_qkv * 2));
  CK(cudaMalloc(&out, (size_t)B * N * D * 2));
  fill_qkv<<<(unsigned)((n_qkv + 255) / 256), 256>>>(qkv, n_qkv, 12345u);
  const int nq = (N + 127) / 128, ctas = nq * H * B, nb = (N + 63) / 64;
  unsigned long long* dbg;
  CK(cudaMalloc(&dbg, (size_t)ctas * 16 * 8));
  CK(cudaMemset(dbg, 0, (size_t)ctas * 16 * 8));
  CK(cudaMemcpyToSymbol(dgs::g_att_dbg, &dbg, sizeof(dbg)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int it = 0; it < 6; it++) {
    CK(cudaEventRecord(e0));
    if (dgs::attention_fwd(qkv, out, nullptr, B, N, H, nullptr)) { printf("attention_fwd failed: %s\n", dgs_last_error()); return 1; }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  std::vector<unsigned long long> h((size_t)ctas * 16);
  CK(cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost));
  double s[16] = {};
  for (int c = 0; c < ctas; c++)
    for (int i = 0; i < 16; i++) s[i] += (double)h[(size_t)c * 16 + i];
  for (int i = 0; i < 16; i++) s[i] /= ctas;
  printf("attention N=%d H=%d B=%d: %.1f us (with probe overhead), %.0f TFLOP/s; %d CTAs x %d key blocks; per CTA, average cycles:\n",
         N, H, B, best * 1e3, 4.0 * N * N * D * B / best / 1e9, ctas, nb);
  printf("  softmax thread  loop total %.0f (%.0f per block)   wait S %.0f   TMEM ld S %.0f   wait prev PV (rescale) %.0f   TMEM st P %.0f   => math+other %.0f\n",
         s[8], s[8] / nb, s[4], s[5], s[6], s[7], s[8] - s[4] - s[5] - s[6] - s[7]);
  printf("  MMA issuer      loop total %.0f   wait K %.0f   wait P %.0f   wait V %.0f   issuing S (4 MMA + commit) %.0f (%.0f per block)   issuing PV+L (8 MMA + 2 commits) %.0f (%.0f per block)\n",
         s[9], s[1], s[2], s[3], s[10], s[10] / nb, s[11], s[11] / nb);
  printf("  TMA producer    wait free K/V stage %.0f\n", s[0]);
  return 0;
}
