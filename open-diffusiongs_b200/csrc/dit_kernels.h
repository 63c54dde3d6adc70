// dit_kernels.h -- internal launchers of the DiT denoiser kernels (host side, C++).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace dgs {

enum GemmEpi { EPI_BIAS_BF16 = 0, EPI_BIAS_GELU_BF16 = 1, EPI_GATE_RESID_F32 = 2, EPI_F32 = 3 };

struct GemmEpilogue {
  void* out = nullptr;          // bf16 or fp32 [M, ldc]
  int ldc = 0;
  const float* bias = nullptr;  // [N] or null
  const float* gate = nullptr;  // EPI_GATE_RESID_F32: gate vector of sample b at gate + b * gate_stride
  int gate_stride = 0;
  int rows_per_sample = 1;      // sample index of a row = row / rows_per_sample
};

// C = epi(A[M,K] * W[N,K]^T), bf16 operands, fp32 accumulate (gemm_sm100.cu)
int gemm_bf16(const void* A, const void* W, int M, int N, int K, int epi, const GemmEpilogue& ep, cudaStream_t st);

// CTA-pair (cta_group::2) variant, 256 x 256 tiles, needs N % 256 == 0 (gemm2_sm100.cu); gemm_bf16 dispatches to it
int gemm_bf16_2cta(const void* A, const void* W, int M, int N, int K, int epi, const GemmEpilogue& ep, cudaStream_t st);

// single-CTA 256 x 256 tile variant (gemm3_sm100.cu), needs N % 256 == 0
int gemm_bf16_m256(const void* A, const void* W, int M, int N, int K, int epi, const GemmEpilogue& ep, cudaStream_t st);

// softmax(Q K^T / sqrt(64)) V over qkv [B, N, 3, H, 64] (bf16) -> out [B, N, H*64] (bf16) (attention_sm100.cu)
int attention_fwd(const void* qkv, void* out, int B, int N, int H, cudaStream_t st);

// ---- dit_misc.cu -------------------------------------------------------------------------------
// h[r,:] = (LN(x[r,:]; eps) [* w]) * (1 + scale[b,:]) + shift[b,:]   -> bf16
// rows are gathered: output row r (0..B*rows_out) reads x row  b*rows_in + row_off + (r % rows_out)
// split != 0: the output row is [hi | lo | hi] (3*D bf16) with value = hi + lo (split-bf16 operand)
int ln_modulate(const float* x, const float* ln_weight, const float* shift, const float* scale, int mod_stride,
                __nv_bfloat16* h, int B, int rows_in, int row_off, int rows_out, int D, float eps, int split,
                cudaStream_t st);
// plain LayerNorm with weight (no bias), fp32 -> fp32, in place over [rows, D]
int ln_weight_inplace(float* x, const float* w, int rows, int D, float eps, cudaStream_t st);
// out[b, n] = act_in(in[b,:]) . W[n,:] + bias[n], W fp32 [N,K]; act_in: 0 none, 1 SiLU
int skinny_linear(const float* in, const float* W, const float* bias, float* out, int B, int N, int K,
                  int act_in, int act_out_silu, cudaStream_t st);
// sinusoidal timestep embedding (denoiser.py:44-66): out [B, 256] = [cos(t f), sin(t f)]
int timestep_embedding(const float* t, float* out, int B, int dim, cudaStream_t st);
// Plücker-style posed image + patchify (denoiser.py:312-334, 210-216) -> split-bf16 tokens
// [B*V*hh*ww, 3*p*p*9] = [hi | lo | hi]
int posed_patchify(const float* images, const float* ray_o, const float* ray_d, __nv_bfloat16* tokens, int B, int V,
                   int H, int W, int patch, int plucker_mode, cudaStream_t st);
// x[b, 0:G] = pos_embed ; x[b, G:] = tok[b] ; (then the caller applies the input LayerNorm)
int assemble_tokens(const float* tok, const float* pos_embed, float* x, int B, int G, int T, int D, cudaStream_t st);
// small-M linear for the 2 free Gaussian tokens: out[r, n] = h[r,:] . W[n,:]   (N = 14)
int tiny_linear_bf16(const __nv_bfloat16* h, const __nv_bfloat16* W, float* out, int rows, int N, int K,
                     cudaStream_t st);
// to_gs + pixel alignment (denoiser.py:103-120, 362-413): raw head outputs -> renderer tensors
struct GsOut { float* xyz; float* features; float* scaling; float* rotation; float* opacity; float* img_aligned_xyz; };
int gaussians_epilogue(const float* gs_tokens /*[B,G,14]*/, const float* img_gs /*[B*V*hh*ww, p*p*14]*/,
                       const float* ray_o, const float* ray_d, GsOut out, int B, int G, int V, int H, int W, int patch,
                       int scene_mode, float near_, float far_, cudaStream_t st);
int f32_to_bf16(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t st);

}  // namespace dgs
