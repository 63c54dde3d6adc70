// dit_api.cu -- C ABI of the DiT denoiser forward (include/dgs_b200.h, section B2): orchestrates the
// tcgen05 GEMMs, the tcgen05 attention and the glue kernels into DGSDenoiser.image_to_gaussians
// (diffusionGS/models/denoiser/denoiser.py:306-416).
#include "dgs_internal.h"
#include "dit_kernels.h"

using namespace dgs;

namespace {

struct DitWorkspace {
  __nv_bfloat16* tokens;  // [B*T, 3*p*p*9] split-bf16
  float* tok;             // [B*T, w]
  float* x;               // [B*N, w]   fp32 residual stream
  __nv_bfloat16* h;       // [B*N, w]
  __nv_bfloat16* qkv;     // [B*N, 3w]
  __nv_bfloat16* attn;    // [B*N, w]
  __nv_bfloat16* u;       // [B*N, 4w]
  float* temb0;           // [B, 256]
  float* temb1;           // [B, w]
  float* c;               // [B, w]
  float* mod;             // [B, L*6w + 4w]
  __nv_bfloat16* hg;      // [B*G, 3w] split-bf16
  float* gs_tok;          // [B*G, 14]
  float* img_gs;          // [B*T, p*p*14]
  size_t bytes;
  DitWorkspace(void* base, const dgs_dit_weights* w, int B, int V, int H, int W) {
    const size_t T = (size_t)V * (H / w->patch) * (W / w->patch), N = T + w->n_gaussians, D = w->width;
    Carver c_(base);
    tokens = c_.take<__nv_bfloat16>((size_t)B * T * 3 * w->patch * w->patch * 9);
    tok = c_.take<float>((size_t)B * T * D);
    x = c_.take<float>((size_t)B * N * D);
    h = c_.take<__nv_bfloat16>((size_t)B * N * D);
    qkv = c_.take<__nv_bfloat16>((size_t)B * N * 3 * D);
    attn = c_.take<__nv_bfloat16>((size_t)B * N * D);
    u = c_.take<__nv_bfloat16>((size_t)B * N * w->mlp_hidden);
    temb0 = c_.take<float>((size_t)B * 256);
    temb1 = c_.take<float>((size_t)B * D);
    c = c_.take<float>((size_t)B * D);
    mod = c_.take<float>((size_t)B * ((size_t)w->layers * 6 * D + 4 * D));
    hg = c_.take<__nv_bfloat16>((size_t)B * w->n_gaussians * 3 * D);
    gs_tok = c_.take<float>((size_t)B * w->n_gaussians * 14);
    img_gs = c_.take<float>((size_t)B * T * w->patch * w->patch * 14);
    bytes = c_.bytes();
  }
};

int check_dit(const dgs_dit_weights* w, int B, int V, int H, int W) {
  DGS_REQUIRE(w != nullptr, "weights is NULL");
  DGS_REQUIRE(w->width == 1024 && w->heads * 64 == w->width, "unsupported width/heads %d/%d (1024/16 only)", w->width, w->heads);
  DGS_REQUIRE(w->layers > 0 && w->patch > 0 && w->n_gaussians >= 0 && w->mlp_hidden % 256 == 0, "bad DiT config");
  DGS_REQUIRE(B > 0 && V > 0 && H % w->patch == 0 && W % w->patch == 0, "bad input shape B=%d V=%d H=%d W=%d", B, V, H, W);
  DGS_REQUIRE((w->patch * w->patch * 14) % 32 == 0 && (w->patch * w->patch * 9) % 8 == 0, "patch %d unsupported", w->patch);
  DGS_REQUIRE(w->mlp_hidden >= 3 * w->width, "mlp_hidden must be >= 3*width (decoder head re-uses that buffer)");
  return DGS_OK;
}

#define DGS_TRY(expr)       \
  do {                      \
    int _rc = (expr);       \
    if (_rc) return _rc;    \
  } while (0)

}  // namespace

extern "C" {

size_t dgs_dit_workspace_bytes(const dgs_dit_weights* w, int B, int V, int H, int W) {
  if (check_dit(w, B, V, H, W)) return 0;
  return DitWorkspace(nullptr, w, B, V, H, W).bytes;
}

int dgs_dit_forward(const dgs_dit_weights* w, const dgs_dit_io* io, void* workspace, size_t workspace_bytes,
                    void* stream) {
  DGS_REQUIRE(io != nullptr, "io is NULL");
  DGS_TRY(check_dit(w, io->B, io->V, io->H, io->W));
  DGS_REQUIRE(io->images && io->ray_o && io->ray_d && io->t, "NULL input");
  DGS_REQUIRE(io->xyz && io->features && io->scaling && io->rotation && io->opacity, "NULL output");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = io->B, V = io->V, H = io->H, W = io->W, D = w->width, L = w->layers, G = w->n_gaussians, p = w->patch;
  const int T = V * (H / p) * (W / p), N = T + G, Kin = p * p * 9, Ndec = p * p * 14;
  DitWorkspace ws(workspace, w, B, V, H, W);
  DGS_REQUIRE(workspace && workspace_bytes >= ws.bytes, "workspace too small: %zu < %zu", workspace_bytes, ws.bytes);
  const int mod_stride = L * 6 * D + 4 * D;
  const __nv_bfloat16* bf = nullptr;
  (void)bf;

  // ---- input stage: posed image -> tokens -> tokenizer GEMM -> [pos tokens | image tokens] -> LayerNorm(weight) ----
  if (g_prof_on) prof_begin(st, PROF_DIT_INPUT);
  DGS_TRY(posed_patchify(io->images, io->ray_o, io->ray_d, ws.tokens, B, V, H, W, p, io->plucker_mode, st));
  {
    GemmEpilogue ep;
    ep.out = ws.tok; ep.ldc = D;
    DGS_TRY(gemm_bf16(ws.tokens, w->tokenizer_w, B * T, D, 3 * Kin, EPI_F32, ep, st));  // split-bf16: K = 3*576
  }
  DGS_TRY(assemble_tokens(ws.tok, w->pos_embed, ws.x, B, G, T, D, st));
  DGS_TRY(ln_weight_inplace(ws.x, w->in_ln_w, B * N, D, 1e-5f, st));  // nn.LayerNorm default eps (denoiser.py:234-236)
  if (g_prof_on) { prof_end(st, PROF_DIT_INPUT); prof_begin(st, PROF_DIT_COND); }

  // ---- conditioning: timestep MLP, then the adaLN modulation of ALL blocks and both heads in one launch ----
  DGS_TRY(timestep_embedding(io->t, ws.temb0, B, 256, st));
  DGS_TRY(skinny_linear(ws.temb0, w->t0_w, w->t0_b, ws.temb1, B, D, 256, 0, 1, st));
  DGS_TRY(skinny_linear(ws.temb1, w->t2_w, w->t2_b, ws.c, B, D, D, 0, 0, st));
  DGS_TRY(skinny_linear(ws.c, w->adaln_w, w->adaln_b, ws.mod, B, mod_stride, D, 1, 0, st));
  if (g_prof_on) prof_end(st, PROF_DIT_COND);

  // ---- L x DiTBlock (utils_transformer.py:270-290) ----
  for (int l = 0; l < L; l++) {
    const float* m = ws.mod + (size_t)l * 6 * D;  // shift_msa | scale_msa | gate_msa | shift_mlp | scale_mlp | gate_mlp
    {
      ProfScope ps(st, PROF_DIT_LN);
      DGS_TRY(ln_modulate(ws.x, nullptr, m, m + D, mod_stride, ws.h, B, N, 0, N, D, 1e-6f, 0, st));
    }
    {
      ProfScope ps(st, PROF_DIT_GEMM_QKV);
      GemmEpilogue ep;
      ep.out = ws.qkv; ep.ldc = 3 * D; ep.bias = w->qkv_b + (size_t)l * 3 * D;
      DGS_TRY(gemm_bf16(ws.h, (const __nv_bfloat16*)w->qkv_w + (size_t)l * 3 * D * D, B * N, 3 * D, D, EPI_BIAS_BF16, ep, st));
    }
    {
      ProfScope ps(st, PROF_DIT_ATTN);
      DGS_TRY(attention_fwd(ws.qkv, ws.attn, B, N, w->heads, st));
    }
    {
      ProfScope ps(st, PROF_DIT_GEMM_PROJ);
      GemmEpilogue ep;
      ep.out = ws.x; ep.ldc = D; ep.bias = w->proj_b + (size_t)l * D;
      ep.gate = m + 2 * D; ep.gate_stride = mod_stride; ep.rows_per_sample = N;
      DGS_TRY(gemm_bf16(ws.attn, (const __nv_bfloat16*)w->proj_w + (size_t)l * D * D, B * N, D, D, EPI_GATE_RESID_F32, ep, st));
    }
    {
      ProfScope ps(st, PROF_DIT_LN);
      DGS_TRY(ln_modulate(ws.x, nullptr, m + 3 * D, m + 4 * D, mod_stride, ws.h, B, N, 0, N, D, 1e-6f, 0, st));
    }
    {
      ProfScope ps(st, PROF_DIT_GEMM_FC1);
      GemmEpilogue ep;
      ep.out = ws.u; ep.ldc = w->mlp_hidden; ep.bias = w->fc1_b + (size_t)l * w->mlp_hidden;
      DGS_TRY(gemm_bf16(ws.h, (const __nv_bfloat16*)w->fc1_w + (size_t)l * w->mlp_hidden * D, B * N, w->mlp_hidden, D,
                        EPI_BIAS_GELU_BF16, ep, st));
    }
    {
      ProfScope ps(st, PROF_DIT_GEMM_FC2);
      GemmEpilogue ep;
      ep.out = ws.x; ep.ldc = D; ep.bias = w->fc2_b + (size_t)l * D;
      ep.gate = m + 5 * D; ep.gate_stride = mod_stride; ep.rows_per_sample = N;
      DGS_TRY(gemm_bf16(ws.u, (const __nv_bfloat16*)w->fc2_w + (size_t)l * D * w->mlp_hidden, B * N, D, w->mlp_hidden,
                        EPI_GATE_RESID_F32, ep, st));
    }
  }
  if (io->tokens_out)
    DGS_CUDA_OK(cudaMemcpyAsync(io->tokens_out, ws.x, (size_t)B * N * D * sizeof(float), cudaMemcpyDeviceToDevice, st));

  // ---- heads (denoiser.py:76-164): LN(weight) + modulate + Linear ----
  ProfScope ps_heads(st, PROF_DIT_HEADS);
  const float* mu = ws.mod + (size_t)L * 6 * D;  // upsampler: shift | scale
  const float* md = mu + 2 * D;                  // image_token_decoder: shift | scale
  if (G > 0) {
    DGS_TRY(ln_modulate(ws.x, w->ups_ln_w, mu, mu + D, mod_stride, ws.hg, B, N, 0, G, D, 1e-5f, 1, st));
    DGS_TRY(tiny_linear_bf16(ws.hg, (const __nv_bfloat16*)w->ups_w, ws.gs_tok, B * G, 14, 3 * D, st));
  }
  // the decoder head runs split-bf16 (K = 3*width) so the Gaussian parameters are fp32-accurate functions of the
  // residual stream; its A operand re-uses the (now free) MLP hidden buffer
  __nv_bfloat16* hdec = ws.u;
  DGS_TRY(ln_modulate(ws.x, w->dec_ln_w, md, md + D, mod_stride, hdec, B, N, G, T, D, 1e-5f, 1, st));
  {
    GemmEpilogue ep;
    ep.out = ws.img_gs; ep.ldc = Ndec;
    DGS_TRY(gemm_bf16(hdec, w->dec_w, B * T, Ndec, 3 * D, EPI_F32, ep, st));
  }
  GsOut go;
  go.xyz = io->xyz; go.features = io->features; go.scaling = io->scaling; go.rotation = io->rotation;
  go.opacity = io->opacity; go.img_aligned_xyz = io->img_aligned_xyz;
  DGS_TRY(gaussians_epilogue(ws.gs_tok, ws.img_gs, io->ray_o, io->ray_d, go, B, G, V, H, W, p, io->scene_depth,
                             io->range_near, io->range_far, st));
  return DGS_OK;
}

int dgs_gemm_bf16(const void* A, const void* Wt, const float* bias, const float* gate, void* out, int M, int N, int K,
                  int epi, int ldc, int gate_stride, int rows_per_sample, void* stream) {
  DGS_REQUIRE(A && Wt && out, "NULL pointer");
  DGS_REQUIRE(epi != EPI_GATE_RESID_F32 || (gate && rows_per_sample > 0), "gate epilogue needs gate and rows_per_sample");
  GemmEpilogue ep;
  ep.out = out; ep.ldc = ldc; ep.bias = bias; ep.gate = gate; ep.gate_stride = gate_stride;
  ep.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : 1;
  return gemm_bf16(A, Wt, M, N, K, epi, ep, (cudaStream_t)stream);
}

int dgs_attention_fwd(const void* qkv, void* out, int B, int N, int heads, void* stream) {
  DGS_REQUIRE(qkv && out, "NULL pointer");
  return attention_fwd(qkv, out, B, N, heads, (cudaStream_t)stream);
}

int dgs_ln_modulate(const float* x, const float* ln_w, const float* shift, const float* scale, int mod_stride, void* h,
                    int B, int rows, int width, float eps, void* stream) {
  DGS_REQUIRE(x && shift && scale && h, "NULL pointer");
  return ln_modulate(x, ln_w, shift, scale, mod_stride, (__nv_bfloat16*)h, B, rows, 0, rows, width, eps, 0,
                     (cudaStream_t)stream);
}

}  // extern "C"
