// diffusion_steps.cu -- the elementwise callers either side of the hot path:
//   rays_from_cameras : TransformInput            (diffusionGS/systems/utils.py:621-757, patch_size=None path)
//   q_sample          : GaussianDiffusion.q_sample (diffusionGS/models/diffusion/gaussian_diffusion.py:268-284)
//   p_sample_step     : q_posterior_mean_variance + the ancestral step of p_sample, x0-prediction, FIXED_LARGE
//                       variance (gaussian_diffusion.py:291-312, 380-392, 505-516)
// The schedule tables stay resident on the device (the reference re-uploads its fp64 numpy table on every call,
// _extract_into_tensor, gaussian_diffusion.py:853-865).  All HBM-bound, float4-vectorised where aligned.
#include "dgs_internal.h"

namespace dgs {

__global__ void __launch_bounds__(256) rays_kernel(const float* __restrict__ c2w, const float* __restrict__ fxfycxcy,
                                                   int n_views, int H, int W, float* __restrict__ ray_o,
                                                   float* __restrict__ ray_d) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long plane = (long long)H * W;
  if (idx >= (long long)n_views * plane) return;
  const int v = (int)(idx / plane);
  const int pix = (int)(idx - (long long)v * plane);
  const int y = pix / W, x = pix - y * W;
  const float* m = c2w + 16 * v;
  const float fx = fxfycxcy[4 * v], fy = fxfycxcy[4 * v + 1], cx = fxfycxcy[4 * v + 2], cy = fxfycxcy[4 * v + 3];
  const float dx = ((float)x + 0.5f - cx) / fx, dy = ((float)y + 0.5f - cy) / fy, dz = 1.0f;
  // ray_d = d_cam @ R^T  (world = R d_cam)
  float wx = dx * m[0] + dy * m[1] + dz * m[2];
  float wy = dx * m[4] + dy * m[5] + dz * m[6];
  float wz = dx * m[8] + dy * m[9] + dz * m[10];
  const float inv = 1.0f / sqrtf(wx * wx + wy * wy + wz * wz);
  float* od = ray_d + (size_t)v * 3 * plane + pix;
  float* oo = ray_o + (size_t)v * 3 * plane + pix;
  od[0] = wx * inv; od[plane] = wy * inv; od[2 * plane] = wz * inv;
  oo[0] = m[3]; oo[plane] = m[7]; oo[2 * plane] = m[11];
}

// out = a[t[b]] * x + c[t[b]] * y (+ optional (t != 0) * exp(0.5 * lv[t[b]]) * z)
__global__ void __launch_bounds__(256) schedule_axpby_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                             const float* __restrict__ z, const float* __restrict__ ta,
                                                             const float* __restrict__ tc, const float* __restrict__ tlv,
                                                             const long long* __restrict__ t, long long per_sample,
                                                             long long total, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long tb = t[i / per_sample];
  float r = ta[tb] * x[i] + tc[tb] * y[i];
  if (z) r += (tb != 0 ? 1.0f : 0.0f) * expf(0.5f * tlv[tb]) * z[i];
  out[i] = r;
}

}  // namespace dgs

using namespace dgs;

extern "C" {

int dgs_rays_from_cameras(const float* c2w, const float* fxfycxcy, int n_views, int H, int W, float* ray_o,
                          float* ray_d, void* stream) {
  DGS_REQUIRE(c2w && fxfycxcy && ray_o && ray_d && n_views > 0 && H > 0 && W > 0, "rays: bad arguments");
  const long long total = (long long)n_views * H * W;
  rays_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c2w, fxfycxcy, n_views, H, W, ray_o, ray_d);
  DGS_POST_LAUNCH();
  return DGS_OK;
}

int dgs_q_sample(const float* x_start, const float* noise, const float* sqrt_alphas_cumprod,
                 const float* sqrt_one_minus_alphas_cumprod, const long long* t, int B, long long per_sample,
                 float* out, void* stream) {
  DGS_REQUIRE(x_start && noise && sqrt_alphas_cumprod && sqrt_one_minus_alphas_cumprod && t && out && B > 0 && per_sample > 0,
              "q_sample: bad arguments");
  const long long total = (long long)B * per_sample;
  schedule_axpby_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x_start, noise, nullptr, sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod, nullptr, t, per_sample, total, out);
  DGS_POST_LAUNCH();
  return DGS_OK;
}

int dgs_p_sample_step(const float* pred_xstart, const float* x_t, const float* noise, const float* posterior_mean_coef1,
                      const float* posterior_mean_coef2, const float* model_log_variance, const long long* t, int B,
                      long long per_sample, float* out, void* stream) {
  DGS_REQUIRE(pred_xstart && x_t && noise && posterior_mean_coef1 && posterior_mean_coef2 && model_log_variance && t && out &&
                  B > 0 && per_sample > 0, "p_sample_step: bad arguments");
  const long long total = (long long)B * per_sample;
  schedule_axpby_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      pred_xstart, x_t, noise, posterior_mean_coef1, posterior_mean_coef2, model_log_variance, t, per_sample, total, out);
  DGS_POST_LAUNCH();
  return DGS_OK;
}

}  // extern "C"
