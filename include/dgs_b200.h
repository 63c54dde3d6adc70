/*
 * dgs_b200.h -- C ABI of libdgs_b200.so: the B200-native (sm_100a) hot path of Open-DiffusionGS.
 *
 * Plain C: raw DEVICE pointers, sizes, a cudaStream_t passed as void*.  No torch types, no C++
 * exceptions cross this boundary; every entry point returns a status code and
 * dgs_last_error() gives the message.  Each entry point cites the reference interface it
 * replaces (paths relative to the reference repo; DGR = submodules/diff-gaussian-rasterization).
 *
 * All kernels are enqueued on the caller's stream.  Only the rasterizer forward synchronises
 * that stream once (to read the instance count R that sizes the binning arena), exactly like
 * the reference's cudaMemcpy at DGR/cuda_rasterizer/rasterizer_impl.cu:281 -- but ONCE PER BATCH
 * of (sample, view) pairs instead of once per view.
 */
#ifndef DGS_B200_H_INCLUDED
#define DGS_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGS_VERSION 100

typedef enum {
  DGS_OK = 0,
  DGS_ERR_INVALID_ARGUMENT = 1, /* replaces AT_ERROR / std::runtime_error, rasterize_points.cu:57-59 */
  DGS_ERR_CUDA = 2,             /* replaces CHECK_CUDA's throw, auxiliary.h:166-173 */
  DGS_ERR_ALLOC = 3,            /* allocator callback returned NULL */
  DGS_ERR_OVERFLOW = 4          /* instance count does not fit 32 bits */
} dgs_status;

/* Arena allocator callback: must return DEVICE memory of at least `bytes` bytes (256-B aligned)
 * that stays valid until the matching backward has run.  Replaces the reference's
 * std::function<char*(size_t)> resize callbacks (DGR/cuda_rasterizer/rasterizer.h:31-34,
 * DGR/rasterize_points.cu:27-33). */
typedef void* (*dgs_alloc_fn)(size_t bytes, void* user);

int dgs_version(void);
const char* dgs_last_error(void);
/* Kernels of this library launched so far by the calling process (for bench.py's `gpu_launches`). */
unsigned long long dgs_kernel_launch_count(void);
/* Optional per-kernel-family device timing: when enabled, every entry point records CUDA events on the
 * caller's stream around each kernel family it launches.  dgs_profile_read() waits for the recorded spans,
 * writes the summed milliseconds and span counts per family (order: raster project, scan, emit_keys, sort,
 * tile_ranges, blend_fwd, blend_bwd, geometry_bwd; dit input, conditioning, ln_modulate, gemm_qkv, attention,
 * gemm_proj, gemm_fc1, gemm_fc2, heads), clears the record and returns the number of families. */
int dgs_profile_enable(int on);
int dgs_profile_read(float* ms_sum, int* span_count, int n_families);

/* ------------------------------------------------------------------------------------------------
 * B1a. Single-view rasterizer: what `_C.rasterize_gaussians`, `_C.rasterize_gaussians_backward`
 * and `_C.mark_visible` bind (DGR/ext.cpp:15-19; DGR/rasterize_points.h:18-66;
 * CudaRasterizer::Rasterizer::{forward,backward,markVisible}, DGR/cuda_rasterizer/rasterizer.h:24-85).
 * Inputs are the ACTIVATED tensors the reference binding receives (post-exp scales, normalised
 * quaternions, post-sigmoid opacities).  NULL == "not provided" (the reference's empty tensor).
 * viewmatrix / projmatrix are the 16 floats of the reference's transposed matrices
 * (element [4*c + r] = row r, column c; DGR/cuda_rasterizer/auxiliary.h:58-77).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int P;            /* number of Gaussians                     */
  int D;            /* active SH degree (0..3)                 */
  int M;            /* SH coefficients per Gaussian in `shs`   */
  int W, H;         /* image size                              */
  const float* background;     /* [3]      */
  const float* means3D;        /* [P,3]    */
  const float* shs;            /* [P,M,3] or NULL */
  const float* colors_precomp; /* [P,3]   or NULL */
  const float* opacities;      /* [P]      */
  const float* scales;         /* [P,3]   or NULL */
  const float* rotations;      /* [P,4]   or NULL */
  const float* cov3D_precomp;  /* [P,6]   or NULL */
  const float* viewmatrix;     /* [16]     */
  const float* projmatrix;     /* [16]     */
  const float* campos;         /* [3]      */
  float scale_modifier;
  float tan_fovx, tan_fovy;
  int prefiltered;  /* accepted for signature parity; culled points are skipped either way */
  int debug;        /* non-zero: synchronise and check for CUDA errors after every stage */
} dgs_raster_args;

/* Sizes of the three opaque arenas (geometry: per Gaussian, binning: per instance, image: per pixel)
 * -- the equivalents of required<GeometryState/BinningState/ImageState>() (rasterizer_impl.h:67-73). */
size_t dgs_raster_geom_bytes(int n_views, int P);
size_t dgs_raster_binning_bytes(long long R);
size_t dgs_raster_image_bytes(int n_views, int W, int H);

/* Forward (rasterizer_impl.cu:198-336).  out_color [3,H,W] and radii [P] are caller-allocated;
 * the three arenas are obtained through the callbacks and must be handed back to the backward.
 * *num_rendered receives R (the reference's return value). */
int dgs_raster_forward(const dgs_raster_args* args, dgs_alloc_fn geom_alloc, void* geom_user,
                       dgs_alloc_fn binning_alloc, void* binning_user, dgs_alloc_fn image_alloc,
                       void* image_user, float* out_color, int* radii, int* num_rendered,
                       void* stream);

/* Backward (rasterizer_impl.cu:340-434).  The nine gradient buffers are caller-allocated and
 * ZERO-initialised (rasterize_points.cu:151-159): dL_dmean2D [P,3], dL_dconic [P,2,2],
 * dL_dopacity [P,1], dL_dcolor [P,3], dL_dmean3D [P,3], dL_dcov3D [P,6], dL_dsh [P,M,3],
 * dL_dscale [P,3], dL_drot [P,4]. */
int dgs_raster_backward(const dgs_raster_args* args, int R, const int* radii, const void* geom_buffer,
                        const void* binning_buffer, const void* image_buffer, const float* dL_dpix,
                        float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                        float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                        float* dL_drot, void* stream);

/* markVisible (rasterizer_impl.cu:54-66,141-153): present[i] = view-space z > 0.2 */
int dgs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream);

/* ------------------------------------------------------------------------------------------------
 * B1b. Batched renderer: one call for ALL (sample, view) pairs of `Renderer.forward`
 * (diffusionGS/models/gsrenderer/renderer.py:34-92) = DeferredGaussianRender.forward/backward
 * (gs_core.py:949-1060) + render_opencv_cam (874-945) + Camera (277-316) + the GaussianModel
 * activations (330-334,545-570), fused.  Inputs are the RAW per-sample parameter tensors
 * (fp32, contiguous): xyz [B,P,3], features [B,P,M,3], scaling [B,P,3], rotation [B,P,4],
 * opacity [B,P,1], C2W [B,V,4,4] row-major, fxfycxcy [B,V,4].  Output [B,V,3,H,W].
 * The forward keeps its sorted tile lists in the arenas; the backward re-uses them (no re-render,
 * unlike gs_core.py:1041-1056) and returns gradients w.r.t. the raw tensors, summed over views.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int B, V, P, M, D, W, H;
  const float* xyz;
  const float* features;
  const float* scaling;
  const float* rotation;
  const float* opacity;
  const float* c2w;
  const float* fxfycxcy;
  float scale_modifier; /* 1.0 when the reference passes scaling_modifier=None */
  float bg[3];          /* (1,1,1) in render_opencv_cam, gs_core.py:880 */
  int debug;
  int near_log2;        /* 0: one binning pass over all instances.  k < 0: adaptive (1/8, or 1/16 when those near lists still
                           average >= 2048 entries per tile).  k > 0: two-phase binning -- phase A bins and blends
                           only the nearest P >> k Gaussians of every view; if every pixel saturates there (dense scenes)
                           the remaining instances are never emitted or sorted, otherwise phase B continues from the
                           saved per-pixel state.  Same images, final_T, n_contrib and gradients either way. */
} dgs_render_batch_args;

int dgs_render_batch_forward(const dgs_render_batch_args* args, dgs_alloc_fn geom_alloc, void* geom_user,
                             dgs_alloc_fn binning_alloc, void* binning_user, dgs_alloc_fn image_alloc,
                             void* image_user, float* out_images, long long* num_rendered,
                             long long* chunk_instances /* out [2]: instances binned in phase A / phase B */,
                             void* stream);

/* d_* are caller-allocated, same shapes as the inputs; they are fully overwritten.  scratch_alloc
 * provides the per-(view, Gaussian) screen-space gradient records (44 B each), free after the call. */
int dgs_render_batch_backward(const dgs_render_batch_args* args, long long R,
                              const long long* chunk_instances /* [2] from the forward */, const void* geom_buffer,
                              const void* binning_buffer /* 1st binning_alloc result */,
                              const void* binning_buffer_b /* 2nd (phase B) or NULL */, const void* image_buffer,
                              const float* dL_dimages, float* d_xyz, float* d_features,
                              float* d_scaling, float* d_rotation, float* d_opacity,
                              dgs_alloc_fn scratch_alloc, void* scratch_user, void* stream);

/* The same pair with the image-space MSE of the training loss fused in (LossComputer.forward's l2 term,
 * diffusionGS/utils/losses.py:261-284: l2_loss[b] = mean over (v,3,h,w) of (rendering - target)^2, later averaged over b and
 * weighted by lambda_diffusion, diffusion_gs_system.py:94-124).  Forward: the blend kernel adds sum (render - target)^2 of
 * sample b into loss_sum[b] (caller-zeroed, fp64) while the pixel is in registers.  Backward: the blend-backward kernel forms
 * dL/dpix = dL_dimages (may be NULL) + coef[b] * (render - target) itself, with coef[b] = dL/dl2_loss[b] * 2 / (V*3*H*W)
 * computed by the caller on the device -- no per-element loss or gradient image exists.  mse == NULL: the plain pair. */
typedef struct {
  const float* target;   /* [B,V,target_channels,H,W] fp32 in (0,1)                                       */
  int target_channels;   /* 3, or 4 = rgb + mask (the mask plane is skipped, losses.py:274-276)           */
  double* loss_sum;      /* forward:  out [B]                                                              */
  const float* coef;     /* backward: in  [B] (device)                                                     */
  const float* images;   /* backward: in  the forward's out_images                                         */
} dgs_render_mse;
int dgs_render_batch_forward_mse(const dgs_render_batch_args* args, dgs_alloc_fn geom_alloc, void* geom_user,
                                 dgs_alloc_fn binning_alloc, void* binning_user, dgs_alloc_fn image_alloc,
                                 void* image_user, float* out_images, long long* num_rendered, long long* chunk_instances,
                                 const dgs_render_mse* mse, void* stream);
int dgs_render_batch_backward_mse(const dgs_render_batch_args* args, long long R, const long long* chunk_instances,
                                  const void* geom_buffer, const void* binning_buffer, const void* binning_buffer_b,
                                  const void* image_buffer, const float* dL_dimages, const dgs_render_mse* mse,
                                  float* d_xyz, float* d_features, float* d_scaling, float* d_rotation, float* d_opacity,
                                  dgs_alloc_fn scratch_alloc, void* scratch_user, void* stream);

/* Introspection used by the parity tests: copies of per-(view, Gaussian) / per-pixel forward state
 * out of the opaque arenas into caller DEVICE buffers (any may be NULL):
 * xy [N,2], depth [N], conic_opacity [N,4], rgb [N,3], tiles_touched [N] (N = n_views*P),
 * point_list [R], ranges [n_views*tiles,2], final_T / n_contrib [n_views*H*W].  point_list / ranges describe a
 * single-pass binning (near_log2 == 0); the per-pixel and per-Gaussian outputs are valid in both modes. */
int dgs_raster_export_state(int n_views, int P, int W, int H, long long R, const void* geom_buffer,
                            const void* binning_buffer, const void* image_buffer, float* xy,
                            float* depth, float* conic_opacity, float* rgb, uint32_t* tiles_touched,
                            uint32_t* point_list, uint32_t* ranges, float* final_T,
                            uint32_t* n_contrib, void* stream);

/* ------------------------------------------------------------------------------------------------
 * B2. DiT denoiser forward: DGSDenoiser.image_to_gaussians
 * (diffusionGS/models/denoiser/denoiser.py:306-416; scene twin denoiser_scene.py:314-429), i.e.
 * posed-image patchify + tokenizer -> +2 learned tokens -> LayerNorm -> L x DiTBlock
 * (utils_transformer.py:246-290 around timm Attention/Mlp) -> GaussiansUpsampler / ImageTokenDecoder
 * heads -> to_gs + pixel alignment.  GEMMs and attention run on tcgen05 (bf16 in, fp32 accumulate),
 * the residual stream, LayerNorm statistics and softmax stay fp32.
 * Weights: GEMM matrices are bf16 row-major [out, in] (nn.Linear layout), per-layer tensors stacked
 * along a leading L axis; vectors fp32.  state_dict key -> field mapping is in dgs_b200/denoiser.py.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int width, heads, layers, patch, n_gaussians, mlp_hidden; /* 1024, 16, 24, 8, 2, 4096 */
  const void* tokenizer_w;  /* image_tokenizer.1.weight, split-bf16 [width, 3*patch*patch*9] = [hi|hi|lo] */
  const float* pos_embed;   /* gaussians_pos_embedding             fp32 [n_gaussians, width]          */
  const float* in_ln_w;     /* transformer_input_layernorm.weight  fp32 [width]                       */
  const float* t0_w; const float* t0_b; /* t_embedder.mlp.0  fp32 [width,256], fp32 [width]          */
  const float* t2_w; const float* t2_b; /* t_embedder.mlp.2  fp32 [width,width], fp32 [width]        */
  const float* adaln_w;     /* all adaLN_modulation.1.weight stacked: L x [6*width] rows, then upsampler
                               [2*width], then image_token_decoder [2*width]; fp32 [.., width]        */
  const float* adaln_b;     /* same stacking, fp32                                                    */
  const void* qkv_w;  const float* qkv_b;   /* transformer.{i}.attn.qkv   bf16 [L,3w,w], fp32 [L,3w] */
  const void* proj_w; const float* proj_b;  /* transformer.{i}.attn.proj  bf16 [L,w,w],  fp32 [L,w]  */
  const void* fc1_w;  const float* fc1_b;   /* transformer.{i}.mlp.fc1    bf16 [L,4w,w], fp32 [L,4w] */
  const void* fc2_w;  const float* fc2_b;   /* transformer.{i}.mlp.fc2    bf16 [L,w,4w], fp32 [L,w]  */
  const float* ups_ln_w; const void* ups_w; /* upsampler.layernorm.weight fp32 [w]; upsampler.linear.weight split-bf16 [14,3w] */
  const float* dec_ln_w; const void* dec_w; /* image_token_decoder.*: fp32 [w]; split-bf16 [patch*patch*14, 3w]
     "split-bf16 [n, 3k] = [hi|hi|lo]": hi = bf16(W), lo = bf16(W - hi); paired with activations laid out
     [hi|lo|hi] the bf16 MMA then yields x_hi W_hi + x_lo W_hi + x_hi W_lo (fp32-accurate) -- used for the two
     small GEMMs at the ends of the network whose rounding would otherwise dominate the output error. */
} dgs_dit_weights;

typedef struct {
  int B, V, H, W;
  int plucker_mode;          /* 0 = 'relative_plk' (object model), 1 = 'plk' (scene model)            */
  int scene_depth;           /* 0: depth=(2s-1)*1.8 + (-o.d) (object model, 'relative_plk'); 1: depth=s*(far-near)+near
                                (scene model); 2: depth=s (object model with ray_pe_type 'plk', denoiser.py:381-388) */
  float range_near, range_far;
  const float* images;       /* [B,V,3,H,W] fp32 (view 0 clean, others noised)                        */
  const float* ray_o;        /* [B,V,3,H,W]                                                           */
  const float* ray_d;        /* [B,V,3,H,W]                                                           */
  const float* t;            /* [B] timesteps as fp32                                                 */
  float* xyz;                /* out [B,P,3],  P = n_gaussians + V*H*W                                 */
  float* features;           /* out [B,P,1,3]                                                         */
  float* scaling;            /* out [B,P,3]                                                           */
  float* rotation;           /* out [B,P,4]                                                           */
  float* opacity;            /* out [B,P,1]                                                           */
  float* img_aligned_xyz;    /* out [B,V,3,H,W] or NULL                                               */
  float* tokens_out;         /* optional debug out: final residual stream [B,N,width] fp32, or NULL   */
  void* train_state;         /* NULL: inference.  Else a caller buffer of dgs_dit_train_state_bytes[_ex](): what the
                                backward needs from the forward lives there; it must stay untouched, together with
                                `workspace`, until dgs_dit_backward has run.                              */
  int train_mode;            /* DGS_TRAIN_STORE (0): every activation is kept (4 GB per sample at N=4098; forward + 2x
                                backward FLOPs).  DGS_TRAIN_RECOMPUTE (1): only the fp32 residual stream entering each
                                block is kept (0.6 GB per sample) and the backward re-runs each block's forward before
                                differentiating it -- the reference's torch.utils.checkpoint around every block
                                (denoiser.py:348-354; grad_checkpoint_every = 1), for the yaml batch sizes.           */
} dgs_dit_io;
#define DGS_TRAIN_STORE 0
#define DGS_TRAIN_RECOMPUTE 1

size_t dgs_dit_workspace_bytes(const dgs_dit_weights* w, int B, int V, int H, int W);
int dgs_dit_forward(const dgs_dit_weights* w, const dgs_dit_io* io, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ---- training: backward of dgs_dit_forward (what torch autograd derives for denoiser.py:306-416 in the reference) ----
 * dgrad GEMMs read TRANSPOSED bf16 copies of the block weights (W^T, [in, out] row-major) so that every GEMM of the
 * backward runs on the same K-major tcgen05 kernel as the forward; the caller refreshes them after each optimizer step
 * (dgs_transpose_bf16).  Gradients are fp32, in the state_dict layout of the fp32 master parameters. */
typedef struct {
  const void* qkv_wT;   /* bf16 [L, w, 3w]   */
  const void* proj_wT;  /* bf16 [L, w, w]    */
  const void* fc1_wT;   /* bf16 [L, w, 4w]   */
  const void* fc2_wT;   /* bf16 [L, 4w, w]   */
  const void* dec_wT;   /* bf16 [w, patch*patch*14]  (image_token_decoder.linear.weight^T)             */
  const float* ups_w;   /* fp32 [14, w]      (upsampler.linear.weight, master copy)                    */
} dgs_dit_weights_t;

typedef struct {  /* all fp32, OVERWRITTEN by dgs_dit_backward */
  float* tokenizer_w;  /* [w, patch*patch*9] */
  float* pos_embed;    /* [n_gaussians, w]   */
  float* in_ln_w;      /* [w]                */
  float* t0_w; float* t0_b; float* t2_w; float* t2_b;
  /* per-block tensors: the pointer addresses block 0; block l lives layer_stride floats further.  This is the layout
     of a flat gradient arena in module.parameters() order (dgs_b200/dist.py GradArena: one contiguous bucket per
     block, all-reduced in reverse order as the backward finishes them); a stacked [L, ...] layout is not expressible */
  long long layer_stride;
  float* qkv_w; float* qkv_b; float* proj_w; float* proj_b; float* fc1_w; float* fc1_b; float* fc2_w; float* fc2_b;
  float* adaln_w; float* adaln_b;          /* block 0: [6w, w], [6w] */
  float* ups_ln_w; float* ups_w;           /* [w], [14, w]               */
  float* ups_adaln_w; float* ups_adaln_b;  /* [2w, w], [2w]              */
  float* dec_ln_w; float* dec_w;           /* [w], [patch*patch*14, w]   */
  float* dec_adaln_w; float* dec_adaln_b;  /* [2w, w], [2w]              */
} dgs_dit_grads;

typedef struct {  /* gradients w.r.t. the outputs of dgs_dit_forward (what dgs_render_batch_backward returns) */
  const float* d_xyz; const float* d_features; const float* d_scaling; const float* d_rotation; const float* d_opacity;
} dgs_dit_out_grads;

size_t dgs_dit_train_state_bytes(const dgs_dit_weights* w, int B, int V, int H, int W); /* DGS_TRAIN_STORE */
size_t dgs_dit_train_state_bytes_ex(const dgs_dit_weights* w, int B, int V, int H, int W, int train_mode);
/* io / workspace / io->train_state: exactly what the matching dgs_dit_forward call was given. */
int dgs_dit_backward(const dgs_dit_weights* w, const dgs_dit_weights_t* wT, const dgs_dit_io* io,
                     const dgs_dit_out_grads* dout, const dgs_dit_grads* grads, void* workspace, size_t workspace_bytes,
                     void* stream);
/* Same, with a completion hook for the data-parallel gradient exchange (what torch DDP's autograd hooks give the
 * reference, diffusionGS_rel.yaml:80): block_done[l], l = layers-1 .. 0, is recorded on `stream` as soon as EVERY
 * parameter gradient of transformer block l is final (the blocks are differentiated in reverse order), block_done[layers]
 * when the remaining gradients (tokenizer, embedder, heads) are.  A side stream that waits on block_done[l]
 * (dgs_stream_wait_event) can all-reduce block l's contiguous bucket while blocks l-1 .. 0 are still running.
 * Entries may be NULL; events come from dgs_event_create. */
typedef struct {
  void** block_done;   /* NULL or [layers + 1] events */
} dgs_dit_bwd_opts;
int dgs_dit_backward_ex(const dgs_dit_weights* w, const dgs_dit_weights_t* wT, const dgs_dit_io* io,
                        const dgs_dit_out_grads* dout, const dgs_dit_grads* grads, const dgs_dit_bwd_opts* opts,
                        void* workspace, size_t workspace_bytes, void* stream);
int dgs_event_create(void** event);                 /* cudaEventCreateWithFlags(DisableTiming) */
int dgs_event_destroy(void* event);
int dgs_stream_wait_event(void* stream, void* event);
/* out[c, m] = bf16(in[m, c]) for in [M, C] (fp32 if in_is_f32 else bf16), out [C, round_up(M, 64)] zero padded;
 * colsum (optional, fp32 [C]) += column sums.  Used for the transposed weight copies and inside the backward. */
int dgs_transpose_bf16(const void* in, int in_is_f32, int M, int C, void* out, float* colsum, void* stream);
/* fused AdamW on fp32 master parameters (torch.optim.AdamW semantics; diffusionGS_rel.yaml:57-62), step >= 1.
 * The gradient is multiplied by grad_scale * (grad_scale_dev ? *grad_scale_dev : 1): the clip factor stays on the device */
int dgs_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, float grad_scale, const float* grad_scale_dev,
                   void* stream);
/* Same update with the exponential moving average of the parameters folded in (the reference's EMA callback,
 * diffusionGS/utils/ema.py:82-101 with decay 0.9999, launch.py:227): after the AdamW update of p,
 * ema = ema_decay * ema + (1 - ema_decay) * p, in the same pass over the arena.  ema == NULL: plain dgs_adamw_step. */
int dgs_adamw_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, size_t n, float lr,
                       float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                       const float* grad_scale_dev, float ema_decay, void* stream);
/* weight refresh after an optimizer step: `batch` fp32 matrices [M, C] (matrix i at in + i*in_batch_stride floats) ->
 * bf16 copies [batch, M, C] (optional) and transposed bf16 copies [batch, C, M]; M, C multiples of 64 */
int dgs_cast_transpose_f32(const float* in, long long in_batch_stride, int batch, int M, int C, void* out_bf16,
                           void* outT_bf16, void* stream);

/* Building blocks, exported for the unit parity tests (same kernels dgs_dit_forward launches).
 * epi: 0 = bias -> bf16, 1 = bias + GELU(tanh) -> bf16, 2 = out(fp32) += gate[row / rows_per_sample] * (acc + bias),
 *      3 = (acc + bias) -> fp32.   A [M,K], W [N,K] bf16 row-major. */
int dgs_gemm_bf16(const void* A, const void* W, const float* bias, const float* gate, void* out, int M, int N,
                  int K, int epi, int ldc, int gate_stride, int rows_per_sample, void* stream);
/* qkv [B,N,3,heads,64] bf16 -> out [B,N,heads*64] bf16 = softmax(q k^T / 8) v */
int dgs_attention_fwd(const void* qkv, void* out, int B, int N, int heads, void* stream);
/* training pair: the forward also writes lse2 [B, heads, round_up(N,128)] fp32; the backward turns (qkv, out, lse2,
 * dout [B,N,heads*64] bf16) into dqkv [B,N,3,heads,64] bf16; dsum = scratch of the lse2 size. */
int dgs_attention_fwd_train(const void* qkv, void* out, float* lse2, int B, int N, int heads, void* stream);
int dgs_attention_bwd(const void* qkv, const void* out, const void* dout, float* lse2, float* dsum, void* dqkv, int B,
                      int N, int heads, void* stream);
/* extended GEMM entry (training epilogues): epi 4 = out(bf16) = acc * gelu'(aux);  aux (bf16 [M,ldc]): epi 1/2 also
 * store acc + bias there;  resid: residual source of epi 2 (NULL = in place);  lda/ldb: operand row strides (0 = K) */
int dgs_gemm_bf16_ex(const void* A, const void* W, const float* bias, const float* gate, void* out, void* aux,
                     const float* resid, int M, int N, int K, int lda, int ldb, int epi, int ldc, int gate_stride,
                     int rows_per_sample, void* stream);
/* out[M,N] (fp32, row stride ldc) = A^T W for A [K,M], W [K,N] bf16 row-major with row strides lda / ldb (0 = M / N):
 * the weight-gradient GEMM (K = tokens) on MN-major tcgen05 operands -- no transposed copies */
int dgs_gemm_bf16_tn(const void* A, const void* W, float* out, int M, int N, int K, int lda, int ldb, int ldc, void* stream);
/* backward of dgs_ln_modulate: dx (+)= ..., dshift/dscale [B, mod_stride] += ..., dln_w += ... (NULL where absent);
 * stats = scratch of 2*B*rows floats (per-row mean / rstd handed from the row kernel to the column kernel) */
int dgs_ln_modulate_bwd(const float* x, const void* dh, int dh_is_f32, const float* ln_w, const float* scale,
                        int mod_stride, int B, int rows, int width, float eps, float* dx, int accumulate, float* dshift,
                        float* dscale, float* dln_w, float* stats, void* stream);
/* backward of x_out = x_in + gate[b] * y: dy (bf16 [M,C]), dyT (bf16 [C, round_up(M,64)]), dgate += , dbias += */
int dgs_gate_bwd(const float* dx, const void* y, const float* gate, int gate_stride, int rows_per_sample, int M, int C,
                 void* dy, void* dyT, float* dgate, float* dbias, void* stream);
/* h = (LN(x; eps) [* ln_w]) * (1 + scale[b]) + shift[b] -> bf16 ; x fp32 [B, rows, width] */
int dgs_ln_modulate(const float* x, const float* ln_w, const float* shift, const float* scale, int mod_stride,
                    void* h, int B, int rows, int width, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * B3. The elementwise callers either side of the path.
 * ---------------------------------------------------------------------------------------------- */
/* TransformInput (diffusionGS/systems/utils.py:621-757, patch_size=None): per-pixel world-space rays.
 * c2w [n_views,4,4] row-major, fxfycxcy [n_views,4] -> ray_o, ray_d [n_views,3,H,W] (ray_d normalised). */
int dgs_rays_from_cameras(const float* c2w, const float* fxfycxcy, int n_views, int H, int W, float* ray_o,
                          float* ray_d, void* stream);
/* GaussianDiffusion.q_sample (gaussian_diffusion.py:268-284): out = sqrt_ac[t[b]]*x_start + sqrt_1mac[t[b]]*noise.
 * Tables are DEVICE fp32 arrays indexed by timestep; t is int64 [B]; per_sample = elements per batch item. */
int dgs_q_sample(const float* x_start, const float* noise, const float* sqrt_alphas_cumprod,
                 const float* sqrt_one_minus_alphas_cumprod, const long long* t, int B, long long per_sample,
                 float* out, void* stream);
/* One ancestral step of p_sample with x0-prediction and FIXED_LARGE variance (gaussian_diffusion.py:291-312,
 * 380-392, 505-516): out = coef1[t]*pred_xstart + coef2[t]*x_t + (t != 0) * exp(0.5*log_var[t]) * noise. */
int dgs_p_sample_step(const float* pred_xstart, const float* x_t, const float* noise, const float* posterior_mean_coef1,
                      const float* posterior_mean_coef2, const float* model_log_variance, const long long* t, int B,
                      long long per_sample, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DGS_B200_H_INCLUDED */
