"""Host-side mirror of the reference's `Renderer` (diffusionGS/models/gsrenderer/renderer.py:21-92)
on top of the batched sm_100a rasterizer.

`Renderer.forward(xyz, features, scaling, rotation, opacity, height, width, C2W, fxfycxcy, deferred=True)`
keeps the reference signature and returns [b, v, 3, H, W] fp32.  Where the reference loops over
(sample, view) in Python and RE-RENDERS every view in the backward (gs_core.py:990-1001, 1041-1056),
this issues one batched launch set and keeps the sorted tile lists for the backward.
"""
import copy

import torch
import torch.nn as nn

from . import raster as _raster


class BatchedGaussianRender(torch.autograd.Function):
    """Replaces DeferredGaussianRender (gs_core.py:949-1060); same inputs, same gradient outputs
    (sum over views of d/d{xyz, features, scaling, rotation, opacity}, all w.r.t. the RAW tensors)."""

    @staticmethod
    def forward(ctx, xyz, features, scaling, rotation, opacity, height, width, C2W, fxfycxcy,
                scaling_modifier=None, use_gssplat=False, arena_cache=None):
        needs_bwd = any(ctx.needs_input_grad[:5])
        # Arenas: the inference path re-uses ONE grow-only set (arena_cache["infer"]); a forward that will be
        # differentiated checks a set out of a pool and its backward returns it, so steady-state training does no
        # allocator traffic either (cudaMalloc / cudaFree of multi-GB arenas showed up as 40-70 ms step outliers)
        # while two live forwards (e.g. two renders feeding one loss) never share buffers.
        cache = None
        if arena_cache is not None:
            if needs_bwd:
                pool = arena_cache.setdefault("pool", [])
                cache = pool.pop() if pool else {}
            else:
                cache = arena_cache.setdefault("infer", {})
        with torch.no_grad():
            images, state = _raster.render_batch_forward(xyz, features, scaling, rotation, opacity, height, width,
                                                         C2W, fxfycxcy, scaling_modifier, arena_cache=cache)
        ctx.state = state
        ctx.pool = (arena_cache, cache) if needs_bwd and arena_cache is not None else None
        ctx.in_dtypes = (xyz.dtype, features.dtype, scaling.dtype, rotation.dtype, opacity.dtype)
        ctx.num_rendered = state["R"]
        return images

    @staticmethod
    def backward(ctx, grad_output):
        cache = ctx.pool[1] if ctx.pool else None
        grads = _raster.render_batch_backward(ctx.state, grad_output, arena_cache=cache)
        ctx.state = None  # release the arenas ...
        if ctx.pool:      # ... back into the pool for the next step
            ctx.pool[0]["pool"].append(cache)
            ctx.pool = None
        grads = tuple(g.to(dt) for g, dt in zip(grads, ctx.in_dtypes))
        return (*grads, None, None, None, None, None, None, None)


class BatchedGaussianRenderMSE(torch.autograd.Function):
    """Render + the l2 term of LossComputer.forward (diffusionGS/utils/losses.py:261-284) as ONE node:
    -> (images [b,v,3,H,W], l2_loss [b] = mean over (v,3,h,w) of (images - target[:, :, :3])^2).
    The per-sample sums come out of the blend-forward kernel and the backward forms the MSE part of dL/dpix inside the
    blend-backward kernel from (images, target, dL/dl2_loss) -- no per-element loss or gradient image."""

    @staticmethod
    def forward(ctx, xyz, features, scaling, rotation, opacity, height, width, C2W, fxfycxcy, target,
                scaling_modifier=None, arena_cache=None):
        ctx.set_materialize_grads(False)
        needs_bwd = any(ctx.needs_input_grad[:5])
        cache = None
        if arena_cache is not None:
            if needs_bwd:
                pool = arena_cache.setdefault("pool", [])
                cache = pool.pop() if pool else {}
            else:
                cache = arena_cache.setdefault("infer", {})
        b, v = C2W.shape[0], C2W.shape[1]
        with torch.no_grad():
            loss_sum = torch.zeros(b, dtype=torch.float64, device=xyz.device)
            images, state = _raster.render_batch_forward(xyz, features, scaling, rotation, opacity, height, width, C2W,
                                                         fxfycxcy, scaling_modifier, arena_cache=cache,
                                                         mse_target=target, mse_loss_sum=loss_sum)
            n = v * 3 * int(height) * int(width)
            l2 = (loss_sum / n).float()
        ctx.state, ctx.n = state, n
        ctx.pool = (arena_cache, cache) if needs_bwd and arena_cache is not None else None
        ctx.in_dtypes = (xyz.dtype, features.dtype, scaling.dtype, rotation.dtype, opacity.dtype)
        return images, l2

    @staticmethod
    def backward(ctx, g_images, g_l2):
        cache = ctx.pool[1] if ctx.pool else None
        if g_images is None and g_l2 is None:
            raise RuntimeError("BatchedGaussianRenderMSE.backward without any upstream gradient")
        coef = None if g_l2 is None else (g_l2.float() * (2.0 / ctx.n)).contiguous()
        grads = _raster.render_batch_backward(ctx.state, g_images, arena_cache=cache, mse_coef=coef)
        ctx.state = None
        if ctx.pool:
            ctx.pool[0]["pool"].append(cache)
            ctx.pool = None
        grads = tuple(g.to(dt) for g, dt in zip(grads, ctx.in_dtypes))
        return (*grads, None, None, None, None, None, None, None)


batched_gaussian_render = BatchedGaussianRender.apply
deferred_gaussian_render = batched_gaussian_render  # reference name (gs_core.py:1064)


class GaussianModel:
    """Minimal parameter holder with the reference's activations (gs_core.py:323-373, 545-575); the
    PLY / mesh / filter methods of the reference class are out of scope."""

    def __init__(self, sh_degree: int, scaling_modifier=None):
        self.sh_degree = sh_degree
        self.scaling_modifier = scaling_modifier
        self._xyz = self._features_dc = self._scaling = self._rotation = self._opacity = torch.empty(0)
        self._features_rest = torch.empty(0) if sh_degree > 0 else None

    def empty(self):
        self.__init__(self.sh_degree, self.scaling_modifier)

    def set_data(self, xyz, features, scaling, rotation, opacity):
        self._xyz = xyz
        self._features_dc = features[:, :1, :].contiguous()
        self._features_rest = features[:, 1:, :].contiguous() if self.sh_degree > 0 else None
        self._scaling, self._rotation, self._opacity = scaling, rotation, opacity
        return self

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_scaling(self):
        s = torch.exp(self._scaling)
        return s * self.scaling_modifier if self.scaling_modifier is not None else s

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self._rotation)

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    @property
    def get_features(self):
        if self.sh_degree > 0:
            return torch.cat((self._features_dc, self._features_rest), dim=1)
        return self._features_dc


class Renderer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.scaling_modifier = None
        sh_degree = getattr(config, "gaussians_sh_degree", 0)
        self.gaussians_model = GaussianModel(sh_degree, self.scaling_modifier)
        self._arena_cache = {}  # inference-path arenas, grown on demand, re-used step after step

    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(self, xyz, features, scaling, rotation, opacity, height, width, C2W, fxfycxcy, deferred=True):
        """xyz [b,n,3], features [b,n,(deg+1)^2,3], scaling [b,n,3], rotation [b,n,4], opacity [b,n,1],
        C2W [b,v,4,4], fxfycxcy [b,v,4] -> [b,v,3,height,width] fp32.  `deferred` is accepted for
        signature parity: both reference branches compute the same images; here both map to the
        batched kernel set."""
        out = batched_gaussian_render(xyz, features, scaling, rotation, opacity, height, width, C2W, fxfycxcy,
                                      self.scaling_modifier, getattr(self.config, "use_gssplat", False),
                                      self._arena_cache)
        self.last_num_rendered = _raster.LAST_NUM_RENDERED
        return out

    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward_mse(self, xyz, features, scaling, rotation, opacity, height, width, C2W, fxfycxcy, target):
        """forward() + the image-space MSE of the training loss in the same launch set:
        -> (renderings [b,v,3,H,W], l2_loss [b]) with l2_loss exactly LossComputer.forward's first output
        (losses.py:261-284; target [b,v,3|4,H,W], a 4th mask channel is ignored as there)."""
        out = BatchedGaussianRenderMSE.apply(xyz, features, scaling, rotation, opacity, height, width, C2W, fxfycxcy, target,
                                             self.scaling_modifier, self._arena_cache)
        self.last_num_rendered = _raster.LAST_NUM_RENDERED
        return out

    def new_gaussians_model(self):
        return copy.deepcopy(self.gaussians_model)
