"""Torch-facing host side of the rasterizer C ABI.

`rasterize_gaussians`, `rasterize_gaussians_backward`, `mark_visible` have the exact signatures and
return tuples of the reference's pybind module `_C` (DGR/ext.cpp:15-19, DGR/rasterize_points.cu:35-217);
torch only supplies device memory (the arena allocator callbacks) and the current stream.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ALLOC_FN, DgsError, RasterArgs, RenderBatchArgs, RenderMse, check


LAST_NUM_RENDERED = None  # instance count R of the most recent batched forward (bench/roofline bookkeeping)


class _Arena:
    """dgs_alloc_fn backed by a torch uint8 tensor (the reference's resizeFunctional,
    rasterize_points.cu:27-33)."""

    def __init__(self, device, cache=None, key=None):
        self.device = device
        self.cache, self.key = cache, key   # optional grow-only buffers re-used across calls (inference path)
        self.tensor = torch.empty(0, dtype=torch.uint8, device=device)
        self.tensors = []                   # every buffer handed out (the binning arena is requested once per phase)
        self.cb = ALLOC_FN(self._alloc)

    def _alloc(self, nbytes, _user):
        try:
            if self.cache is not None:
                k = f"{self.key}{len(self.tensors)}"
                t = self.cache.get(k)
                if t is None or t.numel() < nbytes or t.device != self.device:
                    t = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.device)
                    self.cache[k] = t
            else:
                t = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self.tensor = t
            self.tensors.append(t)
            return t.data_ptr()
        except Exception as e:  # noqa: BLE001  (propagated as DGS_ERR_ALLOC by the C side)
            import sys
            print(f"[dgs_b200] arena allocation of {nbytes} bytes failed: {e!r}"[:600], file=sys.stderr)
            return None


def _ptr(t):
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.contiguous().float()
    return t


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t, name):
    if not t.is_cuda:
        raise _lib.DgsError(f"{name} must be a CUDA tensor: libdgs_b200 has no CPU path")


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree,
                        campos, prefiltered, debug):
    """-> (num_rendered, out_color[3,H,W], radii[P] int32, geomBuffer, binningBuffer, imgBuffer)
    (RasterizeGaussiansCUDA, rasterize_points.cu:35-115)."""
    if means3D.ndimension() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:57-59
    _require_cuda(means3D, "means3D")
    dev = means3D.device
    P, H, W = means3D.size(0), int(image_height), int(image_width)
    with torch.cuda.device(dev):
        out_color = torch.zeros(3, H, W, dtype=torch.float32, device=dev)
        radii = torch.zeros(P, dtype=torch.int32, device=dev)
        geom, binning, img = _Arena(dev), _Arena(dev), _Arena(dev)
        if P == 0:
            return 0, out_color, radii, geom.tensor, binning.tensor, img.tensor
        keep = [_f32c(t) for t in (background, means3D, sh, colors, opacity, scales, rotations, cov3D_precomp,
                                   viewmatrix, projmatrix, campos)]
        bg, m3, sh_, col, op, sc, ro, cov, vm, pm, cp = keep
        M = sh_.size(1) if (sh_ is not None and sh_.numel()) else 0
        a = RasterArgs(P=P, D=int(degree), M=M, W=W, H=H, background=_ptr(bg), means3D=_ptr(m3), shs=_ptr(sh_),
                       colors_precomp=_ptr(col), opacities=_ptr(op), scales=_ptr(sc), rotations=_ptr(ro),
                       cov3D_precomp=_ptr(cov), viewmatrix=_ptr(vm), projmatrix=_ptr(pm), campos=_ptr(cp),
                       scale_modifier=float(scale_modifier), tan_fovx=float(tan_fovx), tan_fovy=float(tan_fovy),
                       prefiltered=int(bool(prefiltered)), debug=int(bool(debug)))
        R = C.c_int(0)
        check(_lib.lib().dgs_raster_forward(C.byref(a), geom.cb, None, binning.cb, None, img.cb, None,
                                            out_color.data_ptr(), radii.data_ptr(), C.byref(R), _stream(dev)))
    return R.value, out_color, radii, geom.tensor, binning.tensor, img.tensor


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh,
                                 degree, campos, geomBuffer, R, binningBuffer, imageBuffer, debug,
                                 opacities=None):
    """-> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)
    (RasterizeGaussiansBackwardCUDA, rasterize_points.cu:117-196).  `opacities` is unused (the forward
    state already holds them), kept only so callers may pass it."""
    _require_cuda(means3D, "means3D")
    dev = means3D.device
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    M = sh.size(1) if (sh is not None and sh.numel()) else 0
    with torch.cuda.device(dev):
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)  # noqa: E731
        dm3, dm2, dcol, dcon = z(P, 3), z(P, 3), z(P, 3), z(P, 2, 2)
        dop, dcov, dsh, dsc, dro = z(P, 1), z(P, 6), z(P, M, 3), z(P, 3), z(P, 4)
        if P != 0:
            keep = [_f32c(t) for t in (background, means3D, sh, colors, scales, rotations, cov3D_precomp, viewmatrix,
                                       projmatrix, campos, dL_dout_color)]
            bg, m3, sh_, col, sc, ro, cov, vm, pm, cp, dpix = keep
            a = RasterArgs(P=P, D=int(degree), M=M, W=W, H=H, background=_ptr(bg), means3D=_ptr(m3), shs=_ptr(sh_),
                           colors_precomp=_ptr(col), opacities=None, scales=_ptr(sc), rotations=_ptr(ro),
                           cov3D_precomp=_ptr(cov), viewmatrix=_ptr(vm), projmatrix=_ptr(pm), campos=_ptr(cp),
                           scale_modifier=float(scale_modifier), tan_fovx=float(tan_fovx),
                           tan_fovy=float(tan_fovy), prefiltered=0, debug=int(bool(debug)))
            check(_lib.lib().dgs_raster_backward(
                C.byref(a), int(R), _ptr(radii), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                dpix.data_ptr(), dm2.data_ptr(), dcon.data_ptr(), dop.data_ptr(), dcol.data_ptr(), dm3.data_ptr(),
                dcov.data_ptr(), _ptr(dsh), dsc.data_ptr(), dro.data_ptr(), _stream(dev)))
    return dm2, dcol, dop, dm3, dcov, dsh, dsc, dro


def mark_visible(means3D, viewmatrix, projmatrix):
    """-> bool[P] (markVisible, rasterize_points.cu:198-217)."""
    _require_cuda(means3D, "means3D")
    dev = means3D.device
    P = means3D.size(0)
    present = torch.zeros(P, dtype=torch.bool, device=dev)
    if P:
        with torch.cuda.device(dev):
            m3, vm, pm = _f32c(means3D), _f32c(viewmatrix), _f32c(projmatrix)
            check(_lib.lib().dgs_mark_visible(P, m3.data_ptr(), vm.data_ptr(), pm.data_ptr(), present.data_ptr(),
                                              _stream(dev)))
    return present


# ------------------------------------------------------------------------------------------------
# batched renderer
# ------------------------------------------------------------------------------------------------
import os as _os

# two-phase binning: phase A = the nearest 1/2^k of every view's Gaussians (0 = single pass); DGS_RASTER_NEAR_LOG2 overrides
DEFAULT_NEAR_LOG2 = int(_os.environ.get("DGS_RASTER_NEAR_LOG2", "-1"))  # -1 = adaptive: 1/8, or 1/16 for dense scenes


def _batch_args(xyz, features, scaling, rotation, opacity, C2W, fxfycxcy, H, W, scale_modifier, debug=False,
                near_log2=0):
    B, P = xyz.shape[0], xyz.shape[1]
    V = C2W.shape[1]
    M = features.shape[2]
    D = int(round(M ** 0.5)) - 1  # gs_core.py:978
    a = RenderBatchArgs(B=B, V=V, P=P, M=M, D=D, W=int(W), H=int(H), xyz=xyz.data_ptr(), features=features.data_ptr(),
                        scaling=scaling.data_ptr(), rotation=rotation.data_ptr(), opacity=opacity.data_ptr(),
                        c2w=C2W.data_ptr(), fxfycxcy=fxfycxcy.data_ptr(),
                        scale_modifier=1.0 if scale_modifier is None else float(scale_modifier), debug=int(debug),
                        near_log2=int(near_log2))
    a.bg[0] = a.bg[1] = a.bg[2] = 1.0  # render_opencv_cam's default bg_color, gs_core.py:880
    return a


def render_batch_forward(xyz, features, scaling, rotation, opacity, H, W, C2W, fxfycxcy, scale_modifier=None,
                         arena_cache=None, near_log2=None, mse_target=None, mse_loss_sum=None):
    """All (sample, view) pairs in one launch set -> (images [B,V,3,H,W] fp32, state).
    mse_target [B,V,3|4,H,W] + mse_loss_sum (fp64 [B], zeroed by the caller): the blend kernel also adds
    sum (render - target)^2 of every sample into mse_loss_sum (dgs_render_batch_forward_mse).  When the batch holds more than
    2^31-1 instances (e.g. a random-init denoiser at 512x512: ~6e8 per view) the views are rendered in halves,
    recursively -- the reference renders one view per call anyway (gs_core.py:990-1001); `state` then carries one
    sub-state per chunk and render_batch_backward sums the chunks' gradients."""
    try:
        return _render_batch_forward_one(xyz, features, scaling, rotation, opacity, H, W, C2W, fxfycxcy, scale_modifier,
                                         arena_cache, near_log2, mse_target, mse_loss_sum)
    except DgsError as e:
        V = C2W.shape[1]
        if "exceeds 2^31-1" not in str(e) or V < 2:
            raise
    global LAST_NUM_RENDERED
    outs, subs, total = [], [], 0
    for ci, (v0, v1) in enumerate(((0, V // 2), (V // 2, V))):
        sub_cache = None if arena_cache is None else arena_cache.setdefault(("views", ci), {})
        if mse_loss_sum is not None and ci == 0:
            mse_loss_sum.zero_()  # the failed whole-batch attempt may have counted some tiles already
        o, st = render_batch_forward(xyz, features, scaling, rotation, opacity, H, W, C2W[:, v0:v1].contiguous(),
                                     fxfycxcy[:, v0:v1].contiguous(), scale_modifier, sub_cache, near_log2,
                                     None if mse_target is None else mse_target[:, v0:v1].contiguous(), mse_loss_sum)
        outs.append(o)
        subs.append((v0, v1, st))
        total += st["R"]
    LAST_NUM_RENDERED = total
    return torch.cat(outs, dim=1), dict(sub=subs, R=total, tensors=subs[0][2]["tensors"])


def _render_batch_forward_one(xyz, features, scaling, rotation, opacity, H, W, C2W, fxfycxcy, scale_modifier=None,
                              arena_cache=None, near_log2=None, mse_target=None, mse_loss_sum=None):
    """One launch set over every (sample, view) pair.  `arena_cache` (a dict): re-use grow-only arenas across calls; the caller must not hand the same dict to another
    forward while this call's state is still needed (renderer.py keeps one dict for inference and a pool of dicts for
    differentiated forwards); stream order makes the re-use safe."""
    _require_cuda(xyz, "xyz")
    dev = xyz.device
    tens = [_f32c(t) for t in (xyz, features, scaling, rotation, opacity, C2W, fxfycxcy)]
    B, V = tens[5].shape[0], tens[5].shape[1]
    with torch.cuda.device(dev):
        out = torch.empty(B, V, 3, int(H), int(W), dtype=torch.float32, device=dev)
        geom, binning, img = (_Arena(dev, arena_cache, k) for k in ("geom", "binning", "img"))
        near_log2 = DEFAULT_NEAR_LOG2 if near_log2 is None else near_log2
        a = _batch_args(*tens, H, W, scale_modifier, near_log2=near_log2)
        R = C.c_longlong(0)
        chunks = (C.c_longlong * 2)(0, 0)
        mse = None
        if mse_target is not None:
            mse_target = _f32c(mse_target)
            if tuple(mse_target.shape) not in ((B, V, 3, int(H), int(W)), (B, V, 4, int(H), int(W))):
                raise ValueError(f"mse target shape {tuple(mse_target.shape)} != [{B},{V},3|4,{H},{W}]")
            if mse_loss_sum.dtype != torch.float64 or mse_loss_sum.numel() != B:
                raise ValueError("mse_loss_sum must be a float64 tensor with one entry per sample")
            mse = RenderMse(target=mse_target.data_ptr(), target_channels=mse_target.shape[2],
                            loss_sum=mse_loss_sum.data_ptr(), coef=None, images=None)
        check(_lib.lib().dgs_render_batch_forward_mse(C.byref(a), geom.cb, None, binning.cb, None, img.cb, None,
                                                      out.data_ptr(), C.byref(R), chunks,
                                                      None if mse is None else C.byref(mse), _stream(dev)))
    global LAST_NUM_RENDERED
    LAST_NUM_RENDERED = R.value
    state = dict(mse_target=mse_target, images=out if mse_target is not None else None, tensors=tens, geom=geom.tensor, binning=binning.tensors[0],
                 binning_b=binning.tensors[1] if len(binning.tensors) > 1 else None, img=img.tensor, R=R.value,
                 chunks=(int(chunks[0]), int(chunks[1])), H=int(H), W=int(W), scale_modifier=scale_modifier,
                 near_log2=near_log2)
    return out, state


def render_batch_backward(state, grad_images, arena_cache=None, mse_coef=None):
    """-> (d_xyz, d_features, d_scaling, d_rotation, d_opacity), re-using the forward's sorted lists.
    mse_coef (fp32 [B], device): dL/dpix += mse_coef[b] * (render - target) is formed inside the blend-backward kernel
    (the forward must have been given mse_target); grad_images may then be None."""
    if "sub" in state:  # view-chunked forward: the per-Gaussian gradients are sums over views
        total = None
        for ci, (v0, v1, st) in enumerate(state["sub"]):
            sub_cache = None if arena_cache is None else arena_cache.setdefault(("views", ci), {})
            g = render_batch_backward(st, None if grad_images is None else grad_images[:, v0:v1], sub_cache, mse_coef)
            total = g if total is None else tuple(a + b for a, b in zip(total, g))
        return total
    tens = state["tensors"]
    dev = tens[0].device
    g = _f32c(grad_images)
    if g is None and mse_coef is None:
        raise ValueError("render_batch_backward needs grad_images or mse_coef")
    mse = None
    if mse_coef is not None:
        if state.get("mse_target") is None:
            raise ValueError("mse_coef given but the forward ran without mse_target")
        mse_coef = _f32c(mse_coef)
        mse = RenderMse(target=state["mse_target"].data_ptr(), target_channels=state["mse_target"].shape[2], loss_sum=None,
                        coef=mse_coef.data_ptr(), images=state["images"].data_ptr())
    with torch.cuda.device(dev):
        outs = [torch.empty_like(t) for t in tens[:5]]
        scratch = _Arena(dev, arena_cache, "bwd_scratch")
        a = _batch_args(*tens, state["H"], state["W"], state["scale_modifier"], near_log2=state["near_log2"])
        chunks = (C.c_longlong * 2)(*state["chunks"])
        check(_lib.lib().dgs_render_batch_backward_mse(
            C.byref(a), state["R"], chunks, _ptr(state["geom"]), _ptr(state["binning"]), _ptr(state["binning_b"]),
            _ptr(state["img"]), None if g is None else g.data_ptr(), None if mse is None else C.byref(mse),
            outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), outs[4].data_ptr(),
            scratch.cb, None, _stream(dev)))
    return tuple(outs)


def export_state(n_views, P, W, H, R, geom, binning, img):
    """Debug/test introspection of the opaque arenas -> dict of torch tensors."""
    dev = geom.device
    N = n_views * P
    tiles = n_views * ((W + 15) // 16) * ((H + 15) // 16)
    o = dict(xy=torch.zeros(N, 2, device=dev), depth=torch.zeros(N, device=dev),
             conic_opacity=torch.zeros(N, 4, device=dev), rgb=torch.zeros(N, 3, device=dev),
             tiles_touched=torch.zeros(N, dtype=torch.int32, device=dev),
             point_list=torch.zeros(max(R, 1), dtype=torch.int32, device=dev),
             ranges=torch.zeros(tiles, 2, dtype=torch.int32, device=dev),
             final_T=torch.zeros(n_views * H * W, device=dev),
             n_contrib=torch.zeros(n_views * H * W, dtype=torch.int32, device=dev))
    with torch.cuda.device(dev):
        check(_lib.lib().dgs_raster_export_state(
            n_views, P, W, H, R, _ptr(geom), _ptr(binning), _ptr(img), o["xy"].data_ptr(), o["depth"].data_ptr(),
            o["conic_opacity"].data_ptr(), o["rgb"].data_ptr(), o["tiles_touched"].data_ptr(),
            o["point_list"].data_ptr(), o["ranges"].data_ptr(), o["final_T"].data_ptr(), o["n_contrib"].data_ptr(),
            _stream(dev)))
    o["point_list"] = o["point_list"][:R]
    return o
