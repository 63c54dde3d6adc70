"""Drop-in `diff_gaussian_rasterization` package backed by libdgs_b200.so (sm_100a).

The reference imports this name (diffusionGS/models/gsrenderer/gs_core.py:10-13) and uses
`GaussianRasterizationSettings` + `GaussianRasterizer`; both keep the field order / call signature of
DGR/diff_gaussian_rasterization/__init__.py:157-220, and the autograd contract of its
`_RasterizeGaussians` (44-155): gradients come back in the order
(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, None) and are
taken w.r.t. the ACTIVATED inputs.  `_C` exposes the three functions of DGR/ext.cpp:15-19.
"""
import os
import sys
import types
from typing import NamedTuple

import torch
import torch.nn as nn

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:  # make the sibling host package importable when only this one is on the path
    sys.path.insert(0, _PKG_ROOT)

from dgs_b200 import raster as _raster  # noqa: E402

_C = types.SimpleNamespace(
    rasterize_gaussians=_raster.rasterize_gaussians,
    rasterize_gaussians_backward=_raster.rasterize_gaussians_backward,
    mark_visible=_raster.mark_visible,
)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _snapshot(args, path):
    """debug=True behaviour of the reference: dump the call's inputs when the native call throws."""
    torch.save(tuple(a.detach().cpu().clone() if isinstance(a, torch.Tensor) else a for a in args), path)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        s = raster_settings
        call = (s.bg, means3D, colors_precomp, opacities, scales, rotations, s.scale_modifier, cov3Ds_precomp,
                s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy, s.image_height, s.image_width, sh, s.sh_degree,
                s.campos, s.prefiltered, s.debug)
        try:
            num_rendered, color, radii, geom, binning, img = _C.rasterize_gaussians(*call)
        except Exception:
            if s.debug:
                _snapshot(call, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
            raise
        ctx.raster_settings = s
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning, img)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii):
        s = ctx.raster_settings
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning, img = ctx.saved_tensors
        call = (s.bg, means3D, radii, colors_precomp, scales, rotations, s.scale_modifier, cov3Ds_precomp,
                s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy, grad_out_color, sh, s.sh_degree, s.campos, geom,
                ctx.num_rendered, binning, img, s.debug)
        try:
            g_m2d, g_col, g_op, g_m3d, g_cov, g_sh, g_sc, g_rot = _C.rasterize_gaussians_backward(*call)
        except Exception:
            if s.debug:
                _snapshot(call, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
            raise
        return g_m3d, g_m2d, g_sh, g_col, g_op, g_sc, g_rot, g_cov, None


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            return _C.mark_visible(positions, self.raster_settings.viewmatrix, self.raster_settings.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        has_sr = scales is not None and rotations is not None
        if (not has_sr and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([])  # the reference's "not provided" sentinel
        shs, colors_precomp = (empty if t is None else t for t in (shs, colors_precomp))
        scales, rotations, cov3D_precomp = (empty if t is None else t for t in (scales, rotations, cov3D_precomp))
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, self.raster_settings)
