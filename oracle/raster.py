"""oracle/raster.py -- TEST INFRASTRUCTURE: numpy front-end of the CPU rasterizer oracle.

Mirrors `CudaRasterizer::Rasterizer::forward/backward`
(reference submodules/diff-gaussian-rasterization/cuda_rasterizer/rasterizer_impl.cu:198-434)
and the ATen glue's output conventions (rasterize_points.cu:35-196), on numpy arrays.
The arithmetic lives in raster_oracle.c.  Never imported by the product path.
"""
import ctypes as C

import numpy as np

from . import build as _build

_libs = {}
_F64 = False  # module switch: True -> fp64 self-check build (see set_f64)


def set_f64(on: bool):
    """Route every call through the -DORC_F64 build with float64 arrays (gradcheck only)."""
    global _F64
    _F64 = bool(on)


def lib():
    if _F64 not in _libs:
        L = C.CDLL(_build.build(f64=_F64))
        L.orc_bin_sort.restype = C.c_longlong
        L.orc_num_threads.restype = C.c_int
        _libs[_F64] = L
    return _libs[_F64]


def _ft():
    return np.float64 if _F64 else np.float32


def _cf(v):
    return C.c_double(v) if _F64 else C.c_float(v)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=_ft())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _opt(a):
    """Reference convention: an empty tensor means 'not provided' (nullptr)."""
    if a is None:
        return None
    a = np.asarray(a)
    return None if a.size == 0 else a


def rasterize_forward(bg, means3D, colors_precomp, opacities, scales, rotations, scale_modifier,
                      cov3D_precomp, viewmatrix, projmatrix, tanfovx, tanfovy, image_height,
                      image_width, sh, degree, campos):
    """-> dict with color[3,H,W], radii[P], num_rendered and every intermediate buffer."""
    L = lib()
    means3D = _f(means3D)
    P = means3D.shape[0]
    H, W = int(image_height), int(image_width)
    colors_precomp, scales, rotations = _f(_opt(colors_precomp)), _f(_opt(scales)), _f(_opt(rotations))
    cov3D_precomp, sh = _f(_opt(cov3D_precomp)), _f(_opt(sh))
    opacities = _f(opacities).reshape(-1)
    view, proj, campos, bg = _f(viewmatrix).reshape(16), _f(projmatrix).reshape(16), _f(campos), _f(bg)
    M = 0 if sh is None else sh.shape[1]
    st = dict(P=P, H=H, W=W, M=M, degree=int(degree), mod=float(scale_modifier), tanx=float(tanfovx),
              tany=float(tanfovy), means3D=means3D, scales=scales, rotations=rotations, sh=sh,
              colors_precomp=colors_precomp, cov3D_precomp=cov3D_precomp, view=view, proj=proj,
              campos=campos, bg=bg, opacities=opacities)
    st["radii"] = np.zeros(P, np.int32)
    st["xy"] = np.zeros((P, 2), _ft())
    st["depths"] = np.zeros(P, _ft())
    st["cov3D"] = np.zeros((P, 6), _ft())
    st["rgb"] = np.zeros((P, 3), _ft())
    st["conic_opacity"] = np.zeros((P, 4), _ft())
    st["tiles_touched"] = np.zeros(P, np.uint32)
    st["clamped"] = np.zeros((P, 3), np.uint8)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    st["ranges"] = np.zeros((gx * gy, 2), np.uint32)
    st["color"] = np.zeros((3, H, W), _ft())
    st["final_T"] = np.zeros(H * W, _ft())
    st["n_contrib"] = np.zeros(H * W, np.uint32)
    st["num_rendered"] = 0
    st["point_list"] = np.zeros(0, np.uint32)
    st["keys"] = np.zeros(0, np.uint64)
    st["visited_pairs"] = 0
    if P == 0:
        return st
    L.orc_preprocess(C.c_int(P), C.c_int(int(degree)), C.c_int(M), _p(means3D), _p(scales),
                     _cf(scale_modifier), _p(rotations), _p(opacities), _p(sh), _p(cov3D_precomp),
                     _p(colors_precomp), _p(view), _p(proj), _p(campos), C.c_int(W), C.c_int(H),
                     _cf(tanfovx), _cf(tanfovy), _p(st["radii"]), _p(st["xy"]),
                     _p(st["depths"]), _p(st["cov3D"]), _p(st["rgb"]), _p(st["conic_opacity"]),
                     _p(st["tiles_touched"]), _p(st["clamped"]))
    args = (C.c_int(P), _p(st["radii"]), _p(st["xy"]), _p(st["depths"]), _p(st["tiles_touched"]),
            C.c_int(W), C.c_int(H))
    R = int(L.orc_bin_sort(*args, None, None, None))
    st["num_rendered"] = R
    st["point_list"] = np.zeros(max(R, 1), np.uint32)
    st["keys"] = np.zeros(max(R, 1), np.uint64)
    L.orc_bin_sort(*args, _p(st["point_list"]), _p(st["keys"]), _p(st["ranges"]))
    st["point_list"] = st["point_list"][:R]
    st["keys"] = st["keys"][:R]
    feat = colors_precomp if colors_precomp is not None else st["rgb"]
    vis = C.c_ulonglong(0)
    L.orc_render_fwd(C.c_int(W), C.c_int(H), _p(st["ranges"]), _p(st["point_list"]), _p(st["xy"]),
                     _p(feat), _p(st["conic_opacity"]), _p(bg), _p(st["color"]), _p(st["final_T"]),
                     _p(st["n_contrib"]), C.byref(vis))
    st["visited_pairs"] = int(vis.value)
    return st


def rasterize_backward(st, dL_dout_color):
    """-> dict of the eight gradient arrays `_C.rasterize_gaussians_backward` returns
    (rasterize_points.cu:151-195), same shapes."""
    L = lib()
    P, H, W, M = st["P"], st["H"], st["W"], st["M"]
    g = dict(dL_dmeans2D=np.zeros((P, 3), _ft()), dL_dcolors=np.zeros((P, 3), _ft()),
             dL_dopacity=np.zeros((P, 1), _ft()), dL_dmeans3D=np.zeros((P, 3), _ft()),
             dL_dcov3D=np.zeros((P, 6), _ft()), dL_dsh=np.zeros((P, M, 3), _ft()),
             dL_dscales=np.zeros((P, 3), _ft()), dL_drotations=np.zeros((P, 4), _ft()),
             dL_dconic=np.zeros((P, 2, 2), _ft()))
    if P == 0:
        return g
    dpix = _f(dL_dout_color)
    dm2 = np.zeros((P, 2), np.float64)
    dcon = np.zeros((P, 3), np.float64)
    dop = np.zeros(P, np.float64)
    dcol = np.zeros((P, 3), np.float64)
    feat = st["colors_precomp"] if st["colors_precomp"] is not None else st["rgb"]
    L.orc_render_bwd(C.c_int(W), C.c_int(H), _p(st["ranges"]), _p(st["point_list"]), _p(st["xy"]),
                     _p(feat), _p(st["conic_opacity"]), _p(st["bg"]), _p(st["final_T"]),
                     _p(st["n_contrib"]), _p(dpix), _p(dm2), _p(dcon), _p(dop), _p(dcol))
    dm2f, dconf = dm2.astype(_ft()), dcon.astype(_ft())
    g["dL_dmeans2D"][:, :2] = dm2f
    g["dL_dconic"][:, 0, 0] = dconf[:, 0]
    g["dL_dconic"][:, 0, 1] = dconf[:, 1]
    g["dL_dconic"][:, 1, 1] = dconf[:, 2]
    g["dL_dopacity"][:, 0] = dop.astype(_ft())
    g["dL_dcolors"][:] = dcol.astype(_ft())
    cov3d = st["cov3D_precomp"] if st["cov3D_precomp"] is not None else st["cov3D"]
    L.orc_preprocess_bwd(C.c_int(P), C.c_int(st["degree"]), C.c_int(M), _p(st["means3D"]),
                         _p(st["radii"]), _p(st["sh"]), _p(st["clamped"]), _p(st["scales"]),
                         _p(st["rotations"]), _cf(st["mod"]), _p(cov3d), _p(st["view"]),
                         _p(st["proj"]), _p(st["campos"]), C.c_int(W), C.c_int(H), _cf(st["tanx"]),
                         _cf(st["tany"]), _p(dm2f), _p(dconf), _p(g["dL_dcolors"]),
                         _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]),
                         _p(g["dL_drotations"]))
    return g


def mark_visible(means3D, viewmatrix):
    means3D = _f(means3D)
    out = np.zeros(means3D.shape[0], np.uint8)
    lib().orc_mark_visible(C.c_int(means3D.shape[0]), _p(means3D), _p(_f(viewmatrix).reshape(16)), _p(out))
    return out.astype(bool)
