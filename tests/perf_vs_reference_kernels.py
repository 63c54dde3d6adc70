"""Rasterizer throughput on the GPU box: ours (batched, one launch set) vs the UNMODIFIED reference kernels
(oracle/_ref/dgr_ref_C.so, rebuilt for sm_100a) driven the way the reference drives them -- one call per view
(gs_core.py:990-1001) and, for the backward, the reference's re-render + backward per view (gs_core.py:1041-1056).
BASELINE config C5 (stress sweep) subset + the obj-256 denoise-step shape.  Not a pytest file (name perf_*).

    python tests/perf_vs_reference_kernels.py gpurun_out/perf_raster.json [--quick]

Reports views/s, R, algorithmic bytes (B_fwd = 159 P + 84 R + 20 N_pix, B_bwd = 263 P + 76 R + 20 N_pix)
and GB/s against the measured HBM peak.  Parity on the same inputs is asserted (colour rel-L2 <= 1e-4)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "open-diffusiongs_b200")):
    sys.path.insert(0, p)

from dgs_b200 import raster, synth  # noqa: E402
from oracle import build_ref  # noqa: E402

DEV = "cuda:0"


def T(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device=DEV)


def timeit(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def case(ref, P, res, V, dist, iters):
    g = synth.make_gaussians(P, 0, dist)
    c2w, fx = synth.orbit_cameras(V, res, res, az_step=360.0 / V)
    raw = [T(g[k][None]) for k in ("xyz", "features", "scaling", "rotation", "opacity")]
    c2w_t, fx_t = T(c2w[None]), T(fx[None])
    out = {}

    cache = {}  # grow-only arenas re-used call after call, as dgs_b200.renderer.Renderer does (r1 timed the allocator too)

    def ours_fwd():
        return raster.render_batch_forward(*raw, res, res, c2w_t, fx_t, arena_cache=cache)
    img, state = ours_fwd()
    R = state["R"]
    gimg = torch.randn_like(img)
    t_f = timeit(lambda: ours_fwd(), iters)
    t_fb = timeit(lambda: raster.render_batch_backward(ours_fwd()[1], gimg, arena_cache=cache), iters)
    npix = res * res
    b_fwd = 159 * P * V + 84 * R + 20 * npix * V
    b_bwd = 263 * P * V + 76 * R + 20 * npix * V
    out.update(P=P, res=res, views=V, dist=dist, R=int(R), ours_fwd_ms=t_f, ours_fwdbwd_ms=t_fb,
               ours_fwd_views_per_s=V / t_f * 1e3, ours_fwdbwd_views_per_s=V / t_fb * 1e3,
               alg_bytes_fwd=b_fwd, alg_bytes_bwd=b_bwd, ours_fwd_GBps=b_fwd / t_f / 1e6,
               ours_fwdbwd_GBps=(b_fwd + b_bwd) / t_fb / 1e6)
    if ref is not None:
        act = synth.activate(g)
        m3, sh, op = T(act["means3D"]), T(act["shs"]), T(act["opacities"])
        sc, ro = T(act["scales"]), T(act["rotations"])
        e = torch.empty(0, device=DEV)
        cams = [synth.camera_matrices(c2w[v], fx[v], res, res) for v in range(V)]
        cams = [(T(c[0]), T(c[1]), T(c[2]), float(c[3]), float(c[4])) for c in cams]
        bg = T(np.ones(3))

        def ref_fwd_view(c):
            return ref.rasterize_gaussians(bg, m3, e, op, sc, ro, 1.0, e, c[0], c[1], c[3], c[4], res, res, sh, 0, c[2],
                                           False, False)

        def ref_fwd():
            return [ref_fwd_view(c) for c in cams]

        def ref_fwdbwd():  # forward (no_grad) + the reference's backward = re-render + backward, per view
            ref_fwd()
            for v, c in enumerate(cams):
                Rr, col, radii, gb, bb, ib = ref_fwd_view(c)
                ref.rasterize_gaussians_backward(bg, m3, radii, e, sc, ro, 1.0, e, c[0], c[1], c[3], c[4], gimg[0, v], sh, 0,
                                                 c[2], gb, Rr, bb, ib, False)
        r0 = ref_fwd()
        err = float(((img[0, 0] - r0[0][1]).norm() / r0[0][1].norm()))
        assert err < 1e-4, err
        assert sum(r[0] for r in r0) == R or abs(sum(r[0] for r in r0) - R) < 1e-4 * R
        t_rf = timeit(lambda: ref_fwd(), max(2, iters // 2))
        t_rfb = timeit(lambda: ref_fwdbwd(), max(2, iters // 2))
        out.update(ref_fwd_ms=t_rf, ref_fwdbwd_ms=t_rfb, ref_fwd_views_per_s=V / t_rf * 1e3,
                   ref_fwdbwd_views_per_s=V / t_rfb * 1e3, speedup_fwd=t_rf / t_f, speedup_fwdbwd=t_rfb / t_fb,
                   colour_rel_l2_vs_ref=err)
    print(json.dumps(out), flush=True)
    return out


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/perf_raster.json"
    quick = "--quick" in sys.argv
    ref = build_ref.load_module()
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    cases = [(10000, 256, 1, "trained"), (50000, 256, 8, "trained"), (262146, 256, 4, "init"), (200000, 512, 8, "trained"),
             (200000, 512, 8, "fine"), (1000000, 512, 8, "fine")]
    if not quick:
        cases += [(1000000, 1024, 8, "trained"), (2000000, 1024, 8, "fine"), (500000, 256, 8, "init")]
    res = []
    t0 = time.time()
    for (P, r, V, d) in cases:
        res.append(case(ref, P, r, V, d, iters=5))
    json.dump(dict(cases=res, hbm_peak_gbs=peaks.get("hbm_gbs"), wall_s=time.time() - t0,
                   reference_kernels="oracle/_ref/dgr_ref_C.so" if ref is not None else None), open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
