"""BASELINE config C5 -- rasterizer stress sweep on ONE B200 (not a pytest file):

    P in {50k, 100k, 200k, 500k, 1M, 2M}  x  res in {256, 512, 1024}^2  x  8 orbit views, forward and forward+backward,
    distributions "fine" and "trained" ("init" only with --init, up to 512^2: its instance count explodes).

    python scripts/raster_sweep.py gpurun_out/raster_sweep.json [--quick] [--init]

Per case: R (reference-rule instance count, as returned by the forward), ms, views/s, and the achieved ALGORITHMIC
bandwidth  B_fwd = 159 P + 84 R + 20 N_pix,  B_bwd = 263 P + 76 R + 20 N_pix  bytes per view against the
measured HBM peak (MEASURED_PEAKS.json, else the profiling guide's fallback).  Timing: CUDA events on the current stream,
median of `iters` launches after 2 warm-ups; the operands of the larger cases exceed the 126 MB L2, the small ones are
L2-resident (flagged `l2_resident`).  The whole sweep is bounded (~2 min) so that it fits a short gpurun call."""
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

DEV = "cuda:0"
VIEWS = 8


def T(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device=DEV)


def timeit(fn, iters, warmup=2):
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


def case(P, res, dist, iters, peak):
    g = synth.make_gaussians(P, 0, dist)
    c2w, fx = synth.orbit_cameras(VIEWS, res, res, az_step=360.0 / VIEWS)
    raw = [T(g[k][None]) for k in ("xyz", "features", "scaling", "rotation", "opacity")]
    c2w_t, fx_t = T(c2w[None]), T(fx[None])

    cache = {}  # grow-only arenas re-used call after call, as dgs_b200.renderer.Renderer does: no allocator traffic in the timed region

    def fwd():
        return raster.render_batch_forward(*raw, res, res, c2w_t, fx_t, arena_cache=cache)
    img, state = fwd()
    R = int(state["R"])
    gimg = torch.randn_like(img)
    t_f = timeit(fwd, iters)
    t_fb = timeit(lambda: raster.render_batch_backward(fwd()[1], gimg, arena_cache=cache), iters)
    npix = res * res
    b_fwd = 159 * P * VIEWS + 84 * R + 20 * npix * VIEWS
    b_bwd = 263 * P * VIEWS + 76 * R + 20 * npix * VIEWS
    out = dict(P=P, res=res, views=VIEWS, dist=dist, R=R, instances_per_gaussian_view=R / (P * VIEWS),
               fwd_ms=t_f, fwdbwd_ms=t_fb, fwd_views_per_s=VIEWS / t_f * 1e3, fwdbwd_views_per_s=VIEWS / t_fb * 1e3,
               alg_bytes_fwd=b_fwd, alg_bytes_bwd=b_bwd, fwd_GBps=b_fwd / t_f / 1e6, fwdbwd_GBps=(b_fwd + b_bwd) / t_fb / 1e6,
               fwd_frac_of_hbm_peak=b_fwd / t_f / 1e6 / peak, fwdbwd_frac_of_hbm_peak=(b_fwd + b_bwd) / t_fb / 1e6 / peak,
               l2_resident=bool(b_fwd < 126e6), mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
    # size-independent properties (the oracle cannot run these sizes in seconds): colours bounded by [0, 1] + background,
    # gradients finite
    assert torch.isfinite(img).all() and float(img.min()) >= -1e-5 and float(img.max()) <= 1.0 + 1e-4
    print(json.dumps(out), flush=True)
    del img, state, gimg, raw, cache
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    return out


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "gpurun_out/raster_sweep.json"
    quick = "--quick" in sys.argv
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, src = (json.load(open(pk))["hbm_gbs"], "measured") if os.path.exists(pk) else (6650.0, "fallback")
    Ps = [50000, 100000, 200000, 500000, 1000000, 2000000]
    ress = [256, 512, 1024]
    dists = ["fine", "trained"]
    if quick:
        Ps, ress = [50000, 500000, 2000000], [256, 1024]
    cases = [(P, r, d) for d in dists for r in ress for P in Ps]
    if "--init" in sys.argv:
        cases += [(P, r, "init") for r in (256, 512) for P in Ps if P * r * r <= 500000 * 512 * 512]
    res = []
    t0 = time.time()
    budget_s = float(os.environ.get("DGS_SWEEP_BUDGET_S", "170"))
    skipped = []
    for (P, r, d) in cases:
        if time.time() - t0 > budget_s:
            skipped.append((P, r, d))
            continue
        res.append(case(P, r, d, iters=3 if P * r >= 500000 * 512 else 5, peak=peak))
    json.dump(dict(config="C5 rasterizer stress sweep, 8 views, fwd and fwd+bwd, 1 GPU", cases=res, skipped=skipped,
                   hbm_peak_gbs=peak, hbm_peak_source=src, wall_s=time.time() - t0), open(out_path, "w"), indent=1)
    print(f"[sweep] {len(res)} cases, {len(skipped)} skipped (time budget), {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
