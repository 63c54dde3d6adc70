"""Per-stage device time of the rasterizer forward / backward on a few cases (not a pytest file): where the time goes
in the launch-bound and the sparse regimes.   python scripts/raster_breakdown.py [out.json]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "open-diffusiongs_b200")):
    sys.path.insert(0, p)
from dgs_b200 import _lib, raster, synth  # noqa: E402

DEV = "cuda:0"
CASES = [(262146, 256, 4, "init"), (1000000, 512, 8, "fine"), (10000, 256, 1, "trained"), (10000, 256, 1, "fine"), (50000, 256, 8, "fine"), (500000, 256, 8, "fine"),
         (500000, 1024, 8, "fine"), (2000000, 512, 8, "fine"), (50000, 256, 8, "trained"), (500000, 512, 8, "trained")]


def main():
    L = _lib.lib()
    out = []
    for P, res, V, dist in CASES:
        g = synth.make_gaussians(P, 0, dist)
        c2w, fx = synth.orbit_cameras(V, res, res)
        raw = [torch.tensor(g[k][None], device=DEV) for k in ("xyz", "features", "scaling", "rotation", "opacity")]
        c2w_t, fx_t = torch.tensor(c2w[None], device=DEV), torch.tensor(fx[None], device=DEV)
        cache = {}
        img, st = raster.render_batch_forward(*raw, res, res, c2w_t, fx_t, arena_cache=cache)
        gimg = torch.randn_like(img)
        for _ in range(2):
            raster.render_batch_backward(raster.render_batch_forward(*raw, res, res, c2w_t, fx_t, arena_cache=cache)[1], gimg, cache)
        torch.cuda.synchronize()
        iters = 5
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        tf = tb = 0.0
        for _ in range(iters):
            e0.record()
            _, s2 = raster.render_batch_forward(*raw, res, res, c2w_t, fx_t, arena_cache=cache)
            e1.record()
            raster.render_batch_backward(s2, gimg, cache)
            e2.record()
            torch.cuda.synchronize()
            tf += e0.elapsed_time(e1) / iters
            tb += e1.elapsed_time(e2) / iters
        L.dgs_profile_enable(1)
        _lib.profile_read()
        for _ in range(iters):
            _, s2 = raster.render_batch_forward(*raw, res, res, c2w_t, fx_t, arena_cache=cache)
            raster.render_batch_backward(s2, gimg, cache)
        torch.cuda.synchronize()
        fam = {k: round(v[0] / iters, 4) for k, v in _lib.profile_read().items() if v[1]}
        L.dgs_profile_enable(0)
        rec = dict(P=P, res=res, views=V, dist=dist, R=int(st["R"]), chunks=st["chunks"], fwd_ms=round(tf, 4), bwd_ms=round(tb, 4),
                   families_ms=fam, sum_families_fwd=round(sum(v for k, v in fam.items() if "bwd" not in k), 4))
        out.append(rec)
        print(json.dumps(rec), flush=True)
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/raster_breakdown.json"
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
