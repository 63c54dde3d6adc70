"""Generates tests/golden/ref_*.npz by running the UNMODIFIED reference kernels (oracle/_ref/dgr_ref_C.so,
built by oracle/build_ref.py from /root/reference) on a B200, on the seeded C1 scenes of tests/util.py.
Run on the GPU box:   python tests/golden/make_golden.py gpurun_out/golden
then copy gpurun_out/golden/*.npz to tests/golden/ and commit them.  Stored fp16-free, fp32 arrays,
small scenes so the fixtures stay < 1 MB each."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "open-diffusiongs_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from oracle import build_ref  # noqa: E402
from util import scene_c1  # noqa: E402

CASES = [  # (name, P, dist, W, H, seed, degree)
    ("trained_2k_128", 2000, "trained", 128, 128, 0, 0),
    ("init_1k_96x64", 1000, "init", 96, 64, 1, 0),
    ("fine_4k_128", 4000, "fine", 128, 128, 2, 0),
    ("trained_1k_sh2", 1000, "trained", 64, 64, 3, 2),
]


def main(out_dir):
    ref = build_ref.load_module()
    assert ref is not None, "oracle/_ref/dgr_ref_C.so missing"
    os.makedirs(out_dir, exist_ok=True)
    dev = "cuda:0"
    T = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device=dev)  # noqa: E731
    e = torch.empty(0, device=dev)
    for name, P, dist, W, H, seed, deg in CASES:
        sc = scene_c1(P=P, dist=dist, W=W, H=H, seed=seed)
        a = sc["act"]
        rng = np.random.default_rng(100 + seed)
        sh = a["shs"] if deg == 0 else rng.normal(0, 0.4, (P, (deg + 1) ** 2, 3)).astype(np.float32)
        args = (T(np.ones(3)), T(a["means3D"]), e, T(a["opacities"]), T(a["scales"]), T(a["rotations"]), 1.0, e,
                T(sc["view"]), T(sc["proj"]), float(sc["tanx"]), float(sc["tany"]), H, W, T(sh), deg,
                T(sc["campos"]), False, False)
        R, color, radii, geom, binning, img = ref.rasterize_gaussians(*args)
        dpix = rng.normal(0, 1, (3, H, W)).astype(np.float32)
        g = ref.rasterize_gaussians_backward(args[0], args[1], radii, e, args[4], args[5], 1.0, e, args[8], args[9],
                                             args[10], args[11], T(dpix), args[14], deg, args[16], geom, R, binning,
                                             img, False)
        names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
                 "dL_drotations"]
        out = dict(P=P, dist=dist, W=W, H=H, seed=seed, degree=deg, sh=sh if deg else np.zeros(0, np.float32),
                   num_rendered=R, color=color.cpu().numpy(), radii=radii.cpu().numpy(), dL_dcolor=dpix)
        out.update({n: t.cpu().numpy() for n, t in zip(names, g)})
        np.savez_compressed(os.path.join(out_dir, f"ref_{name}.npz"), **out)
        print(name, "R =", R, "mean colour", float(color.mean()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
