"""Writes tests/golden/dit_ref_<case>.npz: outputs of the REFERENCE's own denoiser code (denoiser.py /
denoiser_scene.py / utils_transformer.py executed from /root/reference through tests/golden/ref_import.py) on seeded
parameters and inputs.  Parameters and inputs are regenerated from the seed on any box (ref_import.seeded_*), so the
fixtures hold only the outputs and a digest of the gradients.

    python tests/golden/make_dit_golden.py        # needs /root/reference (this container)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import as ri  # noqa: E402


def main():
    torch.set_num_threads(8)
    for name in ri.DIT_CASES:
        model, _, outs, grads, _ = ri.reference_dit_case(name)
        rec = {"out/" + k: v.numpy() for k, v in outs.items()}
        # gradient digest: per-parameter L2 norm and a fixed random projection (a full copy would be 100 MB at width 1024)
        rng = np.random.default_rng(7)
        for k, g in grads.items():
            gg = g.double().numpy().ravel()
            rec["gnorm/" + k] = np.float64(np.linalg.norm(gg))
            rec["gproj/" + k] = np.float64(gg @ rng.standard_normal(gg.size))
        path = os.path.join(HERE, f"dit_ref_{name}.npz")
        np.savez_compressed(path, **rec)
        print(name, {k: tuple(v.shape) for k, v in outs.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
