"""Training-step timing on the GPU box (not a pytest file): DiT forward (training mode, activations stored) + batched
render of V_render views + MSE loss + raster backward + dgs_dit_backward + AdamW, obj-256 shapes (BASELINE configs[2]:
4 input views 256x256, N = 4098 tokens, 24 layers), per-GPU batch B.

    python tests/perf_train.py gpurun_out/perf_train.json [B] [V_render] [steps]

Reports ms per phase (CUDA events), per-kernel-family ms from the library's profiling hooks, and achieved TFLOP/s of the
DiT part against F_train = 3 x F_fwd (fwd + 2 x bwd; no recompute)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "open-diffusiongs_b200")):
    sys.path.insert(0, p)

from dgs_b200 import _lib, synth  # noqa: E402
from dgs_b200.denoiser import DGSDenoiser  # noqa: E402
from dgs_b200.train import DitTrainer  # noqa: E402

DEV = "cuda:0"


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/perf_train.json"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    VR = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    layers = int(os.environ.get("DGS_LAYERS", "24"))
    torch.manual_seed(0)
    model = DGSDenoiser(dict(patch_size=8, num_layers=layers)).to(DEV)
    trainer = DitTrainer(model)
    model.train()
    V, H, W = 4, 256, 256
    g = torch.Generator(DEV).manual_seed(0)
    images = torch.rand(B, V, 3, H, W, device=DEV, generator=g)
    c2w, fx = synth.orbit_cameras(VR, H, W, az_step=36.0)
    c2w = torch.tensor(np.repeat(c2w[None], B, 0), device=DEV)
    fx = torch.tensor(np.repeat(fx[None], B, 0), device=DEV)
    from dgs_b200.diffusion import transform_input
    ray_o, ray_d = transform_input(images, c2w[:, :V].contiguous(), fx[:, :V].contiguous())
    t = torch.randint(0, 1000, (B,), device=DEV, generator=g)
    target = torch.rand(B, VR, 3, H, W, device=DEV, generator=g)
    L = _lib.lib()
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    rec = []
    for step in range(steps + 2):
        prof = step >= 2
        if prof:
            L.dgs_profile_enable(1)
            _lib.profile_read()
        e = [ev() for _ in range(5)]
        e[0].record()
        out, _ = model.image_to_gaussians(images, ray_o, ray_d, t)
        e[1].record()
        renders = model.render_gaussians(out, c2w, fx, H, W)
        loss = ((renders - target) ** 2).mean()
        e[2].record()
        trainer.zero_grad()
        loss.backward()
        e[3].record()
        trainer.optimizer_step(allreduce=False)
        e[4].record()
        torch.cuda.synchronize()
        if prof:
            fam = {k: round(v[0], 3) for k, v in _lib.profile_read().items() if v[1]}
            L.dgs_profile_enable(0)
            rec.append(dict(dit_fwd_ms=e[0].elapsed_time(e[1]), render_loss_ms=e[1].elapsed_time(e[2]),
                            backward_ms=e[2].elapsed_time(e[3]), optimizer_ms=e[3].elapsed_time(e[4]),
                            step_ms=e[0].elapsed_time(e[4]), loss=float(loss), families=fam))
            print(json.dumps(rec[-1]), flush=True)
    N = 2 + V * (H // 8) * (W // 8)
    D = 1024
    f_fwd = layers * (24 * N * D * D + 4 * N * N * D + 12 * D * D) + 2 * (N - 2) * D * (576 + 896)
    med = lambda k: float(np.median([r[k] for r in rec]))  # noqa: E731
    bwd_dit = float(np.median([sum(v for k, v in r["families"].items() if k.startswith("dit.bwd")) for r in rec]))
    res = dict(B=B, V_render=VR, layers=layers, steps=rec, step_ms=med("step_ms"), samples_per_s=B / med("step_ms") * 1e3,
               dit_fwd_tflops=B * f_fwd / med("dit_fwd_ms") / 1e9, dit_bwd_ms=bwd_dit,
               dit_bwd_tflops=2 * B * f_fwd / bwd_dit / 1e9 if bwd_dit else None,
               mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
    json.dump(res, open(out_path, "w"), indent=1)
    print({k: v for k, v in res.items() if k != "steps"})


if __name__ == "__main__":
    main()
