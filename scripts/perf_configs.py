"""The other shipped configurations of the hot path on ONE B200 (not a pytest file, not the bench line):

    obj-512   pipline_obj.py demo / diffusionGS_rel_512.yaml : 4 views 512x512 -> N = 16,386 tokens, P = 1,048,578 Gaussians
    scene-256 diffusionGS_scene.yaml      : scene denoiser (Pluecker 'plk', depth = sigmoid * 500), 4 views 256x256, 7 render views
    scene-512 diffusionGS_scene_512.yaml  : same at 512x512
    obj-256   (bench.py's default workload, repeated here so that all rows come from one run)

    python scripts/perf_configs.py gpurun_out/perf_configs.json [--train] [--only obj-512,scene-256]

Inference row = 1 denoise step (DiT forward + V_in-view render), per-GPU batch 1.  --train adds one training step per
config (DiT fwd + V_render-view render + MSE + raster bwd + DiT bwd + AdamW) at a batch that fits comfortably.
Timing: CUDA events, median of 5 steps after 2 warm-ups, a 256 MB buffer written between steps (L2 flush).
FLOPs: F_fwd(N) of SURVEY 8d."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "open-diffusiongs_b200")):
    sys.path.insert(0, p)

from dgs_b200 import _lib, synth  # noqa: E402
from dgs_b200.denoiser import DGSDenoiser, DGSDenoiserScene  # noqa: E402
from dgs_b200.diffusion import transform_input  # noqa: E402

DEV = "cuda:0"
D, LAYERS, PATCH, V_IN = 1024, 24, 8, 4
CONFIGS = {  # name: (scene, H = W, V_render in training, training batch used here)
    "obj-256": (False, 256, 10, 4),
    "obj-512": (False, 512, 10, 1),
    "scene-256": (True, 256, 7, 8),
    "scene-512": (True, 512, 7, 1),
}


# --yaml: the per-GPU batch sizes of the shipped training yamls (data.batch_size: diffusionGS_rel.yaml:14 = 4,
# diffusionGS_rel_512.yaml:14 = 4, diffusionGS_scene.yaml:16 = 24, diffusionGS_scene_512.yaml:16 = 12), run with activation
# recompute (the reference checkpoints every block) as (micro-batch, micro-batches per optimizer step)
YAML_BATCH = {"obj-256": (4, 1), "obj-512": (2, 2), "scene-256": (8, 3), "scene-512": (4, 3)}


def f_fwd(n):
    return LAYERS * (24 * n * D * D + 4 * n * n * D + 12 * D * D) + 2 * (n - 2) * D * (576 + 896)


def make_inputs(B, res, v_render, seed=0):
    g = torch.Generator(DEV).manual_seed(seed)
    images = torch.rand(B, V_IN, 3, res, res, device=DEV, generator=g)
    images[:, 1:] = torch.randn(B, V_IN - 1, 3, res, res, device=DEV, generator=g)
    c2w, fx = synth.orbit_cameras(max(v_render, V_IN), res, res, az_step=36.0 if v_render > V_IN else None)
    c2w = torch.tensor(np.repeat(c2w[None], B, 0), device=DEV)
    fx = torch.tensor(np.repeat(fx[None], B, 0), device=DEV)
    ray_o, ray_d = transform_input(images, c2w[:, :V_IN].contiguous(), fx[:, :V_IN].contiguous())
    t = torch.full((B,), 500, device=DEV, dtype=torch.int64)
    target = torch.rand(B, c2w.shape[1], 3, res, res, device=DEV, generator=g)
    return dict(images=images, ray_o=ray_o, ray_d=ray_d, t=t, c2w=c2w, fx=fx, target=target)


def timed(fn, flush, steps=5, warmup=2, profile=True):
    L = _lib.lib()
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    fam = {}
    if profile:
        L.dgs_profile_enable(1)
        _lib.profile_read()
        fn()
        torch.cuda.synchronize()
        fam = {k: round(v[0], 3) for k, v in _lib.profile_read().items() if v[1]}
        L.dgs_profile_enable(0)
    return ms, fam


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "gpurun_out/perf_configs.json"
    train = "--train" in sys.argv
    only = None
    if "--only" in sys.argv:
        only = sys.argv[sys.argv.index("--only") + 1].split(",")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    rows = []
    t_start = time.time()
    models = {}
    for name, (scene, res, v_render, b_train) in CONFIGS.items():
        if only and name not in only:
            continue
        if scene not in models:
            torch.manual_seed(0)
            cfg = dict(patch_size=PATCH, ray_pe_type="plk" if scene else "relative_plk")
            models[scene] = (DGSDenoiserScene if scene else DGSDenoiser)(cfg).to(DEV)
        model = models[scene]
        n_tok = 2 + V_IN * (res // PATCH) ** 2
        # ---- inference: one denoise step, batch 1 ----
        model.eval()
        b = make_inputs(1, res, V_IN)

        def step():
            out, _ = model.image_to_gaussians(b["images"], b["ray_o"], b["ray_d"], b["t"])
            return model.render_gaussians(out, b["c2w"][:, :V_IN], b["fx"][:, :V_IN], res, res)
        img = step()
        assert torch.isfinite(img).all()
        ms, fam = timed(step, flush)
        dit_ms = sum(v for k, v in fam.items() if k.startswith("dit."))
        row = dict(config=name, mode="denoise", batch=1, res=res, tokens=n_tok, gaussians=2 + V_IN * res * res,
                   ms_per_step=ms, steps_per_s=1e3 / ms, views_per_s=V_IN * 1e3 / ms,
                   instances=int(getattr(model.gs_renderer, "last_num_rendered", 0) or 0),
                   dit_ms=dit_ms, raster_ms=sum(v for k, v in fam.items() if k.startswith("raster.")),
                   dit_tflops=f_fwd(n_tok) / dit_ms / 1e9 if dit_ms else None, families=fam,
                   mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
        print(json.dumps(row), flush=True)
        rows.append(row)
        if "--yaml" in sys.argv:
            # ---- training step at the yaml batch size: recompute + micro-batches with gradient accumulation ----
            from dgs_b200 import losses
            from dgs_b200.train import DitTrainer
            mb, k = YAML_BATCH[name]
            torch.manual_seed(0)
            cfg = dict(patch_size=PATCH, ray_pe_type="plk" if scene else "relative_plk")
            m2 = (DGSDenoiserScene if scene else DGSDenoiser)(cfg).to(DEV)
            tr = DitTrainer(m2, lr=1e-5, betas=(0.9, 0.99), eps=1e-8, clip=0.5, recompute=True, accumulate_grad_batches=k,
                            ema_decay=0.9999)
            m2.train()
            bts = [make_inputs(mb, res, v_render, seed=10 + i) for i in range(k)]

            def ystep():
                tot = 0.0
                for bt in bts:
                    out, _ = m2.image_to_gaussians(bt["images"], bt["ray_o"], bt["ray_d"], bt["t"])
                    res_, _ = losses.fused_render_and_loss(m2, out, bt["c2w"], bt["fx"], res, res, bt["target"])
                    res_["loss"].backward()
                    tot = tot + res_["loss"].detach()
                tr.optimizer_step(allreduce=False)
                return tot / k
            torch.cuda.reset_peak_memory_stats()
            loss = ystep()
            assert torch.isfinite(loss)
            ms, fam = timed(ystep, flush, steps=2, warmup=1)
            dit_ms = sum(v for kk, v in fam.items() if kk.startswith("dit."))
            row = dict(config=name, mode="train-yaml-batch", batch=mb * k, micro_batch=mb, micro_batches=k, activations="recompute",
                       res=res, tokens=n_tok, render_views=v_render, ms_per_step=ms, samples_per_s=mb * k * 1e3 / ms, dit_ms=dit_ms,
                       raster_ms=sum(v for kk, v in fam.items() if kk.startswith("raster.")),
                       dit_train_tflops_algorithmic=3 * mb * k * f_fwd(n_tok) / dit_ms / 1e9 if dit_ms else None, loss=float(loss),
                       families=fam, mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
            print(json.dumps(row), flush=True)
            rows.append(row)
            m2._trainer = None
            del tr, m2, bts
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            continue
        if not train:
            continue
        # ---- training step ----
        from dgs_b200.train import DitTrainer
        if getattr(model, "_trainer", None) is None:
            DitTrainer(model, lr=1e-5, betas=(0.9, 0.99), eps=1e-8, clip=0.5)
        trainer = model._trainer
        model.train()
        B = b_train
        bt = make_inputs(B, res, v_render, seed=1)

        def tstep():
            out, _ = model.image_to_gaussians(bt["images"], bt["ray_o"], bt["ray_d"], bt["t"])
            renders = model.render_gaussians(out, bt["c2w"], bt["fx"], res, res)
            loss = ((renders - bt["target"]) ** 2).mean()
            trainer.zero_grad()
            loss.backward()
            trainer.optimizer_step(allreduce=False)
            return loss
        torch.cuda.reset_peak_memory_stats()
        loss = tstep()
        assert torch.isfinite(loss)
        ms, fam = timed(tstep, flush, steps=3, warmup=1)
        dit_ms = sum(v for k, v in fam.items() if k.startswith("dit."))
        row = dict(config=name, mode="train", batch=B, res=res, tokens=n_tok, render_views=v_render, ms_per_step=ms,
                   samples_per_s=B * 1e3 / ms, dit_ms=dit_ms,
                   raster_ms=sum(v for k, v in fam.items() if k.startswith("raster.")),
                   dit_train_tflops=3 * B * f_fwd(n_tok) / dit_ms / 1e9 if dit_ms else None, loss=float(loss), families=fam,
                   mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
        print(json.dumps(row), flush=True)
        rows.append(row)
        trainer._state = None  # release this config's stored activations before the next one
        del bt
        torch.cuda.empty_cache()
    json.dump(dict(rows=rows, wall_s=time.time() - t_start), open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
