"""Input-stage rows (a1 TransformInput, a2 q_sample, a18 p_sample step): host tables on CPU, kernels on GPU."""
import numpy as np
import pytest
import torch

from dgs_b200 import diffusion as dd
from oracle import diffusion as od


def test_schedule_tables_match_oracle_and_known_answers():
    for resp in (None, "30", [10, 5]):
        a, b = dd.create_diffusion(resp), od.Tables(resp)
        assert a.timestep_map == b.timestep_map and a.num_timesteps == len(b.betas)
        for k, v in a.tables_f64.items():
            assert np.array_equal(v, getattr(b, k)), k
    d = dd.create_diffusion(None)
    import math
    f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
    # alpha_bar_t = f((t+1)/T) / f(0) until the 0.999 beta cap engages (only the very last steps)
    assert abs(d.tables_f64["sqrt_alphas_cumprod"][499] ** 2 - f(0.5) / f(0.0)) < 1e-12
    assert d.num_timesteps == 1000 and dd.create_diffusion("30").num_timesteps == 30
    assert dd.create_diffusion("30").timestep_map[0] == 0 and dd.create_diffusion("30").timestep_map[-1] == 999


@pytest.mark.gpu
def test_q_sample_and_p_sample_step_match_oracle():
    dev = "cuda:0"
    for resp in (None, "30"):
        d, tab = dd.create_diffusion(resp), od.Tables(resp)
        g = torch.Generator(dev).manual_seed(0)
        x0 = torch.rand(4, 3, 3, 64, 64, device=dev, generator=g)
        noise = torch.randn(4, 3, 3, 64, 64, device=dev, generator=g)
        t = torch.tensor([0, 1, d.num_timesteps // 2, d.num_timesteps - 1], device=dev)
        a = d.q_sample(x0, t, noise)
        b = od.q_sample(tab, x0, t, noise)
        assert float((a - b).abs().max()) <= 1e-6
        xt = torch.randn(4, 3, 3, 64, 64, device=dev, generator=g)
        a = d.p_sample_step(x0, xt, t, noise)
        b = od.p_sample_step(tab, x0, xt, t, noise)
        assert float((a - b).abs().max()) <= 2e-6 * float(b.abs().max())
        assert torch.equal(d.map_timesteps(torch.tensor([0, d.num_timesteps - 1], device=dev)).cpu(),
                           torch.tensor([tab.timestep_map[0], tab.timestep_map[-1]]))


@pytest.mark.gpu
def test_transform_input_matches_oracle():
    from dgs_b200 import synth
    dev = "cuda:0"
    c2w, fx = synth.orbit_cameras(4, 96, 64)
    c2w = torch.tensor(np.stack([c2w, c2w[::-1].copy()]), device=dev)
    fx = torch.tensor(np.stack([fx, fx * 1.1]), device=dev)
    img = torch.zeros(2, 4, 3, 64, 96, device=dev)
    ro, rd = dd.transform_input(img, c2w, fx)
    ro2, rd2 = od.transform_input(img, c2w, fx)
    assert ro.shape == (2, 4, 3, 64, 96) and torch.equal(ro, ro2)
    assert float((rd - rd2).abs().max()) < 2e-6


@pytest.mark.gpu
def test_sampler_loop_matches_oracle_loop():
    """p_sample_loop_progressive against the oracle's restatement of the reference loop, same fake
    denoiser on both sides (a cheap deterministic function of the conditioning view, x_t and the MAPPED timestep), same
    noise: every x_{t-1} of a 30-step respaced chain, the dict keys, and the timestep mapping."""
    dev = "cuda:0"
    d, tab = dd.create_diffusion("30"), od.Tables("30")
    g = torch.Generator(dev).manual_seed(3)
    B, V, H, W = 2, 4, 32, 48
    cond = torch.rand(B, 1, 3, H, W, device=dev, generator=g)
    x_T = torch.randn(B, V - 1, 3, H, W, device=dev, generator=g)
    noises = torch.randn(30, B, V - 1, 3, H, W, device=dev, generator=g)
    seen = []

    class Fake(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(1, device=dev))

        def forward(self, input_batch, timesteps):
            seen.append(timesteps.clone())
            img = input_batch["image"]
            assert img.shape == (B, V, 3, H, W)
            k = (timesteps.float() / 1000.0).view(B, 1, 1, 1, 1)
            renders = torch.tanh(0.7 * img + 0.2 * img.roll(1, dims=1)) * (1.0 - 0.5 * k) + 0.1
            return renders, ["gaussians"] * B
    fake = Fake()
    nf = lambda i, like: noises[i]  # noqa: E731
    ours = list(d.p_sample_loop_progressive(fake, x_T.shape, dict(image=cond.clone(), image_noisy=x_T.clone()),
                                            clip_denoised=False, noise_fn=nf))
    t_ours = [t.cpu() for t in seen]
    seen.clear()
    ref = list(od.p_sample_loop_progressive(tab, fake, x_T.shape, dict(image=cond.clone(), image_noisy=x_T.clone()),
                                            clip_denoised=False, noise_fn=nf))
    assert len(ours) == len(ref) == 30
    assert all(torch.equal(a, b.cpu()) for a, b in zip(t_ours, seen)) and int(t_ours[0][0]) == 999 and int(t_ours[-1][0]) == 0
    for a, b in zip(ours, ref):
        assert set(a) == set(b) == {"sample", "pred_xstart", "input_batch", "denoiser_output_dict"}
        assert float((a["sample"] - b["sample"]).abs().max()) <= 5e-6 * max(1.0, float(b["sample"].abs().max()))
    final = d.p_sample_loop(fake, x_T.shape, dict(image=cond.clone(), image_noisy=x_T.clone()), clip_denoised=False,
                            progress=False, noise_fn=nf)
    assert torch.equal(final["sample"], ours[-1]["sample"])
    # the last step (t = 0) adds no noise: x_0 = posterior mean = pred_xstart (coef1 = 1, coef2 = 0 at t = 0)
    assert float((final["sample"] - final["pred_xstart"]).abs().max()) < 1e-5


def test_sampler_loop_host_logic_without_gpu():
    """The loop's host side (step order, timestep mapping, dict plumbing) with the device kernel stubbed out."""
    d = dd.create_diffusion("30")
    calls = []

    class Stub(dd.GaussianDiffusionB200):
        pass
    d.__class__ = Stub
    Stub.p_sample_step = lambda self, p, x, t, noise=None: 0.5 * (p + x)
    Stub.map_timesteps = lambda self, ts: torch.as_tensor(self.timestep_map)[ts]

    class Fake(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(1))

        def forward(self, input_batch, timesteps):
            calls.append(int(timesteps[0]))
            return input_batch["image"] * 0.9, [None]
    x = torch.ones(1, 3, 3, 8, 8)
    outs = list(d.p_sample_loop_progressive(Fake(), x.shape, dict(image=torch.zeros(1, 1, 3, 8, 8), image_noisy=x),
                                            clip_denoised=True))
    assert len(outs) == 30 and calls == d.timestep_map[::-1]
    assert outs[-1]["input_batch"]["image"].shape == (1, 4, 3, 8, 8)
    assert float(outs[0]["sample"].max()) == pytest.approx(0.5 * (0.9 + 1.0))
