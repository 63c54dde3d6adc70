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
