"""oracle/diffusion.py -- TEST INFRASTRUCTURE: restatement of the reference's diffusion math and ray generation.

Follows diffusionGS/models/diffusion/gaussian_diffusion.py:122-167 (squaredcos_cap_v2 schedule), 183-243 (tables),
268-284 (q_sample), 291-312 (posterior), 380-392 (FIXED_LARGE variance), 505-516 (ancestral step),
respace.py:16-66,77-102 (timestep respacing) and diffusionGS/systems/utils.py:621-757 (TransformInput, patch_size=None).
numpy fp64 tables, torch fp32 tensors, exactly like the reference.  Never imported by the product path."""
import math

import numpy as np
import torch


def cosine_betas(n=1000, max_beta=0.999):
    ab = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
    return np.array([min(1 - ab((i + 1) / n) / ab(i / n), max_beta) for i in range(n)], dtype=np.float64)


def space_timesteps(num_timesteps, section_counts):
    if isinstance(section_counts, str):
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start, steps = 0, []
    for i, cnt in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        stride = 1 if cnt <= 1 else (size - 1) / (cnt - 1)
        cur = 0.0
        for _ in range(cnt):
            steps.append(start + round(cur))
            cur += stride
        start += size
    return set(steps)


class Tables:
    def __init__(self, respacing=None, n=1000):
        base = cosine_betas(n)
        use = set(range(n)) if not respacing else space_timesteps(n, respacing)
        ac = np.cumprod(1.0 - base)
        last, betas, self.timestep_map = 1.0, [], []
        for i, a in enumerate(ac):
            if i in use:
                betas.append(1 - a / last)
                last = a
                self.timestep_map.append(i)
        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas)
        prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.posterior_variance = betas * (1.0 - prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef1 = betas * np.sqrt(prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self.model_log_variance = np.log(np.append(self.posterior_variance[1], betas[1:]))  # FIXED_LARGE


def _extract(arr, t, shape):
    res = torch.from_numpy(arr).to(t.device)[t].float()
    while res.dim() < len(shape):
        res = res[..., None]
    return res


def q_sample(tab, x_start, t, noise):
    return _extract(tab.sqrt_alphas_cumprod, t, x_start.shape) * x_start + \
        _extract(tab.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise


def p_sample_step(tab, pred_xstart, x_t, t, noise):
    mean = _extract(tab.posterior_mean_coef1, t, x_t.shape) * pred_xstart + \
        _extract(tab.posterior_mean_coef2, t, x_t.shape) * x_t
    nz = (t != 0).float().view(-1, *([1] * (x_t.dim() - 1)))
    return mean + nz * torch.exp(0.5 * _extract(tab.model_log_variance, t, x_t.shape)) * noise


def transform_input(image, c2w, fxfycxcy):
    b, v, c, h, w = image.shape
    fx = fxfycxcy.reshape(b * v, 4)
    m = c2w.reshape(b * v, 4, 4)
    y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    x = x[None].expand(b * v, -1, -1).reshape(b * v, -1).to(m.device)
    y = y[None].expand(b * v, -1, -1).reshape(b * v, -1).to(m.device)
    x = (x + 0.5 - fx[:, 2:3]) / fx[:, 0:1]
    y = (y + 0.5 - fx[:, 3:4]) / fx[:, 1:2]
    d = torch.stack([x, y, torch.ones_like(x)], dim=2)
    d = torch.bmm(d.to(m), m[:, :3, :3].transpose(1, 2))
    d = d / torch.norm(d, dim=2, keepdim=True)
    o = m[:, :3, 3][:, None, :].expand_as(d)
    return (o.reshape(b, v, h, w, 3).permute(0, 1, 4, 2, 3).contiguous(),
            d.reshape(b, v, h, w, 3).permute(0, 1, 4, 2, 3).contiguous())


def p_sample_loop_progressive(tab, model, shape, input_batch, clip_denoised=False, noise_fn=None):
    """Restatement of p_sample_loop_progressive / p_sample / p_mean_variance (gaussian_diffusion.py:560-603, 479-518,
    316-459) and of `_WrappedModel.__call__` (respace.py:121-137: the model sees timestep_map[t]) for the START_X /
    FIXED_LARGE configuration.  `model(input_batch, mapped_t) -> (render_imgs [b, v, 3, h, w], gaussians)`."""
    n = len(tab.betas)
    tmap = torch.tensor(tab.timestep_map)
    for i in list(range(n))[::-1]:
        x = input_batch["image_noisy"]
        t = torch.tensor([i] * shape[0], device=x.device)
        input_batch["image"] = torch.cat([input_batch["image"][:, 0:1], input_batch["image_noisy"]], dim=1)
        render_imgs, gaussians = model(input_batch, tmap.to(x.device)[t])
        pred_xstart = render_imgs[:, 1:]
        if clip_denoised:
            pred_xstart = pred_xstart.clamp(-1, 1)
        noise = torch.randn_like(x) if noise_fn is None else noise_fn(i, x)
        sample = p_sample_step(tab, pred_xstart, x, t, noise)
        input_batch["image_noisy"] = sample
        yield dict(sample=sample, pred_xstart=pred_xstart, input_batch=input_batch,
                   denoiser_output_dict=dict(render_images=render_imgs, pred_gaussians=gaussians))
