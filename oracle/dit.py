"""oracle/dit.py -- TEST INFRASTRUCTURE: plain-PyTorch fp32 restatement of the reference denoiser.

Follows diffusionGS/models/denoiser/denoiser.py:21-22,26-72,76-164,199-253,306-416 (object model),
denoiser_scene.py:232-263,314-429 (scene differences) and
diffusionGS/models/transformers/utils_transformer.py:26-36,246-290 (DiTBlock).
The block's arithmetic lives in a third-party dependency that is NOT under /root/reference:
timm==0.9.16 (requirement.txt:25) `timm.models.vision_transformer.Attention` and `Mlp`; their published
forward is restated here (qkv Linear -> reshape [B,N,3,H,hd] -> softmax(q k^T / sqrt(hd)) v -> proj;
fc1 -> GELU(tanh) -> fc2; dropout 0; qk_norm off).
PINNED (round 2): tests/test_oracle_dit_vs_reference.py executes the reference's OWN denoiser.py /
denoiser_scene.py / utils_transformer.py (loaded by path, tests/golden/ref_import.py; only the absent third-party
packages are stubbed -- timm's Attention there is an independent restatement over F.scaled_dot_product_attention)
and holds this file to it at 1e-6 (object + scene model, both ray_pe_type values); the same run's inputs/outputs are
committed as tests/golden/dit_ref_*.npz (tests/golden/make_dit_golden.py) for boxes without /root/reference.
Same module tree / state_dict keys as the reference so a reference checkpoint loads with strict=True.
Never imported by the product path.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


class Attention(nn.Module):  # timm 0.9.16 semantics
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        att = (q * self.head_dim ** -0.5) @ k.transpose(-2, -1)
        x = att.softmax(dim=-1) @ v
        return self.proj(x.transpose(1, 2).reshape(B, N, C))


class Mlp(nn.Module):  # timm 0.9.16 semantics, act = GELU(tanh)
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x), approximate="tanh"))


class DiTBlock(nn.Module):
    def __init__(self, dim, heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.attn = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(dim, 6 * dim, bias=True))

    def forward(self, x, c):
        s1, c1, g1, s2, c2, g2 = self.adaLN_modulation(c).chunk(6, dim=1)
        x = x + g1.unsqueeze(1) * self.attn(modulate(self.norm1(x), s1, c1))
        x = x + g2.unsqueeze(1) * self.mlp(modulate(self.norm2(x), s2, c2))
        return x


class TimestepEmbedder(nn.Module):
    def __init__(self, hidden, freq=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(freq, hidden), nn.SiLU(), nn.Linear(hidden, hidden))
        self.freq = freq

    def forward(self, t):
        half = self.freq // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
        args = t[:, None].float() * freqs[None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        return self.mlp(emb.to(self.mlp[0].weight.dtype))


class _Head(nn.Module):
    def __init__(self, dim, out):
        super().__init__()
        self.layernorm = nn.LayerNorm(dim, bias=False)
        self.linear = nn.Linear(dim, out, bias=False)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(dim, 2 * dim, bias=True))

    def forward(self, x, c):
        shift, scale = self.adaLN_modulation(c).chunk(2, dim=1)
        return self.linear(modulate(self.layernorm(x), shift, scale))


def _init_linear(m):
    if isinstance(m, nn.Linear):
        nn.init.normal_(m.weight, mean=0.0, std=0.02)
        if m.bias is not None:
            nn.init.zeros_(m.bias)


class DenoiserOracle(nn.Module):
    def __init__(self, width=1024, heads=16, layers=24, patch=8, n_gaussians=2, scene=False, near=0.0, far=500.0,
                 ray_pe_type=None):
        super().__init__()
        self.width, self.patch, self.G, self.scene, self.near, self.far = width, patch, n_gaussians, scene, near, far
        # yaml defaults: object configs leave the class default 'relative_plk' (denoiser.py:186), scene configs set 'plk'
        self.ray_pe_type = ray_pe_type or ("plk" if scene else "relative_plk")
        self.t_embedder = TimestepEmbedder(width)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        self.image_tokenizer = nn.Sequential(nn.Identity(), nn.Linear(9 * patch * patch, width, bias=False))
        self.image_tokenizer.apply(_init_linear)
        shape = (1, n_gaussians, width) if scene else (n_gaussians, width)
        self.gaussians_pos_embedding = nn.Parameter(torch.randn(*shape))
        nn.init.trunc_normal_(self.gaussians_pos_embedding, std=0.02)
        self.transformer_input_layernorm = nn.LayerNorm(width, bias=False)
        self.transformer = nn.ModuleList([DiTBlock(width, heads) for _ in range(layers)])
        self.transformer.apply(_init_linear)
        self.upsampler = _Head(width, 14)
        self.upsampler.apply(_init_linear)
        self.image_token_decoder = _Head(width, patch * patch * 14)
        self.image_token_decoder.apply(_init_linear)

    def image_to_gaussians(self, images, ray_o, ray_d, t, return_tokens=False):
        p = self.patch
        o_dot_d = torch.sum(-ray_o * ray_d, dim=2, keepdim=True)
        if self.ray_pe_type == "relative_plk":  # denoiser.py:312-322 == denoiser_scene.py:319-331
            posed = torch.cat([images[:, :, :3] * 2.0 - 1.0, ray_d, ray_o + o_dot_d * ray_d], dim=2)
        else:
            posed = torch.cat([images[:, :, :3] * 2.0 - 1.0, torch.cross(ray_o, ray_d, dim=2), ray_d], dim=2)
        b, v, c, h, w = posed.shape
        # "b v c (hh ph) (ww pw) -> (b v) (hh ww) (ph pw c)"
        tok = posed.reshape(b, v, c, h // p, p, w // p, p).permute(0, 1, 3, 5, 4, 6, 2).reshape(b * v, -1, p * p * c)
        tok = self.image_tokenizer(tok).reshape(b, -1, self.width)
        temb = self.t_embedder(t)
        pos = self.gaussians_pos_embedding.reshape(self.G, self.width).expand(b, -1, -1)
        x = self.transformer_input_layernorm(torch.cat((pos, tok), dim=1))
        for blk in self.transformer:
            x = blk(x, temb)
        tokens = x
        g_tok, i_tok = x.split([self.G, x.shape[1] - self.G], dim=1)
        gaussians = self.upsampler(g_tok, temb)
        img_g = self.image_token_decoder(i_tok, temb).reshape(b, -1, 14)
        allg = torch.cat((gaussians, img_g), dim=1)
        xyz, features, scaling, rotation, opacity = allg.split([3, 3, 3, 4, 1], dim=2)
        features = features.reshape(b, -1, 1, 3)
        scaling = (scaling - 2.3).clamp(max=-1.20)
        opacity = opacity - 2.0
        n_img = img_g.shape[1]
        ia = xyz[:, -n_img:, :].reshape(b, v, h // p, w // p, p, p, 3).permute(0, 1, 6, 2, 4, 3, 5).reshape(b, v, 3, h, w)
        ia = ia.mean(dim=2, keepdim=True)
        if self.scene:  # denoiser_scene.py:263,406-410 (range_func, whatever ray_pe_type is)
            depth = torch.sigmoid(ia) * (self.far - self.near) + self.near
        elif self.ray_pe_type == "relative_plk":  # denoiser.py:381-388
            depth = (2.0 * torch.sigmoid(ia) - 1.0) * 1.8 + o_dot_d
        else:
            depth = torch.sigmoid(ia)
        ia = ray_o + depth * ray_d
        ia_flat = ia.reshape(b, v, 3, h // p, p, w // p, p).permute(0, 1, 3, 5, 4, 6, 2).reshape(b, -1, 3)
        xyz = torch.cat((xyz[:, :-n_img, :], ia_flat), dim=1)
        out = dict(xyz=xyz, features=features, scaling=scaling, rotation=rotation, opacity=opacity)
        return (out, ia, tokens) if return_tokens else (out, ia)
