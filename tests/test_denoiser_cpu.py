"""CPU tests of the denoiser host logic: registry names, state_dict surface, weight packing, and the
oracle's own consistency (fp32 vs fp64).  No GPU compute."""
import pytest
import torch

from dgs_b200 import denoiser as dn
from oracle.dit import DenoiserOracle


def test_registry_and_state_dict_surface():
    assert dn.find("diffusion-gs-model") is dn.DGSDenoiser
    assert dn.find("diffusion-gs-model-scene") is dn.DGSDenoiserScene
    m = dn.DGSDenoiser(dict(patch_size=8, num_layers=3))
    sd = m.state_dict()
    expect = {"t_embedder.mlp.0.weight": (1024, 256), "t_embedder.mlp.2.weight": (1024, 1024),
              "image_tokenizer.1.weight": (1024, 576), "gaussians_pos_embedding": (2, 1024),
              "transformer_input_layernorm.weight": (1024,), "transformer.0.attn.qkv.weight": (3072, 1024),
              "transformer.0.attn.qkv.bias": (3072,), "transformer.2.attn.proj.weight": (1024, 1024),
              "transformer.1.mlp.fc1.weight": (4096, 1024), "transformer.1.mlp.fc2.weight": (1024, 4096),
              "transformer.0.adaLN_modulation.1.weight": (6144, 1024), "upsampler.layernorm.weight": (1024,),
              "upsampler.linear.weight": (14, 1024), "upsampler.adaLN_modulation.1.bias": (2048,),
              "image_token_decoder.linear.weight": (896, 1024), "image_token_decoder.adaLN_modulation.1.weight": (2048, 1024)}
    for k, shp in expect.items():
        assert tuple(sd[k].shape) == shp, k
    assert "_dummy" not in sd and not any("norm1" in k or "norm2" in k for k in sd)
    assert dn.DGSDenoiserScene(dict(patch_size=8, num_layers=1)).state_dict()["gaussians_pos_embedding"].shape == (1, 2, 1024)
    o = DenoiserOracle(layers=3)
    assert o.load_state_dict(sd, strict=True)


def test_full_model_parameter_count():
    m = dn.DGSDenoiser(dict(patch_size=8, num_layers=1))
    per_block = sum(p.numel() for p in m.transformer[0].parameters())
    other = sum(p.numel() for p in m.parameters()) - per_block
    assert per_block == 18_889_728 and other + 24 * per_block == 460_391_424


def test_packing_layout_and_cache():
    m = dn.DGSDenoiser(dict(patch_size=8, num_layers=2))
    w, t = m.packed_weights()
    assert t["adaln_w"].shape == (2 * 6144 + 2 * 2048, 1024) and t["adaln_w"].dtype == torch.float32
    assert torch.equal(t["adaln_w"][6144:12288], m.transformer[1].adaLN_modulation[1].weight.detach())
    # split-bf16 [hi | hi | lo] reconstructs the fp32 weight to ~2^-17
    tw, w0 = t["tokenizer_w"].float(), m.image_tokenizer[1].weight.detach()
    assert tw.shape == (1024, 3 * 576) and torch.equal(tw[:, :576], tw[:, 576:1152])
    assert float(((tw[:, :576] + tw[:, 1152:]) - w0).abs().max() / w0.abs().max()) < 2e-5
    assert t["dec_w"].shape == (896, 3072) and t["ups_w"].shape == (14, 3072)
    assert torch.equal(t["adaln_b"][-2048:], m.image_token_decoder.adaLN_modulation[1].bias.detach())
    assert t["qkv_w"].shape == (2, 3072, 1024) and t["fc2_w"].shape == (2, 1024, 4096)
    assert w.layers == 2 and w.heads == 16 and w.patch == 8 and w.mlp_hidden == 4096
    assert m.packed_weights()[1] is t  # cached
    with torch.no_grad():
        m.transformer[0].attn.qkv.bias.add_(1.0)
    assert m.packed_weights()[1] is not t  # in-place update invalidates the cache


def test_cpu_call_fails_loudly():
    m = dn.DGSDenoiser(dict(patch_size=8, num_layers=1))
    z = torch.zeros(1, 2, 3, 16, 16)
    with pytest.raises(Exception, match="CUDA device only"):
        m.image_to_gaussians(z, z, z, torch.tensor([1]))


def test_oracle_fp32_vs_fp64():
    torch.manual_seed(0)
    o = DenoiserOracle(layers=2)
    img, ro = torch.rand(1, 2, 3, 16, 16), torch.randn(1, 2, 3, 16, 16)
    rd = torch.nn.functional.normalize(torch.randn(1, 2, 3, 16, 16), dim=2)
    t = torch.tensor([321])
    a, _ = o.image_to_gaussians(img, ro, rd, t)
    b, _ = o.double().image_to_gaussians(img.double(), ro.double(), rd.double(), t)
    for k in a:
        assert float((a[k].double() - b[k]).norm() / b[k].norm()) < 1e-5, k


def test_reference_yaml_config_builds_and_loads_a_576_column_tokenizer():
    """every shipped reference yaml sets in_channels: 9 and the tokenizer is
    Linear(in_channels * patch^2, width) (denoiser.py:216-221) -> image_tokenizer.1.weight is [1024, 576] in released
    and Lightning checkpoints.  Build from the yaml's values and strictly load such a state_dict."""
    import pytest
    import torch
    from dgs_b200 import denoiser as dn
    yaml_obj = dict(width=1024, in_channels=9, patch_size=8, n_gaussians=2, dim_heads=64, num_layers=1,
                    prior_distribution="gaussian", use_flash=True, use_checkpoint=True)       # diffusionGS_rel.yaml:26-36
    yaml_scene = dict(yaml_obj, range_setting_near=0, range_setting_far=500, ray_pe_type="plk")  # diffusionGS_scene.yaml:28-40
    for cls, cfg in ((dn.DGSDenoiser, yaml_obj), (dn.DGSDenoiserScene, yaml_scene)):
        m = cls(cfg)
        assert tuple(m.image_tokenizer[1].weight.shape) == (1024, 576)
        sd = {k: torch.zeros_like(v) for k, v in m.state_dict().items()}
        assert sd["image_tokenizer.1.weight"].shape == (1024, 9 * 64)
        m.load_state_dict(sd, strict=True)
    with pytest.raises(ValueError, match="in_channels"):
        dn.DGSDenoiser(dict(yaml_obj, in_channels=3))
    with pytest.raises(ValueError, match="ray_pe_type"):
        dn.DGSDenoiser(dict(yaml_obj, ray_pe_type="abs"))
