"""Checkpoint-format compatibility (reference: pipline_obj.py:66-71, systems/base.py:51-57,
utils/misc.py:40-70, denoiser.py:259-268): pure host logic, no GPU."""
import io

import pytest
import torch

from dgs_b200 import checkpoint as ck
from dgs_b200.denoiser import DGSDenoiser, DGSDenoiserScene
from oracle.dit import DenoiserOracle


def _model(seed, scene=False):
    torch.manual_seed(seed)
    return (DGSDenoiserScene if scene else DGSDenoiser)(dict(patch_size=8, num_layers=2))


def test_three_layouts_reduce_to_the_same_state_dict():
    src = _model(0)
    bare = {k: v.clone() for k, v in src.state_dict().items()}
    lightning = {"state_dict": {**{"shape_model." + k: v for k, v in bare.items()},
                                "loss_computer.lpips_loss_module.net.slice1.0.weight": torch.zeros(3)},
                 "epoch": 7, "global_step": 1234}
    release = {"model": {**{"denoiser." + k: v for k, v in bare.items()}, "denoiser.loss_computer.w": torch.zeros(1)}}
    for obj, nign in ((bare, 0), (lightning, 1), (release, 1)):
        sd, meta, ignored = ck.extract_denoiser_state_dict(obj)
        assert set(sd) == set(bare) and len(ignored) == nign
        assert all(torch.equal(sd[k], bare[k]) for k in bare)
    assert ck.extract_denoiser_state_dict(lightning)[1] == {"epoch": 7, "global_step": 1234}
    with pytest.raises(ValueError):
        ck.extract_denoiser_state_dict({})


def test_round_trip_through_the_lightning_layout_and_the_oracle():
    src, dst = _model(1), _model(2)
    assert not torch.equal(src.transformer[0].attn.qkv.weight, dst.transformer[0].attn.qkv.weight)
    buf = io.BytesIO()
    torch.save(ck.system_checkpoint(src, epoch=3, global_step=99, extra_state_dict={"loss_computer.x": torch.ones(2)}), buf)
    buf.seek(0)
    obj = torch.load(buf, weights_only=False)
    assert all(k.startswith(("shape_model.", "loss_computer.")) for k in obj["state_dict"])
    meta = ck.load_checkpoint(dst, obj)
    assert meta == {"epoch": 3, "global_step": 99}
    for (ka, a), (kb, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert ka == kb and torch.equal(a, b)
    # the reference-side readers: load_module_weights(module_name="shape_model") and a strict load into the restated model
    sd, epoch, step = ck.module_weights(obj, "shape_model")
    assert (epoch, step) == (3, 99)
    DenoiserOracle(layers=2).load_state_dict(sd, strict=True)


def test_half_precision_release_and_strictness():
    src, dst = _model(3, scene=True), _model(4, scene=True)
    half = {"model": {"denoiser." + k: v.half() for k, v in src.state_dict().items()}}
    ck.load_checkpoint(dst, half)
    assert dst.transformer[1].mlp.fc1.weight.dtype == torch.float32
    assert torch.equal(dst.transformer[1].mlp.fc1.weight, src.transformer[1].mlp.fc1.weight.half().float())
    broken = dict(src.state_dict())
    broken.pop("upsampler.linear.weight")
    with pytest.raises(RuntimeError, match="missing"):
        ck.load_checkpoint(dst, broken)
    ck.load_checkpoint(dst, broken, strict=False)
    # object model (pos embedding [2, w]) vs scene model ([1, 2, w]) are NOT interchangeable: shapes are checked by torch
    with pytest.raises(RuntimeError):
        ck.load_checkpoint(_model(5, scene=False), src.state_dict())
