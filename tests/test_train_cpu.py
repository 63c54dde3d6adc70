"""CPU checks of the training-side host logic (dgs_b200/train.py, denoiser packing): no GPU, no compute calls."""
import pytest
import torch


def _small():
    from dgs_b200.denoiser import DGSDenoiser
    torch.manual_seed(0)
    return DGSDenoiser(dict(patch_size=8, num_layers=3))


def test_trainer_requires_cuda_no_fallback():
    from dgs_b200 import _lib
    from dgs_b200.train import DitTrainer
    with pytest.raises(_lib.DgsError):
        DitTrainer(_small())


def test_pack_dict_fields_match_the_c_struct():
    from dgs_b200 import _lib
    m = _small()
    t = m._pack_dict()
    names = [n for n, _ in _lib.DitWeights._fields_][6:]
    assert sorted(t) == sorted(names)
    c = m.cfg
    assert t["qkv_w"].shape == (3, 3 * c.width, c.width) and t["qkv_w"].dtype == torch.bfloat16
    assert t["adaln_w"].shape == (3 * 6 * c.width + 4 * c.width, c.width) and t["adaln_w"].dtype == torch.float32
    assert t["dec_w"].shape == (c.patch_size ** 2 * 14, 3 * c.width)  # split-bf16 [hi|hi|lo]
    small = m._pack_dict(skip=("qkv_w", "proj_w", "fc1_w", "fc2_w"))
    assert "qkv_w" not in small and "qkv_b" in small


def test_blocks_share_one_stride_in_parameter_order():
    """dgs_dit_grads addresses block l at block 0 + l * layer_stride: true for a flat arena in module.parameters()
    order because every block owns the same parameter shapes in the same order."""
    from dgs_b200.dist import GradArena
    m = _small()
    arena = GradArena(m)
    T = m.transformer
    base = arena.flat.data_ptr()
    offs = [[(p.grad.data_ptr() - base) // 4 for p in blk.parameters()] for blk in T]
    strides = {tuple(b - a for a, b in zip(offs[0], offs[i])) for i in range(1, len(T))}
    assert all(len(set(s)) == 1 for s in strides)              # every tensor of a block moves by the same amount
    assert len({s[0] // i for i, s in enumerate(sorted(strides), 1)}) == 1
    per_block = sum(p.numel() for p in T[0].parameters())
    assert offs[1][0] - offs[0][0] == per_block


def test_grad_struct_field_order_matches_header():
    import re
    import os
    from dgs_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "dgs_b200.h")).read()
    body = hdr[hdr.index("typedef struct {  /* all fp32, OVERWRITTEN by dgs_dit_backward */"):hdr.index("} dgs_dit_grads;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(?:float\*|long long)\s+(\w+);", body)
    assert fields == [n for n, _ in _lib.DitGrads._fields_]
