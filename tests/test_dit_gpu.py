"""GPU parity tests of the DiT kernels (tcgen05 GEMM + epilogues, tcgen05 attention, LN+modulate) and of the
whole DGSDenoiser.image_to_gaussians against the fp32 PyTorch oracle (oracle/dit.py).
Tolerance (BASELINE north_star): 1e-3 relative in bf16, measured norm-wise against fp32 on the SAME
(bf16-representable where the kernel consumes bf16) inputs; the exact bound per check is written below."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(A, W, bias, epi, gate=None, x=None, rows_per_sample=1, gate_stride=0):
    from dgs_b200 import _lib
    M, K = A.shape
    N = W.shape[0]
    if epi in (0, 1):
        out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    elif epi == 2:
        out = x.clone()
    else:
        out = torch.empty(M, N, dtype=torch.float32, device=DEV)
    _lib.check(_lib.lib().dgs_gemm_bf16(A.data_ptr(), W.data_ptr(), None if bias is None else bias.data_ptr(),
                                        None if gate is None else gate.data_ptr(), out.data_ptr(), M, N, K, epi, N,
                                        gate_stride, rows_per_sample, stream()))
    return out


SHAPES = [(4098, 3072, 1024), (4098, 1024, 1024), (4098, 4096, 1024), (4098, 1024, 4096), (4096, 896, 1024),
          (4096, 1024, 576), (8196, 3072, 1024), (130, 128, 64), (1, 32, 8), (257, 160, 200)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_fp32_out_exact_products(M, N, K):
    """bf16 x bf16 products are exact in fp32, so only the accumulation order differs from torch: ~1e-6."""
    g = torch.Generator(DEV).manual_seed(M + N + K)
    A = torch.randn(M, K, device=DEV, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV, generator=g)
    ref = A.float() @ W.float().t() + bias
    out = gemm(A, W, bias, 3)
    torch.cuda.synchronize()
    e = rel(out, ref)
    print(f"gemm {M}x{N}x{K} fp32-out rel={e:.2e}")
    assert e < 2e-5
    out_nb = gemm(A, W, None, 3)
    assert rel(out_nb, ref - bias) < 2e-5


@pytest.mark.parametrize("M,N,K", [(4098, 3072, 1024), (4098, 4096, 1024), (300, 256, 128)])
def test_gemm_bf16_epilogues(M, N, K):
    g = torch.Generator(DEV).manual_seed(1)
    A = torch.randn(M, K, device=DEV, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    acc = A.float() @ W.float().t() + bias
    o0 = gemm(A, W, bias, 0).float()
    o1 = gemm(A, W, bias, 1).float()
    ref1 = torch.nn.functional.gelu(acc, approximate="tanh")
    # outputs are rounded to bf16 (relative step 2^-8): norm-wise error ~ 2^-9/sqrt(3) ~ 1.1e-3 at most
    print(f"bias->bf16 rel={rel(o0, acc):.2e}  gelu->bf16 rel={rel(o1, ref1):.2e}")
    assert rel(o0, acc) < 2.5e-3 and rel(o1, ref1) < 2.5e-3
    assert rel(o0, acc.to(torch.bfloat16).float()) < 2e-4  # equal to rounding the fp32 result, up to ties


def test_gemm_gate_residual_epilogue():
    B, Nt, K, N = 3, 1370, 512, 1024
    g = torch.Generator(DEV).manual_seed(2)
    A = torch.randn(B * Nt, K, device=DEV, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    x = torch.randn(B * Nt, N, device=DEV, generator=g)
    mod = torch.randn(B, 3 * N, device=DEV, generator=g)
    gate = mod[:, N:2 * N]
    ref = x + gate.repeat_interleave(Nt, 0) * (A.float() @ W.float().t() + bias)
    # gate pointer inside a wider row (as in the adaLN table): stride = full row, offset = one chunk
    out = gemm(A, W, bias, 2, gate=mod[:, N:], x=x, rows_per_sample=Nt, gate_stride=mod.stride(0))
    print(f"gate+residual rel={rel(out, ref):.2e}")
    assert rel(out, ref) < 2e-5


# (1, 2050, 20) / (1, 1500, 16) / (3, 4098, 16): ragged query/key tails at other head counts and batch sizes
# (1, 16386, 2): the 512x512 configurations (obj-512 / scene-512 / the pipline_obj.py demo), two heads bound the fp32 reference
@pytest.mark.parametrize("B,N,H", [(1, 4098, 16), (2, 1026, 16), (1, 128, 2), (1, 130, 1), (2, 77, 4), (1, 1, 1), (1, 16386, 2),
                                   (1, 2050, 20), (1, 1500, 16), (3, 4098, 16)])
def test_attention_vs_fp32_softmax(B, N, H):
    from dgs_b200 import _lib
    g = torch.Generator(DEV).manual_seed(N)
    qkv = (torch.randn(B, N, 3, H, 64, device=DEV, generator=g) * 1.5).to(torch.bfloat16)
    out = torch.zeros(B, N, H * 64, dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().dgs_attention_fwd(qkv.data_ptr(), out.data_ptr(), B, N, H, stream()))
    q, k, v = [t.float().permute(0, 2, 1, 3) for t in qkv.unbind(2)]  # [B,H,N,64]
    ref = torch.softmax((q @ k.transpose(-1, -2)) * 0.125, dim=-1) @ v
    ref = ref.permute(0, 2, 1, 3).reshape(B, N, H * 64)
    e = rel(out.float(), ref)
    print(f"attention B={B} N={N} H={H}: rel={e:.2e}")
    # Operator-level bound, not the north-star one: the inputs here are N(0, 1.5^2) q/k (logit std 2.25, peaky rows) and
    # BOTH P (before the PV MMA) and the output are rounded to bf16 (2^-9 relative each, uncorrelated): ~2e-3 norm-wise.
    # Against the fp32 result rounded to bf16 the kernel must be within the P rounding alone.
    assert e < 3e-3
    assert rel(out.float(), ref.to(torch.bfloat16).float()) < 2.5e-3


def test_ln_modulate():
    from dgs_b200 import _lib
    B, R, D = 2, 515, 1024
    g = torch.Generator(DEV).manual_seed(5)
    x = torch.randn(B, R, D, device=DEV, generator=g) * 3 + 1
    mod = torch.randn(B, 6 * D, device=DEV, generator=g)
    lnw = torch.randn(D, device=DEV, generator=g)
    h = torch.empty(B, R, D, dtype=torch.bfloat16, device=DEV)
    for w_, eps in ((None, 1e-6), (lnw, 1e-5)):
        _lib.check(_lib.lib().dgs_ln_modulate(x.data_ptr(), None if w_ is None else w_.data_ptr(), mod.data_ptr(),
                                              mod[:, D:].data_ptr(), 6 * D, h.data_ptr(), B, R, D, eps, stream()))
        ln = torch.nn.functional.layer_norm(x, (D,), w_, None, eps)
        ref = ln * (1 + mod[:, None, D:2 * D]) + mod[:, None, :D]
        assert rel(h.float(), ref) < 2.5e-3
        assert rel(h.float(), ref.to(torch.bfloat16).float()) < 3e-4


def _inputs(B, V, H, W, seed=0):
    g = torch.Generator(DEV).manual_seed(seed)
    images = torch.rand(B, V, 3, H, W, device=DEV, generator=g)
    images[:, 1:] = torch.randn(B, V - 1, 3, H, W, device=DEV, generator=g)
    ray_o = torch.randn(B, V, 3, 1, 1, device=DEV, generator=g).expand(B, V, 3, H, W).contiguous() * 1.5
    ray_d = torch.nn.functional.normalize(torch.randn(B, V, 3, H, W, device=DEV, generator=g), dim=2)
    t = torch.randint(0, 1000, (B,), device=DEV, generator=g)
    return images, ray_o, ray_d, t


def _compare_models(model, oracle, B, V, H, W, tag):
    images, ray_o, ray_d, t = _inputs(B, V, H, W)
    with torch.no_grad():
        ref, ref_xyz, ref_tok = oracle.image_to_gaussians(images, ray_o, ray_d, t, return_tokens=True)
        out, xyz_img, tok = model.image_to_gaussians(images, ray_o, ray_d, t, return_tokens=True)
    torch.cuda.synchronize()
    errs = {k: rel(out[k], ref[k]) for k in ref}
    errs["tokens"] = rel(tok, ref_tok)
    errs["img_aligned_xyz"] = rel(xyz_img, ref_xyz)
    print(f"[{tag}] " + "  ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    return errs


@pytest.mark.parametrize("scene", [False, True])
def test_denoiser_small_vs_oracle(scene):
    from dgs_b200.denoiser import DGSDenoiser, DGSDenoiserScene
    from oracle.dit import DenoiserOracle
    torch.manual_seed(0)
    cfg = dict(patch_size=8, num_layers=2, ray_pe_type="plk" if scene else "relative_plk")
    model = (DGSDenoiserScene if scene else DGSDenoiser)(cfg).to(DEV)
    oracle = DenoiserOracle(layers=2, scene=scene).to(DEV)
    oracle.load_state_dict(model.state_dict(), strict=True)
    errs = _compare_models(model, oracle, 2, 4, 64, 64, f"small scene={scene}")
    assert all(v < 1e-3 for v in errs.values()), errs  # north-star bound (measured r1: 1e-4 .. 3e-4)


def test_denoiser_full_depth_obj256_vs_oracle():
    """BASELINE configs[1] model: 24 layers, 4 views at 256x256 (N = 4098 tokens, P = 262,146 Gaussians),
    random-init weights by the reference's init rules; bf16 tensor-core path vs the fp32 oracle."""
    from dgs_b200.denoiser import DGSDenoiser
    from oracle.dit import DenoiserOracle
    torch.manual_seed(0)
    model = DGSDenoiser(dict(patch_size=8)).to(DEV)
    oracle = DenoiserOracle().to(DEV)
    oracle.load_state_dict(model.state_dict(), strict=True)
    errs = _compare_models(model, oracle, 1, 4, 256, 256, "obj-256 x24")
    # north_star: DiT outputs within 1e-3 rel in bf16 (vs the fp32 oracle with the same fp32 master weights).
    # The two GEMMs at the ends of the network run split-bf16 and the conditioning runs fp32, so what is left is
    # the bf16 operand rounding inside the 24 blocks, entering through the gated residual updates (measured r1: 2.2e-4).
    assert all(v < 1e-3 for v in errs.values()), errs
    # the hot path's final product: the rendered views from both sets of Gaussians
    from dgs_b200 import synth
    c2w, fx = synth.orbit_cameras(4, 256, 256)
    c2w, fx = torch.tensor(c2w[None], device=DEV), torch.tensor(fx[None], device=DEV)
    images, ray_o, ray_d, t = _inputs(1, 4, 256, 256)
    with torch.no_grad():
        ref, _ = oracle.image_to_gaussians(images, ray_o, ray_d, t)
        out, _ = model.image_to_gaussians(images, ray_o, ray_d, t)
        r_ref = model.gs_renderer(ref["xyz"], ref["features"], ref["scaling"], ref["rotation"], ref["opacity"], 256,
                                  256, c2w, fx)
        r_out = model.render_gaussians(out, c2w, fx, 256, 256)
    e = rel(r_out, r_ref)
    print(f"[obj-256 x24] rendered views (ours DiT vs oracle DiT, same rasterizer): rel={e:.2e}")
    assert e < 1e-3
