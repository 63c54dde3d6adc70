"""Micro-benchmarks of the DiT building blocks through the C ABI (not a pytest file).
    python tests/perf_kernels.py            # attention + the four DiT GEMM shapes at N=4098, B=1
Prints one JSON line per kernel: median ms over 20 launches (CUDA events, L2-hot operands), TFLOP/s."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-diffusiongs_b200"))
from dgs_b200 import _lib  # noqa: E402

DEV = "cuda:0"


def st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2]


def main():
    L = _lib.lib()
    tag = {k: os.environ.get(k) for k in ("DGS_ATT_POLY", "DGS_GEMM_2CTA", "DGS_GEMM_M256") if os.environ.get(k)}
    B, N, H, D = 1, 4098, 16, 1024
    qkv = (torch.randn(B, N, 3, H, 64, device=DEV) * 1.5).to(torch.bfloat16)
    out = torch.empty(B, N, D, dtype=torch.bfloat16, device=DEV)
    ms = timeit(lambda: _lib.check(L.dgs_attention_fwd(qkv.data_ptr(), out.data_ptr(), B, N, H, st())))
    print(json.dumps(dict(kernel="attention", ms=ms, tflops=4 * N * N * D * B / ms / 1e9, **tag)))
    if "--gemm-sweep" in sys.argv:
        for (m, n, k) in [(8192, 4096, 4096), (8192, 4096, 1024), (4096, 4096, 8192), (16384, 1024, 1024)]:
            A = torch.randn(m, k, device=DEV).to(torch.bfloat16)
            W = (torch.randn(n, k, device=DEV) * 0.03).to(torch.bfloat16)
            o = torch.zeros(m, n, dtype=torch.bfloat16, device=DEV)
            ms = timeit(lambda: _lib.check(L.dgs_gemm_bf16(A.data_ptr(), W.data_ptr(), None, None, o.data_ptr(), m, n, k, 0, n, 0,
                                                           1, st())))
            print(json.dumps(dict(kernel=f"gemm_{m}x{n}x{k}", ms=ms, tflops=2 * m * n * k / ms / 1e9, **tag)))
        return
    if "--attn-bwd" in sys.argv or "--all" in sys.argv:
        Bb = int(os.environ.get("DGS_PERF_B", "1"))
        qkvb = (torch.randn(Bb, N, 3, H, 64, device=DEV) * 1.5).to(torch.bfloat16)
        outb = torch.empty(Bb, N, D, dtype=torch.bfloat16, device=DEV)
        dout = torch.randn(Bb, N, D, device=DEV).to(torch.bfloat16)
        Np = (N + 127) // 128 * 128
        lse = torch.empty(Bb, H, Np, device=DEV)
        dsum = torch.empty(Bb, H, Np, device=DEV)
        dqkv = torch.empty_like(qkvb)
        _lib.check(L.dgs_attention_fwd_train(qkvb.data_ptr(), outb.data_ptr(), lse.data_ptr(), Bb, N, H, st()))
        ms = timeit(lambda: _lib.check(L.dgs_attention_bwd(qkvb.data_ptr(), outb.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                                           dsum.data_ptr(), dqkv.data_ptr(), Bb, N, H, st())))
        print(json.dumps(dict(kernel="attention_bwd", B=Bb, ms=ms, tflops_alg=2.5 * 4 * N * N * D * Bb / ms / 1e9, **tag)))
        if "--attn-bwd" in sys.argv:
            return
    for name, (n, k, epi) in dict(qkv=(3 * D, D, 0), proj=(D, D, 2), fc1=(4 * D, D, 1), fc2=(D, 4 * D, 2)).items():
        A = torch.randn(B * N, k, device=DEV).to(torch.bfloat16)
        W = (torch.randn(n, k, device=DEV) * 0.03).to(torch.bfloat16)
        bias = torch.randn(n, device=DEV)
        gate = torch.randn(B, n, device=DEV)
        o = torch.zeros(B * N, n, dtype=torch.float32 if epi == 2 else torch.bfloat16, device=DEV)
        ms = timeit(lambda: _lib.check(L.dgs_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), gate.data_ptr(),
                                                       o.data_ptr(), B * N, n, k, epi, n, n, N, st())))
        print(json.dumps(dict(kernel="gemm_" + name, ms=ms, tflops=2 * B * N * n * k / ms / 1e9, **tag)))


if __name__ == "__main__":
    main()
