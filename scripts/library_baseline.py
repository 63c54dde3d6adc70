"""Same-box library baselines for the DiT kernels: cuDNN / flash SDPA and cuBLAS(Lt) matmul on the
shapes the denoise / training steps use.  Not part of the product path and not a pytest file.

    python scripts/library_baseline.py [out.json]

Times with CUDA events, median of 20 after 5 warm-up launches, operands L2-hot like tests/perf_kernels.py so the two
are comparable.  Writes one JSON document (default profiles/r2_library_baseline.json)."""
import json
import sys

import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

DEV = "cuda:0"


def timeit(fn, iters=20, warmup=5):
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
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "attention": [], "gemm": []}
    for (B, N) in [(1, 4098), (4, 4098), (1, 16386)]:
        H, Dh = 16, 64
        q, k, v = [(torch.randn(B, H, N, Dh, device=DEV) * 1.5).to(torch.bfloat16).requires_grad_(True) for _ in range(3)]
        do = torch.randn(B, H, N, Dh, device=DEV).to(torch.bfloat16)
        flops = 4 * N * N * H * Dh * B
        for name, be in [("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION),
                         ("efficient", SDPBackend.EFFICIENT_ATTENTION)]:
            rec = dict(backend=name, B=B, N=N, H=H, Dh=Dh)
            try:
                with sdpa_kernel(be):
                    with torch.no_grad():
                        ms = timeit(lambda: F.scaled_dot_product_attention(q, k, v))
                    rec.update(fwd_ms=ms, fwd_tflops=flops / ms / 1e9)
                    o = F.scaled_dot_product_attention(q, k, v)
                    ms = timeit(lambda: torch.autograd.grad(o, (q, k, v), do, retain_graph=True))
                    rec.update(bwd_ms=ms, bwd_tflops_alg=2.5 * flops / ms / 1e9)
            except Exception as e:  # noqa: BLE001
                rec["error"] = str(e).splitlines()[0][:200]
            out["attention"].append(rec)
            print(json.dumps(rec), flush=True)
    try:
        from flash_attn import flash_attn_func
        for (B, N) in [(1, 4098), (1, 16386)]:
            q, k, v = [(torch.randn(B, N, 16, 64, device=DEV) * 1.5).to(torch.bfloat16) for _ in range(3)]
            ms = timeit(lambda: flash_attn_func(q, k, v))
            rec = dict(backend="flash_attn2_pkg", B=B, N=N, fwd_ms=ms, fwd_tflops=4 * N * N * 1024 * B / ms / 1e9)
            out["attention"].append(rec)
            print(json.dumps(rec), flush=True)
    except Exception as e:  # noqa: BLE001
        print("flash_attn pkg:", str(e)[:200])
    D = 1024
    for B in (1, 4):
        M = B * 4098
        for name, (n, k) in dict(qkv=(3 * D, D), proj=(D, D), fc1=(4 * D, D), fc2=(D, 4 * D)).items():
            A = torch.randn(M, k, device=DEV).to(torch.bfloat16)
            W = (torch.randn(n, k, device=DEV) * 0.03).to(torch.bfloat16)
            bias = torch.randn(n, device=DEV).to(torch.bfloat16)
            ms = timeit(lambda: F.linear(A, W, bias))
            rec = dict(shape=name, M=M, N=n, K=k, ms=ms, tflops=2 * M * n * k / ms / 1e9)
            out["gemm"].append(rec)
            print(json.dumps(rec), flush=True)
    path = sys.argv[1] if len(sys.argv) > 1 else "profiles/r2_library_baseline.json"
    with open(path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
