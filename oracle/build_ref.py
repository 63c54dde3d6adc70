"""oracle/build_ref.py -- TEST INFRASTRUCTURE: compile the UNMODIFIED reference rasterizer.

Recipe (no reference source is copied into this repo): nvcc/g++ compile the reference's own
files where they lie under /root/reference/submodules/diff-gaussian-rasterization
({ext.cpp, rasterize_points.cu, cuda_rasterizer/{rasterizer_impl,forward,backward}.cu}, glm from
its third_party/) for sm_100a into  oracle/_ref/dgr_ref_C.so  (git-ignored; it travels to the GPU
box with the gpurun snapshot).  `-include cstdint` is needed because rasterizer_impl.h uses
std::uintptr_t / uint32_t without including <cstdint> (gcc 13).

The module exports exactly the reference's pybind functions (DGR/ext.cpp:15-19):
rasterize_gaussians, rasterize_gaussians_backward, mark_visible.  It is used ONLY
 (a) by tests (-m gpu) as the GPU-side ground truth for parity,
 (b) by tests/golden/make_golden.py to generate the committed golden vectors,
 (c) by bench.py to time "the reference kernels on the same box" next to ours (reported, not shipped).
/root/reference does not exist on the GPU box: there this script only *loads* the prebuilt .so.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "dgr_ref_C"
REF = "/root/reference/submodules/diff-gaussian-rasterization"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(force: bool = False):
    """Build if the reference tree is present; returns the .so path or None."""
    if os.path.exists(so_path()) and not force:
        return so_path()
    if not os.path.isdir(REF):
        return None
    os.makedirs(OUT, exist_ok=True)
    from torch.utils.cpp_extension import load
    srcs = [os.path.join(REF, "ext.cpp"), os.path.join(REF, "rasterize_points.cu"),
            os.path.join(REF, "cuda_rasterizer", "rasterizer_impl.cu"),
            os.path.join(REF, "cuda_rasterizer", "forward.cu"),
            os.path.join(REF, "cuda_rasterizer", "backward.cu")]
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    load(name=NAME, sources=srcs, build_directory=OUT, is_python_module=False, verbose=False,
         extra_include_paths=[os.path.join(REF, "third_party", "glm"), REF],
         extra_cflags=["-O3", "-include", "cstdint"],
         extra_cuda_cflags=["-O3", "-include", "cstdint", "-gencode", "arch=compute_100a,code=sm_100a",
                            "-Xcompiler", "-fno-gnu-unique"])
    return so_path() if os.path.exists(so_path()) else None


def load_module():
    """Import the prebuilt reference extension (needs torch imported first); None if absent."""
    p = so_path()
    if not os.path.exists(p):
        return None
    import torch  # noqa: F401  (the .so links against libtorch)
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[NAME] = mod
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
