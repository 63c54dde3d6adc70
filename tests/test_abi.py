"""The C-ABI library loads and exports every symbol include/dgs_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from dgs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "dgs_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dgs_[a-z0-9_]+)\s*\(", txt)) - {"dgs_alloc_fn"})


def _ensure_built():
    if not os.path.exists(_lib.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location(
            "dgs_build", os.path.join(ROOT, "open-diffusiongs_b200", "csrc", "build.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        m.build()


def test_library_exports_every_declared_symbol():
    _ensure_built()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/dgs_b200.h but not exported"
    assert sorted(_lib.EXPORTED) == names


def test_version_and_error_string():
    _ensure_built()
    L = _lib.lib()
    assert L.dgs_version() == 100
    assert isinstance(L.dgs_last_error(), bytes)


def test_argument_validation_without_gpu():
    """Error convention of the boundary: invalid arguments -> status code + message, no exception, no GPU."""
    _ensure_built()
    L = _lib.lib()
    a = _lib.RasterArgs(P=10, D=5, M=1, W=16, H=16)
    rc = L.dgs_raster_backward(ctypes.byref(a), 0, *([None] * 15))
    assert rc == 1 and b"degree" in L.dgs_last_error()
    assert L.dgs_raster_geom_bytes(1, 1000) > 1000 * 56
    assert L.dgs_raster_image_bytes(1, 256, 256) >= 256 * 256 * 8
