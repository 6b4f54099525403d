"""CPU: the C-ABI libraries load and export every symbol their headers declare (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(ml[h]?_[a-z_0-9]+)\s*\(", text)))


def test_gpu_library_exports_header_symbols():
    from machline_b200 import build
    lib_path = build.build_gpu()
    L = C.CDLL(str(lib_path))
    names = [n for n in _declared(ROOT / "include" / "machline_gpu.h") if n.startswith("ml_")]
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"libmachline_gpu.so does not export {n}"
    L.ml_abi_version.restype = C.c_int
    assert L.ml_abi_version() == 1


def test_host_library_exports_header_symbols():
    from machline_b200 import build
    L = C.CDLL(str(build.build_host()))
    for n in _declared(ROOT / "include" / "machline_host.h"):
        if n.startswith("mlh_"):
            assert hasattr(L, n), f"libmachline_host.so does not export {n}"


def test_context_creation_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device ml_ctx_create reports ML_CUDA_ERROR."""
    import torch
    from machline_b200 import gpu
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(gpu.GpuError):
        gpu.Context(0)


def test_product_does_not_reference_oracle():
    """The oracle is test infrastructure: nothing under machline_b200/ or include/ may mention it."""
    for p in list((ROOT / "machline_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".cpp", ".hpp", ".h"}:
            text = p.read_text(errors="ignore")
            assert "liboracle" not in text and "oracle_binding" not in text and "orc_" not in text, p
