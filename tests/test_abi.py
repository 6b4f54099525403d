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


def test_headers_are_plain_c_and_match_the_ctypes_mirrors(tmp_path):
    """The boundary is a C ABI (a Fortran bind(C) module binds it): both headers must compile as C99, and the ctypes mirrors of the
    structures that cross it (machline_b200/_abi.py) must have the sizes the C compiler gives them."""
    import shutil
    import subprocess
    from machline_b200 import _abi
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    structs = {"ml_flow": _abi.MlFlow, "ml_panel_soa": _abi.MlPanelSoa, "ml_system_map": _abi.MlSystemMap,
               "ml_solver_opts": _abi.MlSolverOpts, "ml_solve_info": _abi.MlSolveInfo, "ml_profile": _abi.MlProfile,
               "ml_post_tables": _abi.MlPostTables, "ml_post_flow": _abi.MlPostFlow, "ml_post_out": _abi.MlPostOut,
               "mlh_cp_table": _abi.MlhCpTable, "mlh_solver_settings": _abi.MlhSolverSettings, "mlh_results": _abi.MlhResults,
               "mlh_mesh_info": _abi.MlhMeshInfo}
    src = tmp_path / "abi.c"
    lines = ['#include <stdio.h>', '#include "machline_gpu.h"', '#include "machline_host.h"', "int main(void) {"]
    lines += [f'    printf("{n} %zu\\n", sizeof({n}));' for n in structs]
    lines += ["    return 0;", "}"]
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    sizes = dict(line.split() for line in out.strip().splitlines())
    for n, cls in structs.items():
        assert int(sizes[n]) == C.sizeof(cls), f"{n}: C {sizes[n]} bytes, ctypes {C.sizeof(cls)}"
