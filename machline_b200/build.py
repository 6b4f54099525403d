"""Build recipes for the in-tree native libraries.

* ``libmachline_host.so``  -- host-side setup (g++, -ffp-contract=off so that rounding matches the
  reference's gfortran -O2 build; see csrc/host/).
* ``libmachline_gpu.so``   -- the CUDA kernels + C ABI (nvcc, sm_100a only; see csrc/gpu/).
* ``machline_b200.exe``    -- the ``main.f90``-shaped driver (links both).

Everything is built in-tree so the artefacts travel with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
INCLUDE = ROOT / "include"

HOST_LIB = PKG / "libmachline_host.so"
GPU_LIB = PKG / "libmachline_gpu.so"
DRIVER_EXE = PKG / "machline_b200.exe"

HOST_SOURCES = ["flow.cpp", "mesh_io.cpp", "panel_setup.cpp", "surface_mesh.cpp", "wake.cpp",
                "solver_setup.cpp", "outputs.cpp", "capi.cpp"]
GPU_SOURCES = ["capi.cu", "aic_kernels.cu", "aic_sub.cu", "aic_sup.cu", "aic_sub_ho.cu", "aic_sup_ho.cu", "solve_kernels.cu", "lu_kernels.cu", "lu_sharded.cu", "seq_solvers.cu", "peaks.cu", "multi.cu", "post.cu"]

# The image exports CXX=/opt/gcc/bin/g++ (a wrapper without libgomp.spec); the system compiler on
# PATH is the complete one.
GXX = shutil.which("g++") or "g++"
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# The supersonic assembly kernel evaluates the reference's predicates operation by operation: no FMA
# contraction there (csrc/gpu/pair_influence.cuh).  The solver kernels use explicit fma().
PER_FILE_FLAGS = {"aic_sup.cu": ["-fmad=false"], "aic_sup_ho.cu": ["-fmad=false"], "post.cu": ["-fmad=false"]}


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def _run(cmd, **kw):
    res = subprocess.run([str(c) for c in cmd], capture_output=True, text=True, **kw)
    if res.returncode != 0:
        sys.stderr.write(" ".join(str(c) for c in cmd) + "\n" + res.stdout + res.stderr)
        raise RuntimeError(f"build failed: {cmd[0]} (exit {res.returncode})")
    return res


def _nccl_dirs():
    """Header/library directories of the NCCL that torch bundles (nvidia-nccl-cu12 wheel)."""
    try:
        import nvidia.nccl as _n  # type: ignore
        base = Path(list(_n.__path__)[0])
    except Exception:
        return None, None
    inc, lib = base / "include", base / "lib"
    if (inc / "nccl.h").exists() and (lib / "libnccl.so.2").exists():
        return inc, lib
    return None, None


def build_host(force: bool = False) -> Path:
    srcs = [CSRC / "host" / s for s in HOST_SOURCES]
    deps = srcs + list((CSRC / "host").glob("*.hpp")) + list(INCLUDE.glob("*.h"))
    if force or _newer(HOST_LIB, deps):
        _run([GXX, "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-Wall", "-Wextra",
              "-o", HOST_LIB, *srcs, "-lquadmath"])
    return HOST_LIB


def gpu_compile_flags():
    flags = [*NVCC_ARCH, "-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
             "-Xptxas", "-v", "-I", INCLUDE]
    inc, _ = _nccl_dirs()
    if inc is not None:
        flags += ["-DML_HAVE_NCCL=1", "-I", inc]
    return flags


def build_gpu(force: bool = False) -> Path:
    gdir = CSRC / "gpu"
    srcs = [gdir / s for s in GPU_SOURCES if (gdir / s).exists()]
    deps = srcs + list(gdir.glob("*.cuh")) + list(gdir.glob("*.h")) + list(INCLUDE.glob("*.h")) + [CSRC / "host" / "pressure_rules.hpp"]
    if force or _newer(GPU_LIB, deps):
        objdir = gdir / "build"
        objdir.mkdir(exist_ok=True)
        objs = [objdir / (s.stem + ".o") for s in srcs]
        todo = [(s, o) for s, o in zip(srcs, objs) if force or _newer(o, deps)]

        def compile_one(so):
            s, o = so
            extra = PER_FILE_FLAGS.get(s.name, [])
            res = _run([NVCC, *gpu_compile_flags(), *extra, "-ccbin", GXX, "-c", s, "-o", o])
            return f"== {s.name}\n{res.stderr}"

        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as pool:   # one nvcc per translation unit
            logs = list(pool.map(compile_one, todo))
        (objdir / "ptxas.log").write_text("\n".join(logs))
        link = [NVCC, *NVCC_ARCH, "-shared", "-ccbin", GXX, "-o", GPU_LIB, *objs, "-lcudart_static"]
        _, lib = _nccl_dirs()
        if lib is not None:
            # link against the exact soname torch loads; resolved at run time through RPATH
            link += [f"-L{lib}", "-l:libnccl.so.2", "-Xlinker", f"-rpath={lib}"]
        _run(link)
    return GPU_LIB


def build_driver(force: bool = False) -> Path:
    src = CSRC / "host" / "main.cpp"
    if not src.exists():
        return DRIVER_EXE
    build_host(force)
    build_gpu(force)
    if force or _newer(DRIVER_EXE, [src, HOST_LIB, GPU_LIB]):
        _run([GXX, "-std=c++17", "-O2", "-o", DRIVER_EXE, src, f"-I{INCLUDE}", f"-L{PKG}",
              "-lmachline_host", "-lmachline_gpu", "-Wl,-rpath,$ORIGIN"])
    return DRIVER_EXE


def build_all(force: bool = False):
    build_host(force)
    build_gpu(force)
    build_driver(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", HOST_LIB, GPU_LIB)
