"""Kernel experiments: build libmachline_gpu_<tag>.so with extra -D flags on some translation units (everything else is linked
from the regular build's objects).   usage: python scripts/build_variant.py <tag> <file.cu> [-DFLAG ...]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from machline_b200 import build as b  # noqa: E402

tag, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
b.build_gpu()
gdir = b.CSRC / "gpu"
obj = gdir / "build" / f"{Path(src).stem}_{tag}.o"
extra = b.PER_FILE_FLAGS.get(src, [])
res = b._run([b.NVCC, *b.gpu_compile_flags(), *extra, *flags, "-ccbin", b.GXX, "-c", gdir / src, "-o", obj])
print("\n".join(l for l in res.stderr.splitlines() if "aic_assemble" in l or "Used" in l or "spill" in l))
objs = [obj if s == src else gdir / "build" / (Path(s).stem + ".o") for s in b.GPU_SOURCES]
out = b.PKG / f"libmachline_gpu_{tag}.so"
link = [b.NVCC, *b.NVCC_ARCH, "-shared", "-ccbin", b.GXX, "-o", out, *objs, "-lcudart_static"]
_, lib = b._nccl_dirs()
if lib is not None:
    link += [f"-L{lib}", "-l:libnccl.so.2", "-Xlinker", f"-rpath={lib}"]
b._run(link)
print(out)
