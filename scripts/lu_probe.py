"""LU timing probe on the bench workload family (not a bench line): measured DFMA / DMMA peaks, then assemble + LU solve.
usage: python scripts/lu_probe.py [n_chord n_span] [--reps K]"""
import argparse
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from machline_b200 import gpu, host, meshgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("dims", nargs="*", type=int, default=[96, 52])
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
nc, ns = args.dims
tmp = tempfile.mkdtemp(prefix="machline_lu_")
pts, tris = meshgen.swept_wing_half(nc, ns)
meshgen.write_vtk(f"{tmp}/w.vtk", pts, tris)
case = host.Case(meshgen.wing_input("w.vtk", mach=0.5, matrix_solver="LU"), base_dir=tmp)
ctx = gpu.Context(0)
fp64, _ = ctx.measure_peaks(hbm=False)
dmma = ctx.measure_dmma_peak()
ctx.set_case(case)
ctx.assemble()
N = case.n_unknown
best = 1e30
for _ in range(args.reps):
    ctx.assemble_resident()
    x, info = ctx.solve(case.solver_opts(), case.BC)
    best = min(best, info.solve_ms)
flops = 2.0 / 3.0 * N ** 3
print(f"N={N} DFMA peak {fp64:.1f} TF/s, DMMA peak {dmma:.1f} TF/s; LU solve best {best:.2f} ms = "
      f"{flops / best / 1e9:.2f} TF/s ({flops / best / 1e9 / dmma:.1%} of DMMA peak), res_norm {info.res_norm:.2e}")
ctx.close()
