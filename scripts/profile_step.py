"""One hot-path step on the bench workload for ncu captures (not a bench: numbers under a profiler are never reported).
usage: python scripts/profile_step.py [n_chord n_span] [--solver GMRES|LU] [--max-iter K]"""
import argparse
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from machline_b200 import gpu, host, meshgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("dims", nargs="*", type=int, default=[96, 52])
ap.add_argument("--solver", default="GMRES")
ap.add_argument("--max-iter", type=int, default=1000)
ap.add_argument("--mach", type=float, default=0.5)
args = ap.parse_args()
nc, ns = args.dims
tmp = tempfile.mkdtemp(prefix="machline_prof_")
pts, tris = meshgen.swept_wing_half(nc, ns)
meshgen.write_vtk(f"{tmp}/w.vtk", pts, tris)
inp = meshgen.wing_input("w.vtk", mach=args.mach, matrix_solver=args.solver)
inp["solver"]["max_iterations"] = args.max_iter
case = host.Case(inp, base_dir=tmp)
ctx = gpu.Context(0)
ctx.set_case(case)
ctx.assemble()
ms = ctx.assemble_resident()
x, info = ctx.solve(case.solver_opts(), case.BC)
print(f"N={case.n_unknown} pairs={ctx.pair_count} assemble_ms={ms:.3f} solve_ms={info.solve_ms:.2f} iters={info.iterations}")
ctx.close()
