"""Repeat the solve on the resident bench system (development aid: timing spread, wall vs device)."""
import sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from machline_b200 import gpu, host, meshgen, _abi  # noqa: E402
nc, ns = 96, 52
solver = sys.argv[1] if len(sys.argv) > 1 else "GMRES"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
tmp = tempfile.mkdtemp(prefix="machline_rep_")
pts, tris = meshgen.swept_wing_half(nc, ns)
meshgen.write_vtk(f"{tmp}/w.vtk", pts, tris)
case = host.Case(meshgen.wing_input("w.vtk", matrix_solver=solver), base_dir=tmp)
ctx = gpu.Context(0)
ctx.set_case(case)
ctx.assemble()
opts = case.solver_opts()
ms, wall = [], []
for i in range(reps):
    t0 = time.perf_counter()
    x, info = ctx.solve(opts, case.BC)
    wall.append(1e3 * (time.perf_counter() - t0))
    ms.append(info.solve_ms)
print(solver, "N", case.n_unknown, "iters", info.iterations, "res %.2e" % info.res_norm, "solve ms:", " ".join(f"{m:.1f}" for m in ms))
print("   wall ms:", " ".join(f"{m:.1f}" for m in wall))
ctx.close()
