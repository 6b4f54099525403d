"""Wall-clock breakdown of the end-to-end path on the bench workload (development aid)."""
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from machline_b200 import gpu, host, meshgen  # noqa: E402

nc, ns = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (96, 52)
tmp = tempfile.mkdtemp(prefix="machline_e2e_")
pts, tris = meshgen.swept_wing_half(nc, ns)
meshgen.write_vtk(f"{tmp}/w.vtk", pts, tris)
t0 = time.perf_counter()
case = host.Case(meshgen.wing_input("w.vtk"), base_dir=tmp)
print(f"host setup {1e3 * (time.perf_counter() - t0):.1f} ms  N={case.n_unknown}")
ctx = gpu.Context(0)
opts = case.solver_opts()
for it in range(4):
    t = [time.perf_counter()]
    ctx.set_case(case); t.append(time.perf_counter())
    ctx.assemble(); t.append(time.perf_counter())
    x, info = ctx.solve(opts, case.BC); t.append(time.perf_counter())
    print(f"iter {it}: set_case {1e3*(t[1]-t[0]):.1f}  assemble(prepare+H2D+kernels) {1e3*(t[2]-t[1]):.1f} (device {info.assemble_ms:.1f})  "
          f"solve {1e3*(t[3]-t[2]):.1f} (device {info.solve_ms:.1f}, {info.iterations} its)")
for it in range(3):
    t0 = time.perf_counter(); ms = ctx.assemble_resident(); t1 = time.perf_counter()
    x, info = ctx.solve(opts, case.BC); t2 = time.perf_counter()
    print(f"resident {it}: assemble {1e3*(t1-t0):.1f} (device {ms:.1f})  solve {1e3*(t2-t1):.1f} (device {info.solve_ms:.1f})")
ctx.close()
