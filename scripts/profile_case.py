"""One hot-path step on a study case (tests/golden/study_meshes.npz) for ncu captures -- not a bench: numbers under a profiler
are never reported.   usage: python scripts/profile_case.py <onera_m6|cone|sears_haack|agard_b> [--solver S] [--max-iter K] [--no-solve]"""
import argparse
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from machline_b200 import gpu  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--solver", default="GMRES")
ap.add_argument("--max-iter", type=int, default=1000)
ap.add_argument("--no-solve", action="store_true")
ap.add_argument("--repeat", type=int, default=1)
args = ap.parse_args()
tmp = tempfile.mkdtemp(prefix="machline_prof_")
case, desc = bench.build_case(1, tmp, args.solver, None, args.workload)
ctx = gpu.Context(0)
ctx.set_case(case)
ctx.assemble()
ms = [ctx.assemble_resident() for _ in range(args.repeat)]
line = f"{args.workload}: N={case.n_unknown} pairs={ctx.pair_count} assemble_ms={min(ms):.3f}"
if case.flow.supersonic:
    line += f" census={ctx.dod_census()}"
if not args.no_solve:
    o = case.solver_opts()
    o.max_iterations = args.max_iter
    x, info = ctx.solve(o, np.array(case.BC))
    line += f" solve_ms={info.solve_ms:.2f} iters={info.iterations} res={info.res_norm:.2e}"
print(line)
ctx.close()
