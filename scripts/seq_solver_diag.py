"""Development aid (GPU box): the reference's other solvers on golden case test_20 (sorted diamond wing, FQRUP) and
their timing on a larger dense system."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import fixtures  # noqa: E402
import oracle_binding as ob  # noqa: E402
from machline_b200 import _abi, gpu  # noqa: E402

ctx = gpu.Context(0)
case, expect, tol = fixtures.make_case("test_20")
ctx.set_case(case)
ctx.assemble()
A_ref, I_ref = ob.assemble(case)
print("test_20 N", case.n_unknown, "tol", tol)
for name in ["FQRUP", "QRUP", "LU", "GMRES", "PURC", "BSSOR"]:
    opts = case.solver_opts()
    opts.matrix_solver = _abi.SOLVERS[name]
    try:
        x, info = ctx.solve(opts, case.BC)
    except Exception as e:  # noqa: BLE001
        print(name, "error", e)
        continue
    res = case.post(x)
    got = [res.C_p_max, res.C_p_min, *res.C_F]
    t0 = time.perf_counter()
    x_or, info_or = ob.solve_system(A_ref, I_ref, case.BC, opts)
    t_or = time.perf_counter() - t0
    print(f"{name}: golden diff {[f'{abs(g - e):.2e}' for g, e in zip(got, expect)]} solve_ms {info.solve_ms:.1f} (oracle {1e3 * t_or:.0f} ms) "
          f"iters {info.iterations}/{info_or.iterations} |x-x_or|/|x| {np.abs(x - x_or).max() / np.abs(x_or).max():.2e} res {info.res_norm:.1e}", flush=True)
case.close()

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
rng = np.random.default_rng(0)
A = rng.standard_normal((n, n)) + 3 * np.sqrt(n) * np.eye(n)
i, j = np.indices((n, n))
A[i - j > n // 4] = 0.0
A = np.asfortranarray(A)
b = rng.standard_normal(n)
for name in ["FQRUP", "QRUP", "PURC", "BSSOR", "BJAC", "LU"]:
    opts = _abi.solver_opts(name, preconditioner="none", rel=0.9)
    x, info = ctx.solve_dense(A, b, opts)
    print(f"N={n} {name}: solve_ms {info.solve_ms:.1f} iters {info.iterations} res {info.res_norm:.1e}", flush=True)
ctx.close()
