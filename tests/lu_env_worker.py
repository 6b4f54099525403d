"""Worker: dense LU solves (heavy pivoting, ties) through ml_solve_dense under whatever MACHLINE_LU_* environment the
parent test set (the variants are selected once per process)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from machline_b200 import _abi, gpu  # noqa: E402

ctx = gpu.Context(0)
for n in [int(v) for v in sys.argv[1:]] or [700, 3000]:
    rng = np.random.default_rng(n)
    A = np.asfortranarray(rng.standard_normal((n, n)))
    A[::5] *= 1e3
    b = rng.standard_normal(n)
    x, info = ctx.solve_dense(A, b, _abi.solver_opts("LU"))
    x_np = np.linalg.solve(A, b)
    err = np.abs(x - x_np).max() / np.abs(x_np).max()
    assert err <= 1e-8, (n, err)
    print(f"n={n} err {err:.2e} solve_ms {info.solve_ms:.2f} OK", flush=True)
ctx.close()
