"""LU A/B (development aid, not a bench line): solve one seeded dense system and print the time and a hash of x, so that two
library configurations (e.g. MACHLINE_LU_PANEL_V1=1 against the default) can be compared bit for bit across processes.
usage: python scripts/lu_ab.py N [--heavy] [--reps K]"""
import argparse
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from machline_b200 import _abi, gpu  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("n", type=int)
ap.add_argument("--heavy", action="store_true", help="no diagonal dominance: an interchange in (almost) every column")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
n = args.n
rng = np.random.default_rng(n)
A = rng.standard_normal((n, n))
if not args.heavy:
    A += 4.0 * np.sqrt(n) * np.eye(n)
A[::5] *= 1e3
A = np.asfortranarray(A)
b = rng.standard_normal(n)
ctx = gpu.Context(0)
ms = []
for _ in range(args.reps):
    x, info = ctx.solve_dense(A, b, _abi.solver_opts("LU"))
    ms.append(info.solve_ms)
r = np.abs(A @ x - b).max()
print(f"N={n} heavy={args.heavy} LU ms min {min(ms):.2f} ({2 / 3 * n ** 3 / min(ms) / 1e9:.2f} TF/s) res_max {r:.2e} "
      f"sha256(x) {hashlib.sha256(x.tobytes()).hexdigest()[:16]}")
ctx.close()
