"""A/B timing of the GMRES solve on a study case under different environment switches (development aid, not a bench line).
usage: python scripts/gmres_ab.py <workload> [--reps K] [--sharded] CONFIG [CONFIG ...]
where CONFIG is a comma-separated list of NAME=VALUE environment settings ("-" = none), e.g.
    python scripts/gmres_ab.py onera_m6 MACHLINE_GMRES_TAIL=1,MACHLINE_GMRES_D2H_COPY=1 - MACHLINE_GEMV_L2_PIN_MB=64"""
import argparse
import os
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
ap.add_argument("configs", nargs="+")
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--solver", default="GMRES")
ap.add_argument("--synthetic", type=int, default=0, help="use the synthetic weak-scaling family sized for this many GPUs instead")
args = ap.parse_args()
tmp = tempfile.mkdtemp(prefix="machline_ab_")
if args.synthetic:
    case, desc = bench.build_case(args.synthetic, tmp, args.solver, None, "synthetic_wing")
else:
    case, desc = bench.build_case(1, tmp, args.solver, None, args.workload)
ctx = gpu.Context(0)
ctx.set_case(case)
ctx.assemble()
opts = case.solver_opts()
BC = np.array(case.BC)
x0 = None
touched = set()
for cfg in args.configs:
    for name in touched:
        os.environ.pop(name, None)
    if cfg != "-":
        for kv in cfg.split(","):
            name, val = kv.split("=", 1)
            os.environ[name] = val
            touched.add(name)
    ms = []
    for _ in range(args.reps):
        ctx.assemble_resident()
        x, info = ctx.solve(opts, BC)
        ms.append(info.solve_ms)
    ctx.set_profiling(True)
    ctx.profile(reset=True)
    ctx.assemble_resident()
    ctx.solve(opts, BC)
    gp = ctx.profile()
    ctx.set_profiling(False)
    if x0 is None:
        x0 = x.copy()
    dx = float(np.abs(x - x0).max() / np.abs(x0).max())
    print(f"{cfg:60s} N={case.n_unknown} solve ms min {min(ms):8.3f} med {sorted(ms)[len(ms) // 2]:8.3f} its {info.iterations} "
          f"res {info.res_norm:.2e} dx {dx:.1e} | gemv {gp.gemv_ms / max(gp.gemv_launches, 1) * 1e3:7.2f} us x {gp.gemv_launches}",
          flush=True)
ctx.close()
