"""Development diagnostics on the GPU box: parity statistics + first timings (not a bench)."""
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import fixtures  # noqa: E402
import oracle_binding as ob  # noqa: E402
from machline_b200 import gpu, host, meshgen  # noqa: E402


def stats(A, A_ref):
    rowmax = np.abs(A_ref).max(axis=1, keepdims=True)
    d = np.abs(A - A_ref)
    nz = A_ref != 0
    rel = np.zeros_like(d)
    rel[nz] = d[nz] / np.abs(A_ref[nz])
    relrow = d / rowmax
    big = nz & (np.abs(A_ref) > 1e-6 * rowmax)
    return dict(max_rel=float(rel.max()), max_rel_over_1e6floor=float(rel[big].max()), max_rowscaled=float(relrow.max()),
                p999_rel=float(np.quantile(rel[nz], 0.999)), frac_rel_gt_1e12=float((rel[nz] > 1e-12).mean()),
                zero_mismatch=int(((A == 0) != (A_ref == 0)).sum()))


def main():
    ctx = gpu.Context(0)
    out = {}
    fp64, hbm = ctx.measure_peaks()
    print(f"peaks: fp64 {fp64:.2f} TFLOP/s, copy {hbm:.0f} GB/s", flush=True)
    out["peaks"] = dict(fp64_tflops=fp64, hbm_gbs=hbm)
    for name in ["test_08", "test_13", "test_01", "test_15", "test_05", "test_20", "test_19"]:
        case, expect, tol = fixtures.make_case(name)
        ctx.set_case(case)
        t = time.time()
        Ik = ctx.assemble()
        A = ctx.get_A()
        A_ref, I_ref = ob.assemble(case)
        s = stats(A, A_ref)
        s["I_known_err"] = float(np.abs(Ik - I_ref).max() / max(1e-300, np.abs(I_ref).max()))
        opts = case.solver_opts()
        if name == "test_20":
            opts.matrix_solver = 3
        try:
            x, info = ctx.solve(opts, case.BC)
            res = case.post(x)
            got = [res.C_p_max, res.C_p_min, *res.C_F]
            s["golden_diff"] = [abs(g - e) for g, e in zip(got, expect)]
            s["iters"] = info.iterations
            s["solve_ms"] = info.solve_ms
            s["assemble_ms"] = info.assemble_ms
        except Exception as e:  # noqa: BLE001
            s["solve_error"] = str(e)
        print(name, json.dumps(s), flush=True)
        out[name] = s
        case.close()
    # timing on synthetic meshes
    tmp = tempfile.mkdtemp(prefix="machline_diag_")
    for label, (nc, ns) in {"wing_3.6k": (40, 22), "wing_14k": (80, 45)}.items():
        pts, tris = meshgen.swept_wing_half(nc, ns)
        meshgen.write_vtk(f"{tmp}/{label}.vtk", pts, tris)
        t = time.time()
        case = host.Case(meshgen.wing_input(f"{label}.vtk"), base_dir=tmp)
        t_setup = time.time() - t
        ctx.set_case(case)
        ctx.assemble()
        ms = [ctx.assemble_resident() for _ in range(5)]
        pairs = ctx.pair_count
        r = dict(panels=case.info.n_body_panels, n=case.n_unknown, pairs=pairs, setup_s=t_setup, assemble_ms=ms,
                 pairs_per_s=pairs / (min(ms) * 1e-3))
        t = time.time()
        x, info = ctx.solve(case.solver_opts(), case.BC)
        r.update(iters=info.iterations, solve_ms=info.solve_ms, res=info.res_norm, solve_wall_s=time.time() - t)
        res = case.post(x)
        r.update(Cz=float(res.C_F[2]), Cp_min=res.C_p_min)
        print(label, json.dumps(r), flush=True)
        out[label] = r
        case.close()
    for level in (3, 4, 5):
        pts, tris = meshgen.icosphere(level)
        meshgen.write_vtk(f"{tmp}/ico{level}.vtk", pts, tris)
        case = host.Case(meshgen.sphere_input(f"ico{level}.vtk"), base_dir=tmp)
        ctx.set_case(case)
        ctx.assemble()
        ms = [ctx.assemble_resident() for _ in range(3)]
        pairs = ctx.pair_count
        x, info = ctx.solve(case.solver_opts(), case.BC)
        res = case.post(x)
        r = dict(panels=case.info.n_body_panels, n=case.n_unknown, pairs=pairs, assemble_ms=ms, pairs_per_s=pairs / (min(ms) * 1e-3),
                 iters=info.iterations, solve_ms=info.solve_ms, Cp_max=res.C_p_max, Cp_min=res.C_p_min)
        print(f"ico{level}", json.dumps(r), flush=True)
        out[f"ico{level}"] = r
        case.close()
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "diag.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
