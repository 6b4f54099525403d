"""GPU: the CUDA path (through the C ABI) against the oracle on identical inputs, and against the
reference's golden tuples end to end.  Tolerances: AIC entries 1e-12 relative (north star); surface
Cp / force coefficients the reference's own test tolerances (1e-12 .. 1e-8)."""
import numpy as np
import pytest

import fixtures
import oracle_binding as ob

pytestmark = pytest.mark.gpu

AIC_CASES = ["test_08", "test_13", "test_01", "test_15", "test_05", "test_20", "test_19",
             # higher-order panels (quadratic doublets, linear sources): subsonic Morino / source-free, supersonic, supersonic + wake
             "test_04", "test_02", "test_16", "test_17"]


def _rel_err(A, A_ref, S):
    """|dA_ij| / S_ij, S_ij = the running-error scale of the reference algorithm for that entry (oracle_aic.cpp
    phi_d_abs): the sum of the |terms| that are added up into it -- per panel the three edge angles of hH113 (each with
    the cancelled products behind its atan2 arguments), the F111 edge terms, times |T_mu| -- over all panels feeding
    the column.  Why not plain |dA|/|A|: an entry is a sum of O(1) angles that cancel to the panel's small solid angle
    and of 6-12 panel terms that cancel in a vertex's far field, so a 1-ulp difference between two libms moves it by
    far more than 1e-12 of ITSELF.  tests/test_oracle_noise_floor.py shows the reference algorithm does this to
    itself: glibc's log/atan2 vs correctly rounded ones differ by up to 7e-12 (sphere) .. 1e-8 (half wing) in plain
    relative terms (4e-5 for a faithfully rounded libm), and by ~1e-16 relative to S.  For entries without
    cancellation S_ij ~ |A_ij| and this IS the plain relative error."""
    den = np.where(S > 0, S, 1.0)
    return np.abs(A - A_ref) / den


@pytest.fixture(scope="module")
def ctx():
    from machline_b200 import gpu
    c = gpu.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", AIC_CASES)
def test_aic_entries_match_oracle(ctx, name):
    case, _, _ = fixtures.make_case(name)
    ctx.set_case(case)
    I_known = ctx.assemble()
    A = ctx.get_A()
    A_ref, I_ref, S = ob.assemble(case, with_scale=True)
    # structural zeros must be exact zeros on both sides
    assert ((A == 0) == (A_ref == 0)).all()
    err = _rel_err(A, A_ref, S)
    # 2e-14 of the summed terms (50x inside the north star's 1e-12; the libm noise floor is ~2e-16)
    assert err.max() < 2e-14, f"max AIC error relative to its terms {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"
    # and 1e-12 of the row's largest entry
    assert (np.abs(A - A_ref) / np.abs(A_ref).max(axis=1, keepdims=True)).max() < 1e-12
    # entries whose terms do not cancel (|A_ij| > S_ij / 4) carry 1e-12 relative to themselves
    sel = np.abs(A_ref) > 0.25 * S
    assert sel.any()
    plain = np.abs(A - A_ref)[sel] / np.abs(A_ref)[sel]
    assert plain.max() < 1e-12
    scale = max(1e-300, np.abs(I_ref).max())
    assert np.abs(I_known - I_ref).max() / scale < 1e-13
    # Entries that miss PLAIN relative 1e-12 (cancelling sums): their fraction must be the one any faithfully rounded
    # log/atan2 produces on the reference algorithm itself -- the oracle re-run with its libm results moved by one ulp in
    # 3/8 of the calls (tests/test_oracle_noise_floor.py) -- not more.
    nz = A_ref != 0
    frac_gpu = float((np.abs(A - A_ref)[nz] > 1e-12 * np.abs(A_ref)[nz]).mean())
    ob.lib().orc_set_exact_libm(2)
    try:
        A_noise, _ = ob.assemble(case)
    finally:
        ob.lib().orc_set_exact_libm(0)
    frac_noise = float((np.abs(A_noise - A_ref)[nz] > 1e-12 * np.abs(A_ref)[nz]).mean())
    print(f"{name}: entries beyond plain-relative 1e-12: GPU {frac_gpu:.4f}, one-ulp libm noise on the oracle {frac_noise:.4f}")
    if not case.flow.supersonic:
        # same order as the model (it perturbs 3/8 of the calls by one ulp, a real libm all of them by up to one)
        assert frac_gpu <= 3.0 * frac_noise + 1e-3
    else:
        # supersonic: the reference evaluates the hH113 arguments in binary128 (src/panel.f90:2553-2565), the device in
        # binary64: ~1e-16 ABSOLUTE per edge angle, inside the metric above (2e-14 of the summed terms) but visible in
        # plain-relative terms wherever the three O(1) angles cancel.  Measured 0.15-0.18 of the entries (13x the
        # one-ulp model); bounded here so that a regression shows.
        # Higher-order entries (tests 16, 17) cancel more: the one-ulp model itself rises from 0.013 to 0.08 there and the
        # device sits at 0.30 (3.8x the model).
        assert frac_gpu <= max(0.25, 5.0 * frac_noise)
    case.close()


# Cases whose system is numerically singular or very ill-conditioned: the asymmetric mirrored half wing (tests 01, 03,
# 12: cond(A) ~ 1e17 -- strength-matching rows; GMRES picks one solution out of a near-null space) and the sorted
# full diamond wing (test 20: cond(A) = 4e6, golden produced by the fast-Givens QRUP).  There a 1e-16 perturbation
# of A (another libm) or another solver moves the force coefficients by 1e-9 ... 4e-7 (measured on the CPU with
# the oracle: numpy LU on the oracle's own matrix misses test 20's golden Cz by 3.9e-7).  The oracle reproduces these
# goldens at the reference's own tolerance because it repeats the reference's operations; the GPU is held to the
# reference tolerance times the slack below (Cp columns, force columns).  Test 20 runs the GPU's own FQRUP, which is
# bit-identical to the reference's on the same matrix (tests/test_gpu_solvers.py): what is left (Cz off by 9e-11 for
# a 1e-12 tolerance, measured; 3.9e-7 with LU) is cond(A) times the 1e-16 libm difference in A itself.
# Tests 15 and 18 (supersonic wake; cond 4e17 / 4e5 with a 1e-9 / 1e-10 force tolerance) sit at cond * eps as well.
ILL_CONDITIONED = {"test_01": (20., 20.), "test_03": (20., 20.), "test_12": (20., 20.), "test_20": (1., 400.),
                   "test_15": (1., 10.), "test_18": (1., 10.),
                   # the same meshes and flows with higher-order panels (test 02's matrix is singular: fixtures oracle_slack)
                   "test_02": (20., 20.), "test_04": (20., 20.), "test_17": (1., 10.)}


@pytest.mark.parametrize("name", fixtures.golden_case_names())
def test_reference_goldens_through_gpu(ctx, name):
    """host setup -> ml_assemble -> ml_solve -> host post == the reference's golden tuple."""
    case, expect, tol = fixtures.make_case(name)
    opts = case.solver_opts()   # the input's own matrix_solver: GMRES, BJAC (test 19) or FQRUP (test 20)
    ctx.set_case(case)
    ctx.assemble()
    x, info = ctx.solve(opts, case.BC)
    # Neumann formulations (test 21): the inner velocity of every panel comes from a sweep of velocity influences
    v_inner = None if case.dirichlet else ctx.velocities_at(case, case.inner_points(), x)
    res = case.post(x, v_inner)
    s_cp, s_f = ILL_CONDITIONED.get(name, (1., 1.))
    fixtures.check_tuple(res, expect, [tol[0] * s_cp, tol[1] * s_cp, tol[2] * s_f, tol[3] * s_f, tol[4] * s_f])
    case.close()


@pytest.mark.parametrize("name", fixtures.comparison_names())
def test_reference_comparisons_through_gpu(ctx, name):
    """Reference tests that compare two runs with each other (test 10: full wing vs mirrored half wing) through the CUDA path."""
    a, b, tol = fixtures.comparison_cases(name)
    tuples = []
    for case in (a, b):
        ctx.set_case(case)
        ctx.assemble()
        x, info = ctx.solve(case.solver_opts(), case.BC)
        r = case.post(x)
        tuples.append([r.C_p_max, r.C_p_min, *[float(v) for v in r.C_F]])
        case.close()
    for x, y, t in zip(tuples[0], tuples[1], tol):
        assert abs(x - y) < t, tuples


@pytest.mark.parametrize("name", ["test_08", "test_13", "test_05"])
def test_gmres_iterations_and_solution_match_oracle(ctx, name):
    case, _, _ = fixtures.make_case(name)
    ctx.set_case(case)
    ctx.assemble()
    x, info = ctx.solve(case.solver_opts(), case.BC)
    A_ref, I_ref = ob.assemble(case)
    x_ref, info_ref = ob.solve_system(A_ref, I_ref, case.BC, case.solver_opts())
    assert abs(info.iterations - info_ref.iterations) <= 1
    assert np.abs(x - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
    assert info.res_norm < 1e-10
    case.close()


def test_neumann_rows_and_least_squares_match_oracle(ctx):
    """Reference test 21 (neumann-mass-flux on the supersonic full wing: control points at the panel centroids, rows
    n . B v of the doublet velocity influences, overdetermined least squares through A^T A): the CUDA matrix against the
    oracle's entry by entry, the least-squares solution, and the velocity sweep that feeds the post-processing."""
    case, _, _ = fixtures.make_case("test_21")
    assert case.n_cp > case.n_unknown
    ctx.set_case(case)
    I_known = ctx.assemble()
    A = ctx.get_A()
    A_ref, I_ref, S = ob.assemble(case, with_scale=True)
    assert A.shape == A_ref.shape == (case.n_cp, case.n_unknown)
    assert ((A == 0) == (A_ref == 0)).all()
    assert (np.abs(A - A_ref) / np.abs(A_ref).max(axis=1, keepdims=True)).max() < 1e-12
    assert np.abs(I_known - I_ref).max() <= 1e-13 * max(1e-300, np.abs(I_ref).max())
    x, info = ctx.solve(case.solver_opts(), case.BC)
    x_ref, info_ref = ob.solve_system(A_ref, I_ref, case.BC, case.solver_opts())
    assert abs(info.iterations - info_ref.iterations) <= max(2, info_ref.iterations // 50)
    assert abs(info.res_norm - info_ref.res_norm) < 1e-9          # the least-squares residual itself (not zero)
    assert np.abs(x - x_ref).max() <= 1e-7 * np.abs(x_ref).max()   # cond(A^T A) = cond(A)^2
    pts = case.inner_points()
    v = ctx.velocities_at(case, pts, x_ref)
    v_ref = ob.velocities_at(case, pts, x_ref)
    assert np.abs(v - v_ref).max() < 1e-12
    case.close()


def test_row_shard_assembles_the_same_rows(ctx):
    """Multi-GPU partitioning is by control-point rows: a shard's rows equal the same rows of the full matrix."""
    case, _, _ = fixtures.make_case("test_13")
    ctx.set_case(case)
    ctx.assemble()
    A_full = ctx.get_A()
    r0, nr = 100, 171
    ctx.set_case(case, row0=r0, nrows=nr)
    ctx.assemble()
    A_part = ctx.get_A()
    np.testing.assert_allclose(A_part, A_full[r0:r0 + nr], rtol=1e-13, atol=1e-18)
    case.close()


def _synthetic_case(kind, tmp_path):
    from machline_b200 import host, meshgen
    if kind == "sears_haack":      # BASELINE configs[2]: supersonic slender body, every pair DoD-tested, source-free
        pts, tris = meshgen.sears_haack(28, 14)
        meshgen.write_vtk(tmp_path / "m.vtk", pts, tris)
        return host.Case(meshgen.sears_haack_input("m.vtk", mach=2.0), base_dir=str(tmp_path))
    if kind == "wing":             # BASELINE configs[1] family (the bench workload, small): mirrored, wake, M = 0.5
        pts, tris = meshgen.swept_wing_half(20, 10)
        meshgen.write_vtk(tmp_path / "m.vtk", pts, tris)
        return host.Case(meshgen.wing_input("m.vtk", mach=0.5), base_dir=str(tmp_path))
    pts, tris = meshgen.icosphere(2)   # BASELINE configs[0] family
    meshgen.write_vtk(tmp_path / "m.vtk", pts, tris)
    return host.Case(meshgen.sphere_input("m.vtk"), base_dir=str(tmp_path))


@pytest.mark.parametrize("kind", ["sears_haack", "wing", "sphere"])
def test_synthetic_bench_families_match_oracle(ctx, kind, tmp_path):
    """The synthetic meshes bench.py / scripts/scale_run.py are run on, at sizes the oracle finishes in seconds: AIC entries,
    I_known, GMRES iteration count and the solution against the oracle."""
    case = _synthetic_case(kind, tmp_path)
    ctx.set_case(case)
    I_known = ctx.assemble()
    A = ctx.get_A()
    A_ref, I_ref, S = ob.assemble(case, with_scale=True)
    assert ((A == 0) == (A_ref == 0)).all()
    assert _rel_err(A, A_ref, S).max() < 2e-14
    assert (np.abs(A - A_ref) / np.abs(A_ref).max(axis=1, keepdims=True)).max() < 1e-12
    assert np.abs(I_known - I_ref).max() <= 1e-13 * max(1e-300, np.abs(I_ref).max())
    x, info = ctx.solve(case.solver_opts(), case.BC)
    x_ref, info_ref = ob.solve_system(A_ref, I_ref, case.BC, case.solver_opts())
    assert abs(info.iterations - info_ref.iterations) <= 1
    assert np.abs(x - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
    res, res_ref = case.post(x), case.post(x_ref)
    assert abs(res.C_p_max - res_ref.C_p_max) < 1e-9 and abs(res.C_p_min - res_ref.C_p_min) < 1e-9
    assert np.abs(np.array(res.C_F) - np.array(res_ref.C_F)).max() < 1e-9
    case.close()


def _offbody_cases():
    import json
    from pathlib import Path
    return json.loads((Path(__file__).resolve().parent / "golden" / "offbody_potentials.json").read_text())["cases"]


@pytest.mark.parametrize("c", _offbody_cases(), ids=lambda c: c["name"])
def test_reference_offbody_potentials_through_gpu(ctx, c):
    """The reference's stored off-body potentials (tests/test_oracle_offbody.py explains what they pin) through the CUDA
    path: ml_assemble + ml_solve on the surface, then ml_assemble with the 400 field points as rows."""
    from machline_b200 import host
    case = host.Case(c["input"], base_dir=fixtures.mesh_root())
    ctx.set_case(case)
    ctx.assemble()
    x, info = ctx.solve(case.solver_opts(), case.BC)
    U = float(np.linalg.norm(c["input"]["flow"]["freestream_velocity"]))
    pts = np.array(c["points"])
    phi_d, phi_s = ctx.potentials_at(case, pts, x)
    gold_d, gold_s = np.array(c["phi_d"]), np.array(c["phi_s"])
    assert np.abs(phi_s * U - gold_s).max() < 2e-11
    # and the influence rows themselves against the oracle's, entry by entry
    A_ref, I_ref = ob.assemble_at_points(case, pts)
    A = ctx.get_A()
    assert ((A == 0) == (A_ref == 0)).all()
    # far-field rows: every entry is a sum of O(1) angles cancelling to ~1e-6 (see _rel_err), so 1e-12 of the row maximum
    # is the libm noise floor here (measured: 97 of 563k entries between 1e-12 and 3e-12); the potentials they sum to agree
    # to 1e-13 absolute
    assert (np.abs(A - A_ref) / np.abs(A_ref).max(axis=1, keepdims=True).clip(1e-300)).max() < 1e-11
    assert np.abs(A @ x - A_ref @ x).max() * U < 1e-11
    assert np.abs(I_ref * U - gold_s).max() < 2e-11 and np.abs(phi_s - I_ref).max() * U < 1e-12
    if "supersonic" in c["name"]:
        assert np.abs(phi_d * U - gold_d).max() < 1e-10
    case.close()


@pytest.mark.parametrize("c", _offbody_cases(), ids=lambda c: c["name"])
def test_offbody_export_reproduces_the_reference_table(ctx, c, tmp_path):
    """output.offbody_points through the CUDA path (vtk_out.export_off_body_points = panel_solver_export_off_body_points,
    src/panel_solver.f90:2771-2895): the 24-column CSV against the table the reference itself wrote for the same input --
    phi_s and the three v_s columns at print precision, phi_d / v_d as far as the case's singular system allows
    (tests/test_oracle_offbody.py), and the velocity-influence rows entry by entry against the oracle."""
    from machline_b200 import host, vtk_out
    case = host.Case(c["input"], base_dir=fixtures.mesh_root())
    ctx.set_case(case)
    ctx.assemble()
    x, info = ctx.solve(case.solver_opts(), case.BC)
    pts = np.array(c["points"])
    pfile, ofile = tmp_path / "points.csv", tmp_path / "out" / "offbody.csv"
    pfile.write_text("x,y,z\n" + "".join("%.17g,%.17g,%.17g\n" % tuple(p) for p in pts))
    assert vtk_out.export_off_body_points(case, ctx, x, pfile, ofile) == len(pts)
    lines = ofile.read_text().split("\n")
    assert lines[0].strip().startswith("x,y,z,phi_inf,phi_d,phi_s,phi,Phi,v_inf_x") and len(lines[1]) == 24 * 20 + 23
    tab = np.genfromtxt(ofile, delimiter=",", skip_header=1)
    assert tab.shape == (len(pts), 24)
    supersonic = "supersonic" in c["name"]
    assert np.abs(tab[:, 5] - np.array(c["phi_s"])).max() < 2e-11
    gold_vs, gold_vd = np.array(c["v_s"]), np.array(c["v_d"])
    assert np.abs(tab[:, 14:17] - gold_vs).max() < 1e-10 * max(1.0, np.abs(gold_vs).max())
    if supersonic:
        assert np.abs(tab[:, 4] - np.array(c["phi_d"])).max() < 1e-10
        assert np.abs(tab[:, 12] - gold_vd[:, 1]).max() < 1e-9
        assert np.abs(tab[:, 11:14] - gold_vd).max() < 1e-4 * np.abs(gold_vd).max()
    # internal consistency of the derived columns
    assert np.abs(tab[:, 6] - (tab[:, 4] + tab[:, 5])).max() < 1e-11 * max(1.0, np.abs(tab[:, 6]).max())
    assert np.abs(tab[:, 23] - np.linalg.norm(tab[:, 20:23], axis=1)).max() < 1e-10 * np.abs(tab[:, 23]).max()
    # velocity-influence rows against the oracle's
    V_ref, vs_ref = ob.velocity_influences_at(case, pts, with_wake=False)
    v_d, v_s = ctx.velocity_parts_at(case, pts, x, with_wake=False)
    vd_ref = np.stack([V_ref[k] @ x for k in range(3)], axis=1)
    assert np.abs(v_s - vs_ref).max() < 1e-12 * max(1.0, np.abs(vs_ref).max())
    assert np.abs(v_d - vd_ref).max() < 1e-9 * max(1.0, np.abs(vd_ref).max())
    case.close()


def test_run_case_writes_the_reference_outputs(tmp_path):
    """solver.run_case = `machline.exe input.json`: report.json, the body results VTK and the iteration history appear where
    the input asks for them, and the numbers in them are the solved ones."""
    import json
    from machline_b200 import solver
    inp, expect, tol = fixtures.golden_input("test_08")
    inp = json.loads(json.dumps(inp))
    inp["solver"]["iterative_solver_output"] = str(tmp_path / "iterations.csv")
    inp["output"] = {"report_file": str(tmp_path / "report.json"), "body_file": str(tmp_path / "results" / "body.vtk"),
                     "wake_file": str(tmp_path / "results" / "wake.vtk"), "control_point_file": str(tmp_path / "results" / "cp.vtk"),
                     "mirrored_body_file": str(tmp_path / "results" / "mirror.vtk"),
                     "offbody_points": {"points_file": str(tmp_path / "pts.csv"), "output_file": str(tmp_path / "results" / "off.csv")}}
    (tmp_path / "pts.csv").write_text("x,y,z\n2.0,0.0,0.0\n0.0,3.0,0.5\n")
    res = solver.run_case(inp, base_dir=fixtures.mesh_root())
    assert (tmp_path / "results" / "cp.vtk").exists() and len((tmp_path / "results" / "off.csv").read_text().split("\n")) == 4
    assert not (tmp_path / "results" / "wake.vtk").exists() and not (tmp_path / "results" / "mirror.vtk").exists()   # sphere: neither
    cpl = (tmp_path / "results" / "cp.vtk").read_text().split("\n")
    i0 = cpl.index("SCALARS residual float 1") + 2
    assert max(abs(float(v)) for v in cpl[i0:i0 + 10]) < 1e-11
    assert abs(res.C_p_max - expect[0]) < tol[0] and abs(res.C_p_min - expect[1]) < tol[1]
    rep = json.loads((tmp_path / "report.json").read_text())
    assert rep["solver_results"]["iterations"] == res.iterations
    hist = (tmp_path / "iterations.csv").read_text().split("\n")
    assert hist[1] == " GMRES" and len([ln for ln in hist[4:] if ln]) == res.iterations
    vtk = (tmp_path / "results" / "body.vtk").read_text().split("\n")
    i0 = vtk.index("SCALARS C_p_inc float 1") + 2
    cp = np.array([float(v) for v in vtk[i0:i0 + len(res.C_p)]])
    assert abs(cp.max() - res.C_p_max) < 1e-10 and abs(cp.min() - res.C_p_min) < 1e-10


def _sv_cases():
    import json
    from pathlib import Path
    return json.loads((Path(__file__).resolve().parent / "golden" / "aic_singular_values.json").read_text())["cases"]


@pytest.mark.parametrize("c", _sv_cases(), ids=lambda c: c["name"])
def test_gpu_aic_has_the_reference_singular_values(ctx, c):
    """Extreme singular values of the reference's own AIC matrix (tests/test_oracle_singular_values.py) from the matrix the
    CUDA assembly builds: pins the entries of the GPU matrix against the reference directly, not through the oracle."""
    from machline_b200 import host
    case = host.Case(c["input"], base_dir=fixtures.mesh_root())
    ctx.set_case(case)
    ctx.assemble()
    S = np.linalg.svd(ctx.get_A(), compute_uv=False)
    assert abs(S[0] - c["S_max"]) < 3e-13 and abs(S[-1] - c["S_min"]) < 3e-13
    case.close()


def test_check_system_statuses(ctx):
    """ml_check_system = panel_solver_check_system (src/panel_solver.f90:1709-1764): status 0 on a healthy system, 2 when a row
    (control point not influenced) or a column (unknown without influence) is entirely zero, 1 on a NaN in A or b."""
    case, _, _ = fixtures.make_case("test_08")
    ctx.set_case(case)
    ctx.assemble()
    assert ctx.check_system(case.BC) == (0, 0, 0)
    A = ctx.get_A()
    B = A.copy()
    B[5, :] = 0.
    B[11, :] = 0.
    B[:, 7] = 0.
    ctx.set_A(B)
    assert ctx.check_system(case.BC) == (2, 2, 1)
    B = A.copy()
    B[3, 4] = np.nan
    ctx.set_A(B)
    assert ctx.check_system(case.BC)[0] == 1
    ctx.set_A(A)
    bc = np.array(case.BC, dtype=np.float64)
    bc[2] = np.nan
    assert ctx.check_system(bc)[0] == 1
    assert ctx.check_system(case.BC) == (0, 0, 0)
    case.close()
