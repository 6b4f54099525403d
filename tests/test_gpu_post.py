"""Post-processing on the device (ml_post_process, csrc/gpu/post.cu): cell velocities, pressure rules, forces
and moments of lower-order panels (src/panel_solver.f90:2030-2615, src/panel.f90:3415-3512, src/flow.f90:313-585).

CPU part: the tables the host library prepares for the kernel (mlh_case_post_tables) carry everything Case::post uses -- a numpy
model of the kernel on those tables reproduces the host post-processing for a random x.  GPU part: the kernel against the host
library per cell and per rule, and -- through the whole CUDA path, x never leaving the device between solve and post -- against
the reference's golden tuples."""
import numpy as np
import pytest

import fixtures
from machline_b200 import _abi

HIGHER_ORDER = {"test_02", "test_04", "test_06", "test_11", "test_16", "test_17"}
LOWER_ORDER_CASES = [n for n in fixtures.golden_case_names() if n not in HIGHER_ORDER]


def _arr(ptr, shape, dtype=np.float64):
    return np.ctypeslib.as_array(ptr, shape=shape).astype(dtype, copy=True)


def _model(t, f, x):
    """The arithmetic of post_cells_kernel / post_moments_sums_kernel in numpy (inc and isentropic rules only)."""
    n = t.n_cells
    mi = _arr(t.mu_index, (n, 3), np.int64)
    T = _arr(t.T_mu, (n, 3, 3))
    A = _arr(t.A_g_to_ls, (n, 3, 3))
    s_dir = _arr(t.s_dir, (n, 3))
    si = _arr(t.sigma_index, (n,), np.int64)
    sk = _arr(t.sigma_known, (n,))
    vin = _arr(t.v_inner, (n, 3))
    ng = _arr(t.n_g, (n, 3))
    area = _arr(t.area, (n,))
    ce = _arr(t.centr, (n, 3))
    fc = _arr(t.force_cell, (n,), np.int64)
    mu_v = np.where(mi >= 0, x[np.maximum(mi, 0)], 0.)
    mu_p = np.einsum("nrk,nk->nr", T, mu_v)
    dv = A[:, 0, :] * mu_p[:, 1:2] + A[:, 1, :] * mu_p[:, 2:3]
    sg = np.where(si >= 0, x[np.maximum(si, 0)], sk)
    dv = dv + sg[:, None] * s_dir
    V = f.U * (vin + dv)
    cp_inc = 1. - (V * V).sum(axis=1) * f.U_inv * f.U_inv
    cps = {"incompressible": cp_inc}
    if f.rules & 2:
        c = f.a_ise * (np.power(1. + f.b_ise * cp_inc, f.c_ise) - 1.)
        cps["isentropic"] = np.where(np.isnan(c), f.C_P_vac, c)
    cpf = cps[_abi.RULES[f.force_rule]]
    dCf = (-cpf * area)[:, None] * ng
    CF = dCf.sum(axis=0) / f.S_ref
    CM = np.cross(ce - np.array(f.CG[:]), dCf[fc]).sum(axis=0) / f.l_ref
    if f.mirrored_symmetric:
        k = f.mirror_plane - 1
        CF = 2. * CF
        CF[k] = 0.
        CM = np.array([2. * CM[i] if i == k else 0. for i in range(3)])
    return V, cps, dCf, CF, CM


@pytest.mark.parametrize("name", ["test_07", "test_01", "test_13", "test_05", "test_19", "test_21", "test_15"])
def test_post_tables_reproduce_host_post(name):
    case, _, _ = fixtures.make_case(name)
    rng = np.random.default_rng(7)
    x = rng.standard_normal(case.n_unknown) * 0.1
    # Neumann formulations (test 21): the inner velocity of every cell is an input (a GPU sweep in a real run)
    v_inner = None if case.dirichlet else rng.standard_normal((len(case.inner_points()), 3)) * 0.05
    res = case.post(x, v_inner)
    t, f = case.post_tables(v_inner)
    if _abi.RULES[f.force_rule] not in ("incompressible", "isentropic"):
        pytest.skip("the numpy model covers the incompressible and isentropic rules")
    V, cps, dCf, CF, CM = _model(t, f, x)
    sc = max(1., np.abs(res.V_cells).max())
    assert np.abs(V - res.V_cells).max() < 1e-13 * sc
    rep = "incompressible" if f.rules & 1 else "isentropic"
    ref_cp = case.result_array(rep)
    ok = np.isfinite(ref_cp) & (np.abs(ref_cp) < 1e6)
    assert np.abs(cps[rep] - ref_cp)[ok].max() < 1e-11 * max(1., np.abs(ref_cp[ok]).max())
    assert np.abs(dCf - case.result_array("dC_f")).max() < 1e-11 * max(1., np.abs(dCf).max())
    assert np.abs(CF - res.C_F).max() < 1e-11 * max(1., np.abs(res.C_F).max())
    assert np.abs(CM - res.C_M).max() < 1e-11 * max(1., np.abs(res.C_M).max())
    case.close()


def test_post_tables_refuse_higher_order():
    from machline_b200 import host
    case, _, _ = fixtures.make_case("test_06")
    with pytest.raises(host.MachLineError, match="lower-order"):
        case.post_tables()
    case.close()


@pytest.fixture(scope="module")
def ctx():
    from machline_b200 import gpu
    c = gpu.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", LOWER_ORDER_CASES)
def test_device_post_matches_host_and_reference_goldens(ctx, name):
    """host setup -> ml_assemble -> ml_solve -> ml_post_process (x stays on the device) == host post-processing per cell and per
    rule, and == the reference's golden tuple at the tolerance of tests/test_gpu_parity.py."""
    from test_gpu_parity import ILL_CONDITIONED
    case, expect, tol = fixtures.make_case(name)
    ctx.set_case(case)
    ctx.assemble()
    x, info = ctx.solve(case.solver_opts(), case.BC)
    # Neumann formulations: the velocity sweep replaces the resident tables but not the solution the solve left on the device
    v_inner = None if case.dirichlet else ctx.velocities_at(case, case.inner_points(), x)
    dev = ctx.post_process(case, v_inner)
    res = case.post(x, v_inner)
    sc = max(1., np.abs(res.V_cells).max())
    assert np.abs(dev["V_cells"] - res.V_cells).max() <= 1e-14 * sc
    for rule, got in dev["C_p"].items():
        ref = case.result_array(rule)
        assert ref.shape == got.shape, rule
        fin = np.isfinite(ref)
        assert (np.isfinite(got) == fin).all(), rule
        # identical operations except pow (isentropic) and the order of nothing: 1e-13 of the scale
        assert np.abs(got - ref)[fin].max() <= 1e-13 * max(1., np.abs(ref[fin]).max()), rule
    assert np.abs(dev["dC_f"] - case.result_array("dC_f")).max() <= 1e-13 * max(1e-300, np.abs(dev["dC_f"]).max())
    # the sums run in another order (strided partial sums + one tree against the host's sequential loop): a few ulp of the
    # sum of the absolute contributions (the sphere of tests 07-09 has radius 60 with S_ref = l_ref = 1: contributions of 1e6
    # cancel to coefficients of 1e-3)
    t, f = case.post_tables(v_inner)
    n = t.n_cells
    ce = np.ctypeslib.as_array(t.centr, shape=(n, 3))
    fc = np.ctypeslib.as_array(t.force_cell, shape=(n,))
    dCf = case.result_array("dC_f")
    mult = 2. if f.mirrored_symmetric else 1.
    scale_F = mult * np.abs(dCf).sum(axis=0).max() / f.S_ref
    scale_M = mult * np.abs(np.cross(ce - np.array(f.CG[:]), dCf[fc])).sum(axis=0).max() / f.l_ref
    assert np.abs(dev["C_F"] - res.C_F).max() <= 1e-14 * max(scale_F, 1e-300)
    assert np.abs(dev["C_M"] - res.C_M).max() <= 1e-14 * max(scale_M, 1e-300)
    assert abs(dev["C_p_max"] - res.C_p_max) <= 1e-13 * max(1., abs(res.C_p_max))
    assert abs(dev["C_p_min"] - res.C_p_min) <= 1e-13 * max(1., abs(res.C_p_min))
    # the reference's golden tuple from the device results alone
    s_cp, s_f = ILL_CONDITIONED.get(name, (1., 1.))
    got = [dev["C_p_max"], dev["C_p_min"], *[float(v) for v in dev["C_F"]]]
    # (force columns: the reference's tolerance, or the summation-order bound above where that is larger -- the sphere)
    floor = [0., 0., 1e-14 * scale_F, 1e-14 * scale_F, 1e-14 * scale_F]
    for g, e, tl, sl, fl, lab in zip(got, expect, tol, [s_cp, s_cp, s_f, s_f, s_f], floor, ["C_p_max", "C_p_min", "Cx", "Cy", "Cz"]):
        assert abs(g - e) < max(tl * sl, fl), f"{lab}: got {g!r}, reference {e!r}"
    case.close()


@pytest.mark.gpu
def test_device_post_needs_a_solution_and_checks_its_tables(ctx):
    from machline_b200 import gpu
    case, _, _ = fixtures.make_case("test_07")
    fresh = gpu.Context(0)
    with pytest.raises(gpu.GpuError) as e:
        fresh.post_process(case)
    assert e.value.status == 10   # ML_BAD_ARGUMENT: no solution on the device
    x = np.zeros(case.n_unknown)
    out = fresh.post_process(case, x=x)   # zero strengths: the cell velocity is the inner flow + sigma s_dir
    assert np.isfinite(out["V_cells"]).all()
    fresh.close()
    case.close()


@pytest.mark.parametrize("compressible", [True, False])
def test_every_pressure_rule_against_the_reference_formulas(compressible):
    """The rules no reference golden exercises (second-order, slender-body, linear) and the three subsonic corrections, evaluated by
    the host library (pressure_rules.hpp, the header the device kernel compiles too) against an independent numpy statement of
    src/flow.f90:313-508 on the cell velocities: a compressible subsonic case with the five rules, and an incompressible one with
    the corrections (which the reference allows at freestream_mach_number = 0 only)."""
    import copy
    from machline_b200 import host
    inp, _, _ = fixtures.golden_input("test_05")
    inp = copy.deepcopy(inp)
    pp = inp.setdefault("post_processing", {})
    if compressible:
        inp["flow"]["freestream_mach_number"] = 0.5
        pp["pressure_rules"] = {"incompressible": True, "isentropic": True, "second-order": True, "slender-body": True, "linear": True}
        pp["pressure_for_forces"] = "second-order"
        expect_rules = {1, 2, 3, 4}   # the reference drops the incompressible rule in a compressible flow
    else:
        pp["pressure_rules"] = {"incompressible": True, "second-order": True, "slender-body": True, "linear": True}
        pp["subsonic_pressure_correction"] = {"correction_mach_number": 0.4, "prandtl-glauert": True, "karman-tsien": True, "laitone": True}
        pp["pressure_for_forces"] = "karman-tsien"
        expect_rules = {0, 2, 3, 4, 5, 6, 7}
    case = host.Case(inp, base_dir=fixtures.mesh_root())
    x = np.random.default_rng(11).standard_normal(case.n_unknown) * 0.05
    res = case.post(x)
    t, f = case.post_tables()
    assert {r for r in range(8) if f.rules & (1 << r)} == expect_rules
    V = res.V_cells
    U_inv, M, g, Mc = f.U_inv, f.M_inf, f.gamma, f.M_inf_corr
    A = np.array(f.A_g_to_c[:]).reshape(3, 3)
    vp = (V - np.array(f.v_inf[:])) @ A.T                                  # flow_get_v_pert_c :345-356
    clip = lambda c: np.minimum(np.maximum(c, f.C_P_vac), f.C_P_stag)      # restrict_pressure
    inc = 1. - (V * V).sum(axis=1) * U_inv * U_inv                         # :313-324
    with np.errstate(invalid="ignore", divide="ignore"):
        ise = f.a_ise * (np.power(1. + f.b_ise * inc, f.c_ise) - 1.)       # :327-342
        ise = np.where(np.isnan(ise), f.C_P_vac, ise)
        lin = clip(-2. * vp[:, 0] * U_inv)                                 # :416-432
        sln = clip(lin - (vp[:, 1] ** 2 + vp[:, 2] ** 2) * U_inv ** 2)     # :397-413
        snd = clip(sln - (1. - M * M) * vp[:, 0] ** 2 * U_inv ** 2)        # :359-394
        s = np.sqrt(1. - Mc * Mc)
        pg = inc / s                                                       # :453-466
        kt = inc / (s + 0.5 * (Mc * Mc / (1. + s)) * inc)                  # :469-487
        lai = inc / (s + (Mc * Mc * (1. + 0.5 * (g - 1.) * Mc * Mc) / (2. * s)) * inc)   # :490-508
    want = {"incompressible": inc, "isentropic": ise, "second-order": snd, "slender-body": sln, "linear": lin,
            "prandtl-glauert": pg, "karman-tsien": kt, "laitone": lai}
    for r, (rule, ref) in enumerate(want.items()):
        assert _abi.RULES[r] == rule
        if r not in expect_rules:
            continue
        got = case.result_array(rule)
        assert got.shape == ref.shape, rule
        assert np.abs(got - ref).max() <= 1e-13 * max(1., np.abs(ref).max()), rule
    n = t.n_cells
    ng = np.ctypeslib.as_array(t.n_g, shape=(n, 3))
    area = np.ctypeslib.as_array(t.area, shape=(n,))
    dCf = (-(snd if compressible else kt) * area)[:, None] * ng
    assert np.abs(dCf - case.result_array("dC_f")).max() <= 1e-13 * np.abs(dCf).max()
    case.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["test_07", "test_13", "test_21"])
def test_run_case_device_post(name):
    """solver.run_case(device_post=True): the public entry point with the device-resident post-processing next to the host's."""
    from machline_b200 import solver
    inp, _, _ = fixtures.golden_input(name)
    r = solver.run_case(inp, base_dir=fixtures.mesh_root(), device_post=True)
    d = r.device_post
    assert d is not None and len(d["C_p"]) >= 1
    assert abs(d["C_p_max"] - r.C_p_max) <= 1e-13 * max(1., abs(r.C_p_max))
    assert abs(d["C_p_min"] - r.C_p_min) <= 1e-13 * max(1., abs(r.C_p_min))
    rep = "incompressible" if "incompressible" in d["C_p"] else "isentropic"
    assert np.abs(d["C_p"][rep] - r.C_p).max() <= 1e-13 * max(1., np.abs(r.C_p).max())
