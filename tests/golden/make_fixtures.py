"""Generates the committed golden fixtures from the read-only reference tree.

Run in the BUILD container only (`python tests/golden/make_fixtures.py`): /root/reference does not
exist on the GPU box, and nothing under tests/ reads it at test time.

Outputs (all under tests/golden/):
  meshes.npz                -- the raw point / triangle arrays of the reference's test meshes
                               (test/meshes/*.vtk, *.stl), pre-duplicate-collapse, so that the host
                               loader sees exactly the same vertex list and ordering.
  reference_goldens.json    -- the (C_p_max, C_p_min, Cx, Cy, Cz) tuples and tolerances asserted by
                               test/test_machline.py, with the input each test builds.
  offbody_potentials.json   -- the reference's stored off-body results test/input_files/half_wing_{inc,supersonic}_offbody_points_correct.csv
                               (400 field points in the root xz plane; written by panel_solver_export_off_body_points,
                               src/panel_solver.f90:2771-2895, format e20.13): x, y, z, phi_d, phi_s and the inputs that the
                               (commented-out) tests 22 / 23 of test/test_machline.py:636-730 build.  They pin the potential
                               integrals at arbitrary field points.
  solver_histories.json     -- the reference's stored per-iteration histories studies/matrix_solvers/iterations/*_prec_history.csv
                               (written by its own GMRES / block_jacobi_solve / block_ssor_solve through
                               solver.iterative_solver_output, '(i6, a, ES10.3)') for the cone (coarse, medium) and the
                               diamond wing (coarse), preconditioner DIAG / none, sorted / unsorted, with the inputs of
                               studies/matrix_solvers/matrix_solver_study.py:69-100, 246-252.  They pin assembly + sort +
                               "preconditioner" + each iterative solver, iteration by iteration.
  aic_singular_values.json  -- largest and smallest singular value of the reference's own AIC matrix (numpy SVD of the A_mat.txt
                               it wrote with solver.write_A_and_b; studies/matrix_solvers/*_condition_data.csv,
                               matrix_conditions.py:20-36) for the cone and diamond-wing study meshes: pins the matrix ENTRIES.
  prototype_integrals.json  -- known-answer H(1,1,1) / hH(1,1,3) / F(1,1,1) values computed by
                               IMPORTING the reference's Python prototype dev/unit_tests/panel.py
                               (quadrilateral panels in local coordinates) at fixed points.
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def read_vtk_v3(path: Path):
    lines = path.read_text().split("\n")
    assert "Version 3" in lines[0], lines[0]
    toks = " ".join(lines[4:]).split()
    assert lines[4].split()[0] == "POINTS"
    n = int(lines[4].split()[1])
    body = " ".join(lines[5:]).split()
    pts = np.array(body[: 3 * n], dtype=np.float64).reshape(n, 3)
    rest = body[3 * n:]
    assert rest[0] == "POLYGONS"
    m = int(rest[1])
    vals = np.array(rest[3: 3 + 4 * m], dtype=np.int64).reshape(m, 4)
    assert (vals[:, 0] == 3).all()
    # keep the decimal strings: they are the ground truth the Fortran reader parses
    pts_txt = np.array(body[: 3 * n]).reshape(n, 3)
    return pts, vals[:, 1:4].astype(np.int32), pts_txt


def read_stl(path: Path):
    pts, txt = [], []
    for line in path.read_text().split("\n"):
        w = line.split()
        if len(w) == 4 and w[0] == "vertex":
            pts.append([float(w[1]), float(w[2]), float(w[3])])
            txt.append(w[1:4])
    return np.array(pts, dtype=np.float64), np.array(txt)


def make_meshes():
    out = {}
    for name in ["sphere", "diamond_half_wing", "half_wing_right", "commercial_fuselage", "full_wing"]:
        pts, tris, txt = read_vtk_v3(REF / "test" / "meshes" / f"{name}.vtk")
        # round-trip check: repr() of the parsed double parses back to the same double
        assert all(float(repr(float(v))) == float(v) for v in pts.ravel()[:100])
        out[f"{name}.vtk:points"] = pts
        out[f"{name}.vtk:triangles"] = tris
    pts, txt = read_stl(REF / "test" / "meshes" / "diamond_full_wing.stl")
    out["diamond_full_wing.stl:facet_vertices"] = pts
    # meshes of studies/matrix_solvers (the stored iteration histories below were produced on them)
    study = REF / "studies" / "matrix_solvers" / "meshes"
    for name in ["cone_10_deg_coarse", "cone_10_deg_medium"]:
        pts, tris, txt = read_vtk_v3(study / f"{name}.vtk")
        out[f"{name}.vtk:points"] = pts
        out[f"{name}.vtk:triangles"] = tris
    pts, txt = read_stl(study / "diamond_5_deg_full_coarse.stl")
    out["diamond_5_deg_full_coarse.stl:facet_vertices"] = pts
    np.savez_compressed(OUT / "meshes.npz", **out)
    print("meshes.npz:", {k: v.shape for k, v in out.items()})


def read_vtk_v5(path: Path):
    """ASCII VTK v5 POLYDATA (OFFSETS / CONNECTIVITY); triangles only.  As src/vtk.f90:560-640 reads it the panel count is
    the POLYGONS count minus one (the count on that line is the number of offsets)."""
    toks = path.read_text().split()
    i = toks.index("POINTS")
    n = int(toks[i + 1])
    pts = np.array(toks[i + 3: i + 3 + 3 * n], dtype=np.float64).reshape(n, 3)
    j = toks.index("POLYGONS") if "POLYGONS" in toks else toks.index("CELLS")   # UNSTRUCTURED_GRID files say CELLS
    m = int(toks[j + 1]) - 1
    k = toks.index("CONNECTIVITY", j)
    tris = np.array(toks[k + 2: k + 2 + 3 * m], dtype=np.int64).reshape(m, 3)
    offs = np.array(toks[toks.index("OFFSETS", j) + 2: toks.index("OFFSETS", j) + 3 + m], dtype=np.int64)
    assert (np.diff(offs) == 3).all()
    return pts, tris.astype(np.int32)


def read_tri(path: Path):
    toks = path.read_text().split()
    n, m = int(toks[0]), int(toks[1])
    pts = np.array(toks[2: 2 + 3 * n], dtype=np.float64).reshape(n, 3)
    tris = np.array(toks[2 + 3 * n: 2 + 3 * n + 3 * m], dtype=np.int64).reshape(m, 3) - 1
    return pts, tris.astype(np.int32)


def make_study_meshes():
    """The meshes of BASELINE.json configs[1]-[3] (SURVEY 8(d) "Concrete inputs"): ONERA M6 fine, cone fine, Sears-Haack
    160x60, AGARD-B coarse / fine.  Kept in a second archive so that the small one stays small."""
    S = REF / "studies"
    out = {}
    for name, path in [("agard_b_coarse.vtk", S / "supersonic_agard_b_wing_body/meshes/agard_b_coarse.vtk"),
                       ("agard_b_fine.vtk", S / "supersonic_agard_b_wing_body/meshes/agard_b_fine.vtk")]:
        pts, tris = read_vtk_v5(path)
        out[f"{name}:points"], out[f"{name}:triangles"] = pts, tris
    pts, tris, _ = read_vtk_v3(S / "supersonic_cone/meshes/cone_10_deg_fine.vtk")
    out["cone_10_deg_fine.vtk:points"], out["cone_10_deg_fine.vtk:triangles"] = pts, tris
    pts, tris = read_tri(S / "sears_haack/meshes/SH_160_60.tri")
    out["SH_160_60.tri:points"], out["SH_160_60.tri:triangles"] = pts, tris
    pts, _ = read_stl(S / "subsonic_onera_m6_wing/meshes/M6_onera_fine.stl")
    out["M6_onera_fine.stl:facet_vertices"] = pts
    np.savez_compressed(OUT / "study_meshes.npz", **out)
    print("study_meshes.npz:", {k: v.shape for k, v in out.items()})


# (C_p_max, C_p_min, Cx, Cy, Cz) and tolerances, transcribed from test/test_machline.py (line numbers
# of the asserts).  "alter" is the list of (dotted key, value) edits the test applies to the base input.
GOLDENS = [
    dict(name="test_01", lines="73-89", input="half_wing_input.json", alter=[],
         expect=[0.749340997284039, -1.280587303416, -0.393828202123999, -0.0468973172312513, 20.6476349003493],
         tol=[1e-10, 1e-9, 1e-9, 1e-9, 1e-8]),
    dict(name="test_03", lines="124-153", input="half_wing_input.json",
         alter=[["solver.formulation", "dirichlet-morino"]],
         expect=[0.749339435828545, -1.28056616763669, -0.393823334747539, -0.0468966684057296, 20.6473533100845],
         tol=[1e-10, 1e-9, 1e-9, 1e-9, 1e-8]),
    dict(name="test_05", lines="189-218", input="half_wing_input.json",
         alter=[["flow.freestream_velocity", [100.0, 0.0, 0.0]], ["solver.formulation", "dirichlet-morino"]],
         expect=[0.221628441564136, -0.427625486100214, 0.301682907690379, 0.0, 3.22319026937329e-12],
         tol=[1e-12, 1e-12, 1e-12, 1e-12, 1e-12]),
    dict(name="test_07", lines="254-283", input="sphere_input.json",
         alter=[["solver.matrix_solver", "BJAC"], ["solver.relaxation", 0.9]],
         expect=[0.99114275830853, -1.23782791839695, 0.0, 0.0, 0.0], tol=[1e-12, 1e-12, 2e-5, 2e-5, 2e-5]),
    dict(name="test_08", lines="286-297", input="sphere_input.json", alter=[],
         expect=[0.991142758308531, -1.23782791839695, 0.0, 0.0, 0.0], tol=[1e-12, 1e-12, 1e-4, 1e-4, 1e-4]),
    # test_09 writes an altered (neumann-mass-flux) input but RUNS the unaltered sphere_input.json (test_machline.py:304-310):
    # what it pins is the default Morino sphere, with its own force tolerances
    dict(name="test_09", lines="300-322", input="sphere_input.json", alter=[],
         expect=[0.991142758308531, -1.23782791839695, 0.0, 0.0, 0.0], tol=[1e-12, 1e-12, 1e-4, 1e-4, 1e-4]),
    dict(name="test_12", lines="416-429", input="compressible_half_wing_input.json", alter=[],
         expect=[0.817285785213847, -1.09376107707523, -0.508549885785561, 0.0102005794126269, 26.2795099006351],
         tol=[1e-9, 1e-9, 1e-9, 1e-9, 1e-8]),
    dict(name="test_13", lines="432-445", input="supersonic_half_wing_input.json", alter=[],
         expect=[0.121697468024553, -0.116094421618336, 0.1424008894449, 0.0, 0.0],
         tol=[1e-12, 1e-12, 1e-12, 1e-12, 1e-12]),
    dict(name="test_14", lines="448-477", input="supersonic_half_wing_input.json",
         alter=[["solver.formulation", "dirichlet-source-free"]],
         expect=[0.121697468024034, -0.116094421600922, 0.142400889444425, 0.0, 0.0],
         tol=[1e-12, 1e-12, 1e-12, 1e-12, 1e-12]),
    dict(name="test_15", lines="480-510", input="supersonic_half_wing_input.json",
         alter=[["flow.freestream_velocity", [100.0, 5.0, 5.0]], ["geometry.wake_model.append_wake", True]],
         expect=[0.194725933694785, -0.298780589679282, 0.142781624853592, 0.000852303050093563, 0.892593299837566],
         tol=[1e-12, 1e-12, 1e-12, 1e-12, 1e-9]),
    dict(name="test_18", lines="578-609", input="supersonic_half_wing_input.json",
         alter=[["flow.freestream_velocity", [100.0, 0.0, 5.0]], ["geometry.wake_model.append_wake", True],
                ["solver.formulation", "dirichlet-source-free"]],
         expect=[0.194950373758414, -0.291642542743292, 0.142905077607286, 0.0, 0.893492282700425],
         tol=[1e-12, 1e-12, 1e-12, 1e-12, 1e-10]),
    dict(name="test_19", lines="612-625", input="fuselage_input.json", alter=[],
         expect=[0.955084903205978, -0.574040916470699, 0.0, 0.0, 0.0], tol=[1e-12, 1e-12, 2.1e-3, 2.1e-3, 2.1e-3]),
    dict(name="test_21", lines="610-633", input="supersonic_full_wing_input.json",
         alter=[["solver.formulation", "neumann-mass-flux"], ["solver.matrix_solver", "GMRES"]],
         expect=[0.205581816085009, -0.265679125725778, 0.0721530311378978, 0.0, 0.431906841113295],
         tol=[1e-9, 1e-9, 1e-9, 1e-11, 1e-9]),
    # higher-order distributions (quadratic doublets, linear sources): geometry.singularity_order = "higher"
    dict(name="test_02", lines="91-119", input="half_wing_input.json", alter=[["geometry.singularity_order", "higher"]],
         expect=[0.749309346263522, -1.11542521362851, -0.376594227257231, -0.0447748437467464, 20.1915074613927],
         tol=[1e-10, 1e-9, 1e-9, 1e-9, 1e-8],
         # The source-free AIC of this case is singular (numpy.linalg.solve refuses it); GMRES stops at ||r|| = 5e-13 on a
         # solution that depends on the rounding of every entry, so two conforming libms give tuples 1e-8 apart.  The oracle
         # (glibc) lands at |dCz| = 1.5e-8 from the reference's (gfortran + its libm) value; all other entries are inside
         # the reference's own tolerance, as are tests 04 / 06 / 16 / 17 (the latter three to 1e-15).
         oracle_slack=2.0),
    dict(name="test_04", lines="153-182", input="half_wing_input.json",
         alter=[["solver.formulation", "dirichlet-morino"], ["geometry.singularity_order", "higher"]],
         expect=[0.750884716128262, -1.10180792847574, -0.379286087291006, -0.045939589132209, 20.2220107466176],
         tol=[1e-10, 1e-9, 1e-9, 1e-9, 1e-8]),
    dict(name="test_06", lines="217-247", input="half_wing_input.json",
         alter=[["flow.freestream_velocity", [100.0, 0.0, 0.0]], ["solver.formulation", "dirichlet-morino"],
                ["geometry.singularity_order", "higher"]],
         expect=[0.217923409759566, -0.432426859568809, 0.308691949491238, 0.0, 0.0],
         tol=[1e-12, 1e-12, 1e-12, 1e-12, 1e-11]),
    dict(name="test_16", lines="496-519", input="supersonic_half_wing_input.json",
         alter=[["solver.formulation", "dirichlet-source-free"], ["geometry.singularity_order", "higher"]],
         expect=[0.119055399649912, -0.112090098429673, 0.142437748460037, 0.0, 0.0],
         tol=[1e-12, 1e-12, 1e-12, 1e-12, 1e-10]),
    dict(name="test_17", lines="522-546", input="supersonic_half_wing_input.json",
         alter=[["flow.freestream_velocity", [100.0, 5.0, 5.0]], ["geometry.wake_model.append_wake", True],
                ["geometry.singularity_order", "higher"]],
         expect=[0.195559054627881, -0.291045728014071, 0.142750535947053, 0.000816008709771361, 0.88869372122009],
         tol=[1e-12, 1e-12, 1e-12, 1e-12, 1e-9]),
    dict(name="test_20", lines="628-641", input="supersonic_full_wing_input.json", alter=[],
         expect=[0.194950351346633, -0.324945660429385, 0.0718540012154408, 0.0, 0.429236847680447],
         tol=[1e-12, 1e-12, 1e-12, 1e-11, 1e-12]),
]


# Tests that compare two runs with each other instead of with stored numbers (test_machline.py:325-363): the full wing against
# the mirrored half wing, lower-order Morino, V = (100, 0, 10); tolerances on |full - half| of (C_p_max, C_p_min, Cx, Cy, Cz).
COMPARISONS = [
    dict(name="test_10", lines="325-363", inputs=["full_wing_input.json", "half_wing_input.json"],
         alter=[["flow.freestream_velocity", [100.0, 0.0, 10.0]], ["solver.formulation", "dirichlet-morino"]],
         tol=[1e-3, 2e-3, 1e-3, 1e-3, 2e-2]),
    dict(name="test_11", lines="366-406", inputs=["full_wing_input.json", "half_wing_input.json"],
         alter=[["flow.freestream_velocity", [100.0, 0.0, 10.0]], ["solver.formulation", "dirichlet-morino"],
                ["geometry.singularity_order", "higher"]],
         tol=[1e-3, 2e-3, 1e-3, 1e-3, 2e-2]),
]


def make_goldens():
    inputs = {}
    for cmp_ in COMPARISONS:
        for nm in cmp_["inputs"]:
            inputs[nm] = json.loads((REF / "test" / "input_files" / nm).read_text())
    for g in GOLDENS:
        if g["input"] not in inputs:
            inputs[g["input"]] = json.loads((REF / "test" / "input_files" / g["input"]).read_text())
    doc = dict(source="usuaero/MachLine test/test_machline.py (golden tuples C_p_max, C_p_min, Cx, Cy, Cz)",
               inputs=inputs, cases=GOLDENS, comparisons=COMPARISONS)
    (OUT / "reference_goldens.json").write_text(json.dumps(doc, indent=1))
    print("reference_goldens.json:", len(GOLDENS), "cases")


def make_prototype_integrals():
    """Import dev/unit_tests/panel.py (matplotlib stubbed) and tabulate its integrals."""
    stub = types.ModuleType("matplotlib")
    stub.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", stub)
    sys.modules.setdefault("matplotlib.pyplot", stub.pyplot)
    sys.path.insert(0, str(REF / "dev" / "unit_tests"))
    import panel as proto  # type: ignore

    cases = []
    # subsonic: rectangular panel centred at the origin, z = 0
    sub = proto.SubsonicPanel(2.0, 1.0)
    for P in ([0.3, 0.2, 0.7], [1.7, -0.4, 0.25], [-2.5, 1.5, -1.1], [0.1, 0.05, 1e-3], [0.9, 0.49, -0.3]):
        P = np.array(P)
        geom = sub.calc_geom(P)
        ints = sub.calc_F_integrals(geom)
        sub.calc_H_integrals(geom, ints)
        cases.append(dict(kind="subsonic", verts=sub.verts.T.tolist(), P=P.tolist(), H111=float(ints.H111),
                          hH113=float(ints.hH113), F111=[float(v) for v in ints.F111]))
    # supersonic subinclined (local-scaled coordinates, freestream along +x)
    verts = np.array([[0.5, -0.5, -0.5, 0.5], [0.5, 0.5, -0.5, -0.5]])
    sup = proto.SupersonicSubinclinedPanel(verts)
    for P in ([5.0, 1.0, 1.0], [3.0, 0.2, 0.4], [2.0, -0.3, -0.6], [1.2, 0.1, 0.2], [4.0, 2.5, 0.5], [0.8, 0.0, 0.1]):
        P = np.array(P)
        geom = sup.calc_geom(P)
        ints = sup.calc_F_integrals(geom)
        sup.calc_H_integrals(geom, ints)
        cases.append(dict(kind="supersonic_subinclined", verts=verts.T.tolist(), P=P.tolist(), H111=float(ints.H111),
                          hH113=float(ints.hH113), F111=[float(v) for v in ints.F111]))
    (OUT / "prototype_integrals.json").write_text(json.dumps(dict(
        source="dev/unit_tests/panel.py (SubsonicPanel, SupersonicSubinclinedPanel): quadrilateral panels",
        note="hH113 of the supersonic prototype carries the opposite sign convention (it subtracts the atan2 "
             "terms, panel.py:703-716, where src/panel.f90:2549,2565 adds them); H111 is convention-free.",
        cases=cases), indent=1))
    print("prototype_integrals.json:", len(cases), "cases")


def make_offbody():
    """The two off-body golden tables and the inputs of the tests that produced them (test_machline.py:636-730)."""
    out = {"source": "test/input_files/*_offbody_points_correct.csv, columns x,y,z,phi_d,phi_s,v_d,v_s (0,1,2,4,5,11-13,14-16 of 24), e20.13",
           "cases": []}
    for name, inp_file, csv_file, vel in [
            ("half_wing_inc", "half_wing_input.json", "half_wing_inc_offbody_points_correct.csv", None),
            ("half_wing_supersonic", "supersonic_half_wing_input.json", "half_wing_supersonic_offbody_points_correct.csv", [100.0, 5.0, 5.0])]:
        inp = json.loads((REF / "test" / "input_files" / inp_file).read_text())
        inp["solver"]["formulation"] = "dirichlet-morino"
        if vel is not None:
            inp["flow"]["freestream_velocity"] = vel
        inp["output"] = {}
        tab = np.genfromtxt(REF / "test" / "input_files" / csv_file, skip_header=1, delimiter=",")
        # columns: x,y,z,phi_inf,phi_d,phi_s,phi,Phi,v_inf(3),v_d(3),v_s(3),v(3),V(3),|V|  (panel_solver.f90:2862-2863)
        out["cases"].append({"name": name, "input": inp, "points": tab[:, :3].tolist(), "phi_d": tab[:, 4].tolist(),
                             "phi_s": tab[:, 5].tolist(), "v_d": tab[:, 11:14].tolist(), "v_s": tab[:, 14:17].tolist()})
    (OUT / "offbody_potentials.json").write_text(json.dumps(out))
    print("offbody_potentials.json:", [(c["name"], len(c["points"])) for c in out["cases"]])


def make_solver_histories():
    """Iteration histories of studies/matrix_solvers (matrix_solver_study.py:246-252: cone V = (-1,0,0), M = 1.5, mirror xy;
    diamond wing V = (1,0,0), M = 2, no mirror; formulation "morino" = today's dirichlet-morino, lower order)."""
    it_dir = REF / "studies" / "matrix_solvers" / "iterations"
    out = {"source": "studies/matrix_solvers/iterations/<mesh>_<solver>_<refinement>_<prec>_<sorted|unsorted>_prec_history.csv",
           "cases": []}
    for root, refinement, mesh, vel, mach, mirror in [
            ("cone_10_deg_", "coarse", "cone_10_deg_coarse.vtk", [-1.0, 0.0, 0.0], 1.5, "xy"),
            ("cone_10_deg_", "medium", "cone_10_deg_medium.vtk", [-1.0, 0.0, 0.0], 1.5, "xy"),
            ("diamond_5_deg_full_", "coarse", "diamond_5_deg_full_coarse.stl", [1.0, 0.0, 0.0], 2.0, None)]:
        for solver in ["GMRES", "BJAC", "BSSOR"]:
            for prec in ["DIAG", "none"]:
                for sort in [True, False]:
                    f = it_dir / f"{root}{solver}_{refinement}_{prec}_{'sorted' if sort else 'unsorted'}_prec_history.csv"
                    if not f.exists():
                        continue
                    rows = f.read_text().split("\n")
                    i0 = next(i for i, r in enumerate(rows) if r.strip().startswith("iteration"))
                    tab = [[float(v) for v in r.split(",")] for r in rows[i0 + 1:] if r.strip()]
                    geom = {"file": f"test/meshes/{mesh}", "spanwise_axis": "+y", "singularity_order": "lower"}
                    if mirror:
                        geom["mirror_about"] = mirror
                    inp = {"flow": {"freestream_velocity": vel, "freestream_mach_number": mach}, "geometry": geom,
                           "solver": {"formulation": "dirichlet-morino", "matrix_solver": solver, "preconditioner": prec,
                                      "sort_system": sort},
                           "post_processing": {}, "output": {}}
                    out["cases"].append({"name": f.stem.replace("_prec_history", ""), "input": inp, "header": rows[:i0 + 1],
                                         "columns": rows[i0].strip().split(","), "rows": tab})
    (OUT / "solver_histories.json").write_text(json.dumps(out))
    print("solver_histories.json:", len(out["cases"]), "histories,", sum(len(c["rows"]) for c in out["cases"]), "rows")


def make_singular_values():
    import csv
    out = {"source": "studies/matrix_solvers/{cone_10_deg,diamond_5_deg_full}_condition_data.csv (S_max, S_min of A_mat.txt)", "cases": []}
    for f, root, ext, vel, mach, mirror in [("cone_10_deg_condition_data.csv", "cone_10_deg_", ".vtk", [-1.0, 0.0, 0.0], 1.5, "xy"),
                                            ("diamond_5_deg_full_condition_data.csv", "diamond_5_deg_full_", ".stl", [1.0, 0.0, 0.0], 2.0, None)]:
        for r in csv.DictReader(open(REF / "studies" / "matrix_solvers" / f)):
            mesh = f"{root}{r['Refinement']}{ext}"
            if not (OUT / "meshes.npz").exists() or f"{mesh}:points" not in np.load(OUT / "meshes.npz").files and \
                    f"{mesh}:facet_vertices" not in np.load(OUT / "meshes.npz").files:
                continue
            geom = {"file": f"test/meshes/{mesh}", "spanwise_axis": "+y"}
            if mirror:
                geom["mirror_about"] = mirror
            inp = {"flow": {"freestream_velocity": vel, "freestream_mach_number": mach}, "geometry": geom,
                   "solver": {"matrix_solver": "GMRES", "preconditioner": r["Preconditioner"], "sort_system": r["Sort System"] == "True"},
                   "post_processing": {}, "output": {}}
            out["cases"].append({"name": f"{root}{r['Refinement']}_{r['Preconditioner']}_{'sorted' if r['Sort System'] == 'True' else 'unsorted'}",
                                 "input": inp, "S_max": float(r["S_max"]), "S_min": float(r["S_min"])})
    (OUT / "aic_singular_values.json").write_text(json.dumps(out, indent=1))
    print("aic_singular_values.json:", len(out["cases"]), "cases")


if __name__ == "__main__":
    make_solver_histories()
    make_offbody()
    make_meshes()
    make_study_meshes()
    make_goldens()
    make_prototype_integrals()
    make_singular_values()     # after make_meshes: only meshes that are in meshes.npz
