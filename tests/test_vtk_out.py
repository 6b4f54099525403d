"""CPU: the result files (host library writers, machline_b200/vtk_out.py) have the reference's layout (src/vtk.f90) and read back."""
import numpy as np

import fixtures
import oracle_binding as ob
from machline_b200 import host, vtk_out


def test_fortran_e_descriptor():
    # gfortran's e20.12: mantissa in [0.1, 1), two-digit exponent, right-aligned in 20
    assert vtk_out.fortran_e(1.0) == "  0.100000000000E+01"
    assert vtk_out.fortran_e(-0.5) == " -0.500000000000E+00"
    assert vtk_out.fortran_e(0.0) == "  0.000000000000E+00"
    assert vtk_out.fortran_e(123456.789012345678) == "  0.123456789012E+06"
    assert vtk_out.fortran_e(9.9999999999996e-8) == "  0.100000000000E-06"      # rounding carries into the exponent
    assert vtk_out.fortran_e(-4.2301340874257765) == " -0.423013408743E+01"
    # the same descriptor wrote the reference's off-body tables (e20.13): spot-check against one of its numbers
    assert vtk_out.fortran_e(-4.579465052707, width=20, digits=13) == "-0.4579465052707E+01"


def test_result_files_layout_and_read_back(tmp_path):
    """Body / mirrored body / wake / control-point files of the host library (csrc/host/outputs.cpp) in the reference's layout."""
    case, _, _ = fixtures.make_case("test_01")            # mirrored half wing, asymmetric flow, wake
    A, I_known = ob.assemble(case)
    x, _ = ob.solve_system(A, I_known, case.BC, case.solver_opts())
    res = case.post(x)
    path = tmp_path / "results" / "body.vtk"
    case.write_body(path)
    lines = path.read_text().split("\n")
    nb, nv = case.info.n_body_panels, case.info.n_body_verts
    assert lines[0] == "# vtk DataFile Version 3.0" and lines[2] == "ASCII" and lines[3] == "DATASET POLYDATA"
    assert lines[4] == "POINTS%20d float" % nv
    assert lines[5 + nv] == "POLYGONS%20d%20d" % (nb, 4 * nb)
    assert lines[6 + nv].startswith("3 ") and len(lines[6 + nv]) == 1 + 3 * 20
    labels = [ln for ln in lines if ln.startswith(("SCALARS", "VECTORS", "NORMALS", "CELL_DATA", "POINT_DATA"))]
    # surface_mesh_write_body, src/surface_mesh.f90:2585-2636
    assert labels == ["CELL_DATA%20d" % nb, "NORMALS normals float", "SCALARS inclination float 1", "SCALARS distribution_order float 1",
                      "SCALARS N_discontinuous_edges float 1", "VECTORS centroid float", "SCALARS C_p_inc float 1", "SCALARS sigma float 1",
                      "VECTORS v float", "VECTORS v_inner float", "VECTORS dC_f float", "POINT_DATA%20d" % nv, "SCALARS mu float 1",
                      "SCALARS Phi_u float 1", "SCALARS convex float 1"]
    i0 = lines.index("SCALARS C_p_inc float 1") + 2
    cp = np.array([float(v) for v in lines[i0:i0 + nb]])
    assert np.abs(cp - np.asarray(res.C_p)[:nb]).max() <= 1e-11 * np.abs(cp).max()
    i0 = lines.index("SCALARS mu float 1") + 2
    mu = np.array([float(v) for v in lines[i0:i0 + nv]])
    assert np.abs(mu - np.asarray(res.mu)[:nv]).max() <= 1e-11 * np.abs(mu).max()
    # the geometry section is a mesh file the host loader accepts: same panel count
    inp = dict(case.input)
    inp["geometry"] = dict(inp["geometry"], file=str(path))
    case2 = host.Case(inp, base_dir="")
    assert case2.info.n_body_panels == nb
    case2.close()
    # mirrored twin: second half of the cell / vertex arrays, panels wound backwards
    mpath = tmp_path / "results" / "mirror.vtk"
    case.write_body(mpath, mirrored=True)
    ml = mpath.read_text().split("\n")
    assert ml[4] == lines[4] and [int(v) for v in ml[6 + nv].split()[1:]] == [int(v) for v in lines[6 + nv].split()[1:]][::-1]
    i0 = ml.index("SCALARS C_p_inc float 1") + 2
    cpm = np.array([float(v) for v in ml[i0:i0 + nb]])
    assert np.abs(cpm - np.asarray(res.C_p)[nb:2 * nb]).max() <= 1e-11 * np.abs(cpm).max()
    y = np.array([[float(ln[20 * k:20 * k + 20]) for k in range(3)] for ln in lines[5:5 + nv]])
    ym = np.array([[float(ln[20 * k:20 * k + 20]) for k in range(3)] for ln in ml[5:5 + nv]])
    assert np.array_equal(ym, y * np.array([1., -1., 1.]))          # mirror_about = xz
    # wake strips with mu = mu(top parent) - mu(bottom parent)
    wpath = tmp_path / "results" / "wake.vtk"
    assert case.write_wake(wpath)
    wl = wpath.read_text().split("\n")
    n_wv = int(wl[4].split()[1])
    assert wl[5 + n_wv] == "POLYGONS%20d%20d" % (case.info.n_wake_panels, 4 * case.info.n_wake_panels)
    assert wl[6 + n_wv + case.info.n_wake_panels] == "POINT_DATA%20d" % n_wv and wl[7 + n_wv + case.info.n_wake_panels] == "SCALARS mu float 1"
    # control points: VERTICES cells, BC_type (int) and the residual
    cpath = tmp_path / "results" / "cp.vtk"
    r = A @ x - (np.asarray(case.BC) - I_known)
    case.write_control_points(cpath, r)
    cl = cpath.read_text().split("\n")
    n_cp = case.n_cp
    assert cl[4] == "POINTS%20d float" % n_cp and cl[5 + n_cp] == "VERTICES%20d%20d" % (n_cp, 2 * n_cp) and cl[6 + n_cp] == "1%20d" % 0
    assert cl[6 + 2 * n_cp] == "POINT_DATA%20d" % n_cp and cl[7 + 2 * n_cp] == "SCALARS BC_type int 1"
    assert {int(v) for v in cl[9 + 2 * n_cp:9 + 3 * n_cp]} == {2, 4}   # source-free potential rows and strength-matching rows
    assert cl[9 + 3 * n_cp] == "SCALARS residual float 1"
    rr = np.array([float(v) for v in cl[11 + 3 * n_cp:11 + 4 * n_cp]])
    assert np.abs(rr - r).max() <= 1e-11 * max(np.abs(r).max(), 1e-300)
    # a case without wake exports none
    sphere, _, _ = fixtures.make_case("test_08")
    sphere.post(np.zeros(sphere.n_unknown))
    assert not sphere.write_wake(tmp_path / "none.vtk") and not (tmp_path / "none.vtk").exists()
    sphere.close()
    case.close()


def test_write_system_round_trips_like_the_reference_study_reads_it(tmp_path):
    """A_mat.txt as studies/matrix_solvers/matrix_conditions.py:26 reads it (np.genfromtxt): 12 significant digits per entry."""
    rng = np.random.default_rng(0)
    A = rng.standard_normal((7, 7)) * 10.0 ** rng.integers(-8, 3, size=(7, 7))
    b = rng.standard_normal(7)
    vtk_out.write_system(A, b, tmp_path / "A_mat.txt", tmp_path / "b_vec.txt")
    lines = (tmp_path / "A_mat.txt").read_text().split("\n")
    assert len(lines[0]) == 7 * 20
    A2 = np.array([[float(ln[20 * j:20 * j + 20]) for j in range(7)] for ln in lines[:7]])
    assert np.abs(A2 - A).max() <= 5e-12 * np.abs(A).max() and (np.abs(A2 - A) <= 5e-12 * np.abs(A)).all()
    assert np.array_equal(np.genfromtxt(tmp_path / "b_vec.txt"), b)
