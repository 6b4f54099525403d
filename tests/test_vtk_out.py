"""CPU: the body results file (machline_b200/vtk_out.py) has the reference's layout (src/vtk.f90) and reads back."""
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


def test_body_file_layout_and_read_back(tmp_path):
    case, _, _ = fixtures.make_case("test_08")            # sphere
    A, I_known = ob.assemble(case)
    x, _ = ob.solve_system(A, I_known, case.BC, case.solver_opts())
    res = case.post(x)
    path = tmp_path / "results" / "sphere.vtk"
    vtk_out.write_body_vtk(path, case, res)
    lines = path.read_text().split("\n")
    nb, nv = case.info.n_body_panels, case.info.n_body_verts
    assert lines[0] == "# vtk DataFile Version 3.0" and lines[2] == "ASCII" and lines[3] == "DATASET POLYDATA"
    assert lines[4] == "POINTS%20d float" % nv
    assert lines[5 + nv] == "POLYGONS%20d%20d" % (nb, 4 * nb)
    assert lines[6 + nv].startswith("3 ") and len(lines[6 + nv]) == 1 + 3 * 20
    labels = [ln for ln in lines if ln.startswith(("SCALARS", "VECTORS", "NORMALS", "CELL_DATA", "POINT_DATA"))]
    assert labels == ["CELL_DATA%20d" % nb, "NORMALS normals float", "SCALARS inclination float 1", "SCALARS distribution_order float 1",
                      "VECTORS centroid float", "SCALARS C_p_inc float 1", "SCALARS sigma float 1", "VECTORS v float",
                      "POINT_DATA%20d" % nv, "SCALARS mu float 1"]
    # numbers read back to 12 significant digits
    i0 = lines.index("SCALARS C_p_inc float 1") + 2
    cp = np.array([float(v) for v in lines[i0:i0 + nb]])
    assert np.abs(cp - np.asarray(res.C_p)[:nb]).max() <= 1e-11 * np.abs(cp).max()
    i0 = lines.index("SCALARS mu float 1") + 2
    mu = np.array([float(v) for v in lines[i0:i0 + nv]])
    assert np.abs(mu - np.asarray(res.mu)[:nv]).max() <= 1e-11 * np.abs(mu).max()
    # the geometry section is a mesh file the host loader accepts: same panel count and the same solution again
    inp = dict(case.input)
    inp["geometry"] = dict(inp["geometry"], file=str(path))
    case2 = host.Case(inp, base_dir="")
    assert case2.info.n_body_panels == nb and case2.info.n_body_verts == nv
    case2.close()
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
