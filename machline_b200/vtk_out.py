"""Text outputs written from Python: the linear system (solver.write_A_and_b) and the off-body CSV (output.offbody_points), with
the reference's Fortran edit descriptors.  The VTK result files (body, mirrored body, wake, control points) are written by the
host library (csrc/host/outputs.cpp; host.Case.write_body / write_wake / write_control_points)."""
from __future__ import annotations

from pathlib import Path

import numpy as np


def fortran_e(v: float, width: int = 20, digits: int = 12) -> str:
    """Fortran `Ew.d` edit descriptor: 0.dddddddddddd E+ee with the mantissa in [0.1, 1)."""
    v = float(v)
    if v != v:
        return "NaN".rjust(width)
    if v in (float("inf"), float("-inf")):
        return ("-Infinity" if v < 0 else "Infinity").rjust(width)
    if v == 0.0:
        s = ("-" if np.signbit(v) else "") + "0." + "0" * digits + "E+00"
        return s.rjust(width)
    m, e = f"{abs(v):.{digits - 1}E}".split("E")      # d.ddddddddddd E+xx, correctly rounded to `digits` figures
    exp = int(e) + 1
    s = ("-" if v < 0 else "") + "0." + m.replace(".", "") + ("E%+03d" % exp if abs(exp) < 100 else "%+04d" % exp)
    return s.rjust(width)


def export_off_body_points(case, ctx, x, points_file, output_file) -> int:
    """output.offbody_points (panel_solver_export_off_body_points, src/panel_solver.f90:2771-2895): potentials and velocities
    induced at the points of `points_file` (a header line, then x,y,z per line) by the solved strengths x, written as the
    reference's 24-column CSV (e20.13).  The influences are evaluated on the GPU (ctx: gpu.Context; ml_assemble with the field
    points as rows).  As in the reference's sum, a wake panel's top and bottom halves cancel (src/panel.f90:3002-3003, 3257):
    the wake contributes nothing to phi_d / v_d here.  Returns the number of points."""
    pts = np.atleast_2d(np.genfromtxt(points_file, delimiter=",", skip_header=1))[:, :3]
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    fs = case.flow
    v_inf = np.array(case.input["flow"]["freestream_velocity"], dtype=np.float64)
    U = float(np.sqrt((v_inf * v_inf).sum()))
    c_hat = np.array(fs.c_hat_g[:])
    phi_d, phi_s = ctx.potentials_at(case, pts, x, with_wake=False)
    v_d, v_s = ctx.velocity_parts_at(case, pts, x, with_wake=False)
    phi_d, phi_s, v_d, v_s = phi_d * U, phi_s * U, v_d * U, v_s * U
    e = lambda v: fortran_e(v, 20, 13)
    Path(output_file).parent.mkdir(parents=True, exist_ok=True)
    with open(output_file, "w") as f:
        f.write(" x,y,z,phi_inf,phi_d,phi_s,phi,Phi,v_inf_x,v_inf_y,v_inf_z,v_d_x,v_d_y,v_d_z,v_s_x,v_s_y,v_s_z,v_x,v_y,v_z,V_x,V_y,V_z,V\n")
        for i in range(len(pts)):
            phi_inf = U * float(pts[i] @ c_hat)
            v = v_d[i] + v_s[i]
            V = v_inf + v
            cols = [pts[i, 0], pts[i, 1], pts[i, 2], phi_inf, phi_d[i], phi_s[i], phi_d[i] + phi_s[i], phi_inf + phi_d[i] + phi_s[i],
                    v_inf[0], v_inf[1], v_inf[2], v_d[i, 0], v_d[i, 1], v_d[i, 2], v_s[i, 0], v_s[i, 1], v_s[i, 2], v[0], v[1], v[2],
                    V[0], V[1], V[2], float(np.sqrt((V * V).sum()))]
            f.write(",".join(e(c) for c in cols) + "\n")
    return len(pts)


def write_system(A, b, a_path="A_mat.txt", b_path="b_vec.txt") -> None:
    """solver.write_A_and_b (panel_solver_write_system, src/panel_solver.f90:1768-1799): A_mat.txt holds one row of A per
    line, entries `e20.12` back to back; b_vec.txt one entry per line (list-directed there; 17 significant digits here).
    This is the channel the reference's own studies read the matrix through (studies/matrix_solvers/matrix_conditions.py)."""
    A = np.asarray(A, dtype=np.float64)
    with open(a_path, "w") as f:
        for row in A:
            f.write("".join(fortran_e(v) for v in row) + "\n")
    with open(b_path, "w") as f:
        for v in np.asarray(b, dtype=np.float64):
            f.write("  %.16E\n" % v)
