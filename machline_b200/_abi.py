"""ctypes mirrors of include/machline_gpu.h and include/machline_host.h (keep in sync)."""
from __future__ import annotations

import ctypes as C

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)

ML_OK = 0
ML_STATUS_NAMES = {
    0: "ML_OK", 1: "ML_NAN_IN_SYSTEM", 2: "ML_UNINFLUENCED", 3: "ML_SINGULAR", 4: "ML_NAN_RESIDUAL",
    10: "ML_BAD_ARGUMENT", 11: "ML_NOT_READY", 12: "ML_UNSUPPORTED", 100: "ML_CUDA_ERROR",
    101: "ML_NCCL_ERROR",
}

SOLVERS = {"LU": 0, "QRUP": 1, "FQRUP": 2, "GMRES": 3, "RGMRES": 4, "PURC": 5, "BSSOR": 6, "BJAC": 7}
PRECONDITIONERS = {"none": 0, "NONE": 0, "DIAG": 1}

BC_ZERO_POTENTIAL, BC_SF_POTENTIAL, BC_ZERO_NORMAL_MF, BC_STRENGTH_MATCHING = 1, 2, 3, 4


class MlFlow(C.Structure):
    _fields_ = [("M_inf", C.c_double), ("B", C.c_double), ("s", C.c_double), ("K_inv", C.c_double),
                ("c_hat_g", C.c_double * 3), ("B_mat_g", C.c_double * 9), ("C_mat_g", C.c_double * 9),
                ("supersonic", C.c_int), ("mirror_plane", C.c_int)]


class MlPanelSoa(C.Structure):
    _fields_ = [("n_panels", C.c_int), ("n_images", C.c_int), ("n_cols", C.c_int), ("in_wake", C.c_int),
                ("centr", c_double_p), ("A_g_to_ls", c_double_p), ("vertices_ls", c_double_p),
                ("n_hat_ls", c_double_p), ("b", c_double_p), ("sqrt_b", c_double_p), ("J", c_double_p),
                ("r", c_int_p), ("area", c_double_p), ("vert_g", c_double_p), ("T_mu", c_double_p),
                ("i_vert_d", c_int_p), ("i_panel_s", c_int_p), ("has_sources", c_ubyte_p),
                ("image_present", c_ubyte_p),
                ("order2", C.c_int), ("order", c_ubyte_p), ("M_dim", c_int_p), ("T_mu6", c_double_p),
                ("S_dim", c_int_p), ("i_panel_s4", c_int_p), ("T_sigma", c_double_p)]


class MlSystemMap(C.Structure):
    _fields_ = [("n_cp", C.c_int), ("n_unknown", C.c_int), ("n_verts", C.c_int), ("n_body_panels", C.c_int),
                ("n_sigma", C.c_int), ("mirrored", C.c_int), ("asym_flow", C.c_int),
                ("P", c_int_p), ("sigma_known", c_ubyte_p), ("i_sigma_in_sys", c_int_p), ("sigma", c_double_p)]


class MlSolverOpts(C.Structure):
    _fields_ = [("matrix_solver", C.c_int), ("preconditioner", C.c_int), ("tol", C.c_double),
                ("rel", C.c_double), ("max_iterations", C.c_int), ("restart_iterations", C.c_int),
                ("block_size", C.c_int), ("iteration_file", C.c_char_p)]


class MlSolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("res_max", C.c_double), ("res_norm", C.c_double),
                ("assemble_ms", C.c_double), ("solve_ms", C.c_double)]


class MlhCpTable(C.Structure):
    _fields_ = [("n_cp", C.c_int), ("loc", c_double_p), ("bc", c_int_p), ("n_g", c_double_p),
                ("row_perm", c_int_p), ("BC", c_double_p)]


class MlhSolverSettings(C.Structure):
    _fields_ = [("opts", MlSolverOpts), ("matrix_solver_name", C.c_char * 16), ("formulation", C.c_char * 48),
                ("sort_system", C.c_int), ("write_A_and_b", C.c_int), ("run_checks", C.c_int)]


class MlhResults(C.Structure):
    _fields_ = [("C_p_max", C.c_double), ("C_p_min", C.c_double), ("C_F", C.c_double * 3),
                ("C_M", C.c_double * 3), ("n_cells", C.c_int), ("n_mu", C.c_int),
                ("mu", c_double_p), ("C_p", c_double_p), ("V_cells", c_double_p)]


RULES = ["incompressible", "isentropic", "second-order", "slender-body", "linear", "prandtl-glauert", "karman-tsien", "laitone"]


class MlPostTables(C.Structure):
    _fields_ = [("n_cells", C.c_int), ("mu_index", c_int_p), ("T_mu", c_double_p), ("A_g_to_ls", c_double_p),
                ("s_dir", c_double_p), ("sigma_index", c_int_p), ("sigma_known", c_double_p), ("v_inner", c_double_p),
                ("n_g", c_double_p), ("area", c_double_p), ("centr", c_double_p), ("force_cell", c_int_p)]


class MlPostFlow(C.Structure):
    _fields_ = [("U", C.c_double), ("U_inv", C.c_double), ("M_inf", C.c_double), ("gamma", C.c_double),
                ("a_ise", C.c_double), ("b_ise", C.c_double), ("c_ise", C.c_double), ("C_P_vac", C.c_double),
                ("C_P_stag", C.c_double), ("M_inf_corr", C.c_double), ("v_inf", C.c_double * 3),
                ("A_g_to_c", C.c_double * 9), ("CG", C.c_double * 3), ("S_ref", C.c_double), ("l_ref", C.c_double),
                ("rules", C.c_int), ("force_rule", C.c_int), ("mirrored_symmetric", C.c_int), ("mirror_plane", C.c_int)]


class MlPostOut(C.Structure):
    _fields_ = [("V_cells", c_double_p), ("C_p", c_double_p * 8), ("dC_f", c_double_p), ("C_F", C.c_double * 3),
                ("C_M", C.c_double * 3), ("C_p_max", C.c_double), ("C_p_min", C.c_double)]


class MlhMeshInfo(C.Structure):
    _fields_ = [("n_body_panels", C.c_int), ("n_body_verts", C.c_int), ("n_wake_panels", C.c_int),
                ("n_wake_strips", C.c_int), ("n_edges", C.c_int), ("n_cp", C.c_int), ("n_unknown", C.c_int),
                ("mirrored", C.c_int), ("asym_flow", C.c_int), ("mirror_plane", C.c_int),
                ("supersonic", C.c_int), ("sort_seconds", C.c_double)]


def solver_opts(matrix_solver="GMRES", preconditioner="DIAG", tol=1e-12, rel=0.8, max_iterations=1000,
                restart_iterations=20, block_size=-1) -> MlSolverOpts:
    """Defaults of panel_solver.f90:173-209; unknown solver names fall back to GMRES (:1969-1973)."""
    o = MlSolverOpts()
    o.matrix_solver = SOLVERS.get(matrix_solver, SOLVERS["GMRES"])
    o.preconditioner = 1 if preconditioner == "DIAG" else 0
    o.tol, o.rel = tol, rel
    o.max_iterations, o.restart_iterations, o.block_size = max_iterations, restart_iterations, block_size
    o.iteration_file = None
    return o


class MlProfile(C.Structure):
    _fields_ = [("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong), ("gemv_launches", C.c_longlong),
                ("gemv_bytes", C.c_longlong), ("gemv_ms", C.c_double), ("assemble_ms", C.c_double), ("comm_ms", C.c_double)]
