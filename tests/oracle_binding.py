"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference's hot paths (oracle/oracle_aic.cpp,
oracle/oracle_linalg.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from machline_b200 import _abi  # noqa: E402  (struct layouts only)

ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "liboracle.so"


class OrcPairOut(C.Structure):
    _fields_ = [("in_dod", C.c_int), ("edges_in_dod", C.c_int * 3), ("F111", C.c_double * 3),
                ("hH113", C.c_double), ("H111", C.c_double), ("H213", C.c_double), ("H123", C.c_double),
                ("h", C.c_double), ("phi_s", C.c_double), ("phi_d", C.c_double * 3),
                ("phi_d_abs", C.c_double * 3), ("v_s", C.c_double * 3), ("v_d", C.c_double * 9),
                ("phi_s_S", C.c_double * 4), ("phi_d_M", C.c_double * 6), ("phi_d_M_abs", C.c_double * 6),
                ("F121", C.c_double * 3), ("F211", C.c_double * 3), ("H211", C.c_double), ("H121", C.c_double),
                ("H313", C.c_double), ("H223", C.c_double), ("H133", C.c_double),
                ("v_s_S", C.c_double * 12), ("v_d_M", C.c_double * 18), ("F113", C.c_double * 3), ("F123", C.c_double * 3),
                ("F133", C.c_double * 3), ("h3H115", C.c_double), ("H125", C.c_double), ("hH135", C.c_double), ("H145", C.c_double),
                ("H215", C.c_double), ("H225", C.c_double), ("H235", C.c_double), ("hH315", C.c_double), ("H325", C.c_double),
                ("H415", C.c_double), ("H113_3rsh2H115", C.c_double)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        srcs = list(ORACLE_DIR.glob("*.cpp")) + list(ORACLE_DIR.glob("*.h")) + [ROOT / "include" / "machline_gpu.h"]
        if not ORACLE_LIB.exists() or any(s.stat().st_mtime > ORACLE_LIB.stat().st_mtime for s in srcs):
            subprocess.run(["make", "-C", str(ORACLE_DIR), "-s"], check=True)
        L = C.CDLL(str(ORACLE_LIB))
        dp, ip = _abi.c_double_p, _abi.c_int_p
        L.orc_pair_influence.argtypes = [C.POINTER(_abi.MlFlow), C.POINTER(_abi.MlPanelSoa), C.c_int, C.c_int, dp,
                                         C.POINTER(OrcPairOut)]
        L.orc_pair_influence.restype = None
        L.orc_assemble.argtypes = [C.POINTER(_abi.MlFlow), C.POINTER(_abi.MlPanelSoa), C.POINTER(_abi.MlPanelSoa),
                                   C.POINTER(_abi.MlSystemMap), C.c_int, dp, ip, ip, C.c_int, C.c_int, dp, C.c_int,
                                   dp, C.c_int, dp]
        L.orc_assemble_n.argtypes = [C.POINTER(_abi.MlFlow), C.POINTER(_abi.MlPanelSoa), C.POINTER(_abi.MlPanelSoa),
                                     C.POINTER(_abi.MlSystemMap), C.c_int, dp, ip, dp, ip, C.c_int, C.c_int, dp, C.c_int,
                                     dp, C.c_int, dp]
        L.orc_solve_system_ls.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.POINTER(_abi.MlSolverOpts), dp, C.POINTER(_abi.MlSolveInfo)]
        L.orc_lu_solve.argtypes = [C.c_int, dp, dp, dp]
        L.orc_gmres.argtypes = [C.c_int, dp, dp, C.c_double, C.c_int, ip, dp, dp]
        L.orc_restarted_gmres.argtypes = [C.c_int, dp, dp, C.c_double, C.c_int, C.c_int, ip, dp]
        L.orc_block_jacobi.argtypes = [C.c_int, dp, dp, C.c_int, C.c_double, C.c_double, C.c_int, ip, dp]
        L.orc_block_ssor.argtypes = [C.c_int, dp, dp, C.c_int, C.c_double, C.c_double, C.c_int, ip, dp]
        L.orc_qr_givens_up.argtypes = [C.c_int, dp, dp, dp]
        L.orc_qr_fast_givens_up.argtypes = [C.c_int, dp, dp, dp]
        L.orc_purcell.argtypes = [C.c_int, dp, dp, dp]
        L.orc_lower_bandwidth.argtypes = [C.c_int, dp]
        L.orc_solve_system.argtypes = [C.c_int, dp, dp, dp, C.POINTER(_abi.MlSolverOpts), dp,
                                       C.POINTER(_abi.MlSolveInfo)]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_set_threads.restype = None
        _lib = L
    return _lib


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


def _dp(a):
    return a.ctypes.data_as(_abi.c_double_p)


def assemble(case, row0: int = 0, nrows: int | None = None, n_threads: int = 0, with_scale: bool = False):
    """Oracle AIC for a machline_b200.host.Case.  Returns (A [n_cp x n_unknown, Fortran order], I_known) and,
    with_scale, also S with S_ij = sum over panels of |contribution to A_ij| (>= |A_ij|)."""
    n_cp, n_u = case.n_cp, case.n_unknown
    if nrows is None:
        nrows = n_cp - row0
    A = np.zeros((nrows, n_u), dtype=np.float64, order="F")   # rows row0..row0+nrows of the permuted system
    I_known = np.zeros(nrows, dtype=np.float64)
    wake = C.byref(case.wake) if case.wake.n_panels > 0 else None
    S = np.zeros((nrows, n_u), dtype=np.float64, order="F") if with_scale else None
    st = lib().orc_assemble_n(C.byref(case.flow), C.byref(case.body), wake, C.byref(case.map), n_cp, case.cps.loc,
                              case.cps.bc, case.cps.n_g, case.cps.row_perm, row0, nrows, _dp(A), max(1, nrows), _dp(I_known), n_threads,
                              _dp(S) if with_scale else None)
    if st != 0:
        raise RuntimeError(f"orc_assemble status {st}")
    if with_scale:
        return A, I_known, S
    return A, I_known


def assemble_at_points(case, points, with_wake: bool = True):
    """Oracle influence matrix of every unknown on arbitrary field points (rows = points, zero-potential rows, unsorted):
    phi_d(points) = A x, phi_s(points) = I_known.  with_wake=False leaves the wake panels out."""
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    n = pts.shape[0]
    A = np.zeros((n, case.n_unknown), dtype=np.float64, order="F")
    I_known = np.zeros(n, dtype=np.float64)
    bc = np.full(n, 1, dtype=np.int32)
    rows = np.arange(n, dtype=np.int32)
    wake = C.byref(case.wake) if (with_wake and case.wake.n_panels > 0) else None
    st = lib().orc_assemble(C.byref(case.flow), C.byref(case.body), wake, C.byref(case.map), n, _dp(pts),
                            bc.ctypes.data_as(_abi.c_int_p), rows.ctypes.data_as(_abi.c_int_p), 0, n, _dp(A), n, _dp(I_known), 0, None)
    if st != 0:
        raise RuntimeError(f"orc_assemble status {st}")
    return A, I_known


def velocities_at(case, points, x):
    """Oracle: induced velocity v_d + v_s (per unit freestream speed) at field points from the solved strengths x."""
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    n = pts.shape[0]
    bc = np.full(n, 5, dtype=np.int32)
    rows = np.arange(n, dtype=np.int32)
    wake = C.byref(case.wake) if case.wake.n_panels > 0 else None
    v = np.zeros((n, 3))
    for k in range(3):
        n_g = np.zeros((n, 3))
        n_g[:, k] = 1.0
        A = np.zeros((n, case.n_unknown), dtype=np.float64, order="F")
        I_known = np.zeros(n, dtype=np.float64)
        st = lib().orc_assemble_n(C.byref(case.flow), C.byref(case.body), wake, C.byref(case.map), n, _dp(pts),
                                  bc.ctypes.data_as(_abi.c_int_p), _dp(n_g), rows.ctypes.data_as(_abi.c_int_p), 0, n, _dp(A), n,
                                  _dp(I_known), 0, None)
        if st != 0:
            raise RuntimeError(f"orc_assemble_n status {st}")
        v[:, k] = A @ np.asarray(x, dtype=np.float64) + I_known
    return v


def velocity_influences_at(case, points, with_wake: bool = True):
    """Oracle velocity influence rows at field points: (V [3][n x n_unknown], v_s [n x 3]) with v_d = V[k] @ x."""
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    n = pts.shape[0]
    bc = np.full(n, 5, dtype=np.int32)
    rows = np.arange(n, dtype=np.int32)
    wake = C.byref(case.wake) if (with_wake and case.wake.n_panels > 0) else None
    V, v_s = [], np.zeros((n, 3))
    for k in range(3):
        n_g = np.zeros((n, 3))
        n_g[:, k] = 1.0
        A = np.zeros((n, case.n_unknown), dtype=np.float64, order="F")
        I_known = np.zeros(n, dtype=np.float64)
        st = lib().orc_assemble_n(C.byref(case.flow), C.byref(case.body), wake, C.byref(case.map), n, _dp(pts),
                                  bc.ctypes.data_as(_abi.c_int_p), _dp(n_g), rows.ctypes.data_as(_abi.c_int_p), 0, n, _dp(A), n,
                                  _dp(I_known), 0, None)
        if st != 0:
            raise RuntimeError(f"orc_assemble_n status {st}")
        V.append(A)
        v_s[:, k] = I_known
    return V, v_s


def pair(case, table, j: int, img: int, P) -> OrcPairOut:
    out = OrcPairOut()
    Pa = np.ascontiguousarray(P, dtype=np.float64)
    lib().orc_pair_influence(C.byref(case.flow), C.byref(table), j, img, _dp(Pa), C.byref(out))
    return out


def solve_system(A, I_known, BC, opts):
    """panel_solver_solve_system on the CPU.  Returns (x, MlSolveInfo)."""
    n_cp, N = A.shape
    A = np.asfortranarray(A, dtype=np.float64)
    x = np.zeros(N)
    info = _abi.MlSolveInfo()
    I_known = np.ascontiguousarray(I_known, dtype=np.float64)
    BC = np.ascontiguousarray(BC, dtype=np.float64)
    if n_cp != N:    # overdetermined least squares (the Neumann formulations), panel_solver.f90:1842-1895
        st = lib().orc_solve_system_ls(n_cp, N, _dp(A), _dp(I_known), _dp(BC), C.byref(opts), _dp(x), C.byref(info))
    else:
        st = lib().orc_solve_system(N, _dp(A), _dp(I_known), _dp(BC), C.byref(opts), _dp(x), C.byref(info))
    if st != 0:
        raise RuntimeError(f"orc_solve_system status {st}")
    return x, info


def run_case(case):
    """Full CPU pipeline: host setup (already done in `case`) -> oracle assemble -> oracle solve -> host post."""
    A, I_known = assemble(case)
    x, info = solve_system(A, I_known, case.BC, case.solver_opts())
    v_inner = None if case.dirichlet else velocities_at(case, case.inner_points(), x)
    return case.post(x, v_inner), info, A
