"""CPU: the oracle against the reference's stored ITERATION HISTORIES.

studies/matrix_solvers/iterations/*_prec_history.csv were written by the reference's own GMRES, block_jacobi_solve and
block_ssor_solve (solver.iterative_solver_output; common/linalg.f90:1273-1316, 659-717, 514-587) on the meshes of
studies/matrix_solvers/meshes for preconditioner DIAG / none and a sorted / unsorted system.  Reproducing a history row by
row pins, in one chain and independently of the surface golden tuples: the supersonic AIC (cone M = 1.5 mirrored, diamond
wing M = 2), the permutation, the "DIAG preconditioner" quirk (uniform 1/A(N,N) scale), GMRES (Arnoldi + Givens + the
residual estimate), block Jacobi and block SSOR including decompose_blocks / lu_decomp / lu_back_sub and the N/5 default
block size -- the solvers SURVEY lists as "parity unpinned" by the reference's test suite.

Bar: identical iteration counts and every printed value (4 significant digits, ES10.3) within print rounding (plus 2e-12 of
the history's first value, which only matters in its last few rows)."""
import ctypes as C
import json
from functools import lru_cache
from pathlib import Path

import numpy as np
import pytest

import fixtures
import oracle_binding as ob
from machline_b200 import _abi, host

DOC = json.loads((Path(__file__).resolve().parent / "golden" / "solver_histories.json").read_text())
PRINT_ROUNDING = 6e-4     # ES10.3: half a unit of the 4th significant digit is at most 5e-4 relative
NOISE_FLOOR = 2e-12       # times the history's first value: the last iterations difference numbers of size 0.1-1: summation order (libgfortran's matmul) shows
                          # (worst case: 2.5e-13 on ||dx|| = 1.0e-11 in the last two rows of the sorted diamond-wing BJAC runs)


@lru_cache(maxsize=None)
def _system(mesh: str, mirror, vel: tuple, mach: float, sort: bool):
    """(case, A, b): A does not depend on the solver or the preconditioner."""
    geom = {"file": mesh, "spanwise_axis": "+y", "singularity_order": "lower"}
    if mirror:
        geom["mirror_about"] = mirror
    inp = {"flow": {"freestream_velocity": list(vel), "freestream_mach_number": mach}, "geometry": geom,
           "solver": {"formulation": "dirichlet-morino", "sort_system": sort}, "post_processing": {}, "output": {}}
    case = host.Case(inp, base_dir=fixtures.mesh_root())
    A, I_known = ob.assemble(case)
    return case, A, I_known


def _oracle_history(c):
    inp = c["input"]
    g, s = inp["geometry"], inp["solver"]
    case, A, I_known = _system(g["file"], g.get("mirror_about"), tuple(inp["flow"]["freestream_velocity"]),
                               inp["flow"]["freestream_mach_number"], bool(s["sort_system"]))
    N = case.n_unknown
    BC = np.array(case.BC)
    opts = _abi.solver_opts(s["matrix_solver"], preconditioner=s["preconditioner"])
    L = ob.lib()
    if s["matrix_solver"] == "GMRES":
        b = BC - I_known
        inv = 1.0 / A[-1, -1] if s["preconditioner"] == "DIAG" else 1.0      # linalg.f90:1813-1816
        As, bs = np.asfortranarray(A * inv), b * inv
        hist, n_it, x = np.zeros(2000), C.c_int(), np.zeros(N)
        L.orc_gmres(N, As.ctypes.data_as(_abi.c_double_p), bs.ctypes.data_as(_abi.c_double_p), opts.tol, opts.max_iterations,
                    C.byref(n_it), x.ctypes.data_as(_abi.c_double_p), hist.ctypes.data_as(_abi.c_double_p))
        return N, n_it.value, {"||err||": hist[:n_it.value]}
    L.orc_set_block_history.argtypes = [_abi.c_double_p, _abi.c_double_p, C.c_int]
    L.orc_set_block_history.restype = None
    dx, err = np.zeros(2000), np.zeros(2000)
    L.orc_set_block_history(dx.ctypes.data_as(_abi.c_double_p), err.ctypes.data_as(_abi.c_double_p), 2000)
    try:
        x, info = ob.solve_system(A, I_known, BC, opts)
    finally:
        L.orc_set_block_history(None, None, 0)
    return N, info.iterations, {"||dx||": dx[:info.iterations], "||err||": err[:info.iterations]}


@pytest.mark.parametrize("c", DOC["cases"], ids=[c["name"] for c in DOC["cases"]])
def test_oracle_reproduces_reference_iteration_history(c):
    ref = np.array(c["rows"])
    N, iters, cols = _oracle_history(c)
    n_line = next((h for h in c["header"] if h.strip().startswith("N=")), None)
    if n_line is not None:
        assert int(n_line.split("=")[1]) == N
    assert iters == len(ref), f"{iters} iterations, the reference's file has {len(ref)}"
    assert (ref[:, 0] == np.arange(1, len(ref) + 1)).all()
    for name, vals in cols.items():
        col = c["columns"].index(name)
        r = ref[:, col]
        ok = np.isfinite(r) & (r > 0)
        assert ok.sum() >= len(r) - 1
        floor = NOISE_FLOOR * r[0]
        assert (r[ok] > 1e3 * floor).sum() >= min(5, len(r) // 2)     # most rows are checked at print rounding alone
        assert (np.abs(vals[ok] - r[ok]) <= PRINT_ROUNDING * r[ok] + floor).all(), \
            f"{name}: worst {np.max(np.abs(vals[ok] - r[ok]) / r[ok]):.2e} at iteration {1 + int(np.argmax(np.abs(vals[ok] - r[ok]) / r[ok]))}"
    if "relaxation" in c["columns"]:
        assert (ref[:, c["columns"].index("relaxation")] == 0.8).all()
