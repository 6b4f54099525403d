"""GPU: the dense solvers behind ml_solve / ml_solve_dense against the oracle restatement of
common/linalg.f90 on the same systems (LU, GMRES, RGMRES, BJAC, BSSOR, QRUP, FQRUP, PURC), including the reference's quirks:
the uniform 1/A(N,N) "DIAG" scale, the N/5 default block size and the invalid-name -> GMRES fallback."""
import numpy as np
import pytest

import fixtures
import oracle_binding as ob
from machline_b200 import _abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from machline_b200 import gpu
    c = gpu.Context(0)
    yield c
    c.close()


def _system(n, seed=0, dominance=4.0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    A += dominance * np.sqrt(n) * np.eye(n)
    # rows of very different scale, as Morino rows vs strength-matching rows (+-1) in the reference
    A[::7] *= 1e-3
    b = rng.standard_normal(n)
    return np.asfortranarray(A), b


@pytest.mark.parametrize("n", [1, 5, 63, 64, 65, 200, 777, 2048])
def test_lu_matches_numpy_and_oracle(ctx, n):
    A, b = _system(n, seed=n)
    x, info = ctx.solve_dense(A, b, _abi.solver_opts("LU"))
    x_np = np.linalg.solve(A, b)
    assert np.abs(x - x_np).max() <= 1e-10 * np.abs(x_np).max()
    x_or, _ = ob.solve_system(A, np.zeros(n), b, _abi.solver_opts("LU"))
    assert np.abs(x - x_or).max() <= 1e-10 * np.abs(x_or).max()
    assert info.iterations == -1 and info.res_norm <= 1e-10 * np.linalg.norm(b) + 1e-13


def test_lu_needs_pivoting(ctx):
    """Zero leading diagonal entries: only a pivoting factorisation gets through."""
    n = 130
    rng = np.random.default_rng(3)
    A = rng.standard_normal((n, n))
    A = A[::-1].copy() + 8 * np.fliplr(np.eye(n))   # large anti-diagonal
    for i in range(0, n, 3):
        A[i, i] = 0.0
    b = rng.standard_normal(n)
    x, info = ctx.solve_dense(A, b, _abi.solver_opts("LU"))
    assert np.abs(A @ x - b).max() < 1e-10


@pytest.mark.parametrize("n", [1500, 4500, 5000, 7100])
def test_lu_heavy_pivoting_across_ctas(ctx, n):
    """No diagonal dominance: (almost) every column interchanges two rows held by different CTAs of the panel kernels
    (n = 4500 / 5000: more than 32 CTAs for the grid-wide kernel, the two-level candidate reduction; 5000 is not a multiple of
    64; n = 7100: the first panels of the cluster kernel keep their last rows in global memory)."""
    rng = np.random.default_rng(n)
    A = np.asfortranarray(rng.standard_normal((n, n)))
    A[::5] *= 1e3          # implicit scaling matters: vv differs by row
    b = rng.standard_normal(n)
    x, info = ctx.solve_dense(A, b, _abi.solver_opts("LU"))
    x_np = np.linalg.solve(A, b)
    assert np.abs(x - x_np).max() <= 1e-8 * np.abs(x_np).max()
    r = A @ x - b
    assert np.abs(r).max() <= 1e-9 * (np.abs(A).sum(axis=1) * np.abs(x).max()).max()


def test_lu_ties_take_the_last_row(ctx):
    """Exact ties of vv*|a| (rows that are sign / permutation copies of each other in a column): the reference keeps the LAST
    maximal row (`>=`, linalg.f90:242).  With a different tie rule the factorisation is still valid, so the check is
    against the oracle's solution to rounding on a matrix where wrong ties change the pivot order drastically."""
    n = 192
    rng = np.random.default_rng(5)
    A = rng.integers(-1, 2, size=(n, n)).astype(float) + 3 * np.eye(n)
    A = np.asfortranarray(A)
    b = rng.standard_normal(n)
    x, _ = ctx.solve_dense(A, b, _abi.solver_opts("LU"))
    x_or, _ = ob.solve_system(A, np.zeros(n), b, _abi.solver_opts("LU"))
    assert np.abs(x - x_or).max() <= 1e-11 * np.abs(x_or).max()


def test_lu_singular_reports_status_3(ctx):
    from machline_b200 import gpu
    A, b = _system(50, seed=9)
    A[17, :] = 0.0
    with pytest.raises(gpu.GpuError) as e:
        ctx.solve_dense(A, b, _abi.solver_opts("LU"))
    assert e.value.status == 3  # ML_SINGULAR: lu_decomp code 1 (linalg.f90:205-208)


@pytest.mark.parametrize("solver", ["GMRES", "RGMRES", "BJAC", "LU"])
@pytest.mark.parametrize("prec", ["DIAG", "none"])
def test_solvers_on_assembled_system_match_oracle(ctx, solver, prec):
    case, _, _ = fixtures.make_case("test_13")   # supersonic half wing, sorted (upper-pentagonal) system
    A_ref, I_ref = ob.assemble(case)
    b = case.BC - I_ref
    opts = _abi.solver_opts(solver, preconditioner=prec, rel=0.9)
    x, info = ctx.solve_dense(A_ref, b, opts)
    x_or, info_or = ob.solve_system(A_ref, np.zeros_like(b), b, opts)
    assert np.abs(x - x_or).max() <= 2e-9 * np.abs(x_or).max()
    if solver in ("GMRES", "RGMRES", "BJAC"):
        assert abs(info.iterations - info_or.iterations) <= 1, (info.iterations, info_or.iterations)
    case.close()


def test_gmres_mgs_mode_matches_oracle_iteration_history(ctx, monkeypatch):
    """MACHLINE_GMRES_MGS=1 orthogonalises in the reference's modified Gram-Schmidt order."""
    monkeypatch.setenv("MACHLINE_GMRES_MGS", "1")
    case, _, _ = fixtures.make_case("test_05")
    A_ref, I_ref = ob.assemble(case)
    b = case.BC - I_ref
    opts = _abi.solver_opts("GMRES")
    x, info = ctx.solve_dense(A_ref, b, opts)
    x_or, info_or = ob.solve_system(A_ref, np.zeros_like(b), b, opts)
    assert info.iterations == info_or.iterations
    assert np.abs(x - x_or).max() <= 1e-10 * np.abs(x_or).max()
    case.close()


def test_gmres_many_iterations_many_blocks(ctx):
    """Hundreds of Arnoldi steps on a system wide enough for every kernel to run many CTAs (the config-2 bench
    case needs ~500 iterations at N ~ 10k): iteration count and solution must follow the oracle."""
    n = 3000
    rng = np.random.default_rng(11)
    A = rng.standard_normal((n, n)) / np.sqrt(n) * 0.9 + np.eye(n)   # spectrum fills a disc of radius ~0.9 around 1
    b = rng.standard_normal(n)
    opts = _abi.solver_opts("GMRES", max_iterations=1000)
    x, info = ctx.solve_dense(A, b, opts)
    x_or, info_or = ob.solve_system(np.asfortranarray(A), np.zeros(n), b, opts)
    assert info_or.iterations > 150
    assert abs(info.iterations - info_or.iterations) <= 2, (info.iterations, info_or.iterations)
    assert np.abs(x - x_or).max() <= 1e-9 * np.abs(x_or).max()
    assert info.res_norm < 1e-9


def test_invalid_solver_name_falls_back_to_gmres(ctx):
    A, b = _system(120, seed=5)
    o = _abi.solver_opts("NOT_A_SOLVER")
    assert o.matrix_solver == _abi.SOLVERS["GMRES"]   # panel_solver.f90:1969-1973
    x, info = ctx.solve_dense(A, b, o)
    assert info.iterations > 0 and np.abs(A @ x - b).max() < 1e-9


def _pentagonal(n, band, seed):
    """Upper-pentagonal system like the reference's sorted AIC (panel_solver.f90:778-1030): zero below the band."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n)) + 3.0 * np.sqrt(n) * np.eye(n)
    i, j = np.indices((n, n))
    A[i - j > band] = 0.0
    A[(i - j > 0) & ((i + 2 * j) % 5 == 0)] = 0.0          # exact zeros inside the band are skipped (linalg.f90:908, 1145)
    A[(i - j == band) & (i % 3 == 0)] = 3e-13               # below the 1e-12 bandwidth threshold, yet still rotated
    return np.asfortranarray(A), rng.standard_normal(n)


@pytest.mark.parametrize("solver", ["QRUP", "FQRUP"])
@pytest.mark.parametrize("n,band", [(1, 0), (2, 1), (3, 2), (65, 7), (300, 299), (700, 90), (1500, 400)])
@pytest.mark.parametrize("prec", ["none", "DIAG"])
def test_givens_solvers_are_bit_identical_to_oracle(ctx, solver, n, band, prec):
    """Every rotation is generated and applied in the reference's order with unfused IEEE operations, so the
    solution equals the oracle's bit for bit (including the never-rotated last column, linalg.f90:914)."""
    if n == 1500 and prec == "DIAG":
        pytest.skip("covered by the unscaled case")
    A, b = _pentagonal(n, band, seed=n + band)
    opts = _abi.solver_opts(solver, preconditioner=prec)
    x, info = ctx.solve_dense(A, b, opts)
    x_or, info_or = ob.solve_system(A, np.zeros(n), b, opts)
    assert np.array_equal(x, x_or), np.abs(x - x_or).max()
    assert info.iterations == -1


def test_givens_zero_diagonal_reports_status_3(ctx):
    from machline_b200 import gpu
    n = 20
    A = np.triu(np.ones((n, n)))
    A[7, 7] = 0.0
    with pytest.raises(gpu.GpuError) as e:
        ctx.solve_dense(np.asfortranarray(A), np.ones(n), _abi.solver_opts("QRUP"))
    assert e.value.status == 3   # "Zero found on the diagonal of R" (linalg.f90:956-961)


@pytest.mark.parametrize("n", [1, 2, 33, 257, 600])
@pytest.mark.parametrize("prec", ["none", "DIAG"])
def test_purcell_is_bit_identical_to_oracle(ctx, n, prec):
    A, b = _system(n, seed=100 + n)
    opts = _abi.solver_opts("PURC", preconditioner=prec)
    x, info = ctx.solve_dense(A, b, opts)
    x_or, _ = ob.solve_system(A, np.zeros(n), b, opts)
    assert np.array_equal(x, x_or), np.abs(x - x_or).max()
    assert np.abs(A @ x - b).max() < 1e-8


@pytest.mark.parametrize("n,bsz", [(200, 40), (777, 0), (1000, 333)])
@pytest.mark.parametrize("prec", ["none", "DIAG"])
def test_block_ssor_matches_oracle(ctx, n, bsz, prec):
    A, b = _system(n, seed=7 * n)
    opts = _abi.solver_opts("BSSOR", preconditioner=prec, block_size=bsz, rel=0.9, tol=1e-11)
    x, info = ctx.solve_dense(A, b, opts)
    x_or, info_or = ob.solve_system(A, np.zeros(n), b, opts)
    assert info_or.iterations > 2
    assert abs(info.iterations - info_or.iterations) <= 1, (info.iterations, info_or.iterations)
    assert np.abs(x - x_or).max() <= 1e-9 * np.abs(x_or).max()


@pytest.mark.parametrize("solver", ["QRUP", "FQRUP", "BSSOR", "PURC"])
def test_sequential_solvers_on_assembled_system(ctx, solver):
    """The sorted supersonic half-wing system of test 13 through ml_solve with each of the reference's other solvers."""
    case, _, _ = fixtures.make_case("test_13")
    ctx.set_case(case)
    ctx.assemble()
    opts = case.solver_opts()
    opts.matrix_solver = _abi.SOLVERS[solver]
    opts.rel = 0.9
    x, info = ctx.solve(opts, case.BC)
    A_ref, I_ref = ob.assemble(case)
    x_or, info_or = ob.solve_system(A_ref, I_ref, case.BC, opts)
    assert np.abs(x - x_or).max() <= 2e-9 * np.abs(x_or).max()
    assert info.res_norm < 1e-9
    case.close()


@pytest.mark.parametrize("solver", ["GMRES", "RGMRES"])
def test_iteration_history_file_matches_reference_format(ctx, solver, tmp_path):
    """solver.iterative_solver_output: the per-iteration residual estimate, written as the reference writes it
    (linalg.f90:1273-1280, 1316 / 1376-1383, 1438): list-directed header lines, then '(i6, a, ES10.3)' rows; the values are the
    oracle's err history."""
    import ctypes as C
    A, b = _system(300, seed=11)
    path = str(tmp_path / "iterations.csv").encode()
    o = _abi.solver_opts(solver)
    o.iteration_file = path
    x, info = ctx.solve_dense(A, b, o)
    lines = open(path.decode()).read().split("\n")
    assert lines[0] == " method" and lines[1] == " GMRES" and lines[2] == " N=%12d" % 300
    rows = [ln for ln in lines[4:] if ln]
    assert len(rows) == info.iterations
    if solver == "GMRES":
        assert lines[3] == " iteration,||err||"
        hist = np.zeros(1000)
        n_it = C.c_int()
        x_ref = np.zeros(300)
        inv = 1.0 / A[-1, -1]        # the DIAG "preconditioner" scales the system by 1/A(N,N)
        As, bs = np.asfortranarray(A * inv), b * inv
        ob.lib().orc_gmres(300, As.ctypes.data_as(_abi.c_double_p), bs.ctypes.data_as(_abi.c_double_p), 1e-12, 1000, C.byref(n_it),
                           x_ref.ctypes.data_as(_abi.c_double_p), hist.ctypes.data_as(_abi.c_double_p))
        assert n_it.value == info.iterations
        for k, row in enumerate(rows):
            assert len(row) == 17 and row[6] == ","
            assert int(row[:6]) == k + 1
            # four printed digits; the device orthogonalises with CGS2 instead of MGS, so close to convergence the estimate
            # differs from the oracle's in more than rounding (the iteration COUNT is the same, asserted above)
            assert abs(float(row[7:]) - hist[k]) <= 0.05 * hist[k] + 1e-12
            assert row[7:] == "%10.3E" % float(row[7:])
    else:
        assert lines[3] == " iteration,outer iteration,inner iteration,||err||"
        tot, outer, inner, err = rows[-1].split(",")
        assert int(tot) == info.iterations and int(inner) <= 20 and int(outer) >= 1 and float(err) < 1e-12


@pytest.mark.parametrize("solver", ["BJAC", "BSSOR"])
def test_block_solver_history_file(ctx, solver, tmp_path):
    """iteration,||dx||,||err||,relaxation rows of block_jacobi_solve / block_ssor_solve (linalg.f90:659-666, 717; 514-520, 587)."""
    A, b = _system(400, seed=5, dominance=8.0)
    path = str(tmp_path / "block_iterations.csv")
    o = _abi.solver_opts(solver, block_size=100)
    o.iteration_file = path.encode()
    x, info = ctx.solve_dense(A, b, o)
    lines = open(path).read().split("\n")
    assert lines[0] == " method" and lines[1] == " " + solver
    k = 2
    if solver == "BJAC":
        assert lines[2] == " N=%12d" % 400
        k = 3
    assert lines[k] == " iteration,||dx||,||err||,relaxation"
    rows = [ln for ln in lines[k + 1:] if ln]
    assert len(rows) == info.iterations and info.iterations > 1
    it, dx, err, rel = rows[-1].split(",")
    assert int(it) == info.iterations
    assert float(err) < 1e-10      # the rounding floor of this system (3e-12) is above tol: it runs to max_iterations
    assert abs(float(rel) - 0.8) < 1e-12
    assert float(dx) < 1e-5
    assert all(len(r) == 39 for r in rows)
    d = [float(r.split(",")[1]) for r in rows]
    assert d[0] > d[-1]
    assert np.abs(A @ x - b).max() < 1e-9
