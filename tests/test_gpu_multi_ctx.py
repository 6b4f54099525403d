"""GPU: the single-process multi-GPU context (ml_ctx_create_multi, include/machline_gpu.h): one host thread drives all
devices through the entry points of a single-device context.  On the 1-GPU box the one-device group exercises the
facade (fan-out, row gathering); with >= 2 devices the same calls run the row-sharded assembly and the sharded GMRES /
RGMRES / LU (peer access inside the process instead of CUDA IPC) and are compared with the single-device context and
the oracle."""
import ctypes as C

import numpy as np
import pytest

import fixtures
import oracle_binding as ob

pytestmark = pytest.mark.gpu


def _n_devices() -> int:
    import torch
    return torch.cuda.device_count()


def _run(ctx, case, opts, cyclic=None):
    ctx.set_case(case, cyclic=cyclic)
    I_known = ctx.assemble()
    A = ctx.get_A()
    x, info = ctx.solve(opts, case.BC)
    return A, I_known, x, info


@pytest.mark.parametrize("n_dev", [1, 2, 4])
@pytest.mark.parametrize("name,solver,cyclic", [("test_08", "GMRES", None), ("test_13", "GMRES", None), ("test_13", "RGMRES", None),
                                                ("test_08", "LU", (128, 0, 0)), ("test_13", "LU", None), ("test_16", "GMRES", None),
                                                ("test_08", "BJAC", None), ("test_13", "BJAC", (64, 0, 0))])
def test_multi_context_matches_single_device(n_dev, name, solver, cyclic):
    if _n_devices() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    from machline_b200 import _abi, gpu
    case, _, _ = fixtures.make_case(name)
    opts = case.solver_opts()
    opts.matrix_solver = _abi.SOLVERS[solver]
    single = gpu.Context(0)
    A1, I1, x1, info1 = _run(single, case, opts)
    single.close()
    multi = gpu.Context(devices=list(range(n_dev)))
    assert gpu.lib().ml_device_count(multi._h) == n_dev
    A, I_known, x, info = _run(multi, case, opts, cyclic=cyclic)
    # the assembly is row-parallel with a fixed summation order per entry: the shards reproduce the single-device matrix exactly
    assert A.shape == A1.shape and (A == A1).all()
    assert (I_known == I1).all()
    # a row window that crosses shard boundaries
    r0, n = case.n_cp // 3, case.n_cp // 2
    assert (multi.get_A(r0, n) == A1[r0:r0 + n]).all()
    scale = np.abs(x1).max()
    if solver == "LU":
        assert np.abs(x - x1).max() <= 1e-10 * scale
    elif solver == "BJAC":   # block Jacobi on the shards repeats the single-device arithmetic row by row
        assert info.iterations == info1.iterations
        assert np.abs(x - x1).max() <= 1e-12 * scale
    else:
        assert abs(info.iterations - info1.iterations) <= 2
        assert np.abs(x - x1).max() <= 1e-8 * scale
    assert info.res_norm < 1e-9
    # against the oracle
    A_ref, I_ref = ob.assemble(case)
    x_ref, _ = ob.solve_system(A_ref, I_ref, case.BC, opts)
    assert np.abs(x - x_ref).max() <= 1e-8 * np.abs(x_ref).max()
    assert multi.check_system(case.BC) == (0, 0, 0)
    B = A.copy()
    B[case.n_cp - 2, :] = 0.
    B[:, 3] = 0.
    multi.set_A(B)
    assert multi.check_system(case.BC) == (2, 1, 1)
    assert multi.pair_count == case.n_pairs
    multi.close()
    case.close()


def test_multi_context_rejects_per_device_calls():
    from machline_b200 import gpu
    multi = gpu.Context(devices=[0])
    L = gpu.lib()
    assert L.ml_set_row_shard(multi._h, 0, 10) == 10          # ML_BAD_ARGUMENT
    assert L.ml_set_communicator(multi._h, C.create_string_buffer(128), 0, 1) == 10
    bad = C.c_void_p()
    ids = (C.c_int * 2)(0, 0)
    assert L.ml_ctx_create_multi(C.byref(bad), ids, 2) == 10  # the same device twice
    multi.close()


@pytest.mark.parametrize("name", ["test_08", "test_13", "test_19"])
def test_block_jacobi_sharded_kernels_on_one_rank(name):
    """MACHLINE_BJAC_SHARDED=1: block_jacobi_sharded (diagonal blocks packed from the row shard, right-hand sides and residuals
    through the exchange) on one rank against the single-device block Jacobi and the oracle: what the 1-GPU box executes."""
    import os
    from machline_b200 import _abi, gpu
    case, _, _ = fixtures.make_case(name)
    opts = case.solver_opts()
    opts.matrix_solver = _abi.SOLVERS["BJAC"]
    ctx = gpu.Context(0)
    A1, I1, x1, info1 = _run(ctx, case, opts)
    os.environ["MACHLINE_BJAC_SHARDED"] = "1"
    try:
        x, info = ctx.solve(opts, case.BC)
    finally:
        del os.environ["MACHLINE_BJAC_SHARDED"]
    assert info.iterations == info1.iterations and info.iterations > 3
    assert np.abs(x - x1).max() <= 1e-13 * np.abs(x1).max()
    x_ref, info_ref = ob.solve_system(*ob.assemble(case), case.BC, opts)
    assert abs(info.iterations - info_ref.iterations) <= 1
    assert np.abs(x - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
    ctx.close()
    case.close()
