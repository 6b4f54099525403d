"""GPU: LU of a row-sharded system (lu_solve_sharded) through ml_solve.

* one rank: MACHLINE_LU_SHARDED=1 runs the whole distributed algorithm (gathered panel, replicated panel factorisation,
  position/slot permutation, U-row exchange, local DMMA update, distributed back substitution) with its collectives
  degenerated to copies -- this runs on the single-GPU test box;
* two ranks under torchrun (NCCL): skipped unless two GPUs are visible."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
WORKER = str(ROOT / "tests" / "mp_lu_worker.py")


def _run(cmd, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=str(ROOT))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return res.stdout


@pytest.mark.parametrize("dims", ["12x6", "40x20"])
def test_sharded_lu_algorithm_on_one_rank(dims):
    out = _run([sys.executable, WORKER, dims], {"MACHLINE_LU_SHARDED": "1"})
    assert "OK" in out


@pytest.mark.parametrize("cyclic", ["0", "128", "64"])
@pytest.mark.parametrize("dims", ["12x6", "40x20"])
def test_sharded_lu_two_ranks_nccl(dims, cyclic):
    """Contiguous row blocks and block-cyclic dealing (MACHLINE_TEST_CYCLIC = block size) through the same kernels."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                "--master-port", "29533", WORKER, dims], {"MACHLINE_TEST_CYCLIC": cyclic})
    assert out.count("OK") == 2


def test_cyclic_shard_on_one_rank():
    """world = 1 block-cyclic = all rows; exercises the row-list plumbing on the single-GPU test box."""
    out = _run([sys.executable, WORKER, "12x6"], {"MACHLINE_LU_SHARDED": "1", "MACHLINE_TEST_CYCLIC": "64"})
    assert "OK" in out


@pytest.mark.parametrize("env", [{"MACHLINE_LU_PANEL_RPC": "512"}, {"MACHLINE_LU_PER_COLUMN": "1"}, {"MACHLINE_LU_GEMM_V1": "1"}])
def test_lu_kernel_variants(env):
    """The other code paths of the factorisation on heavy-pivoting systems: panel rows overflowing shared memory (rows 320..511
    of every CTA stay in global memory), the per-column fallback, the first-generation trailing update."""
    out = _run([sys.executable, str(ROOT / "tests" / "lu_env_worker.py"), "700", "3000"], env)
    assert out.count("OK") == 2
