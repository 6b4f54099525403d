"""GPU: GMRES / restarted GMRES of a row-sharded system through ml_solve -- the path bench.py --gpus N and
the driver's scaling run measure.

* one rank, MACHLINE_GMRES_SHARDED=1: the multi-rank kernels run with the rank as its own peer (p2p_push_kernel,
  p2p_wait_copy_kernel, compact_shards_kernel, arnoldi_tail_sharded_kernel with its peer-memory reductions): this is what
  the single-GPU test box executes;
* two and four ranks under torchrun: CUDA-IPC windows over NVLink, and MACHLINE_NO_P2P=1 (ncclAllGather); contiguous and
  block-cyclic rows; skipped unless that many GPUs are visible.
x and the iteration count are checked against the oracle inside the worker (tests/mp_gmres_worker.py)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
WORKER = str(ROOT / "tests" / "mp_gmres_worker.py")

MODES = {"sharded_basis": {}, "replicated_basis_p2p": {"MACHLINE_GMRES_REPLICATED": "1"}, "no_p2p": {"MACHLINE_NO_P2P": "1"}}


def _run(cmd, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900, cwd=str(ROOT))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return res.stdout


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("what", ["16x8", "40x20"])
def test_sharded_gmres_kernels_on_one_rank(what, mode):
    out = _run([sys.executable, WORKER, what], {"MACHLINE_GMRES_SHARDED": "1", **MODES[mode]})
    assert "OK" in out


def test_sharded_gmres_one_rank_cyclic_rows_and_supersonic_wake():
    out = _run([sys.executable, WORKER, "agard_b_coarse"], {"MACHLINE_GMRES_SHARDED": "1", "MACHLINE_TEST_CYCLIC": "64"})
    assert "OK" in out


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("cyclic", ["0", "128"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_gmres_nccl_ranks(world, cyclic, mode):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                "--master-port", "29541", WORKER, "40x20"], {"MACHLINE_TEST_CYCLIC": cyclic, **MODES[mode]})
    assert out.count("OK") == world
