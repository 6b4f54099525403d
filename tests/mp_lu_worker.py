"""Worker of tests/test_gpu_sharded_lu.py: LU through ml_solve on a row-sharded system, one rank per GPU (torchrun) or one
rank with MACHLINE_LU_SHARDED=1 (the NCCL algorithm with its collectives degenerated to copies).

Two systems per run: the assembled AIC of a small wing (checked against the oracle's lu_solve on the oracle's matrix), and a
random matrix without diagonal dominance written over the resident rows with ml_set_A (every column interchanges rows,
most of them across ranks; checked against numpy)."""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle_binding as ob  # noqa: E402
from machline_b200 import _abi, gpu, host, meshgen, shard  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
dims = sys.argv[1] if len(sys.argv) > 1 else "24x12"
nc, ns = (int(v) for v in dims.split("x"))
if world > 1:
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tmp = tempfile.mkdtemp(prefix=f"machline_lu_r{rank}_")
pts, tris = meshgen.swept_wing_half(nc, ns)
meshgen.write_vtk(f"{tmp}/w.vtk", pts, tris)
case = host.Case(meshgen.wing_input("w.vtk", mach=0.5, matrix_solver="LU"), base_dir=tmp)
N = case.n_cp
row0, nrows = shard.row_shard(N, rank, world)
ctx = gpu.Context(local)
if world > 1:
    uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(gpu.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, src=0)
    ctx.set_communicator(bytes(uid.cpu().numpy().tobytes()), rank, world)
cyclic_block = int(os.environ.get("MACHLINE_TEST_CYCLIC", "0"))
if cyclic_block > 0:
    ctx.set_case(case, cyclic=(cyclic_block, rank, world))
    my_rows = shard.cyclic_rows(N, rank, world, cyclic_block)
else:
    ctx.set_case(case, row0=row0, nrows=nrows)
    my_rows = np.arange(row0, row0 + nrows)
I_loc = ctx.assemble()
opts = case.solver_opts()
BC = np.array(case.BC)

# ---- 1. the assembled system --------------------------------------------------------------------------
x, info = ctx.solve(opts, BC)
A_ref, I_ref = ob.assemble(case)
x_ref, _ = ob.solve_system(A_ref, I_ref, BC, opts)
err = np.abs(x - x_ref).max() / np.abs(x_ref).max()
assert info.iterations == -1
assert err < 1e-9, f"rank {rank}: AIC system, sharded LU vs oracle LU: {err:.2e}"
assert info.res_norm < 1e-10, info.res_norm

# ---- 2. a random matrix over the same shards: heavy pivoting ------------------------------------------------
rng = np.random.default_rng(1234)
A = np.asfortranarray(rng.standard_normal((N, N)))
A[::5] *= 1e3
ctx.set_A(A[my_rows])
I_full = np.zeros(N)
if world > 1:
    parts = [None] * world
    dist.all_gather_object(parts, (my_rows, I_loc))
    for rows_r, I_r in parts:
        I_full[rows_r] = I_r
else:
    I_full[my_rows] = I_loc
b = BC - I_full
x2, info2 = ctx.solve(opts, BC)
x_np = np.linalg.solve(A, b)
err2 = np.abs(x2 - x_np).max() / np.abs(x_np).max()
assert err2 < 1e-7, f"rank {rank}: random system, sharded LU vs numpy: {err2:.2e}"
print(f"rank {rank}/{world}: N={N} {len(my_rows)} rows ({'cyclic ' + str(cyclic_block) if cyclic_block else 'contiguous'}) AIC err {err:.2e} random err {err2:.2e} res {info2.res_norm:.2e} "
      f"solve_ms {info.solve_ms:.1f}/{info2.solve_ms:.1f} OK", flush=True)
ctx.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
