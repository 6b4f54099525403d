"""BASELINE configs[4]-class run: a synthetic refined case (50k-100k+ unknowns) assembled and solved on N GPUs, row-sharded.
Not the bench (one pass, no warm-up repetitions): it records that the large case runs end to end and what each phase costs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
        scripts/scale_run.py --dims 320x160 --solvers GMRES,LU

Rank 0 prints one JSON line per solver."""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from machline_b200 import _abi, gpu, host, meshgen, shard  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dims", default="226x112")
ap.add_argument("--solvers", default="GMRES,LU")
ap.add_argument("--mach", type=float, default=None)
ap.add_argument("--cyclic", type=int, default=0, help="deal rows block-cyclically in blocks of this many rows (LU load balance); 0 = contiguous")
ap.add_argument("--case", default="wing", choices=["wing", "sears_haack", "agard_b"],
                help="wing: mirrored half wing with wake, M = 0.5 (configs[1]/[4]); sears_haack: supersonic slender body, M = 2 (configs[2]); "
                     "agard_b: the reference's AGARD-B study mesh (configs[3]: M = 1.6, mirrored about yz, supersonic wake) refined "
                     "--levels times 1:4 (the north star's 100k-panel class)")
ap.add_argument("--levels", type=int, default=2, help="agard_b: 1:4 refinement levels of agard_b_fine.vtk (4366 panels per half)")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
if world > 1:
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nc, ns = (int(v) for v in args.dims.split("x"))
tmp = tempfile.mkdtemp(prefix=f"machline_scale_r{rank}_")
t0 = time.perf_counter()
if args.case == "wing":
    mach = 0.5 if args.mach is None else args.mach
    pts, tris = meshgen.swept_wing_half(nc, ns)
    meshgen.write_vtk(f"{tmp}/w.vtk", pts, tris)
    case = host.Case(meshgen.wing_input("w.vtk", mach=mach), base_dir=tmp)
    label = f"half wing {nc}x{ns}, M={mach}"
elif args.case == "agard_b":
    z = np.load(ROOT / "tests" / "golden" / "study_meshes.npz")
    pts, tris = meshgen.subdivide(z["agard_b_fine.vtk:points"], z["agard_b_fine.vtk:triangles"], args.levels)
    meshgen.write_vtk(f"{tmp}/agard_b_fine.vtk", pts, tris)
    inp = meshgen.study_input("agard_b")
    if args.mach is not None:
        inp["flow"]["freestream_mach_number"] = args.mach
    mach = inp["flow"]["freestream_mach_number"]
    case = host.Case(inp, base_dir=tmp)
    label = f"AGARD-B wing-body (agard_b_fine.vtk refined {args.levels}x 1:4: {len(tris)} panels x 2 images), M={mach}, mirrored about yz, supersonic wake"
else:
    mach = 2.0 if args.mach is None else args.mach
    pts, tris = meshgen.sears_haack(nc, ns)
    meshgen.write_vtk(f"{tmp}/w.vtk", pts, tris)
    case = host.Case(meshgen.sears_haack_input("w.vtk", mach=mach), base_dir=tmp)
    label = f"Sears-Haack body {nc}x{ns}, M={mach}, source-free"
host_s = time.perf_counter() - t0
N = case.n_cp
row0, nrows = shard.row_shard(N, rank, world)
ctx = gpu.Context(local)
if world > 1:
    uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(gpu.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, src=0)
    ctx.set_communicator(bytes(uid.cpu().numpy().tobytes()), rank, world)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


if args.cyclic > 0:
    ctx.set_case(case, cyclic=(args.cyclic, rank, world))
else:
    ctx.set_case(case, row0=row0, nrows=nrows)
barrier()
t0 = time.perf_counter()
ctx.assemble()
barrier()
asm_wall = time.perf_counter() - t0
asm_ms = ctx.assemble_resident()
BC = np.array(case.BC)
first = None   # (solver, x, C_p) of the first solver: the others are compared with it
for solver in args.solvers.split(","):
    opts = case.solver_opts()
    o = _abi.solver_opts(solver, preconditioner="DIAG" if opts.preconditioner else "none", tol=opts.tol,
                         max_iterations=opts.max_iterations)
    barrier()
    t0 = time.perf_counter()
    x, info = ctx.solve(o, BC)
    barrier()
    wall = time.perf_counter() - t0
    first_ms = None
    if solver != "LU":
        # The first solve of a process pays one-time costs that no kernel owns: NCCL connection set-up on the first collective,
        # mapping the peer windows (cudaIpcOpenMemHandle), the first cudaMalloc of the Krylov basis.  The steady state (what
        # bench.py times after its warm-up steps) is the second solve; both are reported.
        first_ms = info.solve_ms
        barrier()
        t0 = time.perf_counter()
        x, info = ctx.solve(o, BC)
        barrier()
        wall = time.perf_counter() - t0
    prof = None
    if os.environ.get("MACHLINE_SCALE_PROFILE"):   # a second, profiled solve: CUDA events around every matvec and exchange / tail
        ctx.set_profiling(True)
        ctx.profile(reset=True)
        ctx.solve(o, BC)
        pp = ctx.profile()
        ctx.set_profiling(False)
        prof = {"gemv_launches": int(pp.gemv_launches), "gemv_ms": pp.gemv_ms, "exchange_or_tail_ms": pp.comm_ms}
    vals = torch.tensor([asm_ms, info.solve_ms, wall], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    if rank == 0:
        res = case.post(x)
        a_ms, s_ms, w_s = (float(v) for v in vals.cpu())
        cmp_ = {}
        if first is None:
            first = (solver, x.copy(), np.array(res.C_p))
        else:
            cmp_ = {f"max_abs_dCp_vs_{first[0]}": float(np.abs(np.array(res.C_p) - first[2]).max()),
                    f"max_abs_dx_over_max_abs_x_vs_{first[0]}": float(np.abs(x - first[1]).max() / np.abs(first[1]).max())}
        if case.flow.supersonic and world == 1:
            cmp_["pair_classes"] = ctx.dod_census()
        if prof:
            cmp_["profile"] = prof
        if first_ms is not None:
            cmp_["first_solve_ms_incl_one_time_setup"] = first_ms
        print(json.dumps({**cmp_, "case": label, "n_gpus": world, "n_panels": case.info.n_body_panels,
                          "n_unknown": N, "A_bytes": 8.0 * N * N, "pairs": float(case.n_pairs), "matrix_solver": solver, "row_dealing": f"block-cyclic {args.cyclic}" if args.cyclic else "contiguous",
                          "host_setup_s": host_s, "assemble_ms": a_ms, "assemble_first_wall_s": asm_wall,
                          "pairs_per_s": case.n_pairs / (a_ms * 1e-3), "solve_ms": s_ms, "solve_wall_s": w_s,
                          "iterations": int(info.iterations), "res_norm": info.res_norm, "res_max": info.res_max,
                          "lu_tflops": (2.0 / 3.0 * N ** 3 / (s_ms * 1e-3) / 1e12) if solver == "LU" else None,
                          "C_p_max": res.C_p_max, "C_p_min": res.C_p_min, "Cx": float(res.C_F[0]), "Cz": float(res.C_F[2])}),
              flush=True)
ctx.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
