#!/usr/bin/env python
"""bench.py -- headline benchmark of the two hot paths (AIC assembly + dense solve).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (oracle port)

A "step" is one pass of the hot path over one case: ml_assemble (DoD + body + wake influences -> A
resident in HBM) followed by ml_solve (the input's matrix_solver, GMRES by default as in the
reference).  Workload at N=1: BASELINE.json configs[1] on the reference's OWN mesh
studies/subsonic_onera_m6_wing/meshes/M6_onera_fine.stl (committed in tests/golden/study_meshes.npz):
ONERA M6, mirrored about xz, M = 0.5, alpha = 3.06 deg, automatic wake, 14 512 panels x 2 images.
At N > 1 GPUs (weak scaling) the case must grow with N, which a fixed mesh cannot: the synthetic
ONERA-M6-planform family of machline_b200.meshgen is used, refined so that pairs/GPU stays ~constant,
and the line carries the same family's single-GPU value measured in the same run
("scaling_reference") plus a parity block (sharded solve against a single-GPU solve of the same case).

One JSON line is printed by rank 0 (see DESIGN.md "Measurement" for every key).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ALG_FLOPS_PER_PAIR_SUBSONIC = 187.0   # lower-order Dirichlet, subsonic evaluated pair
METRIC = "aic_pair_influences_per_s"
UNIT = "pair-influences/s"


def wing_dims(n_gpus: int):
    """(n_chord, n_span) of the config-2 family; pairs scale ~ N_panels^2, so panels ~ sqrt(n_gpus)."""
    f = n_gpus ** 0.25
    return int(round(96 * f)), int(round(52 * f))


STUDY_NPZ = ROOT / "tests" / "golden" / "study_meshes.npz"
STUDY_MESH = {"onera_m6": "M6_onera_fine.stl", "cone": "cone_10_deg_fine.vtk", "sears_haack": "SH_160_60.tri",
              "agard_b": "agard_b_fine.vtk"}


def default_workload(n_gpus: int) -> str:
    return "onera_m6" if n_gpus == 1 else "synthetic_wing"


def build_case(n_gpus: int, tmpdir: str, matrix_solver: str, dims: str | None = None, workload: str | None = None):
    """(host.Case, description dict).  workload: a study case of meshgen.study_input on the reference's own mesh, or
    "synthetic_wing" (sized for n_gpus; `dims` = "NCxNS" overrides the size, tests only)."""
    from machline_b200 import host, meshgen
    workload = workload or (default_workload(n_gpus) if not dims else "synthetic_wing")
    if workload == "synthetic_wing":
        nc, ns = wing_dims(n_gpus)
        if dims:
            nc, ns = (int(v) for v in dims.lower().split("x"))
        pts, tris = meshgen.swept_wing_half(nc, ns)
        name = f"wing_{nc}x{ns}.vtk"
        meshgen.write_vtk(Path(tmpdir) / name, pts, tris)
        inp = meshgen.wing_input(name, mach=0.5, alpha_deg=3.06, matrix_solver=matrix_solver)
        desc = dict(workload=workload, mesh=f"synthetic {nc}x{ns} (machline_b200.meshgen.swept_wing_half)", data="synthetic")
    else:
        meshgen.materialise_npz(STUDY_NPZ, tmpdir, only=[STUDY_MESH[workload]])
        inp = meshgen.study_input(workload, matrix_solver=matrix_solver)
        desc = dict(workload=workload, mesh=f"{STUDY_MESH[workload]} (the reference's study mesh, tests/golden/study_meshes.npz)",
                    data="reference mesh, synthetic freestream per SURVEY 8(d)")
    t0 = time.perf_counter()
    case = host.Case(inp, base_dir=tmpdir)
    desc["host_setup_s"] = time.perf_counter() - t0
    return case, desc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.dev = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.dev), "-lms", "100"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            w = [x.strip() for x in line.split(",")]
            if len(w) < 9:
                continue
            try:
                sm.append(float(w[1])); mx.append(float(w[2])); power.append(float(w[3]))
            except ValueError:
                continue
            for nm, val in zip(names, w[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus: int):
    if n_gpus <= 1 and "RANK" not in os.environ:
        return None, 0, 1, 0
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return (dist if world > 1 else None), rank, world, local


def _oracle():
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_binding as ob
    return ob


class ReferenceRunner:
    """The reference's own CPU algorithm for the step (oracle port: AIC assembly with OpenMP over control points as
    src/panel_solver.f90:1307, then panel_solver_solve_system with the input's matrix_solver), on all host threads.

    A step is the WHOLE hot path when that fits the time budget; otherwise a bounded sample of it: the same fraction f
    of both phases -- a window of f*N rows of the assembly and the first f*iterations Arnoldi steps of the solve (the
    iteration count comes from one untimed run to convergence) -- and the pair count is scaled by f."""

    def __init__(self, case, budget_s: float, n_steps: int):
        self.ob = _oracle()
        self.case = case
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.ob.set_threads(self.cores)   # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core
        self.per_row = case.n_pairs // case.n_cp
        self.ob.assemble(case, row0=0, nrows=min(case.n_cp, self.cores))   # thread pool / page-in warm-up, untimed
        t0 = time.perf_counter()
        self.A, self.I = self.ob.assemble(case)
        self.t_asm = time.perf_counter() - t0
        self.opts = case.solver_opts()
        t0 = time.perf_counter()
        self.x, info = self.ob.solve_system(self.A, self.I, case.BC, self.opts)
        self.t_sol = time.perf_counter() - t0
        self.iters = int(info.iterations)
        self.res_norm = float(info.res_norm)
        full = self.t_asm + self.t_sol
        self.f = 1.0 if full * n_steps <= budget_s else max(0.02, budget_s / (full * n_steps))
        if self.iters <= 0:   # direct solver: only the assembly can be sampled
            self.f = 1.0 if full * n_steps <= budget_s else self.f

    def step(self, k: int):
        """Returns (pairs processed, seconds)."""
        case, ob = self.case, self.ob
        if self.f >= 1.0:
            t0 = time.perf_counter()
            A, I = ob.assemble(case)
            ob.solve_system(A, I, case.BC, self.opts)
            return case.n_pairs, time.perf_counter() - t0
        rows = max(self.cores, int(self.f * case.n_cp))
        row0 = (k * 7919) % max(1, case.n_cp - rows)
        o = case.solver_opts()
        o.max_iterations = max(2, int(round(self.f * self.iters)) + 1)   # GMRES runs max_iterations - 1 Arnoldi steps
        t0 = time.perf_counter()
        ob.assemble(case, row0=row0, nrows=rows)
        if self.iters > 0:
            ob.solve_system(self.A, self.I, case.BC, o)
        return rows * self.per_row, time.perf_counter() - t0

    def sample_text(self):
        if self.f >= 1.0:
            what = "the whole step (full AIC assembly + full solve)"
        else:
            what = (f"a {self.f:.3f} fraction of the step: {max(self.cores, int(self.f * self.case.n_cp))} of {self.case.n_cp} rows "
                    f"assembled + {int(round(self.f * self.iters))} of {self.iters} Arnoldi steps on the full matrix")
        return (f"{what}; oracle/ C++ restatement of the reference (the Fortran cannot be built: no compiler), OpenMP on "
                f"{self.cores} threads incl. the GMRES matvec (the reference's matmul is single-threaded); one full run: "
                f"assembly {self.t_asm:.2f} s, solve {self.t_sol:.2f} s, {self.iters} iterations, residual {self.res_norm:.1e}")


def cpu_baseline(case):
    """bench.py's cpu_baseline leg: one reference step on the box's host cores (bounded to ~30 s of CPU work)."""
    r = ReferenceRunner(case, budget_s=30.0, n_steps=1)
    if r.f >= 1.0:
        pairs, dt = case.n_pairs, r.t_asm + r.t_sol      # the constructor's own full run is the sample
    else:
        pairs, dt = r.step(0)
    return {"value": pairs / dt, "unit": UNIT, "cores": r.cores, "kind": "port", "sample": r.sample_text(),
            "assemble_only_pairs_per_s": case.n_pairs / r.t_asm}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on this box's host cores, same metric and config."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    tmp = tempfile.mkdtemp(prefix="machline_bench_ref_")
    case, dims = build_case(args.gpus, tmp, args.matrix_solver, args.dims, args.workload)
    n_steps = args.steps + args.warmup
    r = ReferenceRunner(case, budget_s=120.0, n_steps=n_steps)
    times, pairs = [], []
    for i in range(n_steps):
        if r.f >= 1.0 and i == 0:
            p, dt = case.n_pairs, r.t_asm + r.t_sol       # the constructor's full run is the first warm-up step
        else:
            p, dt = r.step(i)
        if i >= args.warmup:
            times.append(dt)
            pairs.append(p)
    dt = float(np.mean(times))
    value = float(np.sum(pairs) / np.sum(times))
    cores = r.cores
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": dims["data"],
            "config": {"workload": workload_name(case, dims), "n_panels": case.info.n_body_panels, "n_unknown": case.n_unknown,
                       "pairs_per_step": float(np.mean(pairs)), "pairs_full_case": case.n_pairs, "matrix_solver": args.matrix_solver},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": r.sample_text()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(case, dims):
    i = case.info
    img = case.body.n_images
    head = {"onera_m6": "BASELINE configs[1]: ONERA M6 wing, mirrored about xz, M=0.5, alpha=3.06deg, automatic wake",
            "synthetic_wing": "BASELINE configs[1] family (weak-scaling stand-in): mirrored swept tapered half wing, ONERA-M6 planform, "
                              "M=0.5, alpha=3.06deg, automatic wake",
            "cone": "BASELINE configs[2]: 10 deg cone, mirrored about xy, M=1.5", "sears_haack": "BASELINE configs[2]: Sears-Haack body, M=2, source-free",
            "agard_b": "BASELINE configs[3]: AGARD-B wing-body, mirrored about yz, M=1.6, supersonic wake"}[dims["workload"]]
    return (f"{head}; lower-order Dirichlet; {i.n_body_panels} body panels x {img} images + {i.n_wake_panels} wake panels, "
            f"{case.n_unknown} unknowns; mesh: {dims['mesh']}")


def supersonic_probe(local: int, fp64_peak: float, tmp: str):
    """BASELINE configs[2] on the reference's SH_160_60.tri (M = 2, every pair DoD-tested): the supersonic instantiation of
    the assembly kernel against the FP64 roofline in SURVEY 8(d)'s flop convention (30 per culled pair, 76 + 44 e per pair
    evaluated with e edges in the domain of dependence; the class counts come from ml_dod_census), outside the timed region."""
    from machline_b200 import gpu
    case, desc = build_case(1, tmp, "GMRES", None, "sears_haack")
    ctx = gpu.Context(local)
    try:
        ctx.set_case(case)
        ctx.assemble()
        ms = min(ctx.assemble_resident() for _ in range(3))
        census = ctx.dod_census()
        flops = 30.0 * census[0] + sum((76.0 + 44.0 * e) * census[e] for e in (1, 2, 3))
        x, info = ctx.solve(case.solver_opts(), np.array(case.BC))
        res = case.post(x)
        pairs = ctx.pair_count
        tfl = flops / (ms * 1e-3) / 1e12
        return {"kernel": "aic_assemble_kernel<supersonic> (DoD test fused, culled pairs skipped)", "bound": "fp64",
                "achieved": tfl, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tfl / fp64_peak if fp64_peak else None, "traffic": None,
                "workload": workload_name(case, desc), "avg_launch_ms": ms, "pairs": pairs, "pairs_per_s": pairs / (ms * 1e-3),
                "pair_classes": {"culled": census[0], "edges_in_dod_1": census[1], "edges_in_dod_2": census[2], "edges_in_dod_3": census[3]},
                "algorithmic_flops": flops, "solve": {"iterations": int(info.iterations), "ms": info.solve_ms, "res_norm": info.res_norm},
                "result_check": {"C_p_max": res.C_p_max, "C_p_min": res.C_p_min, "Cx": float(res.C_F[0])}}
    finally:
        ctx.close()
        case.close()


def single_gpu_run(local: int, case, opts, BC, steps: int):
    """One context, whole system on this GPU: (x, iterations, ms per step with resident tables, pairs)."""
    from machline_b200 import gpu
    ctx = gpu.Context(local)
    try:
        ctx.set_case(case)
        ctx.assemble()
        x, info = ctx.solve(opts, BC)
        t = []
        for _ in range(steps):
            a = ctx.assemble_resident()
            x, info = ctx.solve(opts, BC)
            t.append(a + info.solve_ms)
        return x, info, (float(np.mean(t)) if t else None), ctx.pair_count
    finally:
        ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--matrix-solver", default="GMRES")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lu-probe", action="store_true")
    ap.add_argument("--dims", default=None, help="tests only: NCxNS synthetic mesh instead of the BASELINE-sized one")
    ap.add_argument("--workload", default=None, choices=["onera_m6", "synthetic_wing", "cone", "sears_haack", "agard_b"],
                    help="default: onera_m6 (the reference's mesh) at N=1, synthetic_wing (refined with N) at N>1")
    ap.add_argument("--no-extra-probes", action="store_true", help="skip the supersonic / scaling-reference / parity legs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if os.environ.get("MACHLINE_BENCH_WATCHDOG"):   # debugging aid: dump every thread's stack and exit if the run exceeds N seconds
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["MACHLINE_BENCH_WATCHDOG"]), exit=True)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    from machline_b200 import gpu, shard

    dist, rank, world, local = dist_setup(args.gpus)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    tmp = tempfile.mkdtemp(prefix=f"machline_bench_r{rank}_")
    case, dims = build_case(world, tmp, args.matrix_solver, args.dims, args.workload)
    N = case.n_cp
    row0, nrows = shard.row_shard(N, rank, world)   # contiguous row blocks of the permuted system
    ctx = gpu.Context(local)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(gpu.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        ctx.set_communicator(bytes(uid.cpu().numpy().tobytes()), rank, world)
    opts = case.solver_opts()
    BC = np.array(case.BC)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # rows of the permuted system per rank: contiguous blocks for the Krylov solvers; block-cyclic for the direct solver,
    # where contiguous blocks would retire the ranks one after the other (DESIGN.md "Multi-GPU")
    shard_kw = dict(row0=row0, nrows=nrows)
    if world > 1 and args.matrix_solver == "LU":
        shard_kw = dict(cyclic=(shard.CYCLIC_BLOCK, rank, world))

    # ---- resident-input steps (value): tables already on the device --------------------------------
    ctx.set_case(case, **shard_kw)
    ctx.assemble()                      # H2D of the tables + first assembly (untimed)
    x = None
    for _ in range(args.warmup):
        ctx.assemble_resident()
        x, info = ctx.solve(opts, BC)
    launches0 = ctx.launch_count
    sampler = ClockSampler(local)
    # device timing: CUDA events recorded on the context's own stream (torch.cuda.Event sees only the stream it is
    # recorded on); the host-side Givens/convergence test of GMRES runs while that stream is busy, so the event
    # interval is the whole step.  Wall clock is kept as a cross-check.
    ext = torch.cuda.ExternalStream(ctx.stream_handle, device=torch.device("cuda", local))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    ev[0].record(ext)
    asm_ms, sol_ms = [], []
    for _ in range(args.steps):
        asm_ms.append(ctx.assemble_resident())
        x, info = ctx.solve(opts, BC)
        sol_ms.append(info.solve_ms)
    ev[1].record(ext)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    step_ms_local = ev[0].elapsed_time(ev[1]) / args.steps
    step_wall_ms_local = wall * 1e3 / args.steps
    local_pairs = ctx.pair_count

    # ---- end-to-end steps: host tables in, x out, every step ---------------------------------------
    ctx.profile(reset=True)
    barrier()
    t0 = time.perf_counter()
    ev[2].record(ext)
    for _ in range(args.steps):
        ctx.set_case(case, **shard_kw)   # marks the device tables dirty -> H2D again in assemble()
        ctx.assemble()
        x, info_e = ctx.solve(opts, BC)
    ev[3].record(ext)
    barrier()
    e2e_wall = time.perf_counter() - t0
    prof = ctx.profile()
    # the end-to-end figure includes the host-side packing of the tables (ml_set_*), which no device event sees:
    # wall clock between the two barriers, cross-checked by the event interval
    e2e_ms_local = e2e_wall * 1e3 / args.steps
    e2e_dev_ms_local = ev[2].elapsed_time(ev[3]) / args.steps

    # ---- one profiled step: CUDA-event time of the HBM-bound gemv kernel ------------------------------
    ctx.set_profiling(True)
    ctx.profile(reset=True)
    ctx.assemble_resident()
    ctx.solve(opts, BC)
    gp = ctx.profile()
    ctx.set_profiling(False)

    # ---- LU probe (outside the timed region; N = 1 only): the direct solver on the same resident system, against the
    # measured FP64 tensor-pipe (DMMA) peak -- the third roofline north_star names ------------------------------------
    lu_probe = None
    if world == 1 and not args.no_lu_probe:
        from machline_b200 import _abi
        lu_opts = _abi.solver_opts("LU", preconditioner="DIAG" if opts.preconditioner else "none")
        best = None
        for _ in range(2):
            ctx.assemble_resident()
            x_lu, info_lu = ctx.solve(lu_opts, BC)
            best = info_lu.solve_ms if best is None else min(best, info_lu.solve_ms)
        dmma_peak = ctx.measure_dmma_peak()
        flops = 2.0 / 3.0 * float(N) ** 3
        lu_probe = {"kernel": "blocked LU (lu_panel_cl2_kernel / lu_panel_coop_kernel + lu_trsm_kernel + lu_gemm2_kernel DMMA), whole solve",
                    "bound": "tensor(fp64)", "n": int(N), "solve_ms": best, "achieved": flops / (best * 1e-3) / 1e12,
                    "peak": dmma_peak, "unit": "TFLOP/s", "frac": flops / (best * 1e-3) / 1e12 / dmma_peak,
                    "peak_source": "measured live: register-resident mma.sync.m8n8k4.f64 loop (ml_measure_dmma_peak)",
                    "res_norm": info_lu.res_norm, "max_abs_dx_vs_gmres": float(np.abs(x_lu - x).max()),
                    "note": "at this N the solve is latency-bound on the panel launches; profiles/ holds larger N and the "
                            "8-GPU N = 104k run"}
        ctx.assemble_resident()   # leave the resident system as the timed region left it

    # ---- legs outside the timed region -------------------------------------------------------------------------------
    extra = {}
    if not args.no_extra_probes and not args.dims:
        if world > 1:
            barrier()
            if rank == 0:
                # (1) parity of the sharded path: the same case assembled and solved on ONE GPU
                x1, info1, _, _ = single_gpu_run(local, case, opts, BC, 0)
                r_sh, r_1 = case.post(x), case.post(x1)
                extra["parity"] = {
                    "what": f"row-sharded x{world} assemble + solve against a single-GPU assemble + solve of the same case",
                    "max_abs_dx_over_max_abs_x": float(np.abs(x - x1).max() / np.abs(x1).max()),
                    "iterations": [int(info.iterations), int(info1.iterations)],
                    "res_norm": [float(info.res_norm), float(info1.res_norm)],
                    "max_abs_dCp": float(np.abs(r_sh.C_p - r_1.C_p).max()),
                    "max_abs_dCF": float(np.abs(np.array(r_sh.C_F) - np.array(r_1.C_F)).max())}
                # (2) the weak-scaling baseline of THIS mesh family: its N = 1 member on one GPU, same run
                case1, dims1 = build_case(1, tmp, args.matrix_solver, None, "synthetic_wing")
                _, i1, ms1, pairs1 = single_gpu_run(local, case1, case1.solver_opts(), np.array(case1.BC), args.steps)
                extra["scaling_reference"] = {"what": "same synthetic family at N = 1 (pairs per GPU ~equal), one GPU, resident tables",
                                              "workload": workload_name(case1, dims1), "value": pairs1 / (ms1 * 1e-3), "unit": UNIT,
                                              "ms_per_step": ms1, "iterations": int(i1.iterations)}
                case1.close()
            barrier()
    # ---- sharded direct solve (N > 1): the row-sharded blocked LU on the same case, rows dealt block-cyclically -------------
    lu_sharded = None
    if world > 1 and not args.no_lu_probe and not args.no_extra_probes and not args.dims:
        from machline_b200 import _abi
        barrier()
        ctx.set_case(case, cyclic=(shard.CYCLIC_BLOCK, rank, world))
        ctx.assemble()
        lu_opts = _abi.solver_opts("LU", preconditioner="DIAG" if opts.preconditioner else "none")
        x_lu, info_lu = ctx.solve(lu_opts, BC)
        t_lu = torch.tensor([info_lu.solve_ms], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t_lu, op=dist.ReduceOp.MAX)
        lu_ms = float(t_lu.cpu()[0])
        if rank == 0:
            dmma_peak = ctx.measure_dmma_peak()
            flops = 2.0 / 3.0 * float(N) ** 3
            lu_sharded = {"kernel": f"row-sharded blocked LU over {world} GPUs (lu_sharded.cu: gathered panel, replicated panel factorisation, "
                                    "U-row exchange over NCCL, local DMMA update), whole solve, rows block-cyclic 128",
                          "bound": "tensor(fp64)", "n": int(N), "solve_ms": lu_ms, "achieved": flops / (lu_ms * 1e-3) / 1e12,
                          "peak": dmma_peak * world, "unit": "TFLOP/s", "frac": flops / (lu_ms * 1e-3) / 1e12 / (dmma_peak * world),
                          "peak_source": f"{world} x the DMMA peak measured live on rank 0 (ml_measure_dmma_peak)",
                          "res_norm": info_lu.res_norm, "max_abs_dx_vs_gmres": float(np.abs(x_lu - x).max())}
        barrier()
    # ---- reduce over ranks: max time, sum of pairs ---------------------------------------------------
    vals = torch.tensor([step_ms_local, e2e_ms_local, float(np.mean(asm_ms)), float(np.mean(sol_ms)), step_wall_ms_local,
                         e2e_dev_ms_local, gp.gemv_ms, gp.comm_ms], dtype=torch.float64, device=f"cuda:{local}")
    pairs_t = torch.tensor([float(local_pairs)], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(pairs_t, op=dist.ReduceOp.SUM)
    step_ms, e2e_ms, a_ms, s_ms, step_wall_ms, e2e_dev_ms, gemv_ms_max, comm_ms_max = [float(v) for v in vals.cpu()]
    pairs = float(pairs_t.cpu()[0])

    if rank == 0:
        res = case.post(x)
        fp64_peak, copy_peak = ctx.measure_peaks()
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        hbm_peak, hbm_src = copy_peak, "measured live (ml_measure_peaks copy kernel); MEASURED_PEAKS.json absent"
        if peaks_file.exists():
            try:
                hbm_peak = float(json.loads(peaks_file.read_text())["hbm_gbs"])
                hbm_src = "MEASURED_PEAKS.json hbm_gbs (driver-measured copy)"
            except Exception:
                pass
        traffic = {}
        tf = ROOT / "profiles" / "roofline_traffic.json"
        if tf.exists():
            try:
                traffic = json.loads(tf.read_text())
            except Exception:
                traffic = {}
        gemv_avg_ms = gp.gemv_ms / max(1, gp.gemv_launches)
        gemv_bytes = gp.gemv_bytes / max(1, gp.gemv_launches)
        gemv_gbs = gemv_bytes / (gemv_avg_ms * 1e-3) / 1e9 if gemv_avg_ms > 0 else 0.0
        gemv_share = gp.gemv_ms / max(1e-9, (a_ms + s_ms))
        asm_tflops = ALG_FLOPS_PER_PAIR_SUBSONIC * local_pairs / (a_ms * 1e-3) / 1e12
        roof_gemv = {"kernel": "gemv_n_partial_kernel (GMRES matvec w = A q)", "bound": "hbm", "achieved": gemv_gbs,
                     "peak": hbm_peak, "unit": "GB/s", "frac": gemv_gbs / hbm_peak if hbm_peak else None,
                     "traffic": traffic.get("gemv_n_partial_kernel") if world == 1 else None,
                     "traffic_source": (traffic.get("_source", {}).get("gemv") if world == 1 else "not profiled at this N"),
                     "peak_source": hbm_src,
                     "algorithmic_bytes_per_launch": gemv_bytes, "avg_launch_ms": gemv_avg_ms,
                     "launches_per_step": int(gp.gemv_launches), "share_of_step": gemv_share}
        roof_asm = {"kernel": "aic_assemble_kernel<subsonic>", "bound": "fp64", "achieved": asm_tflops, "peak": fp64_peak,
                    "unit": "TFLOP/s", "frac": asm_tflops / fp64_peak if fp64_peak else None,
                    "traffic": traffic.get("aic_assemble_kernel") if world == 1 else None,
                    "traffic_source": (traffic.get("_source", {}).get("aic") if world == 1 else "not profiled at this N"),
                    "ncu": traffic.get("_aic_ncu") if world == 1 else None,
                    "peak_source": "measured live: register-resident DFMA loop on all SMs (ml_measure_peaks); "
                                   "MEASURED_PEAKS.json carries no FP64 figure",
                    "algorithmic_flops_per_pair": ALG_FLOPS_PER_PAIR_SUBSONIC, "avg_launch_ms": a_ms,
                    "share_of_step": a_ms / max(1e-9, (a_ms + s_ms))}
        dominant, other = (roof_gemv, roof_asm) if gemv_share >= roof_asm["share_of_step"] else (roof_asm, roof_gemv)
        sup_probe = None
        if world == 1 and not args.no_extra_probes and not args.dims:
            sup_probe = supersonic_probe(local, fp64_peak, tmp)
        line = {
            "metric": METRIC, "value": pairs / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": dims["data"],
            "config": {"workload": workload_name(case, dims), "n_panels": case.info.n_body_panels, "n_unknown": case.n_unknown,
                       "pairs_per_step": pairs, "matrix_solver": args.matrix_solver, "parallelism": f"row-sharded x{world}" + (" (block-cyclic 128)" if "cyclic" in shard_kw else ""),
                       "l2": "inputs larger than L2: A (8*N^2 bytes) is rewritten by every assembly and streamed by every matvec"},
            "assemble": {"ms": a_ms, "pairs_per_s": pairs / (a_ms * 1e-3)},
            "solve": {"ms": s_ms, "iterations": int(info.iterations), "res_norm": info.res_norm, "res_max": info.res_max,
                      "breakdown_ms": {"matvec_kernels": gemv_ms_max, "exchange_allgather": comm_ms_max,
                                       "orthogonalisation_and_host": max(0.0, s_ms - gemv_ms_max - comm_ms_max),
                                       "how": "CUDA events per launch in one separately profiled step, max over ranks"}},
            "end_to_end_solve_ms": step_ms,
            "timing": {"ms_per_step": "CUDA events on the context's stream, max over ranks", "wall_ms_per_step": step_wall_ms,
                       "e2e": "wall clock between barriers (includes host-side table packing), max over ranks",
                       "e2e_device_event_ms_per_step": e2e_dev_ms},
            "result_check": {"C_p_max": res.C_p_max, "C_p_min": res.C_p_min, "Cx": float(res.C_F[0]), "Cz": float(res.C_F[2])},
            "e2e": {"value": pairs / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": prof.h2d_bytes / args.steps, "d2h_bytes_per_step": prof.d2h_bytes / args.steps,
                    "path": "host tables (caller buffers) -> ml_set_* -> packed into pinned staging -> async H2D -> ml_assemble -> ml_solve -> x on host (C ABI)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": dominant,
            "roofline_other": [other] + ([lu_probe] if lu_probe else []) + ([sup_probe] if sup_probe else []) + ([lu_sharded] if lu_sharded else []),
            "host_setup_s": dims["host_setup_s"],
            # the `main`-equivalent wall time of SURVEY 8(d) M2: mesh / wake / control points / panel tables on the host, then
            # tables -> device -> assembly -> solve -> x on the host (post-processing and file output excluded)
            "main_equivalent_s": dims["host_setup_s"] + e2e_ms * 1e-3,
        }
        if extra:
            line.update(extra)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(case)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
