"""Public entry point: run one MachLine input end to end.

Mirrors ``program main`` / ``panel_solver%solve`` of the reference (src/main.f90:102-176,
src/panel_solver.f90:1033-1101): host setup -> AIC assembly on the GPU -> dense solve on the GPU
-> host post-processing -> report.json.  The two hot paths only exist as CUDA kernels
(libmachline_gpu.so); without a GPU this raises.
"""
from __future__ import annotations

import time
from dataclasses import dataclass

import numpy as np

from . import _abi, gpu, host, vtk_out


@dataclass
class RunResult:
    C_p_max: float
    C_p_min: float
    C_F: np.ndarray
    C_M: np.ndarray
    mu: np.ndarray
    C_p: np.ndarray
    iterations: int
    res_max: float
    res_norm: float
    assemble_ms: float
    solve_ms: float
    n_pairs: int
    gpu_launches: int
    total_s: float
    device_post: dict | None = None   # run_case(device_post=True): ml_post_process on the device-resident solution


class SolverStatus(RuntimeError):
    """The solve ended with a non-zero solver_stat of the reference (1..4); report.json has been written with the code."""

    def __init__(self, status: int, total_s: float):
        super().__init__(f"solver status {status} ({_abi.ML_STATUS_NAMES.get(status, status)})")
        self.status = status
        self.total_s = total_s


def run_case(inp, base_dir=None, device: int = 0, report_file: str | None = None,
             matrix_solver: str | None = None, device_post: bool = False) -> RunResult:
    """`inp`: dict, JSON text or path of a MachLine input file.  device_post: also run the lower-order post-processing on the
    device from the solution ml_solve left there (gpu.Context.post_process; velocities, pressure rules, forces, moments)."""
    t0 = time.perf_counter()
    case = host.Case(inp, base_dir=base_dir)
    ctx = gpu.Context(device)
    try:
        ctx.set_case(case)
        I_known = ctx.assemble()
        opts = case.solver_opts()
        if matrix_solver is not None:
            opts.matrix_solver = _abi.SOLVERS.get(matrix_solver, _abi.SOLVERS["GMRES"])
        if case.settings.write_A_and_b:           # panel_solver.f90:1834, before the solve; files in the working directory
            vtk_out.write_system(ctx.get_A(), np.asarray(case.BC) - I_known)
        if report_file is None:
            report_file = case.input.get("output", {}).get("report_file")
        # solver_stat of the reference (panel_solver.f90:1722-1761, 2006-2010; main.f90:140-176): 1 NaN in the system, 2 zero
        # row / column (both from check_system, run when solver.run_checks is set), 3 singular (lu_decomp), 4 NaN residual.
        # A non-zero status skips the post-processing and is written to the report, as the reference does.
        solver_stat, info = 0, _abi.MlSolveInfo()
        if case.settings.run_checks:
            solver_stat = ctx.check_system(case.BC)[0]
        if solver_stat == 0:
            try:
                x, info = ctx.solve(opts, case.BC)
            except gpu.GpuError as e:
                if e.status not in (1, 2, 3, 4):
                    raise
                solver_stat = e.status
        if solver_stat != 0:
            total = time.perf_counter() - t0
            if report_file and report_file != "none":
                case.write_report(report_file, info, solver_stat, total)
            raise SolverStatus(solver_stat, total)
        v_inner = None if case.dirichlet else ctx.velocities_at(case, case.inner_points(), x)   # panel_solver.f90:2063-2066
        # device-resident results first: a velocity sweep (Neumann) does not touch the solution kept by the solve
        dev = ctx.post_process(case, v_inner) if device_post else None
        res = case.post(x, v_inner)
        total = time.perf_counter() - t0
        if report_file and report_file != "none":
            case.write_report(report_file, info, 0, total)
        # surface_mesh_output_results (src/surface_mesh.f90:2503-2544) and the off-body sweep (main.f90:118-119, 170)
        out = case.input.get("output", {})
        wanted = lambda key: out.get(key) if out.get(key) and out.get(key) != "none" else None
        if wanted("body_file"):
            case.write_body(wanted("body_file"))
        if wanted("mirrored_body_file") and case.info.asym_flow:
            case.write_body(wanted("mirrored_body_file"), mirrored=True)
        if wanted("wake_file"):
            case.write_wake(wanted("wake_file"))
        if wanted("control_point_file"):
            case.write_control_points(wanted("control_point_file"), ctx.residual(case.BC, x))
        off = out.get("offbody_points", {}) or {}
        if off.get("points_file", "none") != "none" and off.get("output_file", "none") != "none":
            vtk_out.export_off_body_points(case, ctx, x, off["points_file"], off["output_file"])
        return RunResult(res.C_p_max, res.C_p_min, res.C_F, res.C_M, res.mu, res.C_p, info.iterations,
                         info.res_max, info.res_norm, info.assemble_ms, info.solve_ms, ctx.pair_count,
                         ctx.launch_count, total, dev)
    finally:
        ctx.close()
        case.close()
