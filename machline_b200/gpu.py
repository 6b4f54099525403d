"""ctypes wrapper over libmachline_gpu.so (include/machline_gpu.h).

The library is the product's only compute path: there is no CPU fallback.  Loading it on a box
without the built extension, or creating a context without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi, build


class GpuError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"{_abi.ML_STATUS_NAMES.get(status, status)}: {msg}")
        self.status = status


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        import os
        from pathlib import Path
        path = Path(os.environ["MACHLINE_GPU_LIB"]) if os.environ.get("MACHLINE_GPU_LIB") else build.GPU_LIB   # kernel experiments: a variant build
        if not path.exists():
            raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the CUDA extension is the only compute path; there is no fallback)")
        L = C.CDLL(str(path))
        vp, dp, ip = C.c_void_p, _abi.c_double_p, _abi.c_int_p
        L.ml_abi_version.restype = C.c_int
        L.ml_ctx_create.argtypes = [C.POINTER(vp), C.c_int]
        L.ml_ctx_destroy.argtypes = [vp]
        L.ml_ctx_destroy.restype = None
        L.ml_last_error.argtypes = [vp]
        L.ml_last_error.restype = C.c_char_p
        L.ml_set_flow.argtypes = [vp, C.POINTER(_abi.MlFlow)]
        L.ml_set_panels.argtypes = [vp, C.POINTER(_abi.MlPanelSoa), C.POINTER(_abi.MlPanelSoa)]
        L.ml_set_control_points.argtypes = [vp, C.c_int, dp, ip, dp, ip]
        L.ml_set_system_map.argtypes = [vp, C.POINTER(_abi.MlSystemMap)]
        L.ml_set_row_shard.argtypes = [vp, C.c_int, C.c_int]
        L.ml_set_row_shard_cyclic.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.ml_local_rows.argtypes = [vp, ip, ip]
        L.ml_set_communicator.argtypes = [vp, C.c_void_p, C.c_int, C.c_int]
        L.ml_assemble.argtypes = [vp, dp]
        L.ml_assemble_resident.argtypes = [vp, dp]
        L.ml_get_A.argtypes = [vp, C.c_int, C.c_int, dp, C.c_int]
        L.ml_set_A.argtypes = [vp, C.c_int, C.c_int, dp, C.c_int]
        L.ml_pair_count.argtypes = [vp]
        L.ml_pair_count.restype = C.c_longlong
        L.ml_launch_count.argtypes = [vp]
        L.ml_launch_count.restype = C.c_longlong
        L.ml_solve.argtypes = [vp, C.POINTER(_abi.MlSolverOpts), dp, dp, C.POINTER(_abi.MlSolveInfo)]
        L.ml_solve_dense.argtypes = [vp, C.c_int, dp, dp, C.POINTER(_abi.MlSolverOpts), dp,
                                     C.POINTER(_abi.MlSolveInfo)]
        L.ml_dod_census.argtypes = [vp, C.POINTER(C.c_longlong)]
        L.ml_check_system.argtypes = [vp, dp, ip, ip]
        L.ml_residual.argtypes = [vp, dp, dp, dp]
        L.ml_ctx_create_multi.argtypes = [C.POINTER(vp), ip, C.c_int]
        L.ml_multi_set_dealing.argtypes = [vp, C.c_int]
        L.ml_device_count.argtypes = [vp]
        L.ml_device_count.restype = C.c_int
        L.ml_device_system.argtypes = [vp, C.POINTER(dp), ip, ip, ip]
        L.ml_measure_peaks.argtypes = [vp, dp, dp]
        L.ml_measure_dmma_peak.argtypes = [vp, dp]
        L.ml_device_stream.argtypes = [vp, C.POINTER(vp)]
        L.ml_nccl_unique_id.argtypes = [C.c_void_p]
        L.ml_set_profiling.argtypes = [vp, C.c_int]
        L.ml_get_profile.argtypes = [vp, C.POINTER(_abi.MlProfile)]
        L.ml_reset_profile.argtypes = [vp]
        L.ml_post_process.argtypes = [vp, C.POINTER(_abi.MlPostTables), C.POINTER(_abi.MlPostFlow), dp, C.POINTER(_abi.MlPostOut)]
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_abi.c_double_p)


class Context:
    """One ml_ctx: bound to one CUDA device (device = id), or -- devices = [ids] -- the single-process multi-GPU context of
    ml_ctx_create_multi, driven through the same methods from this one thread."""

    def __init__(self, device: int = 0, devices: list[int] | None = None):
        self._h = C.c_void_p()
        if devices is not None:
            ids = (C.c_int * len(devices))(*devices)
            st = lib().ml_ctx_create_multi(C.byref(self._h), ids, len(devices))
            device = devices[0]
        else:
            st = lib().ml_ctx_create(C.byref(self._h), device)
        if st != 0:
            raise GpuError(st, "ml_ctx_create failed (no CUDA device visible?)")
        self.multi = devices is not None
        self.device = device
        self.row0 = 0
        self.nrows = None
        self.n_unknown = 0

    def _check(self, st: int):
        if st != 0:
            raise GpuError(st, lib().ml_last_error(self._h).decode())

    # ---- inputs ----------------------------------------------------------------------------------
    def set_case(self, case, row0: int = 0, nrows: int | None = None, cyclic: tuple[int, int, int] | None = None):
        """Stage the tables of a machline_b200.host.Case (host -> device happens in assemble()).  Rows of the permuted
        system owned by this context: [row0, row0 + nrows), or with cyclic = (block, rank, world) the blocks
        b = rank (mod world) of `block` rows."""
        L = lib()
        self._check(L.ml_set_flow(self._h, C.byref(case.flow)))
        wake = C.byref(case.wake) if case.wake.n_panels > 0 else None
        self._check(L.ml_set_panels(self._h, C.byref(case.body), wake))
        self._check(L.ml_set_control_points(self._h, case.cps.n_cp, case.cps.loc, case.cps.bc, case.cps.n_g,
                                            case.cps.row_perm))
        self._check(L.ml_set_system_map(self._h, C.byref(case.map)))
        self.n_unknown = case.n_unknown
        self.n_cp = case.n_cp
        if self.multi:      # the library deals the rows to its devices; `cyclic` = (block, ...) selects block-cyclic dealing
            assert row0 == 0 and nrows is None
            self._check(L.ml_multi_set_dealing(self._h, cyclic[0] if cyclic is not None else -1))
            self.row0, self.nrows = 0, case.n_cp
            self.local_rows = np.arange(case.n_cp, dtype=np.int32)
            return
        if cyclic is not None:
            from . import shard
            block, rank, world = cyclic
            self._check(L.ml_set_row_shard_cyclic(self._h, block, rank, world))
            self.local_rows = shard.cyclic_rows(case.n_cp, rank, world, block)
            self.row0, self.nrows = (int(self.local_rows[0]) if len(self.local_rows) else 0), len(self.local_rows)
            return
        if nrows is None:
            nrows = case.n_cp - row0
        self._check(L.ml_set_row_shard(self._h, row0, nrows))
        self.row0, self.nrows = row0, nrows
        self.local_rows = np.arange(row0, row0 + nrows, dtype=np.int32)

    def set_points(self, case, points: np.ndarray, direction=None, with_wake: bool = True):
        """Stage `case` with the rows of the system replaced by arbitrary field points (boundary condition "zero
        potential", no sorting): ml_assemble then builds the influence matrix of every unknown on those points, which is
        what the reference's off-body sweep evaluates (surface_mesh_get_induced_potentials_at_point,
        src/surface_mesh.f90:2282-2385)."""
        L = lib()
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        n = pts.shape[0]
        bc = np.full(n, 1, dtype=np.int32)            # ML_BC_ZERO_POTENTIAL
        n_g = None
        if direction is not None:                     # rows = the induced velocity along `direction` (ML_BC_ZERO_NORMAL_VEL)
            bc[:] = 5
            n_g = np.ascontiguousarray(np.tile(np.asarray(direction, dtype=np.float64), (n, 1)))
        rows = np.arange(n, dtype=np.int32)
        m = _abi.MlSystemMap()
        C.memmove(C.byref(m), C.byref(case.map), C.sizeof(m))
        m.n_cp = n
        self._check(L.ml_set_flow(self._h, C.byref(case.flow)))
        wake = C.byref(case.wake) if (with_wake and case.wake.n_panels > 0) else None
        self._check(L.ml_set_panels(self._h, C.byref(case.body), wake))
        self._check(L.ml_set_control_points(self._h, n, _dp(pts), bc.ctypes.data_as(_abi.c_int_p), _dp(n_g) if n_g is not None else None,
                                            rows.ctypes.data_as(_abi.c_int_p)))
        self._check(L.ml_set_system_map(self._h, C.byref(m)))
        self._check(L.ml_set_row_shard(self._h, 0, n))
        self.n_unknown, self.n_cp = case.n_unknown, n
        self.row0, self.nrows = 0, n
        self.local_rows = np.arange(n, dtype=np.int32)

    def potentials_at(self, case, points: np.ndarray, x: np.ndarray, with_wake: bool = True):
        """(phi_d, phi_s) induced at `points` by the solved strengths x (per unit freestream speed): phi_d = A_points x,
        phi_s = the known-source sum the assembly returns as I_known.  Wake panels contribute to phi_d with their
        (top - bottom) strengths, as in the AIC rows (with_wake=False leaves them out)."""
        self.set_points(case, points, with_wake=with_wake)
        phi_s = self.assemble()
        phi_d = self.get_A() @ np.asarray(x, dtype=np.float64)
        return phi_d, phi_s

    def velocity_parts_at(self, case, points: np.ndarray, x: np.ndarray, with_wake: bool = True):
        """(v_d, v_s) induced at `points` (per unit freestream speed), separately: the doublet part A_points x and the known-source
        part, for each of the three global directions."""
        x = np.asarray(x, dtype=np.float64)
        v_d, v_s = np.zeros((len(points), 3)), np.zeros((len(points), 3))
        for k in range(3):
            e = np.zeros(3)
            e[k] = 1.0
            self.set_points(case, points, direction=e, with_wake=with_wake)
            v_s[:, k] = self.assemble()
            v_d[:, k] = self.get_A() @ x
        return v_d, v_s

    def velocities_at(self, case, points: np.ndarray, x: np.ndarray) -> np.ndarray:
        """Induced velocity v_d + v_s at `points` (per unit freestream speed) from the solved strengths x: three assemblies with
        the field points as rows and the unit vectors as projection directions (the kernels of the Neumann rows), what
        surface_mesh_get_induced_velocities_at_point sums panel by panel (src/surface_mesh.f90:2388-2500)."""
        x = np.asarray(x, dtype=np.float64)
        v = np.zeros((len(points), 3))
        for k in range(3):
            e = np.zeros(3)
            e[k] = 1.0
            self.set_points(case, points, direction=e)
            v_s = self.assemble()
            v[:, k] = self.get_A() @ x + v_s
        return v

    def set_communicator(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(unique_id, len(unique_id))
        self._check(lib().ml_set_communicator(self._h, buf, rank, world))

    # ---- hot path 1 --------------------------------------------------------------------------------
    def assemble(self) -> np.ndarray:
        """ml_assemble: builds A on the device, returns I_known for the local rows."""
        I_known = np.zeros(self.nrows, dtype=np.float64)
        self._check(lib().ml_assemble(self._h, _dp(I_known)))
        return I_known

    def assemble_resident(self) -> float:
        """Re-run the assembly kernels with inputs resident; returns device milliseconds."""
        ms = C.c_double()
        self._check(lib().ml_assemble_resident(self._h, C.byref(ms)))
        return ms.value

    @staticmethod
    def _runs(rows):
        """(start index, first row, length) of the maximal runs of consecutive global rows in `rows`."""
        rows = np.asarray(rows)
        if len(rows) == 0:
            return []
        cut = np.flatnonzero(np.diff(rows) != 1) + 1
        starts = np.concatenate([[0], cut])
        ends = np.concatenate([cut, [len(rows)]])
        return [(int(s), int(rows[s]), int(e - s)) for s, e in zip(starts, ends)]

    def get_A(self, row0: int | None = None, nrows: int | None = None) -> np.ndarray:
        if row0 is None and nrows is None:   # every local row, in local order (block-cyclic shards: run by run)
            A = np.zeros((self.nrows, self.n_unknown), dtype=np.float64, order="F")
            for s, r0, n in self._runs(self.local_rows):
                part = np.zeros((n, self.n_unknown), dtype=np.float64, order="F")
                self._check(lib().ml_get_A(self._h, r0, n, _dp(part), n))
                A[s:s + n] = part
            return A
        row0 = self.row0 if row0 is None else row0
        nrows = self.nrows if nrows is None else nrows
        A = np.zeros((nrows, self.n_unknown), dtype=np.float64, order="F")
        self._check(lib().ml_get_A(self._h, row0, nrows, _dp(A), nrows))
        return A

    def set_A(self, A_rows: np.ndarray, row0: int | None = None):
        """Overwrite rows [row0, row0 + len(A_rows)) of the resident system (ml_set_A)."""
        A_rows = np.asarray(A_rows, dtype=np.float64)
        if row0 is None:                      # one row per local row, in local order
            for s, r0, n in self._runs(self.local_rows):
                part = np.asfortranarray(A_rows[s:s + n])
                self._check(lib().ml_set_A(self._h, r0, n, _dp(part), n))
            return
        A_rows = np.asfortranarray(A_rows)
        self._check(lib().ml_set_A(self._h, row0, A_rows.shape[0], _dp(A_rows), A_rows.shape[0]))

    def dod_census(self) -> list[int]:
        """[culled pairs, evaluated pairs with 1, 2, 3 edges in the domain of dependence] of the assembled case."""
        out = (C.c_longlong * 4)()
        self._check(lib().ml_dod_census(self._h, out))
        return [int(v) for v in out]

    @property
    def pair_count(self) -> int:
        return int(lib().ml_pair_count(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib().ml_launch_count(self._h))

    # ---- hot path 2 --------------------------------------------------------------------------------
    def solve(self, opts: _abi.MlSolverOpts, BC: np.ndarray):
        BC = np.ascontiguousarray(BC, dtype=np.float64)
        x = np.zeros(self.n_unknown, dtype=np.float64)
        info = _abi.MlSolveInfo()
        st = lib().ml_solve(self._h, C.byref(opts), _dp(BC), _dp(x), C.byref(info))
        self._check(st)
        return x, info

    def check_system(self, BC: np.ndarray) -> tuple[int, int, int]:
        """panel_solver_check_system (src/panel_solver.f90:1709-1764) on the resident system.  Returns (status, zero rows,
        zero columns): status 0, 1 (NaN in A or b) or 2 (a control point not influenced / an unknown without influence)."""
        BC = np.ascontiguousarray(BC, dtype=np.float64)
        zr, zc = C.c_int(0), C.c_int(0)
        st = lib().ml_check_system(self._h, _dp(BC), C.byref(zr), C.byref(zc))
        if st not in (0, 1, 2):
            self._check(st)
        return st, zr.value, zc.value

    def residual(self, BC: np.ndarray, x: np.ndarray) -> np.ndarray:
        """R_cp = A x - (BC - I_known) on this context's rows (panel_solver.f90:1992)."""
        BC = np.ascontiguousarray(BC, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = np.zeros(self.nrows, dtype=np.float64)
        self._check(lib().ml_residual(self._h, _dp(BC), _dp(x), _dp(r)))
        return r

    def solve_dense(self, A: np.ndarray, b: np.ndarray, opts: _abi.MlSolverOpts):
        A = np.asfortranarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        N = A.shape[0]
        assert A.shape == (N, N) and b.shape == (N,)
        x = np.zeros(N, dtype=np.float64)
        info = _abi.MlSolveInfo()
        self._check(lib().ml_solve_dense(self._h, N, _dp(A), _dp(b), C.byref(opts), _dp(x), C.byref(info)))
        return x, info

    def post_process(self, case, v_inner: np.ndarray | None = None, x: np.ndarray | None = None) -> dict:
        """Lower-order post-processing on the device (ml_post_process) from the solution the last solve() left there (or from x):
        {"V_cells", "C_p": {rule: array}, "dC_f", "C_F", "C_M", "C_p_max", "C_p_min"}.  Neumann formulations need v_inner as
        host.Case.post does."""
        t, f = case.post_tables(v_inner)
        n = t.n_cells
        out = _abi.MlPostOut()
        V, dCf = np.zeros((n, 3)), np.zeros((n, 3))
        out.V_cells, out.dC_f = _dp(V), _dp(dCf)
        cps = {}
        for r, name in enumerate(_abi.RULES):
            if f.rules & (1 << r):
                cps[name] = np.zeros(n)
                out.C_p[r] = _dp(cps[name])
        xp = None
        if x is not None:
            x = np.ascontiguousarray(x, dtype=np.float64)
            xp = _dp(x)
        self._check(lib().ml_post_process(self._h, C.byref(t), C.byref(f), xp, C.byref(out)))
        return {"V_cells": V, "C_p": cps, "dC_f": dCf, "C_F": np.array(out.C_F[:]), "C_M": np.array(out.C_M[:]),
                "C_p_max": out.C_p_max, "C_p_min": out.C_p_min}

    def set_profiling(self, on: bool):
        self._check(lib().ml_set_profiling(self._h, int(on)))

    def profile(self, reset: bool = False) -> _abi.MlProfile:
        p = _abi.MlProfile()
        self._check(lib().ml_get_profile(self._h, C.byref(p)))
        if reset:
            self._check(lib().ml_reset_profile(self._h))
        return p

    @property
    def stream_handle(self) -> int:
        """cudaStream_t of this context (for CUDA-event timing by the caller)."""
        h = C.c_void_p()
        self._check(lib().ml_device_stream(self._h, C.byref(h)))
        return int(h.value or 0)

    def measure_peaks(self, hbm: bool = True):
        """(FP64 DFMA TFLOP/s, copy GB/s) measured on this device."""
        f, h = C.c_double(), C.c_double()
        self._check(lib().ml_measure_peaks(self._h, C.byref(f), C.byref(h) if hbm else None))
        return f.value, h.value

    def measure_dmma_peak(self) -> float:
        """FP64 tensor-pipe (DMMA) TFLOP/s measured on this device."""
        f = C.c_double()
        self._check(lib().ml_measure_dmma_peak(self._h, C.byref(f)))
        return f.value

    def close(self):
        if self._h:
            lib().ml_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    st = lib().ml_nccl_unique_id(buf)
    if st != 0:
        raise GpuError(st, "ml_nccl_unique_id")
    return buf.raw
