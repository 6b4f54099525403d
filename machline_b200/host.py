"""Host-side case setup (mesh, wake, control points, panel tables, post-processing).

Thin ctypes wrapper over libmachline_host.so (csrc/host/), which restates what MachLine's
``main`` does around the two hot paths (src/main.f90:102-160).  No influence or solver code lives
here; the tables produced are the arguments of the GPU library's ``ml_set_*`` calls.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from . import _abi, build


class MachLineError(RuntimeError):
    """Raised where the reference prints '!!! ...' and stops."""


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = build.HOST_LIB
        if not path.exists():
            build.build_host()
        L = C.CDLL(str(path))
        L.mlh_last_error.restype = C.c_char_p
        L.mlh_case_create.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.mlh_case_destroy.argtypes = [C.c_void_p]
        L.mlh_case_info.argtypes = [C.c_void_p, C.POINTER(_abi.MlhMeshInfo)]
        L.mlh_case_tables.argtypes = [C.c_void_p, C.POINTER(_abi.MlFlow), C.POINTER(_abi.MlPanelSoa),
                                      C.POINTER(_abi.MlPanelSoa), C.POINTER(_abi.MlSystemMap),
                                      C.POINTER(_abi.MlhCpTable)]
        L.mlh_case_solver_settings.argtypes = [C.c_void_p, C.POINTER(_abi.MlhSolverSettings)]
        L.mlh_case_post.argtypes = [C.c_void_p, _abi.c_double_p, C.POINTER(_abi.MlhResults)]
        L.mlh_case_post2.argtypes = [C.c_void_p, _abi.c_double_p, _abi.c_double_p, C.POINTER(_abi.MlhResults)]
        L.mlh_case_inner_points.argtypes = [C.c_void_p, _abi.c_double_p, C.POINTER(C.c_int)]
        L.mlh_case_write_report.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(_abi.MlSolveInfo), C.c_int,
                                            C.c_double]
        L.mlh_case_write_body.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.mlh_case_write_wake.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.mlh_case_write_control_points.argtypes = [C.c_void_p, C.c_char_p, _abi.c_double_p]
        L.mlh_case_result_array.argtypes = [C.c_void_p, C.c_int, _abi.c_double_p, C.POINTER(C.c_int)]
        L.mlh_case_post_tables.argtypes = [C.c_void_p, _abi.c_double_p, C.POINTER(_abi.MlPostTables), C.POINTER(_abi.MlPostFlow)]
        _lib = L
    return _lib


@dataclass
class Results:
    C_p_max: float
    C_p_min: float
    C_F: np.ndarray
    C_M: np.ndarray
    mu: np.ndarray
    C_p: np.ndarray
    V_cells: np.ndarray


class Case:
    """One MachLine input (dict / JSON text / path) after mesh + flow + solver initialisation."""

    def __init__(self, inp, base_dir: str | os.PathLike | None = None):
        if isinstance(inp, (str, os.PathLike)) and os.path.exists(str(inp)):
            text = Path(inp).read_text()
            if base_dir is None:
                base_dir = "."
        elif isinstance(inp, dict):
            text = json.dumps(inp)
        else:
            text = str(inp)
        self.input = json.loads(text)
        self._h = C.c_void_p()
        rc = lib().mlh_case_create(text.encode(), str(base_dir or "").encode(), C.byref(self._h))
        if rc != 0:
            raise MachLineError(lib().mlh_last_error().decode())
        self.flow = _abi.MlFlow()
        self.body = _abi.MlPanelSoa()
        self.wake = _abi.MlPanelSoa()
        self.map = _abi.MlSystemMap()
        self.cps = _abi.MlhCpTable()
        rc = lib().mlh_case_tables(self._h, C.byref(self.flow), C.byref(self.body), C.byref(self.wake),
                                   C.byref(self.map), C.byref(self.cps))
        if rc != 0:
            raise MachLineError(lib().mlh_last_error().decode())
        self.info = _abi.MlhMeshInfo()
        lib().mlh_case_info(self._h, C.byref(self.info))
        self.settings = _abi.MlhSolverSettings()
        lib().mlh_case_solver_settings(self._h, C.byref(self.settings))

    # -- convenience views (numpy, no copies) ---------------------------------------------------
    @property
    def n_cp(self) -> int:
        return self.map.n_cp

    @property
    def n_unknown(self) -> int:
        return self.map.n_unknown

    @property
    def BC(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.cps.BC, shape=(self.n_cp,))

    @property
    def cp_loc(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.cps.loc, shape=(self.n_cp, 3))

    @property
    def P(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.map.P, shape=(self.n_unknown,))

    @property
    def n_pairs(self) -> int:
        """Pair count of SURVEY 8(d): N_cp x (body records + evaluated wake records)."""
        nb = self.body.n_panels * self.body.n_images
        nw = self.wake.n_panels
        if self.wake.n_panels and self.wake.n_images == 2:
            pres = np.ctypeslib.as_array(self.wake.image_present, shape=(self.wake.n_panels,))
            nw += int(pres.sum())
        return self.n_cp * (nb + nw)

    def solver_opts(self) -> _abi.MlSolverOpts:
        o = _abi.MlSolverOpts()
        C.memmove(C.byref(o), C.byref(self.settings.opts), C.sizeof(o))
        return o

    @property
    def dirichlet(self) -> bool:
        return str(self.input.get("solver", {}).get("formulation", "dirichlet-morino")).startswith("dirichlet")

    def inner_points(self) -> np.ndarray:
        """The points just inside every panel where the non-Dirichlet formulations evaluate the induced velocity
        (panel_solver_calc_cell_velocities, src/panel_solver.f90:2063, 2080)."""
        n = C.c_int()
        lib().mlh_case_inner_points(self._h, None, C.byref(n))
        pts = np.zeros((n.value, 3), dtype=np.float64)
        if lib().mlh_case_inner_points(self._h, pts.ctypes.data_as(_abi.c_double_p), C.byref(n)) != 0:
            raise MachLineError(lib().mlh_last_error().decode())
        return pts

    def post(self, x: np.ndarray, v_inner: np.ndarray | None = None) -> Results:
        """Post-processing of the solved strengths x.  Neumann formulations need v_inner = the induced velocity (per unit
        freestream speed) at inner_points() -- gpu.Context.velocities_at(case, case.inner_points(), x)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.shape == (self.n_unknown,)
        r = _abi.MlhResults()
        vp = None
        if v_inner is not None:
            v_inner = np.ascontiguousarray(v_inner, dtype=np.float64)
            vp = v_inner.ctypes.data_as(_abi.c_double_p)
        rc = lib().mlh_case_post2(self._h, x.ctypes.data_as(_abi.c_double_p), vp, C.byref(r))
        if rc != 0:
            raise MachLineError(lib().mlh_last_error().decode())
        return Results(
            C_p_max=r.C_p_max, C_p_min=r.C_p_min, C_F=np.array(r.C_F[:]), C_M=np.array(r.C_M[:]),
            mu=np.ctypeslib.as_array(r.mu, shape=(r.n_mu,)).copy(),
            C_p=np.ctypeslib.as_array(r.C_p, shape=(r.n_cells,)).copy(),
            V_cells=np.ctypeslib.as_array(r.V_cells, shape=(r.n_cells, 3)).copy())

    def result_array(self, rule) -> np.ndarray:
        """Per-rule array of the last post(): rule = a name of _abi.RULES (pressure coefficients per cell) or "dC_f"."""
        rid = -1 if rule == "dC_f" else _abi.RULES.index(rule)
        n = C.c_int(0)
        if lib().mlh_case_result_array(self._h, rid, None, C.byref(n)) != 0:
            raise MachLineError(lib().mlh_last_error().decode())
        out = np.zeros(n.value)
        if n.value:
            lib().mlh_case_result_array(self._h, rid, out.ctypes.data_as(_abi.c_double_p), C.byref(n))
        return out.reshape(-1, 3) if rule == "dC_f" else out

    def post_tables(self, v_inner: np.ndarray | None = None):
        """(MlPostTables, MlPostFlow) for gpu.Context.post_process: the lower-order post-processing on the device.  The tables
        live in the case handle until the next call."""
        t, f = _abi.MlPostTables(), _abi.MlPostFlow()
        vp = None
        if v_inner is not None:
            self._v_inner_keep = np.ascontiguousarray(v_inner, dtype=np.float64)
            vp = self._v_inner_keep.ctypes.data_as(_abi.c_double_p)
        if lib().mlh_case_post_tables(self._h, vp, C.byref(t), C.byref(f)) != 0:
            raise MachLineError(lib().mlh_last_error().decode())
        return t, f

    def write_report(self, path, info: _abi.MlSolveInfo, solver_stat: int = 0, runtime: float = 0.0):
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        rc = lib().mlh_case_write_report(self._h, str(path).encode(), C.byref(info), solver_stat, runtime)
        if rc != 0:
            raise MachLineError(lib().mlh_last_error().decode())

    # -- result files of the last post(), in the reference's legacy-VTK layout (csrc/host/outputs.cpp) ------------------
    def write_body(self, path, mirrored: bool = False):
        """output.body_file / output.mirrored_body_file (surface_mesh_write_body / write_body_mirror)."""
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        if lib().mlh_case_write_body(self._h, str(path).encode(), int(mirrored)) != 0:
            raise MachLineError(lib().mlh_last_error().decode())

    def write_wake(self, path) -> bool:
        """output.wake_file (wake_mesh_write_strips); False (and no file) when there is no wake to export."""
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        exported = C.c_int(0)
        if lib().mlh_case_write_wake(self._h, str(path).encode(), C.byref(exported)) != 0:
            raise MachLineError(lib().mlh_last_error().decode())
        return bool(exported.value)

    def write_control_points(self, path, residual: np.ndarray | None = None):
        """output.control_point_file (surface_mesh_write_control_points); residual = gpu.Context.residual(BC, x)."""
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        rp = None
        if residual is not None:
            residual = np.ascontiguousarray(residual, dtype=np.float64)
            assert residual.shape == (self.n_cp,)
            rp = residual.ctypes.data_as(_abi.c_double_p)
        if lib().mlh_case_write_control_points(self._h, str(path).encode(), rp) != 0:
            raise MachLineError(lib().mlh_last_error().decode())

    def close(self):
        if self._h:
            lib().mlh_case_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
