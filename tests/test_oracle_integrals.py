"""CPU: the oracle's per-pair integrals against known answers tabulated from the reference's Python
prototype dev/unit_tests/panel.py (tests/golden/prototype_integrals.json, made by make_fixtures.py).
The prototype integrates quadrilaterals; a quadrilateral's H(1,1,1) and hH(1,1,3) are the sums over
its two triangles (the shared diagonal's edge terms cancel)."""
import ctypes as C
import json

import numpy as np
import pytest

import fixtures
import oracle_binding as ob
from machline_b200 import _abi

CASES = json.loads((fixtures.GOLDEN / "prototype_integrals.json").read_text())["cases"]


def _triangle_table(tris, supersonic):
    """ml_panel_soa for triangles lying in z = 0 with the freestream along +x and B = 1, for which
    panel_calc_g_to_ls_transform (src/panel.f90:403-489) gives A_g_to_ls = identity."""
    n = len(tris)
    arr = dict(centr=np.zeros((n, 3)), A=np.tile(np.eye(3).ravel(), (n, 1)), vls=np.zeros((n, 6)), nh=np.zeros((n, 6)),
               b=np.zeros((n, 3)), sb=np.zeros((n, 3)), J=np.ones(n), r=np.ones(n, dtype=np.int32), area=np.ones(n),
               vg=np.zeros((n, 9)), T=np.zeros((n, 9)), ivd=np.zeros((n, 3), dtype=np.int32),
               ips=np.arange(n, dtype=np.int32), hs=np.ones(n, dtype=np.uint8), ip=np.zeros(n, dtype=np.uint8))
    for j, tri in enumerate(tris):
        tri = np.asarray(tri, dtype=np.float64)  # (3, 2)
        c = tri.mean(axis=0)
        arr["centr"][j, :2] = c
        loc = tri - c
        arr["vls"][j] = loc.ravel()
        arr["vg"][j] = np.column_stack([tri, np.zeros(3)]).ravel()
        S = np.column_stack([np.ones(3), loc[:, 0], loc[:, 1]])
        arr["T"][j] = np.linalg.inv(S).ravel()
        for k in range(3):
            t = loc[(k + 1) % 3] - loc[k]
            t = t / np.hypot(*t)
            nx, ny = t[1], -t[0]
            arr["nh"][j, 2 * k:2 * k + 2] = (nx, ny)
            if supersonic:
                arr["b"][j, k] = (nx - ny) * (nx + ny)
                arr["sb"][j, k] = np.sqrt(abs(arr["b"][j, k]))
            else:
                arr["b"][j, k] = -1.0
                arr["sb"][j, k] = 1.0
    t = _abi.MlPanelSoa()
    t.n_panels, t.n_images, t.n_cols, t.in_wake = n, 1, 3, 0
    dp = lambda a: a.ctypes.data_as(_abi.c_double_p)
    t.centr, t.A_g_to_ls, t.vertices_ls, t.n_hat_ls = dp(arr["centr"]), dp(arr["A"]), dp(arr["vls"]), dp(arr["nh"])
    t.b, t.sqrt_b, t.J, t.area, t.vert_g, t.T_mu = dp(arr["b"]), dp(arr["sb"]), dp(arr["J"]), dp(arr["area"]), dp(arr["vg"]), dp(arr["T"])
    t.r = arr["r"].ctypes.data_as(_abi.c_int_p)
    t.i_vert_d = arr["ivd"].ctypes.data_as(_abi.c_int_p)
    t.i_panel_s = arr["ips"].ctypes.data_as(_abi.c_int_p)
    t.has_sources = arr["hs"].ctypes.data_as(_abi.c_ubyte_p)
    t.image_present = arr["ip"].ctypes.data_as(_abi.c_ubyte_p)
    return t, arr


def _flow(supersonic):
    f = _abi.MlFlow()
    M = np.sqrt(2.0) if supersonic else 0.0
    f.M_inf, f.B, f.s = M, 1.0, (-1.0 if supersonic else 1.0)
    f.K_inv = 1.0 / (2 * np.pi) if supersonic else 1.0 / (4 * np.pi)
    c = np.array([1.0, 0.0, 0.0])
    Bm = np.eye(3) - M * M * np.outer(c, c)
    Cm = (1 - M * M) * np.eye(3) + M * M * np.outer(c, c)
    f.c_hat_g[:] = c
    f.B_mat_g[:] = Bm.ravel()
    f.C_mat_g[:] = Cm.ravel()
    f.supersonic, f.mirror_plane = int(supersonic), 0
    return f


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_prototype_known_answers(idx):
    case = CASES[idx]
    sup = case["kind"] != "subsonic"
    q = np.array(case["verts"])  # (4, 2) counter-clockwise
    # orientation: the prototype's vertex order (0:+x+y, 1:-x+y, 2:-x-y, 3:+x-y) is counter-clockwise seen from +z
    table, keep = _triangle_table([[q[0], q[1], q[2]], [q[0], q[2], q[3]]], sup)
    flow = _flow(sup)
    P = np.array(case["P"], dtype=np.float64)
    H111 = hH113 = 0.0
    for j in range(2):
        out = ob.OrcPairOut()
        ob.lib().orc_pair_influence(C.byref(flow), C.byref(table), j, 0, P.ctypes.data_as(_abi.c_double_p), C.byref(out))
        if out.in_dod:
            H111 += out.H111
            hH113 += out.hH113
    sgn = -1.0 if sup else 1.0  # prototype's supersonic hH113 has the opposite sign convention (see fixture note)
    assert abs(H111 - case["H111"]) < 1e-13 * max(1.0, abs(case["H111"])), (H111, case["H111"])
    assert abs(hH113 - sgn * case["hH113"]) < 1e-13 * max(1.0, abs(case["hH113"])), (hH113, case["hH113"])
