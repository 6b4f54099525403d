"""CPU: the DEVICE pair evaluation (machline_b200/csrc/gpu/pair_influence.cuh) compiled for the host
(tests/device_math/, test infrastructure) against the oracle, pair by pair, on the reference's test meshes:
every control point x every panel image, including the near field (a control point sits 1e-5 under its own
vertex's panels).  Subsonic code is built with FMA contraction as aic_sub.cu is, supersonic without as aic_sup.cu.
What the GPU adds on top of this is only CUDA's libm (log/atan2/sqrt) instead of glibc's."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import fixtures
import oracle_binding as ob
from machline_b200 import _abi

DM_DIR = Path(__file__).resolve().parent / "device_math"


@pytest.fixture(scope="module")
def dm():
    subprocess.run(["make", "-C", str(DM_DIR), "-s"], check=True)
    L = C.CDLL(str(DM_DIR / "libdevice_math_host.so"))
    args = [C.POINTER(_abi.MlFlow), C.POINTER(_abi.MlPanelSoa), C.c_int, _abi.c_double_p, _abi.c_double_p, _abi.c_double_p,
            _abi.c_ubyte_p]
    L.dm_batch_sub.argtypes = args
    L.dm_batch_sup.argtypes = args
    return L


def _compare(dm, case, table, pts):
    n_rec = table.n_panels * table.n_images
    n = len(pts)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    d_ref, d_abs, s_ref = np.zeros((n, n_rec, 3)), np.zeros((n, n_rec, 3)), np.zeros((n, n_rec))
    in_ref = np.zeros((n, n_rec), dtype=np.uint8)
    L = ob.lib()
    L.orc_pair_batch.argtypes = [C.POINTER(_abi.MlFlow), C.POINTER(_abi.MlPanelSoa), C.c_int, _abi.c_double_p, _abi.c_double_p,
                                 _abi.c_double_p, _abi.c_double_p, _abi.c_ubyte_p]
    dp = lambda a: a.ctypes.data_as(_abi.c_double_p)
    up = lambda a: a.ctypes.data_as(_abi.c_ubyte_p)
    L.orc_pair_batch(C.byref(case.flow), C.byref(table), n, dp(pts), dp(d_ref), dp(d_abs), dp(s_ref), up(in_ref))
    d, s, inn = np.zeros_like(d_ref), np.zeros_like(s_ref), np.zeros_like(in_ref)
    fn = dm.dm_batch_sup if case.flow.supersonic else dm.dm_batch_sub
    fn(C.byref(case.flow), C.byref(table), n, dp(pts), dp(d), dp(s), up(inn))
    return d, s, inn, d_ref, d_abs, s_ref, in_ref


@pytest.mark.parametrize("name", ["test_08", "test_05", "test_12", "test_13", "test_15"])
def test_device_pair_math_matches_oracle(dm, name):
    case, _, _ = fixtures.make_case(name)
    pts = case.cp_loc[: min(case.n_cp, 700)]
    for table in [case.body] + ([case.wake] if case.wake.n_panels > 0 else []):
        d, s, inn, d_ref, d_abs, s_ref, in_ref = _compare(dm, case, table, pts)
        if not case.flow.supersonic:
            in_ref = np.ones_like(in_ref)   # the subsonic device code has no DoD test; area > 0 is checked when packing
        else:
            assert (inn == in_ref).all()
        # doublet coefficients: relative to the summed |terms| of the pair (see tests/test_gpu_parity.py)
        err = np.abs(d - d_ref) / np.where(d_abs > 0, d_abs, 1.0)
        assert err.max() < 1e-14, (name, err.max(), np.unravel_index(err.argmax(), err.shape))
        # source coefficient: H111 = sum a F111 - h hH113 against its own terms' scale, approximated by the panel's largest |phi_s|
        # (the oracle reports phi_s only for panels that carry sources; the kernel applies that flag outside the pair function)
        has = s_ref != 0
        if has.any():
            serr = np.abs(s - s_ref)[has] / np.abs(s_ref).max()
            assert serr.max() < 1e-14, (name, serr.max())
    case.close()


@pytest.mark.parametrize("name", ["test_04", "test_02", "test_16", "test_17"])
def test_device_pair_math_higher_order_matches_oracle(dm, name):
    """Quadratic-doublet / linear-source panels (geometry.singularity_order = "higher"): the six strength-space doublet
    influences and the summed source influence of every (control point, panel image) pair, device arithmetic against the oracle."""
    case, _, _ = fixtures.make_case(name)
    table = case.body
    assert table.order2 == 1 and table.n_cols == 6
    pts = np.ascontiguousarray(case.cp_loc[: min(case.n_cp, 500)], dtype=np.float64)
    n, n_rec = len(pts), table.n_panels * table.n_images
    L = ob.lib()
    args = [C.POINTER(_abi.MlFlow), C.POINTER(_abi.MlPanelSoa), C.c_int, _abi.c_double_p, _abi.c_double_p, _abi.c_double_p,
            _abi.c_double_p, _abi.c_ubyte_p]
    L.orc_pair_batch_ho.argtypes = args
    dp = lambda a: a.ctypes.data_as(_abi.c_double_p)
    up = lambda a: a.ctypes.data_as(_abi.c_ubyte_p)
    d_ref, d_abs, s_ref = np.zeros((n, n_rec, 6)), np.zeros((n, n_rec, 6)), np.zeros((n, n_rec))
    in_ref = np.zeros((n, n_rec), dtype=np.uint8)
    L.orc_pair_batch_ho(C.byref(case.flow), C.byref(table), n, dp(pts), dp(d_ref), dp(d_abs), dp(s_ref), up(in_ref))
    d, s, inn = np.zeros_like(d_ref), np.zeros_like(s_ref), np.zeros_like(in_ref)
    fn = dm.dm_batch_sup_ho if case.flow.supersonic else dm.dm_batch_sub_ho
    fn.argtypes = [C.POINTER(_abi.MlFlow), C.POINTER(_abi.MlPanelSoa), C.c_int, _abi.c_double_p, _abi.c_double_p, _abi.c_double_p,
                   _abi.c_ubyte_p]
    fn(C.byref(case.flow), C.byref(table), n, dp(pts), dp(d), dp(s), up(inn))
    if case.flow.supersonic:
        assert (inn == in_ref).all()
    M_dim = np.ctypeslib.as_array(table.M_dim, shape=(table.n_panels,))
    assert M_dim.max() == 6 and M_dim.min() >= 3
    # entries beyond a panel's M_dim are exact zeros (zero-padded T_mu)
    pad = np.tile(np.arange(6)[None, :] >= M_dim[:, None], (table.n_images, 1))
    assert (d[:, pad] == 0.).all()
    err = np.abs(d - d_ref) / np.where(d_abs > 0, d_abs, 1.0)
    assert err.max() < 2e-14, (name, err.max(), np.unravel_index(err.argmax(), err.shape))
    has = s_ref != 0
    if has.any():
        serr = np.abs(s - s_ref)[has] / np.abs(s_ref).max()
        assert serr.max() < 1e-13, (name, serr.max())
    case.close()
