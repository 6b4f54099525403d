"""Host setup on several threads (csrc/host/parallel.hpp: STL scan, vertex normals, panel transforms, distributions): every
iteration writes only its own panel / vertex, so the tables handed to the GPU library must not depend on the thread count."""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import ctypes as C, hashlib, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import numpy as np
import fixtures
case = fixtures.study_case({name!r})
h = hashlib.sha256()
b, w = case.body, case.wake
n_rec = b.n_panels * b.n_images
for ptr, n in [(b.centr, 3 * n_rec), (b.A_g_to_ls, 9 * n_rec), (b.vertices_ls, 6 * n_rec), (b.n_hat_ls, 6 * n_rec), (b.T_mu, 9 * n_rec),
               (b.vert_g, 9 * n_rec), (b.J, n_rec), (b.area, b.n_panels)]:
    h.update(np.ctypeslib.as_array(ptr, shape=(n,)).tobytes())
h.update(np.ctypeslib.as_array(b.i_vert_d, shape=(b.n_panels * b.n_cols,)).tobytes())
h.update(np.asarray(case.BC).tobytes()); h.update(np.asarray(case.cp_loc).tobytes()); h.update(np.asarray(case.P).tobytes())
if w.n_panels:
    h.update(np.ctypeslib.as_array(w.centr, shape=(3 * w.n_panels * w.n_images,)).tobytes())
print(case.n_unknown, h.hexdigest())
"""


@pytest.mark.parametrize("name", ["onera_m6", "agard_b"])
def test_tables_do_not_depend_on_the_thread_count(name):
    outs = []
    for threads in ("1", "3", "8"):
        env = dict(os.environ, MLH_THREADS=threads)
        code = WORKER.format(root=str(ROOT), tests=str(ROOT / "tests"), name=name)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1] == outs[2], outs


def test_non_manifold_edges_keep_the_serial_numbering(tmp_path):
    """A mesh with three and seven panels around an edge: the adjacency search (candidates on the threads, checks on compact arrays
    in panel order, a neighbour table that overflows at the seven-panel edge) must number the edges as the reference's serial double
    loop does (surface_mesh.f90:346-520).  The hashes were produced by the single-threaded library of commit ca427f0 (before the
    setup was parallelised), which the present one reproduces."""
    import numpy as np
    from machline_b200 import host, meshgen
    pts, tris = meshgen.icosphere(2)
    pts, tris = np.asarray(pts), np.asarray(tris)
    a, b = tris[7][1], tris[7][2]
    extra = [pts[tris[0][0]] * 1.5 + 0.1, pts[a] * 1.4, pts[a] * 1.6 + 0.05, pts[a] * 1.8 - 0.05, pts[a] * 2.0 + 0.02, pts[a] * 2.2 - 0.03]
    n0 = len(pts)
    P = np.vstack([pts, extra])
    T = np.vstack([tris, [[tris[0][0], tris[0][1], n0]], [[a, b, n0 + 1]], [[b, a, n0 + 2]], [[a, b, n0 + 3]], [[b, a, n0 + 4]], [[a, b, n0 + 5]]])
    meshgen.write_vtk(str(tmp_path / "m.vtk"), P, T)
    case = host.Case(meshgen.sphere_input("m.vtk"), base_dir=str(tmp_path))
    h = hashlib.sha256()
    bd = case.body
    n_rec = bd.n_panels * bd.n_images
    for ptr, n in [(bd.centr, 3 * n_rec), (bd.T_mu, 9 * n_rec), (bd.vert_g, 9 * n_rec)]:
        h.update(np.ctypeslib.as_array(ptr, shape=(n,)).tobytes())
    h.update(np.ctypeslib.as_array(bd.i_vert_d, shape=(bd.n_panels * bd.n_cols,)).tobytes())
    h.update(np.asarray(case.cp_loc).tobytes())
    h.update(np.asarray(case.P).tobytes())
    h.update(np.asarray(case.BC).tobytes())
    assert (case.n_unknown, case.info.n_edges) == (168, 503)
    assert h.hexdigest()[:16] == "68c0870f2ea49f9b"
    case.close()
