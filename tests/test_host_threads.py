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
