"""Host setup time at scale (no GPU involved): the AGARD-B study mesh refined 1:4 `levels` times, loaded and set
up by the host library with MLH_TIMING laps.   usage: python scripts/host_setup_scale.py LEVELS [LEVELS ...]"""
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("MLH_TIMING", "1")
from machline_b200 import host, meshgen  # noqa: E402

z = np.load(ROOT / "tests" / "golden" / "study_meshes.npz")
for levels in [int(a) for a in sys.argv[1:]] or [2]:
    pts, tris = meshgen.subdivide(z["agard_b_fine.vtk:points"], z["agard_b_fine.vtk:triangles"], levels)
    tmp = tempfile.mkdtemp(prefix="machline_setup_")
    meshgen.write_vtk(f"{tmp}/agard_b_fine.vtk", pts, tris)
    for rep in range(2):
        t = time.perf_counter()
        case = host.Case(meshgen.study_input("agard_b"), base_dir=tmp)
        dt = time.perf_counter() - t
        print(f"levels {levels}: {len(tris)} panels x 2 images, N = {case.n_unknown}: host setup {dt:.3f} s "
              f"({os.cpu_count()} cpus, MLH_THREADS={os.environ.get('MLH_THREADS', 'auto')})", flush=True)
        case.close()
