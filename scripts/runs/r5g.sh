# device post-processing tests (second pass) + run_case with device_post
mkdir -p gpurun_out/r5g
timeout 300 python -m pytest tests/test_gpu_post.py -q > gpurun_out/r5g/pytest_post.log 2>&1
tail -8 gpurun_out/r5g/pytest_post.log
timeout 200 python - > gpurun_out/r5g/run_case.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, "tests")
import numpy as np
import fixtures
from machline_b200 import solver
for name in ["test_07", "test_13", "test_21"]:
    inp, expect, tol = fixtures.golden_input(name)
    r = solver.run_case(inp, base_dir=fixtures.mesh_root(), device_post=True)
    d = r.device_post
    print(name, "host", r.C_p_max, r.C_p_min, r.C_F, "device", d["C_p_max"], d["C_p_min"], d["C_F"], "max|dCp|", float(np.abs(d["C_p"][next(iter(d["C_p"]))] - r.C_p).max()) if len(d["C_p"]) else None)
PY
cat gpurun_out/r5g/run_case.log | tail -5
