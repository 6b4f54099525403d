import sys, json
sys.path.insert(0, "tests")
import numpy as np
import fixtures, oracle_binding as ob
from machline_b200 import host, gpu
doc = json.load(open("tests/golden/offbody_potentials.json"))
ctx = gpu.Context(0)
for c in doc["cases"]:
    case = host.Case(c["input"], base_dir=fixtures.mesh_root())
    ctx.set_case(case); ctx.assemble()
    x, info = ctx.solve(case.solver_opts(), case.BC)
    U = float(np.linalg.norm(c["input"]["flow"]["freestream_velocity"]))
    pts = np.array(c["points"])
    phi_d, phi_s = ctx.potentials_at(case, pts, x)
    A = ctx.get_A()
    A_ref, I_ref = ob.assemble_at_points(case, pts)
    err = np.abs(A - A_ref) / np.abs(A_ref).max(axis=1, keepdims=True).clip(1e-300)
    i, j = np.unravel_index(err.argmax(), err.shape)
    print(c["name"], "max rel-to-rowmax err", err.max(), "at", i, j, pts[i], "A", A[i, j], "ref", A_ref[i, j], "n>1e-12:", (err > 1e-12).sum(), "rows:", np.unique(np.where(err > 1e-12)[0])[:20])
    print("   phi_s err", np.abs(phi_s * U - np.array(c["phi_s"])).max(), "phi_d err vs gold", np.abs(phi_d * U - np.array(c["phi_d"])).max(),
          "phi_d gpu vs oracle", np.abs(A @ x - A_ref @ x).max() * U, "zeros equal", ((A == 0) == (A_ref == 0)).all())
