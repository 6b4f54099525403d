"""Test helpers: materialise the committed mesh fixtures and build golden-case inputs."""
from __future__ import annotations

import copy
import json
import os
import tempfile
from functools import lru_cache
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


@lru_cache(maxsize=1)
def _mesh_dir() -> str:
    """Writes tests/golden/meshes.npz back out as ASCII VTK v3 / STL files (17 significant digits, so the
    loader parses exactly the doubles the reference's files hold) under <tmp>/test/meshes/."""
    root = Path(tempfile.mkdtemp(prefix="machline_fixture_"))
    mdir = root / "test" / "meshes"
    mdir.mkdir(parents=True)
    z = np.load(GOLDEN / "meshes.npz")
    names = sorted({k.split(":")[0] for k in z.files})
    for name in names:
        if name.endswith(".vtk"):
            pts, tris = z[f"{name}:points"], z[f"{name}:triangles"]
            with open(mdir / name, "w") as f:
                f.write("# vtk DataFile Version 3.0\nfixture\nASCII\nDATASET POLYDATA\n")
                f.write(f"POINTS {len(pts)} float\n")
                for p in pts:
                    f.write(f"{float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n")
                f.write(f"POLYGONS {len(tris)} {4 * len(tris)}\n")
                for t in tris:
                    f.write(f"3 {t[0]} {t[1]} {t[2]}\n")
        elif name.endswith(".stl"):
            fv = z[f"{name}:facet_vertices"]
            with open(mdir / name, "w") as f:
                f.write("solid\n")
                for k in range(0, len(fv), 3):
                    f.write(" facet normal 0 0 0\n   outer loop\n")
                    for p in fv[k:k + 3]:
                        f.write(f"     vertex {float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n")
                    f.write("   endloop\n endfacet\n")
                f.write("endsolid\n")
    return str(root)


def mesh_root() -> str:
    return _mesh_dir()


@lru_cache(maxsize=1)
def goldens() -> dict:
    return json.loads((GOLDEN / "reference_goldens.json").read_text())


def golden_case_names():
    return [c["name"] for c in goldens()["cases"]]


def golden_input(name: str):
    """(input dict with the test's alterations applied, expected 5-tuple, tolerances)."""
    doc = goldens()
    case = next(c for c in doc["cases"] if c["name"] == name)
    inp = copy.deepcopy(doc["inputs"][case["input"]])
    for key, val in case["alter"]:
        d = inp
        ks = key.split(".")
        for k in ks[:-1]:
            d = d.setdefault(k, {})
        d[ks[-1]] = val
    inp.setdefault("output", {})["verbose"] = False
    inp["output"].pop("report_file", None)
    return inp, case["expect"], case["tol"]


def make_case(name: str):
    from machline_b200 import host
    inp, expect, tol = golden_input(name)
    return host.Case(inp, base_dir=mesh_root()), expect, tol


def check_tuple(res, expect, tol, slack: float = 1.0):
    got = [res.C_p_max, res.C_p_min, float(res.C_F[0]), float(res.C_F[1]), float(res.C_F[2])]
    for g, e, t, lab in zip(got, expect, tol, ["C_p_max", "C_p_min", "Cx", "Cy", "Cz"]):
        assert abs(g - e) < t * slack, f"{lab}: got {g!r}, reference {e!r}, |diff| {abs(g - e):.3e} >= {t * slack:g}"
