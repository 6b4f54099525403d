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
    """Writes tests/golden/meshes.npz and study_meshes.npz back out as mesh files (17 significant digits, so the loader
    parses exactly the doubles the reference's files hold) under <tmp>/test/meshes/."""
    from machline_b200 import meshgen
    root = Path(tempfile.mkdtemp(prefix="machline_fixture_"))
    mdir = root / "test" / "meshes"
    meshgen.materialise_npz(GOLDEN / "meshes.npz", mdir)
    meshgen.materialise_npz(GOLDEN / "study_meshes.npz", mdir)
    return str(root)


def study_case(name: str, **over):
    """host.Case of one of the BASELINE configs[1]-[3] study inputs (meshgen.study_input) on the committed study mesh."""
    from machline_b200 import host, meshgen
    return host.Case(meshgen.study_input(name, mesh_dir="test/meshes", **over), base_dir=mesh_root())


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


def oracle_slack(name: str) -> float:
    """Factor on the reference's tolerances for the oracle run of a case (1 unless the fixture documents why not)."""
    case = next(c for c in goldens()["cases"] if c["name"] == name)
    return float(case.get("oracle_slack", 1.0))


def comparison_names():
    return [c["name"] for c in goldens().get("comparisons", [])]


def comparison_cases(name: str):
    """(host.Case A, host.Case B, tolerances) of a reference test that compares two runs with each other."""
    from machline_b200 import host
    doc = goldens()
    cmp_ = next(c for c in doc["comparisons"] if c["name"] == name)
    out = []
    for nm in cmp_["inputs"]:
        inp = copy.deepcopy(doc["inputs"][nm])
        for key, val in cmp_["alter"]:
            d = inp
            ks = key.split(".")
            for k in ks[:-1]:
                d = d.setdefault(k, {})
            d[ks[-1]] = val
        inp.setdefault("output", {})["verbose"] = False
        inp["output"] = {"verbose": False}
        out.append(host.Case(inp, base_dir=mesh_root()))
    return out[0], out[1], cmp_["tol"]


def make_case(name: str):
    from machline_b200 import host
    inp, expect, tol = golden_input(name)
    return host.Case(inp, base_dir=mesh_root()), expect, tol


def check_tuple(res, expect, tol, slack: float = 1.0):
    got = [res.C_p_max, res.C_p_min, float(res.C_F[0]), float(res.C_F[1]), float(res.C_F[2])]
    for g, e, t, lab in zip(got, expect, tol, ["C_p_max", "C_p_min", "Cx", "Cy", "Cz"]):
        assert abs(g - e) < t * slack, f"{lab}: got {g!r}, reference {e!r}, |diff| {abs(g - e):.3e} >= {t * slack:g}"
