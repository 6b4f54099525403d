"""Deterministic synthetic surface meshes for benchmarks and scale tests (no RNG, no network).

The reference's benchmark inputs live in its `studies/` tree, which does not travel to the GPU box;
these generators produce meshes of the same families and sizes:

* ``icosphere(level)``       -- unit sphere, 20*4**level panels (config 1 / config 5 family)
* ``swept_wing_half(...)``   -- ONERA-M6-like tapered swept half wing with a sharp trailing edge, root
                                on the xz mirror plane, rounded tip (config 2 family: mirrored, wake)
* ``write_vtk(path, ...)``   -- ASCII VTK v3 POLYDATA, the format src/vtk.f90:480-553 reads
"""
from __future__ import annotations

from pathlib import Path

import numpy as np


def write_vtk(path, points: np.ndarray, triangles: np.ndarray) -> str:
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w") as f:
        f.write("# vtk DataFile Version 3.0\nmachline_b200 synthetic mesh\nASCII\nDATASET POLYDATA\n")
        f.write(f"POINTS {len(points)} float\n")
        np.savetxt(f, points, fmt="%.17g")
        f.write(f"POLYGONS {len(triangles)} {4 * len(triangles)}\n")
        np.savetxt(f, np.column_stack([np.full(len(triangles), 3), triangles]), fmt="%d")
    return str(path)


def icosphere(level: int, radius: float = 1.0):
    """Icosahedron subdivided `level` times and projected on the sphere; outward-oriented triangles."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1)[:, None]
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                  [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    verts = [tuple(p) for p in v]
    for _ in range(level):
        cache = {}
        new_f = []

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = (np.array(verts[a]) + np.array(verts[b])) * 0.5
                m /= np.linalg.norm(m)
                cache[key] = len(verts)
                verts.append(tuple(m))
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new_f += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        f = np.array(new_f, dtype=np.int64)
    return np.array(verts, dtype=np.float64) * radius, f.astype(np.int32)


def _naca4_thickness(x: np.ndarray, t: float) -> np.ndarray:
    """NACA 00xx half-thickness with the closed-trailing-edge coefficient (-0.1036)."""
    return 5.0 * t * (0.2969 * np.sqrt(x) - 0.1260 * x - 0.3516 * x ** 2 + 0.2843 * x ** 3 - 0.1036 * x ** 4)


def swept_wing_half(n_chord: int = 80, n_span: int = 45, root_chord: float = 0.8059, tip_chord: float = 0.4535,
                    semi_span: float = 1.1963, le_sweep_deg: float = 30.0, thickness: float = 0.10, n_cap: int = 4,
                    te_cluster: float = 0.5):
    """Right half (y >= 0) of a swept tapered wing (ONERA-M6 planform); root ring lies on y = 0 (mirror about xz).
    The tip is closed by half a body of revolution (as on the M6): at chord station i the upper and lower surface
    points are joined by a semicircle of radius = local half-thickness in the y-z plane, `n_cap` arcs each.
    Panels: 4*n_chord*n_span on the surface + 2*n_cap*(n_chord-2) + 2*n_cap on the tip cap.
    Chordwise spacing: cosine clustering at the leading edge blended (te_cluster in [0,1]) towards uniform at
    the trailing edge, so trailing-edge panels stay thicker than the control-point offset."""
    nc, ns = n_chord, n_span
    u = np.linspace(0.0, 1.0, nc + 1)
    x_cos = 0.5 * (1.0 - np.cos(np.pi * u))       # clusters at both ends
    x_le = 1.0 - np.cos(0.5 * np.pi * u)            # clusters at the leading edge only
    xc = te_cluster * x_cos + (1.0 - te_cluster) * x_le
    xc[0], xc[-1] = 0.0, 1.0                        # 0 (LE) .. 1 (TE)
    zt = _naca4_thickness(xc, thickness)
    zt[-1] = 0.0                                    # sharp trailing edge
    # ring: TE -> upper surface -> LE -> lower surface -> (back to TE, not repeated): 2*nc points
    ring_x = np.concatenate([xc[::-1], xc[1:-1]])
    ring_z = np.concatenate([zt[::-1], -zt[1:-1]])
    nr = 2 * nc
    eta = np.sin(0.5 * np.pi * np.linspace(0.0, 1.0, ns + 1))   # clusters towards the tip
    tan_le = np.tan(np.radians(le_sweep_deg))
    pts = []
    for e in eta:
        y = e * semi_span
        c = root_chord + (tip_chord - root_chord) * e
        x0 = y * tan_le
        pts.append(np.column_stack([x0 + c * ring_x, np.full(nr, y), c * ring_z]))
    pts = list(np.concatenate(pts))
    tris = []
    for k in range(ns):
        a0, b0 = k * nr, (k + 1) * nr
        for i in range(nr):
            i1 = (i + 1) % nr
            p00, p01, p10, p11 = a0 + i, a0 + i1, b0 + i, b0 + i1
            # ring runs TE->upper->LE->lower (clockwise seen from +y), span index grows with y:
            # (p00, p10, p11) and (p00, p11, p01) have outward normals
            tris.append([p00, p10, p11])
            tris.append([p00, p11, p01])
    # rounded tip cap.  Chord station j = 0 (TE) .. nc (LE); upper point index t0 + j, lower point t0 + (nr - j) % nr.
    t0 = ns * nr
    c_tip = tip_chord
    x_tip0 = semi_span * tan_le
    m = max(2, n_cap)
    cap = {}   # (j, a) -> point index, a = 0 (upper) .. m (lower)
    for j in range(nc + 1):
        up, lo = t0 + j, t0 + (nr - j) % nr
        cap[(j, 0)], cap[(j, m)] = up, lo
        if j == 0 or j == nc:
            for a in range(1, m):
                cap[(j, a)] = up            # zero radius: TE and LE points
            continue
        xs = xc[::-1][j]                    # chord fraction of station j (TE -> LE)
        r = c_tip * zt[::-1][j]
        for a in range(1, m):
            th = np.pi * a / m
            cap[(j, a)] = len(pts)
            pts.append(np.array([x_tip0 + c_tip * xs, semi_span + r * np.sin(th), r * np.cos(th)]))
    for j in range(nc):
        for a in range(m):
            q00, q01, q10, q11 = cap[(j, a)], cap[(j, a + 1)], cap[(j + 1, a)], cap[(j + 1, a + 1)]
            # upper surface runs TE->LE with +z normals as (p00, p10, p11) above; continuing over the tip, the
            # outward orientation is (q00, q01, q11), (q00, q11, q10)
            for tri in ([q00, q01, q11], [q00, q11, q10]):
                if len(set(tri)) == 3:
                    tris.append(tri)
    return np.array(pts), np.array(tris, dtype=np.int32)


def check_outward(points: np.ndarray, triangles: np.ndarray, closed_by_mirror_axis: int | None = None) -> float:
    """Signed volume by the divergence theorem (positive for outward-oriented closed surfaces)."""
    a, b, c = points[triangles[:, 0]], points[triangles[:, 1]], points[triangles[:, 2]]
    return float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)


def wing_input(mesh_file: str, mach: float = 0.5, alpha_deg: float = 3.06, matrix_solver: str = "GMRES",
               formulation: str = "dirichlet-morino") -> dict:
    """MachLine input of the config-2 family: mirrored half wing, automatic wake, Prandtl-Glauert M."""
    a = np.radians(alpha_deg)
    return {
        "flow": {"freestream_velocity": [float(np.cos(a)), 0.0, float(np.sin(a))], "freestream_mach_number": mach},
        "geometry": {"file": mesh_file, "mirror_about": "xz", "spanwise_axis": "+y",
                     # 90 deg (the default) would also flag the square tip-cap edges; only the sharp trailing edge sheds
                     "wake_model": {"wake_present": True, "append_wake": True, "wake_shedding_angle": 120.0},
                     "reference": {"area": 1.0}},
        "solver": {"formulation": formulation, "matrix_solver": matrix_solver, "control_point_offset": 1.1e-5},
        "post_processing": {},
        "output": {"verbose": False},
    }


def sphere_input(mesh_file: str, matrix_solver: str = "GMRES") -> dict:
    """The reference's sphere case (test/input_files/sphere_input.json) on a synthetic icosphere."""
    return {
        "flow": {"freestream_velocity": [0.0, 0.0, 10.0]},
        "geometry": {"file": mesh_file, "wake_model": {"wake_present": False}},
        "solver": {"formulation": "dirichlet-morino", "matrix_solver": matrix_solver, "control_point_offset": 1.1e-5},
        "post_processing": {},
        "output": {"verbose": False},
    }


def sears_haack(n_ax: int = 80, n_theta: int = 30, length: float = 0.6096, rmax_over_length: float = 0.037879):
    """Closed Sears-Haack body of revolution along +x, r(x) = rmax (4 x/L (1 - x/L))^(3/4), pointed nose and tail
    (BASELINE configs[2]/[4] family; the shape of the reference's studies/sears_haack/gen_SH_geometry.py, rebuilt
    here with array operations): n_ax axial stations including both apexes, n_theta points per ring, triangle fans at
    the ends and two triangles per quad in between, oriented outwards.  2 n_theta (n_ax - 2) panels."""
    rmax = rmax_over_length * length
    xs = np.linspace(0.0, length, n_ax)
    r = rmax * (4.0 * (xs / length) * (1.0 - xs / length)) ** 0.75
    th = 2.0 * np.pi * np.arange(n_theta) / n_theta
    rings = np.stack([np.repeat(xs[1:-1, None], n_theta, axis=1),
                      r[1:-1, None] * np.sin(th)[None, :],
                      r[1:-1, None] * np.cos(th)[None, :]], axis=2).reshape(-1, 3)
    pts = np.vstack([[0.0, 0.0, 0.0], rings, [length, 0.0, 0.0]])
    n_rings = n_ax - 2
    ring = lambda i: 1 + i * n_theta + np.arange(n_theta)      # noqa: E731  vertex ids of ring i
    nxt = lambda a: np.roll(a, -1)                              # noqa: E731
    tris = [np.stack([np.zeros(n_theta, dtype=np.int64), ring(0), nxt(ring(0))], axis=1)]
    for i in range(n_rings - 1):
        a, b = ring(i), ring(i + 1)
        tris.append(np.stack([a, b, nxt(b)], axis=1))
        tris.append(np.stack([a, nxt(b), nxt(a)], axis=1))
    last = len(pts) - 1
    tris.append(np.stack([ring(n_rings - 1), np.full(n_theta, last), nxt(ring(n_rings - 1))], axis=1))
    tris = np.vstack(tris).astype(np.int32)
    if check_outward(pts, tris) < 0:
        tris = tris[:, ::-1].copy()
    return pts, tris


def sears_haack_input(mesh_file: str, mach: float = 2.0, matrix_solver: str = "GMRES") -> dict:
    """Supersonic slender body, no wake, source-free formulation, the study's control-point offset
    (studies/sears_haack/sears_haack_input.json); every pair goes through the domain-of-dependence test."""
    return {
        "flow": {"freestream_velocity": [1.0, 0.0, 0.0], "gamma": 1.4, "freestream_mach_number": mach},
        "geometry": {"file": mesh_file, "spanwise_axis": "+y", "wake_model": {"wake_present": False},
                     "reference": {"area": 1.675e-3}},
        "solver": {"formulation": "dirichlet-source-free", "matrix_solver": matrix_solver, "control_point_offset": 1.1e-8},
        "post_processing": {},
        "output": {"verbose": False},
    }


# ---- committed reference meshes (tests/golden/*.npz) -> mesh files ----------------------------------------------------
def materialise_npz(npz_path, out_dir, only=None) -> list[str]:
    """Writes every mesh of an archive made by tests/golden/make_fixtures.py back out as the file type its name says
    (ASCII VTK v3 / STL / Cart3D .tri) with 17 significant digits, so that the host loader parses exactly the doubles the
    reference's own files hold, in the same vertex and panel order."""
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    z = np.load(npz_path)
    names = sorted({k.split(":")[0] for k in z.files})
    if only is not None:
        names = [n for n in names if n in set(only)]
    for name in names:
        path = out_dir / name
        if name.endswith(".vtk"):
            pts, tris = z[f"{name}:points"], z[f"{name}:triangles"]
            with open(path, "w") as f:
                f.write("# vtk DataFile Version 3.0\nfixture\nASCII\nDATASET POLYDATA\n")
                f.write(f"POINTS {len(pts)} float\n")
                f.write("\n".join(f"{float(p[0])!r} {float(p[1])!r} {float(p[2])!r}" for p in pts) + "\n")
                f.write(f"POLYGONS {len(tris)} {4 * len(tris)}\n")
                f.write("\n".join(f"3 {t[0]} {t[1]} {t[2]}" for t in tris) + "\n")
        elif name.endswith(".tri"):
            pts, tris = z[f"{name}:points"], z[f"{name}:triangles"]
            with open(path, "w") as f:
                f.write(f"{len(pts)} {len(tris)}\n")
                f.write("\n".join(f"{float(p[0])!r} {float(p[1])!r} {float(p[2])!r}" for p in pts) + "\n")
                f.write("\n".join(f"{t[0] + 1} {t[1] + 1} {t[2] + 1}" for t in tris) + "\n")
        elif name.endswith(".stl"):
            fv = z[f"{name}:facet_vertices"]
            with open(path, "w") as f:
                f.write("solid\n")
                for k in range(0, len(fv), 3):
                    f.write(" facet normal 0 0 0\n   outer loop\n")
                    for p in fv[k:k + 3]:
                        f.write(f"     vertex {float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n")
                    f.write("   endloop\n endfacet\n")
                f.write("endsolid\n")
    return names


def subdivide(pts: np.ndarray, tris: np.ndarray, levels: int = 1):
    """Uniform 1:4 refinement of a triangle mesh (edge midpoints, flat: the geometry is not smoothed), `levels` times.  Keeps
    the orientation of every panel and produces one shared vertex per edge, so sharp edges (wake-shedding trailing edges) and
    vertices on a mirror plane stay what they were.  Used to grow the reference's study meshes to the 50k-100k panel class of
    BASELINE configs[4]."""
    pts = np.asarray(pts, dtype=np.float64)
    tris = np.asarray(tris, dtype=np.int64)
    for _ in range(levels):
        n = len(pts)
        e = np.sort(np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]]), axis=1)
        key = e[:, 0] * n + e[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // n, uniq % n
        mid = 0.5 * (pts[a] + pts[b])
        m = n + inv.reshape(3, -1).T                     # midpoint vertex of edges (01, 12, 20) of every triangle
        t0, t1, t2 = tris[:, 0], tris[:, 1], tris[:, 2]
        m01, m12, m20 = m[:, 0], m[:, 1], m[:, 2]
        tris = np.concatenate([np.stack([t0, m01, m20], 1), np.stack([m01, t1, m12], 1), np.stack([m20, m12, t2], 1),
                               np.stack([m01, m12, m20], 1)])
        pts = np.concatenate([pts, mid])
    return pts, tris.astype(np.int32)


def study_input(name: str, mesh_dir: str = "", matrix_solver: str = "GMRES", **over) -> dict:
    """Inputs of BASELINE.json configs[1]-[3] on the reference's own study meshes (SURVEY 8(d) "Concrete inputs";
    studies/subsonic_onera_m6_wing/M6_input.json, studies/supersonic_cone/cone_input.json,
    studies/sears_haack/run_study.py:24-55, studies/supersonic_agard_b_wing_body/run_study.py:23-56).  The study scripts
    say "morino" / "source-free"; today's reference spells them dirichlet-morino / dirichlet-source-free
    (src/panel_solver.f90:135).  Lower-order singularities."""
    pre = (mesh_dir.rstrip("/") + "/") if mesh_dir else ""
    a = np.deg2rad(3.06)
    cases = {
        # configs[1]: ONERA M6, Prandtl-Glauert M = 0.5, alpha = 3.06 deg, mirrored about xz, automatic wake
        "onera_m6": {"flow": {"freestream_velocity": [float(np.cos(a)), 0.0, float(np.sin(a))], "freestream_mach_number": 0.5},
                     "geometry": {"file": pre + "M6_onera_fine.stl", "mirror_about": "xz", "spanwise_axis": "+y",
                                  "reference": {"area": 0.7532}},
                     "solver": {"formulation": "dirichlet-morino"}},
        # configs[2]: 10 degree cone, M = 1.5, mirrored about xy, no wake
        "cone": {"flow": {"freestream_velocity": [-1.0, 0.0, 0.0], "gamma": 1.4, "freestream_mach_number": 1.5},
                 "geometry": {"file": pre + "cone_10_deg_fine.vtk", "spanwise_axis": "+z", "mirror_about": "xy",
                              "max_continuity_angle": 45.0, "wake_model": {"append_wake": False}, "reference": {"area": 4.0}},
                 "solver": {"formulation": "dirichlet-morino"}},
        # configs[2]: Sears-Haack body, M = 2, source-free, no wake
        "sears_haack": {"flow": {"freestream_velocity": [1.0, 0.0, 0.0], "gamma": 1.4, "freestream_mach_number": 2.0},
                        "geometry": {"file": pre + "SH_160_60.tri", "spanwise_axis": "+y", "wake_model": {"wake_present": False},
                                     "reference": {"area": 1.675e-3}},
                        "solver": {"formulation": "dirichlet-source-free", "control_point_offset": 1.1e-8}},
        # configs[3]: AGARD-B wing-body, M = 1.6, mirrored about yz, supersonic wake
        "agard_b": {"flow": {"freestream_velocity": [0.0, 0.0, -1.0], "gamma": 1.4, "freestream_mach_number": 1.6},
                    "geometry": {"file": pre + "agard_b_fine.vtk", "spanwise_axis": "+x", "mirror_about": "yz", "wake_model": {},
                                 "reference": {"area": 1.0}},
                    "solver": {"formulation": "dirichlet-morino"}},
        "agard_b_coarse": {"flow": {"freestream_velocity": [0.0, 0.0, -1.0], "gamma": 1.4, "freestream_mach_number": 1.6},
                           "geometry": {"file": pre + "agard_b_coarse.vtk", "spanwise_axis": "+x", "mirror_about": "yz",
                                        "wake_model": {}, "reference": {"area": 1.0}},
                           "solver": {"formulation": "dirichlet-morino"}},
    }
    inp = cases[name]
    inp["geometry"]["singularity_order"] = "lower"
    inp["solver"]["matrix_solver"] = matrix_solver
    inp["post_processing"] = {"pressure_rules": {"isentropic": True}}
    inp["output"] = {"verbose": False}
    for key, val in over.items():
        d = inp
        ks = key.split(".")
        for k in ks[:-1]:
            d = d.setdefault(k, {})
        d[ks[-1]] = val
    return inp
