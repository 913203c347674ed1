"""Stub of libigl's python bindings: only read_obj is functional (text OBJ parser)."""
import numpy as np


def read_obj(path):
    v, vn, fv, fn = [], [], [], []
    with open(path, "r") as fh:
        for line in fh:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                v.append([float(p[1]), float(p[2]), float(p[3])])
            elif p[0] == "vn":
                vn.append([float(p[1]), float(p[2]), float(p[3])])
            elif p[0] == "f":
                vi, ni = [], []
                for tok in p[1:4]:
                    parts = tok.split("/")
                    vi.append(int(parts[0]) - 1)
                    if len(parts) == 3 and parts[2] != "":
                        ni.append(int(parts[2]) - 1)
                fv.append(vi)
                if len(ni) == 3:
                    fn.append(ni)
    v = np.asarray(v, dtype=np.float64).reshape(-1, 3)
    vn = np.asarray(vn, dtype=np.float64).reshape(-1, 3)
    fv = np.asarray(fv, dtype=np.int64).reshape(-1, 3)
    fn = np.asarray(fn, dtype=np.int64).reshape(-1, 3)
    return v, np.zeros((0, 2)), vn, fv, np.zeros((0, 3), dtype=np.int64), fn


def _absent(*a, **k):
    raise NotImplementedError("igl stub: mesh helpers are outside the hot path")


barycenter = doublearea = vertex_triangle_adjacency = per_vertex_normals = _absent
per_face_normals = triangle_triangle_adjacency = _absent
