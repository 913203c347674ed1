"""`Pointcloud` container and OBJ I/O with the interface of the reference's Pointcloud/Modules/Object.py
(:43-162).  The OBJ reader is self-contained (the reference goes through libigl)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from . import _io


def _default_device():
    return "cuda" if torch.cuda.is_available() else "cpu"


def read_obj(file_path: str):
    """v [N,3] f64, vn [M,3] f64, face vertex ids [F,3] i64, face normal ids [Fn,3] i64 (0-based, polygons fan-triangulated).
    Parsed by libngpd_io.so (threads over pieces of the mapped file); the reference goes through igl.read_obj (Object.py:80)."""
    return _io.read_obj(file_path)


def sample_points(pos: torch.Tensor, face: torch.Tensor, num: int):
    """torch_geometric.transforms.SamplePoints(num, include_normals=True) on (pos [V,3], face [3,F]): (points, normals)."""
    pos_max = pos.abs().max()
    pos = pos / pos_max
    v0, v1, v2 = pos[face[0]], pos[face[1]], pos[face[2]]
    area = torch.linalg.cross(v1 - v0, v2 - v0, dim=1).norm(p=2, dim=1).abs() / 2
    prob = area / area.sum()
    sample = torch.multinomial(prob, num, replacement=True)
    face = face[:, sample]
    frac = torch.rand(num, 2, device=pos.device)
    mask = frac.sum(dim=-1) > 1
    frac[mask] = 1 - frac[mask]
    vec1 = pos[face[1]] - pos[face[0]]
    vec2 = pos[face[2]] - pos[face[0]]
    normal = torch.nn.functional.normalize(torch.linalg.cross(vec1, vec2, dim=1), p=2)
    pos_sampled = pos[face[0]]
    pos_sampled = pos_sampled + frac[:, :1] * vec1
    pos_sampled = pos_sampled + frac[:, 1:] * vec2
    return pos_sampled * pos_max, normal


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def read_ply_vertices(file_path: str) -> np.ndarray:
    """x, y, z of the `vertex` element of a PLY file as [N,3] float64 (ascii, binary_little_endian or binary_big_endian;
    other properties and elements are skipped).  The reference reads the same through Open3D (Object.py:128)."""
    with open(file_path, "rb") as fh:
        if fh.readline().strip() != b"ply":
            raise ValueError(f"{file_path}: not a PLY file")
        fmt, elements = None, []                       # elements: [name, count, [(property name, dtype or ("list", count type, item type))]]
        while True:
            line = fh.readline()
            if not line:
                raise ValueError(f"{file_path}: header without end_header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append([tok[1], int(tok[2]), []])
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1][2].append((tok[4], ("list", _PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]])))
                else:
                    elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{file_path}: unknown PLY format {fmt!r}")
        for name, count, props in elements:
            scalar = all(not isinstance(t, tuple) for _, t in props)
            if name != "vertex":
                # skip the element: only possible without parsing when every property has a fixed size
                if fmt == "ascii":
                    for _ in range(count):
                        fh.readline()
                elif scalar:
                    fh.seek(count * sum(np.dtype(t).itemsize for _, t in props), 1)
                else:
                    raise ValueError(f"{file_path}: a list element precedes the vertices")
                continue
            if not scalar:
                raise ValueError(f"{file_path}: list property inside the vertex element")
            names = [n for n, _ in props]
            if not all(a in names for a in ("x", "y", "z")):
                raise ValueError(f"{file_path}: vertex element without x, y, z")
            if fmt == "ascii":
                cols = [names.index(a) for a in ("x", "y", "z")]
                rows = _io.read_table(file_path, max(cols) + 1, offset=fh.tell(), rows=count) if count else np.zeros((0, max(cols) + 1))
                return np.ascontiguousarray(rows[:, cols]).reshape(-1, 3)
            end = "<" if fmt == "binary_little_endian" else ">"
            rec = np.dtype([(n, end + t) for n, t in props])
            data = np.frombuffer(fh.read(count * rec.itemsize), dtype=rec, count=count)
            return np.stack([data[a].astype(np.float64) for a in ("x", "y", "z")], axis=1).reshape(-1, 3)
    raise ValueError(f"{file_path}: no vertex element")


class Pointcloud:
    def __init__(self, v: torch.Tensor, n: torch.Tensor = None) -> None:
        assert v.is_floating_point()
        assert v.dim() == 2
        assert v.size(1) == 3
        if n is not None:
            assert n.is_floating_point()
            assert n.dim() == 2
            assert n.size(1) == 3
            assert v.size(0) == n.size(0)
        self.v = v
        self.n = n
        self.file_path = None

    def hasNormals(self) -> bool:
        return self.n is not None

    def hasFilePath(self) -> bool:
        return self.file_path is not None

    def saveObj(self, file_path: str) -> None:
        """Object.py:58-69: `v x y z` lines then `vn` lines; refuses to overwrite (mode "x")."""
        _io.write_obj(file_path, self.v.detach().cpu().numpy(), None if self.n is None else self.n.detach().cpu().numpy(), exclusive=True)
        self.file_path = file_path

    @classmethod
    def loadObj(cls, file_path: str, device=None) -> "Pointcloud":
        """Object.py:72-89.  Vertex normals are taken from `vn` records: accumulated over faces when the faces
        reference them, used directly when there is one per vertex, otherwise the cloud has no normals."""
        path = Path(file_path)
        assert path.is_file()
        assert path.suffix == ".obj"
        device = device if device is not None else _default_device()
        v, vn, fv, fn = read_obj(file_path)
        vt = torch.tensor(v, dtype=torch.float, device=device)
        if len(vn) > 0 and len(fn) > 0:
            nt = torch.tensor(vn, dtype=torch.float, device=device)
            acc = torch.zeros_like(vt)
            acc.index_add_(0, torch.tensor(fv, device=device).view(-1), nt[torch.tensor(fn, device=device)].view(-1, 3))
            pc = cls(vt, torch.nn.functional.normalize(acc, dim=-1))
        elif len(vn) > 0 and len(vn) == len(v):
            pc = cls(vt, torch.tensor(vn, dtype=torch.float, device=device))
        else:
            pc = cls(vt)
        pc.file_path = file_path
        return pc

    @classmethod
    def loadPly(cls, file_path: str, device=None) -> "Pointcloud":
        """Object.py:119-132: positions of a .ply point cloud (no normals, like the reference); self-contained reader
        instead of Open3D."""
        path = Path(file_path)
        assert path.is_file()
        assert path.suffix == ".ply"
        device = device if device is not None else _default_device()
        pc = cls(torch.tensor(read_ply_vertices(file_path), dtype=torch.float, device=device))
        pc.file_path = file_path
        return pc

    @classmethod
    def sampleObj(cls, file_path: str, num_points: int, device=None) -> "Pointcloud":
        """Object.py:135-156: load an OBJ mesh and sample `num_points` points from its surface with their face normals.
        The reference delegates to torch_geometric.transforms.SamplePoints(num, include_normals=True); its published
        algorithm is restated here with torch ops on `device`: faces drawn with probability
        proportional to area (multinomial with replacement, on coordinates normalised by max |pos|), a uniform point in
        each by two folded barycentric fractions, the face's unit normal."""
        path = Path(file_path)
        assert path.is_file()
        assert path.suffix == ".obj"
        device = device if device is not None else _default_device()
        v, _, fv, _ = read_obj(file_path)
        pos = torch.tensor(v, dtype=torch.float, device=device)
        face = torch.tensor(fv, dtype=torch.long, device=device).T
        assert pos.size(1) == 3
        assert face.size(0) == 3
        pc = cls(*sample_points(pos, face, int(num_points)))
        pc.file_path = file_path
        return pc

    @classmethod
    def loadXYZ(cls, file_path: str, device=None) -> "Pointcloud":
        """Object.py:92-117 (the reference reads into `v_list` but converts an undefined `v`; fixed here)."""
        path = Path(file_path)
        assert path.is_file()
        assert path.suffix in (".xyz", ".clean_xyz")
        device = device if device is not None else _default_device()
        pts = _io.read_table(file_path, 3)
        pc = cls(torch.tensor(pts, dtype=torch.float, device=device))
        pc.file_path = file_path
        return pc
