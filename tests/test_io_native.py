"""libngpd_io.so (include/ngpd_io.h): the threaded OBJ / table readers and the OBJ writer against plain-Python restatements of
what the reference does (Object.py:58-69 saveObj's `str(x)` lines; igl.read_obj's arrays as Object.py:80-89 consumes them)."""
import ctypes
import os
import re

import numpy as np
import pytest

import ngpd_b200
from ngpd_b200 import _io
from ngpd_b200.Object import Pointcloud

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def python_read_obj(path):
    """line-by-line reader (the round-1 implementation of Object.read_obj), kept as the checker"""
    v, vn, fv, fn = [], [], [], []
    with open(path, "r") as fh:
        for raw in fh:
            line = raw.strip()
            if line.startswith("v ") or line.startswith("v\t"):
                v.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("vn ") or line.startswith("vn\t"):
                vn.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f ") or line.startswith("f\t"):
                ids = [t.split("/") for t in line.split()[1:]]

                def vid(t):
                    a = int(t[0])
                    return a - 1 if a > 0 else len(v) + a

                def nid(t):
                    a = int(t[2])
                    return a - 1 if a > 0 else len(vn) + a

                for a in range(1, len(ids) - 1):
                    tri = (ids[0], ids[a], ids[a + 1])
                    fv.append([vid(t) for t in tri])
                    if all(len(t) == 3 and t[2] != "" for t in tri):
                        fn.append([nid(t) for t in tri])
    arr = lambda x, dt: np.asarray(x, dtype=dt).reshape(-1, 3)
    return arr(v, np.float64), arr(vn, np.float64), arr(fv, np.int64), arr(fn, np.int64)


def test_io_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ngpd_io.h")).read()
    declared = set(re.findall(r"\b(ngpd_io_[a-z0-9_]+)\s*\(", header))
    assert declared >= {"ngpd_io_read_obj", "ngpd_io_read_table", "ngpd_io_write_obj", "ngpd_io_write_obj_f64", "ngpd_io_free"}
    lib = ctypes.CDLL(_io.LIB_PATH) if os.path.exists(_io.LIB_PATH) else _io.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_io.SIGNATURES)


def test_write_obj_text_equals_python_str(tmp_path):
    rng = np.random.default_rng(3)
    v = (rng.standard_normal((30000, 3)) * np.array([1.0, 1e-4, 1e6])).astype(np.float32)
    v[:6] = [[0.0, -0.0, 1.0], [1e-5, 1e-4, 1e16], [123456789.0, 1e22, 1e-45], [np.inf, -np.inf, np.nan], [0.1, 0.5, 1e15], [9.999999e15, 1e-4, 3e38]]
    n = rng.standard_normal((30000, 3)).astype(np.float32)
    for arr_v, arr_n, name in ((v, n, "f32.obj"), (v.astype(np.float64) * 1.0000001, None, "f64.obj")):
        path = tmp_path / name
        _io.write_obj(str(path), arr_v, arr_n)
        lines = ["# File made by Ruben Band\n"]
        lines += ["v " + " ".join(str(x) for x in row) + "\n" for row in arr_v.tolist()]
        if arr_n is not None:
            lines += ["vn " + " ".join(str(x) for x in row) + "\n" for row in arr_n.tolist()]
        assert path.read_text() == "".join(lines)
    with pytest.raises(FileExistsError):
        _io.write_obj(str(tmp_path / "f32.obj"), v)
    _io.write_obj(str(tmp_path / "f32.obj"), v[:2], exclusive=False)
    assert len((tmp_path / "f32.obj").read_text().splitlines()) == 3


def test_read_obj_record_forms(tmp_path):
    text = ("# comment\r\n"
            "o thing\n"
            "v 0 0 0\n"
            "  v 1.5 0 0 0.25\n"          # leading blanks, a 4th (weight) column
            "v\t0 +1 0\r\n"
            "v 1e0 1 -2.5E-1\n"
            "vt 0 0\n"
            "vn 0 0 1\n"
            "vn 0 1 0\n"
            "vn 1 0 0\n"
            "f 1 2 3\n"
            "f 1/1 2/1 3/1\n"
            "f 1//1 2//2 3//3\n"
            "f 1/1/1 2/1/2 3/1/3 4/1/1\n"  # quad: two triangles with normals
            "f -4 -3 -2 -1\n"              # relative ids
            "f 1//1 2 3//3\n"              # one corner without a normal: no normal ids for the triangle
            "f -1//-1 -2//-2 -3//-3\n"
            "g group\n"
            "f 1 2\n"                      # degenerate: no triangle
            "s off\n")
    path = tmp_path / "forms.obj"
    path.write_bytes(text.encode())
    got = _io.read_obj(str(path))
    want = python_read_obj(str(path))
    for g, w in zip(got, want):
        assert g.dtype == w.dtype and np.array_equal(g, w)
    assert got[0].shape == (4, 3) and got[2].shape == (9, 3) and got[3].shape == (4, 3)


def test_read_obj_threads_agree_across_piece_boundaries(tmp_path):
    """several MB so that the file is cut into pieces; relative ids reach back across the cuts"""
    rng = np.random.default_rng(5)
    nv = 120000
    v = rng.standard_normal((nv, 3))
    out = []
    rows = v.tolist()
    for i in range(nv):
        x, y, z = rows[i]
        out.append(f"v {x!r} {y!r} {z!r}\n")
        out.append(f"vn {z!r} {x!r} {y!r}\n")
        if i >= 3:
            if i % 3 == 0:
                out.append(f"f {i + 1}//{i + 1} {i}//{i} {i - 1}//{i - 1}\n")
            elif i % 3 == 1:
                out.append("f -1//-1 -2//-2 -3//-3 -4//-4\n")
            else:
                back = min(i + 1, 50000)                       # far enough to leave the piece
                out.append(f"f -1 -{back} -2\n")
    path = tmp_path / "big.obj"
    path.write_text("".join(out))
    assert path.stat().st_size > 6 << 20
    one = _io.read_obj(str(path), threads=1)
    many = _io.read_obj(str(path), threads=7)
    want = python_read_obj(str(path))
    for a, b, w in zip(one, many, want):
        assert np.array_equal(a, b) and np.array_equal(a, w)
    assert np.array_equal(one[0], v)                           # repr round-trips the doubles


def test_read_obj_errors(tmp_path):
    with pytest.raises(OSError):
        _io.read_obj(str(tmp_path / "missing.obj"))
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 x 0\n")
    with pytest.raises(ValueError, match="byte 8"):
        _io.read_obj(str(bad))
    bad.write_text("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(ValueError, match="does not exist"):
        _io.read_obj(str(bad))
    empty = tmp_path / "empty.obj"
    empty.write_text("")
    assert [a.shape for a in _io.read_obj(str(empty))] == [(0, 3)] * 4


def test_read_table(tmp_path):
    rng = np.random.default_rng(7)
    pts = rng.standard_normal((50000, 6))
    path = tmp_path / "cloud.xyz"
    with open(path, "w") as fh:
        fh.write("# x y z nx ny nz\n\n")
        for i, r in enumerate(pts):
            fh.write(" ".join(repr(float(x)) for x in r) + ("\r\n" if i % 2 else "\n"))
            if i % 1000 == 0:
                fh.write("   \n# a comment in between\n")
    assert np.array_equal(_io.read_table(str(path), 3), pts[:, :3])
    assert np.array_equal(_io.read_table(str(path), 6, threads=5), pts)
    assert np.array_equal(_io.read_table(str(path), 3, threads=1), pts[:, :3])
    # a body that starts behind a header and ends before other records
    body = tmp_path / "body.txt"
    header = b"header line\nanother\n"
    with open(body, "wb") as fh:
        fh.write(header)
        for r in pts[:10]:
            fh.write((" ".join(repr(float(x)) for x in r) + "\n").encode())
        fh.write(b"3 0 1 2\n3 1 2 3\n")
    assert np.array_equal(_io.read_table(str(body), 4, offset=len(header), rows=10), pts[:10, :4])
    with pytest.raises(ValueError, match="fewer than 6 numbers"):
        _io.read_table(str(body), 6, offset=len(header))
    with pytest.raises(ValueError, match="ends 5 rows early"):
        _io.read_table(str(body), 3, offset=len(header), rows=17)
    assert _io.read_table(str(body), 3, offset=len(header), rows=0).shape == (0, 3)


def test_pointcloud_files_go_through_the_library(tmp_path):
    import torch
    rng = np.random.default_rng(9)
    v = torch.tensor(rng.standard_normal((500, 3)), dtype=torch.float)
    n = torch.nn.functional.normalize(torch.tensor(rng.standard_normal((500, 3)), dtype=torch.float), dim=-1)
    pc = Pointcloud(v, n)
    path = tmp_path / "pc.obj"
    pc.saveObj(str(path))
    with pytest.raises(FileExistsError):
        pc.saveObj(str(path))
    back = Pointcloud.loadObj(str(path), device="cpu")
    assert torch.equal(back.v, v) and torch.equal(back.n, n)
    xyz = tmp_path / "pc.xyz"
    xyz.write_text("".join(f"{a!r} {b!r} {c!r} 0.5\n" for a, b, c in v.double().tolist()))
    assert torch.equal(Pointcloud.loadXYZ(str(xyz), device="cpu").v, v)
    ply = tmp_path / "pc.ply"
    ply.write_text("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 500\nproperty float nx\nproperty float x\nproperty float y\n"
                   "property float z\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n"
                   + "".join(f"0.25 {a!r} {b!r} {c!r}\n" for a, b, c in v.double().tolist()) + "3 0 1 2\n")
    assert torch.equal(Pointcloud.loadPly(str(ply), device="cpu").v, v)
