"""ctypes binding of libngpd_io.so (include/ngpd_io.h): threaded OBJ / XYZ / ASCII-PLY readers and the OBJ writer.
Host code without CUDA -- it replaces igl.read_obj / Open3D / the Python line loops of the reference's Object.py, not a kernel.
The library is built in-tree by build.py; a missing library is an error (there is no Python parser behind it)."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NGPD_IO_LIBRARY") or os.path.join(_HERE, "libngpd_io.so")

VERTICES, NORMALS, FACES, FACE_NORMALS, TABLE = range(5)
ERR_OPEN, ERR_EXISTS, ERR_PARSE, ERR_ARG = -1, -2, -3, -4

_c = ctypes
SIGNATURES = {
    "ngpd_io_last_error": (_c.c_char_p, []),
    "ngpd_io_read_obj": (_c.c_int, [_c.c_char_p, _c.c_int, _c.POINTER(_c.c_void_p)]),
    "ngpd_io_read_table": (_c.c_int, [_c.c_char_p, _c.c_int64, _c.c_int64, _c.c_int, _c.c_int, _c.POINTER(_c.c_void_p)]),
    "ngpd_io_count": (_c.c_int64, [_c.c_void_p, _c.c_int]),
    "ngpd_io_data": (_c.c_void_p, [_c.c_void_p, _c.c_int]),
    "ngpd_io_free": (None, [_c.c_void_p]),
    "ngpd_io_write_obj": (_c.c_int, [_c.c_char_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_int]),
    "ngpd_io_write_obj_f64": (_c.c_int, [_c.c_char_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_int]),
}
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            from . import build
            build.build_io()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _check(rc: int, path: str):
    if rc == 0:
        return
    msg = (load().ngpd_io_last_error() or b"").decode()
    if rc == ERR_EXISTS:
        raise FileExistsError(msg)
    if rc == ERR_OPEN:
        raise OSError(msg)
    raise ValueError(msg)


def _take(handle, which: int, dtype, cols: int) -> np.ndarray:
    lib = load()
    rows = lib.ngpd_io_count(handle, which)
    if rows == 0:
        return np.zeros((0, cols), dtype=dtype)
    ptr = lib.ngpd_io_data(handle, which)
    buf = (ctypes.c_char * (rows * cols * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(rows, cols).copy()


def read_obj(file_path: str, threads: int = 0):
    """v [N,3] f64, vn [M,3] f64, face vertex ids [F,3] i64, face normal ids [Fn,3] i64 (0-based, polygons fan-triangulated)."""
    lib = load()
    h = ctypes.c_void_p()
    _check(lib.ngpd_io_read_obj(os.fsencode(file_path), threads, ctypes.byref(h)), file_path)
    try:
        return (_take(h, VERTICES, np.float64, 3), _take(h, NORMALS, np.float64, 3),
                _take(h, FACES, np.int64, 3), _take(h, FACE_NORMALS, np.int64, 3))
    finally:
        lib.ngpd_io_free(h)


def read_table(file_path: str, cols: int, offset: int = 0, rows: int = -1, threads: int = 0) -> np.ndarray:
    """The first `cols` numbers of `rows` lines (all when < 0) of a whitespace-separated text file, from byte `offset`."""
    lib = load()
    h = ctypes.c_void_p()
    _check(lib.ngpd_io_read_table(os.fsencode(file_path), offset, rows, cols, threads, ctypes.byref(h)), file_path)
    try:
        return _take(h, TABLE, np.float64, cols)
    finally:
        lib.ngpd_io_free(h)


def write_obj(file_path: str, v: np.ndarray, n: np.ndarray | None = None, exclusive: bool = True, threads: int = 0) -> None:
    """`v x y z` / `vn x y z` lines with the digits Python's str(float) prints for the same values (float32 or float64 input)."""
    dt = np.float64 if np.asarray(v).dtype == np.float64 else np.float32
    v = np.ascontiguousarray(v, dtype=dt).reshape(-1, 3)
    if n is not None:
        n = np.ascontiguousarray(n, dtype=dt).reshape(-1, 3)
        if n.shape != v.shape:
            raise ValueError("write_obj: one normal per vertex")
    fn = load().ngpd_io_write_obj_f64 if dt == np.float64 else load().ngpd_io_write_obj
    _check(fn(os.fsencode(file_path), v.ctypes.data, n.ctypes.data if n is not None else None, v.shape[0], 1 if exclusive else 0, threads),
           file_path)
