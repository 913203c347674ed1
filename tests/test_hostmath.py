"""The product's per-point math headers (csrc/point_math.cuh, csrc/eig3.cuh) compiled for the host with g++ and
checked against the reference's golden vectors: lets the kernels' arithmetic be verified without a GPU.  The
harness (tests/hostmath/) is test-only; the CUDA kernels include the same headers."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

import ngpd_oracle as O
from conftest import ROOT, angle_between

HM = os.path.join(ROOT, "tests", "hostmath")


@pytest.fixture(scope="module")
def hm():
    so = os.path.join(HM, "libhostmath.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", os.path.join(HM, "hostmath.cpp"), "-o", so],
                   check=True)
    return ctypes.CDLL(so)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


XT = None


def xt():
    global XT
    if XT is None:
        XT = ctypes.c_float(float(O.acos_threshold(math.pi * 5 / 12)))
    return XT


@pytest.mark.parametrize("it", [0, 1])
def test_tensor_eig_smooth_labels(hm, fandisk, it):
    t = f"it{it}_"
    n = len(fandisk["pos0"])
    pos, nrm, idx = (np.ascontiguousarray(fandisk[t + k]) for k in ("pos_in", "n_in", "knn16"))
    w = np.zeros((n, 3), np.float32); V = np.zeros((n, 3, 3), np.float32); T = np.zeros((n, 3, 3), np.float32); sw = np.zeros(n, np.int32)
    hm.hm_nvt(P(pos), P(nrm), P(idx), None, ctypes.c_int64(n), 16, xt(), P(w), P(V), P(T), P(sw))
    assert np.array_equal(T, fandisk[t + "T1"])                                   # voting tensors bit-exact
    assert np.abs(w - fandisk[t + "eigval1"]).max() < 1e-6
    gap = np.minimum(np.diff(fandisk[t + "eigval1"], axis=1).min(axis=1), 1.0)
    signs = ((V * fandisk[t + "eigvec1"]).sum(1) > 0).all(1)
    assert signs[gap > 1e-2].all()                                                # LAPACK sign convention reproduced
    assert signs.mean() > 0.99
    good = signs & (gap > 1e-2)
    assert np.abs(V - fandisk[t + "eigvec1"])[good].max() < 1e-4
    # smoothing fed the reference's eigenvectors: bit-exact
    out = np.zeros((n, 3), np.float32)
    gw, gV = np.ascontiguousarray(fandisk[t + "eigval1"]), np.ascontiguousarray(fandisk[t + "eigvec1"])
    hm.hm_smooth(P(gw), P(gV), P(nrm), ctypes.c_int64(n), ctypes.c_float(0.3), ctypes.c_float(3.0), P(out))
    assert np.array_equal(out, fandisk[t + "f_n"])
    # free running (own eigenvectors): within the reference's own 1-ulp noise floor (0.82 % > 1e-4 rad)
    hm.hm_smooth(P(w), P(V), P(nrm), ctypes.c_int64(n), ctypes.c_float(0.3), ctypes.c_float(3.0), P(out))
    assert (angle_between(out, fandisk[t + "f_n"]) > 1e-4).mean() < 0.0082
    # stage 2 + labels
    fn = np.ascontiguousarray(fandisk[t + "f_n"])
    hm.hm_nvt(P(pos), P(fn), P(idx), None, ctypes.c_int64(n), 16, xt(), P(w), P(V), P(T), P(sw))
    assert np.array_equal(T, fandisk[t + "T2"])
    lab = np.zeros(n, np.uint8)
    hm.hm_classify(P(w), ctypes.c_int64(n), ctypes.c_float(0.2), P(lab))
    assert np.array_equal(lab, fandisk[t + "classes"])                            # labels bit-exact from own eigenvalues


@pytest.mark.parametrize("it", [0, 1])
def test_updates(hm, fandisk, it):
    t = f"it{it}_"
    cur = np.ascontiguousarray(fandisk[t + "pos_in"]).copy()
    fn = np.ascontiguousarray(fandisk[t + "f_n"]); cls = fandisk[t + "classes"]; idx8 = fandisk[t + "knn8"]
    edge = np.ascontiguousarray(fandisk[t + "eigvec2"][:, :, 0])
    d = ctypes.c_float(float(np.float32(2) * fandisk["l"]))
    for key, kind, alpha in ((0, 0, 1.0), (1, 1, 0.2), (2, 2, 1.0)):
        rows = np.nonzero(cls == key)[0].astype(np.int32)
        sub = np.ascontiguousarray(idx8[rows])
        _, delta = O.flat_center_delta(cur, sub)
        out = np.zeros((len(rows), 3), np.float32)
        hm.hm_update(kind, P(cur), P(fn), P(edge), P(sub), P(rows), ctypes.c_int64(len(rows)), 8, ctypes.c_float(alpha), d,
                     ctypes.c_float(float(delta)), P(out))
        ref = fandisk[t + f"pos_after_class{key}"]
        assert np.abs(out - ref[rows]).max() / np.abs(ref).max() < 1e-5
        cur = ref.copy()


def test_pca_normals(hm, fandisk):
    n = len(fandisk["pos0"])
    pos, idx = np.ascontiguousarray(fandisk["pos0"]), np.ascontiguousarray(fandisk["knn12_noself"])
    out = np.zeros((n, 3), np.float32)
    hm.hm_pca(P(pos), P(idx), ctypes.c_int64(n), 12, P(out))
    assert angle_between(out, fandisk["n_pca"]).max() < 1e-4                      # sign-consistent with LAPACK's


def test_eigh3_properties(hm):
    """Identities from the reference's PatchGeneration/Tests/test_RotationMatrix.py:86-118 on our solver:
    (n (x) n) has top eigenvector +-n with eigenvalue |n|^2, eigenvectors are orthonormal, A V = V diag(w)."""
    rng = np.random.default_rng(3)
    m = 2000
    v = rng.normal(size=(m, 3)).astype(np.float32)
    T = (v[:, :, None] * v[:, None, :]).astype(np.float32)
    w = np.zeros((m, 3), np.float32); V = np.zeros((m, 3, 3), np.float32)
    hm.hm_eigh3(P(np.ascontiguousarray(T)), ctypes.c_int64(m), P(w), P(V))
    top = V[:, :, 2]
    un = v / np.linalg.norm(v, axis=1, keepdims=True)
    assert np.abs(np.abs((top * un).sum(1)) - 1).max() < 1e-5
    assert np.abs(w[:, 2] - (v * v).sum(1)).max() / (v * v).sum(1).max() < 1e-5
    A = rng.normal(size=(m, 3, 3)).astype(np.float32); A = ((A + A.transpose(0, 2, 1)) / 2).astype(np.float32)
    hm.hm_eigh3(P(np.ascontiguousarray(A)), ctypes.c_int64(m), P(w), P(V))
    assert (np.diff(w, axis=1) >= 0).all()
    I = np.einsum("nij,nik->njk", V, V)
    assert np.abs(I - np.eye(3)).max() < 1e-5
    assert np.abs(np.einsum("nij,njk->nik", A, V) - V * w[:, None, :]).max() < 2e-5
    wr = np.linalg.eigvalsh(A.astype(np.float64))
    assert np.abs(w - wr).max() < 1e-5
