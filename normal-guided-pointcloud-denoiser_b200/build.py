"""Builds libngpd.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python normal-guided-pointcloud-denoiser_b200/build.py [--force]

nvcc cross-compiles without a GPU.  --fmad=false keeps fp32/fp64 multiply-adds un-fused so that the kernels
round like the reference's CPU arithmetic (explicit fmaf() is used where torch itself fuses)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# NGPD_LIB_OUT / NGPD_EXTRA_NVCC_FLAGS: a second library with other macros next to the default one, for A/B runs (NGPD_LIBRARY selects it)
LIB = os.environ.get("NGPD_LIB_OUT") or os.path.join(HERE, "libngpd.so")
SOURCES = ["grid.cu", "knn.cu", "ball.cu", "nbr.cu", "mesh.cu", "orient.cu", "session.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--threads", "4"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def needs_build():
    if not os.path.exists(LIB):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if os.path.isfile(os.path.join(CSRC, f))] + [os.path.join(HERE, "..", "include", "ngpd.h")]
    return _newest(deps) > os.path.getmtime(LIB)


def build(force=False, verbose=False):
    """Safe under torchrun: every rank may call it.  A file lock serialises the ranks (the first one compiles, the others find
    the library up to date) and the library is written to a temporary name and renamed into place, so nobody can dlopen a
    half-written file."""
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            tmp = f"{LIB}.tmp.{os.getpid()}"
            cmd = [nvcc] + NVCC_FLAGS + os.environ.get("NGPD_EXTRA_NVCC_FLAGS", "").split() + ["-shared", "-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB)
            if verbose:
                print(res.stderr)
            return LIB
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


IO_LIB = os.path.join(HERE, "libngpd_io.so")
IO_SOURCE = os.path.join(CSRC, "io", "ngpd_io.cpp")


def build_io(force=False):
    """libngpd_io.so: the host-side file readers / writer (include/ngpd_io.h), plain g++."""
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            deps = [IO_SOURCE, os.path.join(HERE, "..", "include", "ngpd_io.h")]
            if not force and os.path.exists(IO_LIB) and _newest(deps) <= os.path.getmtime(IO_LIB):
                return IO_LIB
            tmp = f"{IO_LIB}.tmp.{os.getpid()}"
            cmd = [os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-pthread", "-o", tmp, IO_SOURCE]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, IO_LIB)
            return IO_LIB
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_io(force="--force" in sys.argv))
