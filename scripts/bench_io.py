"""Host-side file I/O (libngpd_io.so) against the Python line loops it replaces: python scripts/bench_io.py [points] -> markdown."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ngpd_b200
from ngpd_b200 import _io

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rng = np.random.default_rng(1)
v = rng.standard_normal((n, 3)).astype(np.float32)
nrm = rng.standard_normal((n, 3)).astype(np.float32)
d = tempfile.mkdtemp(dir=os.environ.get("TMPDIR", "/tmp"))
path = os.path.join(d, "cloud.obj")
t = time.perf_counter(); _io.write_obj(path, v, nrm); tw = time.perf_counter() - t
size = os.path.getsize(path)
t = time.perf_counter(); V, VN, F, FN = _io.read_obj(path); tr = time.perf_counter() - t
assert np.array_equal(V.astype(np.float32), v) and np.array_equal(VN.astype(np.float32), nrm)
# the Python loops on a 1/20 sample (Object.py:58-69 and a split()-based reader)
m = n // 20
t = time.perf_counter()
with open(os.path.join(d, "py.obj"), "w") as f:
    f.write("# File made by Ruben Band\n")
    for row in v[:m].tolist():
        f.write("v " + " ".join(str(x) for x in row) + "\n")
    for row in nrm[:m].tolist():
        f.write("vn " + " ".join(str(x) for x in row) + "\n")
pw = (time.perf_counter() - t) * 20
t = time.perf_counter()
a, b = [], []
with open(os.path.join(d, "py.obj")) as f:
    for line in f:
        if line.startswith("v "):
            a.append(line.split()[1:4])
        elif line.startswith("vn "):
            b.append(line.split()[1:4])
np.asarray(a, dtype=np.float64); np.asarray(b, dtype=np.float64)
pr = (time.perf_counter() - t) * 20
print(f"| {n} points + normals, {size / 1e6:.0f} MB OBJ, {os.cpu_count()} host threads | write {tw:.2f} s ({size / tw / 1e6:.0f} MB/s) | read {tr:.2f} s ({size / tr / 1e6:.0f} MB/s) | Python loops (x20 from a 1/20 sample): write {pw:.1f} s, read {pr:.1f} s |")
for f in os.listdir(d):
    os.remove(os.path.join(d, f))
os.rmdir(d)
