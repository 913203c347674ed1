"""Where does the host-buffer entry point spend its time?  (run on the GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import ngpd_b200
from ngpd_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda:0")
noisy, nrm = bench.make_input(n, dev)
sess = _lib.Session(noisy, 16)
sess.set_state(noisy, nrm)
s, c = sess.mean_edge_length_parts(6)
params = _lib.make_params(16, 8, None, 0.3, 3.0, 0.2, (0, 1, 2), (1.0, 0.2, 1.0), 2.0 * s / c)
for _ in range(3):
    sess.step(params)
torch.cuda.synchronize()
pos_h = torch.empty((n, 3), dtype=torch.float32).pin_memory()
nrm_h = torch.empty((n, 3), dtype=torch.float32).pin_memory()
pos_o = torch.empty((n, 3), dtype=torch.float32).pin_memory()
nrm_o = torch.empty((n, 3), dtype=torch.float32).pin_memory()
lab_o = torch.empty(n, dtype=torch.uint8).pin_memory()
print("pinned:", pos_h.is_pinned(), nrm_h.is_pinned(), pos_o.is_pinned(), lab_o.is_pinned())
p, q, _ = sess.get_state(False)
pos_h.copy_(p); nrm_h.copy_(q)
d = torch.empty((n, 3), device=dev)
for name, fn in (("torch H2D 120 MB", lambda: d.copy_(pos_h, non_blocking=True)), ("torch D2H 120 MB", lambda: pos_o.copy_(d, non_blocking=True))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {dt*1e3:.2f} ms  {n*12/dt/1e9:.1f} GB/s")
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sess.run_host(params, 1, pos_h, nrm_h, pos_o, nrm_o, lab_o)
    dt = time.perf_counter() - t0
    print(f"run_host {it}: {dt*1e3:.2f} ms   tiers {sess.knn_stats()}")
    pos_h, pos_o = pos_o, pos_h; nrm_h, nrm_o = nrm_o, nrm_h
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): sess.step(params)
torch.cuda.synchronize(); print(f"device step: {(time.perf_counter()-t0)/5*1e3:.2f} ms")
