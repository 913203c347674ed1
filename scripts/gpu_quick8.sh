set -x
mkdir -p gpurun_out
timeout 300 python scripts/bench_small_configs.py > gpurun_out/small_configs.md 2> gpurun_out/small_configs.err; echo "small rc=$?"; cat gpurun_out/small_configs.md; tail -3 gpurun_out/small_configs.err
python - <<'PY'
import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import ngpd_b200 as ng
f = dict(np.load('tests/golden/fandisk_denoise.npz'))
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
pos0, n0 = cu(f["pos0"]), cu(f["n_flip"])
def T(label, fn, reps=10):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        t0=time.perf_counter(); r=fn(); torch.cuda.synchronize(); ts.append(time.perf_counter()-t0)
    ts.sort(); print(f"{label}: median {ts[len(ts)//2]*1e3:.2f} ms  min {ts[0]*1e3:.2f}  max {ts[-1]*1e3:.2f}")
    return r
T("Pointcloud+Processor()", lambda: ng.Processor(ng.Pointcloud(pos0.clone())))
T("Grid create", lambda: ng._lib.Grid(pos0, 16))
T("Session create", lambda: ng._lib.Session(pos0, 16))
p = ng.Processor(ng.Pointcloud(pos0.clone())); p.graph.n = n0.clone()
T("denoise() on a live Processor", lambda: p.denoise())
s = p._get_session()
T("mean_edge_length_parts(6)", lambda: s.mean_edge_length_parts(6))
params = ng._lib.make_params(dmax=0.1)
T("session.step", lambda: s.step(params))
T("set_state+get_state", lambda: (s.set_state(pos0, n0), s.get_state(False)))
PY
