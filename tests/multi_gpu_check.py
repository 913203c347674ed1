"""Run under torchrun on N >= 2 GPUs: the Morton-slab session must reproduce the single-GPU session.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import ngpd_b200
    from ngpd_b200 import _lib, partition, workloads
    n = int(os.environ.get("NGPD_CHECK_POINTS", 400000))
    k_f = int(os.environ.get("NGPD_CHECK_KF", 16))            # 32: BASELINE configs[2]'s neighbourhood size
    clean, normal = workloads.creased_surface(n, 7, dev)
    noisy = workloads.add_noise(clean, 0.3 * workloads.expected_spacing(n))
    slab = partition.SlabSession(noisy, normal, k_f, 8, (1.0, 0.2, 1.0))
    for _ in range(2):
        slab.step()
    pos, nrm, lab = slab.gather_full()
    # single-GPU result of the same run (every rank computes it; cheap at this size)
    sess = _lib.Session(noisy, k_f)
    sess.set_state(noisy, normal)
    s, c = sess.mean_edge_length_parts(6)
    params = _lib.make_params(k_feature=k_f, dmax=2.0 * s / c)
    assert abs(s / c - slab.mean_edge_length) / (s / c) < 1e-9
    for _ in range(2):
        sess.step(params)
    rpos, rnrm, rlab = sess.get_state(True)
    scale = float(rpos.abs().max())
    perr = float((pos - rpos).abs().max()) / scale
    lab_same = float((lab == rlab).float().mean())
    nerr = float((nrm - rnrm).abs().max())
    if rank == 0:
        print(f"world={dist.get_world_size()} n={n} k_f={k_f} owned={slab.n_owned} halo={slab.n_halo} exchanges/2 steps={slab.exchanges} "
              f"max rel position diff={perr:.2e} labels equal={lab_same:.6f} max normal diff={nerr:.2e}")
    assert perr < 1e-6 and lab_same > 0.9999 and nerr < 1e-4, (perr, lab_same, nerr)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK_OK")


if __name__ == "__main__":
    main()
