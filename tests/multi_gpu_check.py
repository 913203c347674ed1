"""Run under torchrun on N >= 2 GPUs: the Morton-slab session must reproduce the single-GPU session bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Covers both plans (replicated input -> SlabPlan, sharded input -> ShardedSlabPlan), both halo transports (peer stores driven
from C = the default, NCCL all-to-all phase by phase), the notebook's clamp mode, the order-independent checksum that bench.py
prints, and the halo-width check (a halo that is too narrow must be reported, not silently give other neighbours).
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import ngpd_b200
    from ngpd_b200 import _lib, partition, workloads as W
    n = int(os.environ.get("NGPD_CHECK_POINTS", 400000))
    k_f = int(os.environ.get("NGPD_CHECK_KF", 16))            # 32: BASELINE configs[2]'s neighbourhood size
    kind = os.environ.get("NGPD_CHECK_SURFACE", "creased")
    sigma = 0.3 * W.mean_knn_distance(kind, n, 6)

    def shard(r, w, n=n, sigma=sigma):
        ps, ns, gs = [], [], []
        for c in W.chunks_of(n, r, w):
            p, q, g = W.surface_chunk(kind, n, c, 7, dev)
            ps.append(W.noise_chunk(p, sigma, c)); ns.append(q); gs.append(g)
        if not ps:
            return torch.empty((0, 3), device=dev), torch.empty((0, 3), device=dev), torch.empty(0, dtype=torch.long, device=dev)
        return torch.cat(ps), torch.cat(ns), torch.cat(gs)

    noisy, normal, _ = shard(0, 1)                              # the whole cloud (every rank; cheap at this size)
    # odd sharding on purpose: rank r gets the points with id % world == r
    mine = torch.arange(rank, n, world, device=dev)
    ok = True

    def single(flags=0, clamp=False, strategy=None):
        sess = _lib.Session(noisy, k_f)
        sess.set_state(noisy, normal)
        s, c = sess.mean_edge_length_parts(6)
        d = 2.0 * s / c
        params = _lib.make_params(k_feature=k_f, dmax=d * (20000.0 if clamp else 1.0), flags=flags, clamp_radius=d if clamp else 0.0,
                                  **({"strategy": strategy} if strategy else {}))
        if clamp:
            sess.set_original(noisy)
        for _ in range(2):
            sess.step(params)
        return sess, s / c

    ref, l_ref = single()
    rpos, rnrm, rlab = ref.get_state(True)
    rsum = ref.checksum()
    scale = float(rpos.abs().max())

    cases = [("replicated/peer", dict(), "peer"), ("replicated/nccl", dict(), "nccl"), ("sharded/peer", dict(shard_ids=mine), "peer")]
    for name, extra, transport in cases:
        if "shard_ids" in extra:
            slab = partition.SlabSession(noisy[mine], normal[mine], k_f, 8, (1.0, 0.2, 1.0), transport=transport, **extra)
        else:
            slab = partition.SlabSession(noisy, normal, k_f, 8, (1.0, 0.2, 1.0), transport=transport)
        assert abs(l_ref - slab.mean_edge_length) / l_ref < 1e-9
        for _ in range(2):
            slab.step()
        pos, nrm, lab = slab.gather_full()
        need = slab.verify_halo()
        csum = slab.checksum()
        perr = float((pos - rpos).abs().max()) / scale
        lab_same = float((lab == rlab).float().mean())
        nerr = float((nrm - rnrm).abs().max())
        same = csum == rsum
        if rank == 0:
            print(f"[{name}] world={world} n={n} k_f={k_f} owned={slab.n_owned} halo={slab.n_halo} exchanges/2 steps={slab.exchanges} "
                  f"max rel position diff={perr:.2e} labels equal={lab_same:.6f} max normal diff={nerr:.2e} "
                  f"checksum equal={same} halo need/width={need:.4g}/{slab.plan.halo_width:.4g}", flush=True)
        ok = ok and perr == 0.0 and lab_same == 1.0 and nerr == 0.0 and same
        del slab

    # the notebook's "Ours" mode: all classes from one snapshot + displacement clamp
    refc, _ = single(_lib.STEP_SNAPSHOT_CLASSES, True, (_lib.STEP_FLAT, _lib.STEP_FEATURE, _lib.STEP_FEATURE))
    slab = partition.SlabSession(noisy[mine], normal[mine], k_f, 8, (1.0, 0.2, 1.0), shard_ids=mine, flags=_lib.STEP_SNAPSHOT_CLASSES,
                                 strategy=(_lib.STEP_FLAT, _lib.STEP_FEATURE, _lib.STEP_FEATURE))
    d = 2.0 * slab.mean_edge_length
    slab.params.dmax, slab.params.clamp_radius = d * 20000.0, d
    slab.session.set_original(slab.tree_local)
    for _ in range(2):
        slab.step()
    same = slab.checksum() == refc.checksum()
    pos, nrm, lab = slab.gather_full()
    cpos, cnrm, clab = refc.get_state(True)
    bad = ((pos != cpos).any(dim=1) | (nrm != cnrm).any(dim=1) | (lab != clab))
    if rank == 0:
        print(f"[sharded/peer clamp+snapshot] checksum equal={same}; rows that differ: {int(bad.sum())} "
              f"(positions {int((pos != cpos).any(dim=1).sum())}, normals {int((nrm != cnrm).any(dim=1).sum())}, labels {int((lab != clab).sum())}; "
              f"by label {[int((bad & (clab == l)).sum()) for l in range(3)]}; max |dpos| {float((pos - cpos).abs().max()):.3e})", flush=True)
    ok = ok and same
    del slab

    # the benchmark's own recipe at a size where a cloud-wide sum has many terms: PCA normals computed per slab, more iterations
    n2 = int(os.environ.get("NGPD_CHECK_POINTS_2", 3000000))
    sigma2 = 0.3 * W.mean_knn_distance(kind, n2, 6)
    ps, ns, gs = shard(rank, world, n2, sigma2)                 # (a rank may get no chunk at all: an empty shard is a valid shard)
    slab = partition.SlabSession(ps, ns, k_f, 8, (1.0, 0.2, 1.0), shard_ids=gs)
    slab.pca_normals(12, orient_like="current")
    for _ in range(6):
        slab.step()
    csum = slab.checksum()
    need = slab.verify_halo()
    del slab, ps, ns, gs
    big, bign, _ = shard(0, 1, n2, sigma2)
    grid = _lib.Grid(big, 12)
    table = grid.knn(big, 12, _lib.KNN_SKIP_SELF | _lib.KNN_QUERY_IS_TREE)
    pn = torch.empty_like(big)
    _lib.check(_lib.load().ngpd_pca_normals(big.data_ptr(), table.data_ptr(), None, n2, 12, pn.data_ptr(), None, None, _lib.stream()), "pca")
    flip = (pn * bign).sum(1) < 0
    pn[flip] *= -1
    del grid, table
    sess = _lib.Session(big, k_f)
    sess.set_state(big, pn)
    s_, c_ = sess.mean_edge_length_parts(6)
    prm = _lib.make_params(k_feature=k_f, dmax=2.0 * s_ / c_)
    for _ in range(6):
        sess.step(prm)
    rs = sess.checksum()
    same = rs == csum
    if rank == 0:
        print(f"[bench recipe, {n2} points, PCA normals per slab, 6 iterations] checksum equal={same} halo need={need:.4g}\n   slab   {csum}\n   single {rs}", flush=True)
    ok = ok and same
    del sess, big, bign, pn

    # a halo that is too narrow must be noticed
    slab = partition.SlabSession(noisy, normal, k_f, 8, (1.0, 0.2, 1.0), halo_width=0.25 * partition.estimate_halo_width(noisy, k_f) / 6.0)
    slab.step()
    try:
        slab.verify_halo()
        caught = False
    except RuntimeError as e:
        caught = True
        if rank == 0:
            print("[narrow halo] reported:", str(e)[:120], flush=True)
    ok = ok and caught
    dist.barrier()
    dist.destroy_process_group()
    assert ok, "slab run differs from the single-GPU run"
    if rank == 0:
        print("MULTI_GPU_CHECK_OK")


if __name__ == "__main__":
    main()
