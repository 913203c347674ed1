"""Markdown tables for profiles/README.md from the bench lines of one round:
    python scripts/r2_tables.py profiles/r2_bench_100M.json profiles/r2_bench_2gpu_100M.json ...   (first = the one-GPU line)"""
import json, sys


def load(f):
    return json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])


lines = [load(f) for f in sys.argv[1:]]
one = lines[0]
print("| GPUs | ms/step | G point-iterations/s | speed-up | efficiency | cold (iterations 1-2) G pt-it/s | e2e G pt-it/s | kernels ms (rank 0) | cross-rank rounds ms (rank 0) | outside kernels ms | checksum (positions) | validated |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for d in lines:
    k = d["kernels"]
    ker = sum(v["ms_per_step"] for n, v in k.items() if n != "halo")
    halo = k.get("halo", {}).get("ms_per_step", 0.0)
    sp = one["ms_per_step"] / d["ms_per_step"]
    v = d.get("validated")
    print(f"| {d['n_gpus']} | {d['ms_per_step']:.3f} | {d['value'] / 1e9:.2f} | {sp:.2f} | {sp / d['n_gpus']:.3f} | {d['cold']['value'] / 1e9:.2f} | "
          f"{d['e2e']['value'] / 1e9:.2f} | {ker:.3f} | {halo:.3f} | {d['roofline']['outside_kernels_ms_per_step']:.3f} | `{d['checksum']['pos_hash']}` | "
          f"{'ok' if v and v['ok'] else ('-' if not v else 'FAILED')} |")
print()
print("| kernel group | algorithmic B/point | " + " | ".join(f"{d['n_gpus']} GPU ms (frac of HBM peak)" for d in lines) + " |")
print("|---|---|" + "---|" * len(lines))
for name in ("knn", "nvt_smooth", "nvt_classify", "flat_scalars", "update", "halo"):
    cells = []
    for d in lines:
        v = d["kernels"].get(name)
        cells.append("-" if not v else (f"{v['ms_per_step']:.3f} ({v['frac']:.3f})" if "frac" in v else f"{v['ms_per_step']:.3f}"))
    ab = one["kernels"].get(name, {}).get("algorithmic_bytes_per_point", "-")
    ab = "-" if ab is None else ab
    print(f"| {name} | {ab} | " + " | ".join(cells) + " |")
print(f"| whole iteration (B_iter = 294 B) | 294 | " + " | ".join(f"{d['ms_per_step']:.3f} ({d['roofline']['iteration_frac']:.3f} per GPU)" for d in lines) + " |")
for d in lines:
    if d.get("kernel_ms_per_step_by_rank"):
        print(f"\nPer-rank kernel time, {d['n_gpus']} GPUs (ms per step):\n")
        print("| group | " + " | ".join(f"rank {r}" for r in range(d["n_gpus"])) + " |")
        print("|---|" + "---|" * d["n_gpus"])
        for kname, vals in d["kernel_ms_per_step_by_rank"].items():
            print(f"| {kname} | " + " | ".join(f"{x:.3f}" for x in vals) + " |")
        if d.get("halo", {}).get("balance"):
            print(f"\nslab sizes after the cost trial: {d['halo']['balance']}")
if one.get("extra_configs"):
    print("\n| configuration | points | ms/step | G pt-it/s | iteration frac of HBM peak | edge / corner fraction | kNN | NVT+smooth | NVT+labels | flat scalars | updates |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for e in one["extra_configs"]:
        km = e["kernels_ms_per_step"]
        h = e["class_histogram"]
        print(f"| {e['name']} | {e['points']} | {e['ms_per_step']:.3f} | {e['value'] / 1e9:.2f} | {e['iteration_frac_of_hbm_peak']:.3f} | {h['edge_fraction']:.3f} / {h['corner_fraction']:.4f} | "
              + " | ".join(f"{km.get(n, 0.0):.3f}" for n in ("knn", "nvt_smooth", "nvt_classify", "flat_scalars", "update")) + " |")
print("\none-GPU line: knn", one.get("knn"), "\ncpu_baseline", one.get("cpu_baseline"), "\ntiers", one.get("knn_tiers_last_step"), "\nclass histogram", one.get("class_histogram"))
