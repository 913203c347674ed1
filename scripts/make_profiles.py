"""Turn one GPU measurement round into the tracked evidence under profiles/:
    python scripts/make_profiles.py <tag> <full.ncu-rep> <launches.csv or -> <bench.json> <points in the ncu run>
writes profiles/<tag>.md (bench line digest + ncu table + launch list digest) and profiles/ncu_traffic.json (DRAM bytes per
point and launch of every session kernel, read by bench.py for roofline.traffic)."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep, launches, bench, points = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
out = [f"# {tag}\n"]
if os.path.exists(bench):
    d = json.loads(open(bench).read().strip().splitlines()[-1])
    out.append(f"## bench.py line ({os.path.basename(bench)})\n")
    out.append(f"* `{d['config']['workload']}`, {d['config']['points']} points, n_gpus {d['n_gpus']}, steps {d['steps']}, warm-up {d['warmup']}")
    out.append(f"* value **{d['value'] / 1e9:.3f} G point-iterations/s** ({d['ms_per_step']:.3f} ms/step, device-resident); "
               f"e2e **{d['e2e']['value'] / 1e6:.1f} M point-iterations/s** through `{d['e2e']['call'].split(' ')[0]}`")
    out.append(f"* clocks {d['clocks']}; launches in the timed region {d['gpu_launches']}")
    out.append(f"* iteration vs HBM roofline (B_iter): {d['roofline']['iteration_achieved_gbs']:.0f} GB/s = {d['roofline']['iteration_frac']:.3f} of {d['roofline']['peak']} GB/s")
    if d.get("cpu_baseline"):
        out.append(f"* cpu_baseline: {d['cpu_baseline']}")
    out.append("\n| kernel group | ms/step | launches/step | algorithmic B/point | achieved GB/s | frac of HBM peak | share of step |\n|---|---|---|---|---|---|---|")
    for k, v in d["kernels"].items():
        if "frac" in v:
            out.append(f"| {k} | {v['ms_per_step']:.3f} | {v['launches_per_step']:.0f} | {v['algorithmic_bytes_per_point']} | {v['achieved_gbs']:.0f} | {v['frac']:.3f} | {v['share_of_step']:.3f} |")
        else:
            out.append(f"| {k} | {v['ms_per_step']:.3f} | {v['launches_per_step']:.0f} | - | - | - | {v['share_of_step']:.3f} |")
out.append(f"\n## ncu --set full ({os.path.basename(rep)}, {points} points, one steady-state iteration, cold-cache serialised replays)\n")
out.append(subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_profile_md.py"), rep], capture_output=True, text=True).stdout)
if launches != "-" and os.path.exists(launches):
    out.append(f"\n## ncu launch list ({os.path.basename(launches)})\n")
    out.append(subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "launch_list_md.py"), launches], capture_output=True, text=True).stdout)
open(os.path.join(ROOT, "profiles", tag + ".md"), "w").write("\n".join(out) + "\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("ngpd::", "").split("<")[0]
    b = sum(float(r[ix[m]].replace(",", "")) * scale[units[ix[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    traffic.setdefault(name, []).append(b / points)
json.dump({"source": os.path.basename(rep), "points": points, "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, divided by the point count; one entry per captured launch",
           "bytes_per_point": traffic}, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print("wrote", tag)
