"""Markdown summary of an `ncu --set full` report, one row per captured launch:
    python scripts/ncu_profile_md.py report.ncu-rep > profiles/<name>.md      (needs ncu on PATH)"""
import csv, io, subprocess, sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, name, scale=1.0, nd=1):
    try:
        v = float(r[ix[name]].replace(",", ""))
    except (KeyError, ValueError):
        return "-"
    u = units[ix[name]]
    if name.startswith("dram__bytes") or name.startswith("gpu__time"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return f"{v * scale:.{nd}f}"


STALLS = ["long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "mio_throttle", "lg_throttle", "branch_resolving",
          "no_instruction", "dispatch_stall", "not_selected", "barrier", "membar", "imc_miss", "drain", "sleeping", "tex_throttle"]
print("| kernel | grid | time us | regs | warps active % | issue active % | warp inst (M) | DRAM rd MB | DRAM wr MB | DRAM % peak | L1 hit % | "
      "L1 thr % | L2 hit % | L2 thr % | top stalls (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("ngpd::", "")
    st = []
    for s in STALLS:
        key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if key in ix:
            try:
                st.append((float(r[ix[key]]), s))
            except ValueError:
                pass
    st.sort(reverse=True)
    stalls = ", ".join(f"{s} {v:.2f}" for v, s in st[:3])
    print("| " + " | ".join([name, f(r, "launch__grid_size", 1, 0), f(r, "gpu__time_duration.sum"), f(r, "launch__registers_per_thread", 1, 0),
                             f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                             f(r, "smsp__inst_executed.sum", 1e-6), f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum"),
                             f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), f(r, "l1tex__t_sector_hit_rate.pct"),
                             f(r, "l1tex__throughput.avg.pct_of_peak_sustained_active"), f(r, "lts__t_sector_hit_rate.pct"),
                             f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"), stalls]) + " |")
