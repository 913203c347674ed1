"""Digest of an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: the kernels of the LAST complete
steady-state iteration (from one re-ranking kernel to the next), their serialised cold-cache times and shares.
    python scripts/launch_list_md.py launches.csv [which_step_from_the_end=3]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
back = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        h, start = r, i
        break
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
seq = []
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0].replace('void ', '').replace('ngpd::', '')
    seq.append((name, float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(r[ui], 1)))
marks = [i for i, (n, v) in enumerate(seq) if 'rerank' in n]
a, b = marks[-back - 1], marks[-back]
step = seq[a:b]
tot = sum(v for n, v in step)
print(f"launches in the list: {len(seq)} (ours: {sum(1 for n, v in seq if n.startswith('session_') or 'knn' in n or 'rs_' in n)}); "
      f"one steady-state iteration = {len(step)} launches, {tot:.1f} us serialised under ncu\n")
print("| # | kernel | us | share |\n|---|---|---|---|")
for j, (n, v) in enumerate(step):
    print(f"| {j} | {n} | {v:.1f} | {v / tot:.1%} |")
