"""Aggregate an `ncu --page source --csv` dump into basic blocks (runs of SASS instructions with the same execution
count) and print the heavy ones: share of warp instructions, share of stall samples, executions per warp."""
import collections
import csv
import sys

path, nwarps = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if 'Source' in r and 'Instructions Executed' in r)
si, ii, sa, ti = (hdr.index(n) for n in ('Source', 'Instructions Executed', '# Samples', 'Thread Instructions Executed'))
data = []
for r in rows:
    try:
        data.append((r[si], float(r[ii]), float(r[sa]), float(r[ti])))
    except (ValueError, IndexError):
        pass
tot = sum(d[1] for d in data); tots = sum(d[2] for d in data)
print(len(data), 'sass instr; warp inst', tot, 'per warp', tot / nwarps, 'samples', tots)
def op(s):
    p = s.split()
    return (p[1] if p[0].startswith('@') else p[0]).split('.')[0]
blocks = []
for s, n, sm, tn in data:
    if blocks and abs(blocks[-1]['n'] - n) <= 0.02 * max(n, 1):
        b = blocks[-1]; b['cnt'] += 1; b['inst'] += n; b['smp'] += sm; b['th'] += tn; b['ops'].append(op(s))
    else:
        blocks.append({'n': n, 'cnt': 1, 'inst': n, 'smp': sm, 'th': tn, 'ops': [op(s)], 'i': len(data)})
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
for b in blocks:
    if b['inst'] / tot > thr:
        c = collections.Counter(b['ops'])
        print(f"{b['inst']/tot:6.2%} inst {b['smp']/tots:6.2%} smp  x{b['n']/nwarps:8.1f}/warp  len {b['cnt']:4d} act {b['th']/max(b['inst'],1):4.1f}  {dict(c.most_common(9))}")
