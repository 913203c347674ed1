"""Dynamic instruction histogram of one kernel from an `ncu --page source --csv` dump: by execution multiplicity per warp
and by opcode.  usage: sass_hist.py dump.csv n_warps"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1]))); nw = float(sys.argv[2])
hdr = next(r for r in rows if 'Source' in r and 'Instructions Executed' in r)
si, ii, sa, ti = (hdr.index(n) for n in ('Source', 'Instructions Executed', '# Samples', 'Thread Instructions Executed'))
data = []
for r in rows:
    try: data.append((r[si], float(r[ii]), float(r[sa]), float(r[ti])))
    except (ValueError, IndexError): pass
data = data[:len(data) // 2] if len(data) % 2 == 0 and data[:len(data)//2] == data[len(data)//2:] else data
tot = sum(d[1] for d in data); tots = sum(d[2] for d in data)
print('instr/warp', round(tot / nw, 1), 'static', len(data), 'samples', tots)
h = collections.defaultdict(lambda: [0, 0, 0.0])
for s, n, sm, tn in data:
    m = round(n / nw, 1); h[m][0] += 1; h[m][1] += n / nw; h[m][2] += sm
for m in sorted(h):
    if h[m][1] > 0.005 * tot / nw: print(f"x{m:6.1f}: static {h[m][0]:5d}  dyn/warp {h[m][1]:8.1f} ({h[m][1]*nw/tot:5.1%})  samples {h[m][2]/tots:6.1%}")
oc = collections.Counter()
for s, n, sm, tn in data:
    p = s.split(); oc[(p[1] if p[0].startswith('@') else p[0]).split('.')[0]] += n / nw
print([(k, round(v)) for k, v in oc.most_common(24)])
