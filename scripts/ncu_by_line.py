"""Attribute ncu per-SASS execution counts / stall samples to source lines via nvdisasm -g line markers.
usage: ncu_by_line.py <src.csv from `ncu --page source --csv --print-source sass`> <nvdisasm -g -c listing> <mangled fn substring> <leaves>"""
import csv, re, sys
src_csv, dis, fn, L = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, iex, ismp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
recs = []
for r in rows[2:]:
    try:
        recs.append((int(r[ia], 16), int(r[iex]), int(r[ismp])))
    except Exception:
        pass
base = recs[0][0]
by_off = {a - base: (e, s) for a, e, s in recs}
cur = None
infn = False
agg = {}
for line in open(dis):
    if line.startswith('.text.'):
        infn = fn in line
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/', line)
    if m and cur:
        off = int(m.group(1), 16)
        if off in by_off:
            e, s = by_off[off]
            a = agg.setdefault(cur, [0, 0, 0])
            a[0] += e; a[1] += s; a[2] += 1
ts = sum(a[1] for a in agg.values())
te = sum(a[0] for a in agg.values())
print(f'total instr/leaf {te/L:.1f}, samples {ts}')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f'{k[0]:18s}:{k[1]:4d}  instr/leaf {a[0]/L:8.1f}  time {100*a[1]/ts:5.1f}%  nsass {a[2]}')
