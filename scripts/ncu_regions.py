import csv, re, sys
src_csv, dis, fn, L = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, iex, ismp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
isrc = hdr.index('Source')
recs = []
for r in rows[2:]:
    try: recs.append((int(r[ia], 16), int(r[iex]), int(r[ismp]), r[isrc]))
    except Exception: pass
base = recs[0][0]
by_off = {a - base: (e, s, t) for a, e, s, t in recs}
cur=None; infn=False
# walk in address order, tracking inline context is not available: use file:line only
seq=[]
for line in open(dis):
    if line.startswith('.text.'):
        infn = fn in line; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), m.group(3)); continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/', line)
    if m and cur:
        off=int(m.group(1),16)
        if off in by_off: seq.append((off,cur)+by_off[off])
# find the sweep loop by address range: contiguous region with exec count == max-ish
import collections
te=sum(x[2] for x in seq); ts=sum(x[3] for x in seq)
# classify by exec count per leaf
hot=[x for x in seq if x[2]/L>3.5]
print("instr/leaf total %.0f; in instructions executed >3.5x per leaf: %.0f (%.1f%% time)"%(te/L,sum(x[2] for x in hot)/L,100*sum(x[3] for x in hot)/ts))
# opcode mix overall and hot
def mix(xs,name):
    c=collections.Counter(); s=collections.Counter()
    for x in xs:
        op=x[4].split()[0] if not x[4].startswith('@') else x[4].split()[1]
        op=op.split('.')[0]
        c[op]+=x[2]; s[op]+=x[3]
    print(name)
    for op,v in c.most_common(22): print("  %-10s %7.1f instr/leaf  %5.1f%% time"%(op,v/L,100*s[op]/ts))
mix(hot,"HOT (sweep)"); mix([x for x in seq if x[2]/L<=3.5],"COLD (rest)")
# cold by file:line top
agg=collections.defaultdict(lambda:[0,0])
for x in seq:
    if x[2]/L<=3.5:
        a=agg[(x[1][0],x[1][1])]; a[0]+=x[2]; a[1]+=x[3]
print("COLD by line")
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:40]:
    print("  %-22s:%4d instr/leaf %6.1f time %4.1f%%"%(k[0],k[1],a[0]/L,100*a[1]/ts))
