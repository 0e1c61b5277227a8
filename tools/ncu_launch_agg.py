"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel and grid size."""
import csv,collections,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit"); gi=hdr.index("Grid Size")
agg=collections.defaultdict(list)
for r in rows[1:]:
    v=float(r[vi].replace(",","")); u=r[ui]
    if u=="ns": v/=1000
    elif u=="ms": v*=1000
    agg[(r[ki][:44], r[gi])].append(v)
for k,v in sorted(agg.items()):
    print(f"{len(v):5d} avg {sum(v)/len(v):6.2f} us min {min(v):6.2f}  grid {k[1]:>6s} {k[0]}")
