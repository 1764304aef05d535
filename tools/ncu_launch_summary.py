import csv,collections,sys
rows=list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
d=collections.defaultdict(list)
for r in rows[1:]:
    v=float(r[vi].replace(",",""))
    if r[ui]=="ns": v/=1e3
    elif r[ui]=="ms": v*=1e3
    d[r[ki][:60]].append(v)
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1])): print(f"{k:60s} n={len(v):3d} mean={sum(v)/len(v):9.1f} us total={sum(v)/1e3:8.2f} ms")
