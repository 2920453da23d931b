#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
# launch lists (direct launches so that every kernel of a step is listed); the bench numbers printed under ncu are not bench values
for c in 3 4; do
  SKIP=$([ $c = 3 ] && echo 5200 || echo 2200)
  PXB_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 240 --csv --log-file $O/r6_launches_c$c.csv python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline > $O/r6_ncu_c$c.log 2>&1
  echo "ncu config $c rc=$?"
  python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$O/r6_launches_c$c.csv")) if len(r)>5]
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
if h:
    hdr=rows[h[0]]; k=hdr.index("Kernel Name"); v=hdr.index("Metric Value")
    acc=collections.OrderedDict(); n=collections.Counter()
    for r in rows[h[0]+1:]:
        try: t=float(r[v].replace(",",""))
        except: continue
        name=r[k].split("(")[0]; acc[name]=acc.get(name,0)+t; n[name]+=1
    tot=sum(acc.values())
    for name,t in sorted(acc.items(), key=lambda x:-x[1])[:14]: print(f"  {name[:60]:60s} {n[name]:4d} launches {t/1e3:10.1f} us {100*t/tot:5.1f}%")
PY
done
