"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: python tools/launch_summary.py launches.csv [top]"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
hdr = rows[h[0]]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value"); u = hdr.index("Metric Unit")
acc, n = collections.OrderedDict(), collections.Counter()
for r in rows[h[0] + 1:]:
    try:
        t = float(r[v].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[u], 1e-3)
    except Exception:
        continue
    name = r[k].split("(")[0].replace("void ", ""); acc[name] = acc.get(name, 0) + t; n[name] += 1
tot = sum(acc.values())
print(f"{len(rows) - h[0] - 1} launches, {tot:.1f} us in total")
for name, t in sorted(acc.items(), key=lambda x: -x[1])[:top]:
    print(f"  {name[:64]:64s} {n[name]:4d} launches {t:10.1f} us {100 * t / tot:5.1f} %  ({t / n[name]:8.1f} us each)")
