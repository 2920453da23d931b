"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum,
smsp__thread_inst_executed_per_inst_executed.ratio] --csv` launch list.
usage: python tools/launch_summary.py launches.csv [top]"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
hdr = rows[h[0]]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value"); u = hdr.index("Metric Unit"); m = hdr.index("Metric Name")
TIME = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc, n, dram, lanes = collections.OrderedDict(), collections.Counter(), collections.Counter(), collections.defaultdict(list)
ids = set()
for r in rows[h[0] + 1:]:
    try:
        x = float(r[v].replace(",", ""))
    except Exception:
        continue
    name = r[k].split("(")[0].replace("void ", "")
    if r[m] == "gpu__time_duration.sum":
        acc[name] = acc.get(name, 0) + x * TIME.get(r[u], 1e-3); n[name] += 1; ids.add(r[0])
    elif r[m].startswith("dram__bytes"):
        dram[name] += x * BYTES.get(r[u], 1.0)
    elif r[m].startswith("smsp__thread_inst_executed_per_inst"):
        lanes[name].append(x)
tot = sum(acc.values())
print(f"{len(ids)} launches, {tot:.1f} us in total" + (f", {sum(dram.values()) / 1e6:.1f} MB of DRAM traffic" if dram else ""))
for name, t in sorted(acc.items(), key=lambda x: -x[1])[:top]:
    extra = ""
    if dram:
        extra += f"  {dram[name] / n[name] / 1e6:9.2f} MB/launch {dram[name] / max(t, 1e-9) / 1e3:7.1f} GB/s"
    if lanes.get(name):
        extra += f"  {sum(lanes[name]) / len(lanes[name]):5.1f} lanes"
    print(f"  {name[:56]:56s} {n[name]:4d} launches {t:10.1f} us {100 * t / tot:5.1f} %  ({t / n[name]:8.1f} us each){extra}")
