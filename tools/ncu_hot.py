"""Summarise an `ncu --page source --csv` (SASS view) dump: hottest instructions by stall samples / executed count.
usage: python tools/ncu_hot.py src.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h) and r[ix["# Samples"]].isdigit()]
S, E = ix["# Samples"], ix["Instructions Executed"]
tot_s = sum(int(r[S]) for r in data); tot_e = sum(int(r[E]) for r in data)
print(f"instructions {len(data)}, samples {tot_s}, warp-instr executed {tot_e}")
print("-- top by samples")
for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][S]))[:top]:
    print(f"{k:5d} {int(r[S]):7d} {100*int(r[S])/max(tot_s,1):5.1f}%  exec {int(r[E]):9d}  {r[ix['Source']].strip()[:90]}")
print("-- by opcode (executed)")
agg = collections.Counter(); ags = collections.Counter()
for r in data:
    op = r[ix["Source"]].strip().split()
    op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
    op = op.split(".")[0]
    agg[op] += int(r[E]); ags[op] += int(r[S])
for op, c in agg.most_common(18):
    print(f"{op:10s} exec {c:10d} {100*c/tot_e:5.1f}%  samples {100*ags[op]/max(tot_s,1):5.1f}%")
