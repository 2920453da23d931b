"""Per-source-line aggregation of an `ncu --page source --csv` SASS dump using nvdisasm -g line info.
usage: python tools/ncu_lines.py src.csv lib.so kernel_substring [topN]
The n-th SASS instruction of the kernel in the ncu dump is matched with the n-th instruction nvdisasm prints."""
import csv, subprocess, sys, tempfile, os, re, collections, glob
src, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, stdout=subprocess.DEVNULL)
txt = []
for cubin in sorted(glob.glob(os.path.join(d, "*.cubin"))):   # the library has one cubin per translation unit: take every one that holds the kernel
    t = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    if kname in t:
        txt += t.splitlines()
lines, cur, inside, pend = [], ("?", 0), False, []
depth = int(os.environ.get("DEPTH", "0"))   # 0 = outermost frame, 1 = one level of inlining below it, ...
for l in txt:
    if l.startswith(".text."):
        inside = kname in l
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]*)", line (\d+)', l)
    if m:
        pend.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        if pend:
            cur = pend[max(0, len(pend) - 1 - depth)]; pend = []
        lines.append(cur)
rows = list(csv.reader(open(src)))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h) and r[ix["# Samples"]].isdigit()]
n = len(lines)
data = data[:n]   # first launch only
print(f"sass instructions: nvdisasm {n}, ncu {len(data)}")
S, E = ix["# Samples"], ix["Instructions Executed"]
agg = collections.defaultdict(lambda: [0, 0])
for (f, ln), r in zip(lines, data):
    agg[(f, ln)][0] += int(r[S]); agg[(f, ln)][1] += int(r[E])
ts = sum(v[0] for v in agg.values()); te = sum(v[1] for v in agg.values())
srcs = {}
def text(f, ln):
    if f not in srcs:
        c = glob.glob(os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)) + glob.glob(os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "**", f), recursive=True)
        srcs[f] = open(c[0]).read().splitlines() if c else []
    return srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ""
print(f"total samples {ts}, executed {te}")
for (f, ln), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/max(ts,1):5.1f}% smp {100*e/max(te,1):5.1f}% exe  {f}:{ln}  {text(f, ln)}")
