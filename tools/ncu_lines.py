"""Join an `ncu --page source --csv` SASS export with nvdisasm -g line info: per source line,
warp-instructions executed and stall samples.  Usage:
  python tools/ncu_lines.py <source.csv> <nvdisasm -g output> <mangled kernel name>"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis, kname = sys.argv[1:4]
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}

# walk the disassembly of the kernel: instruction index -> (file, line)
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kname + ":"))
cur = ("?", 0)
loc = []
inl = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith(".text."):
        break
    m = inl.search(l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        loc.append(cur)
assert len(loc) == len(data), (len(loc), len(data))
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for (f, ln), r in zip(loc, data):
    a = agg[(f, ln)]
    a[0] += int(r[ci["Instructions Executed"]])
    a[1] += int(r[ci["# Samples"]])
    for s in stalls:
        a[2][s] += int(r[ci[s]])
tot_i = sum(a[0] for a in agg.values())
tot_s = sum(a[1] for a in agg.values())
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
for (f, ln), a in sorted(agg.items()):
    if (a[0] * 200 < tot_i and a[1] * 200 < tot_s) and not (len(sys.argv) > 4 and sys.argv[4] in f):
        continue
    top = sorted(a[2].items(), key=lambda x: -x[1])[:3]
    print("%-18s %4d  instr %9d (%4.1f%%)  samples %5d (%4.1f%%)  %s" % (
        f, ln, a[0], 100.0 * a[0] / tot_i, a[1], 100.0 * a[1] / tot_s, " ".join("%s=%d" % (k[6:], v) for k, v in top if v)))

# per-file totals and stall mix
byfile = defaultdict(lambda: [0, 0, defaultdict(int)])
for (f, ln), a in agg.items():
    b = byfile[f]
    b[0] += a[0]
    b[1] += a[1]
    for k, v in a[2].items():
        b[2][k] += v
print("\nper file:")
for f, b in sorted(byfile.items(), key=lambda x: -x[1][1]):
    top = sorted(b[2].items(), key=lambda x: -x[1])[:6]
    print("%-22s instr %9d (%4.1f%%) samples %6d (%4.1f%%)  %s" % (f, b[0], 100.0 * b[0] / tot_i, b[1], 100.0 * b[1] / tot_s,
                                                                 " ".join("%s=%d" % (k[6:], v) for k, v in top if v)))
