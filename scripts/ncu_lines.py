"""Development aid: attribute an ncu report's per-SASS-instruction counters to source lines.
  python scripts/ncu_lines.py <report.ncu-rep> <object.o> <kernel-substring> [units]
`units` = number of (warp x iteration) units to normalise the executed-instruction counts by.
Uses `ncu --page source --csv` for the counters and `nvdisasm -g` on the object for the line table."""
import csv, os, re, subprocess, sys, tempfile
from collections import OrderedDict

rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr, body = rows[hi], rows[hi + 1:]
iS, iE, iP = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], "?", False
for l in sass:
    if l.startswith("\t.section") or l.startswith(".section"):
        on = ".text." in l and kern in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        if inl:
            cur += " <- " + " <- ".join(f"{os.path.basename(a)}:{b}" for a, b in inl)
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(.*?);', l)
    if m:
        lines.append((cur, m.group(1)))
if len(lines) != len(body):
    print(f"warning: {len(lines)} disassembled instructions vs {len(body)} in the report", file=sys.stderr)
agg = OrderedDict()
for (loc, ins), r in zip(lines, body):
    e, p = int(r[iE]), int(r[iP])
    a = agg.setdefault(loc, [0, 0, 0])
    a[0] += e; a[1] += p; a[2] += 1
tot_e = sum(a[0] for a in agg.values()); tot_p = sum(a[1] for a in agg.values())
print(f"total executed/unit {tot_e / units:.2f}, samples {tot_p}")
for loc, (e, p, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{e / units:8.2f} exec/unit {100 * p / max(tot_p, 1):5.1f}% samples {n:4d} sass | {loc}")
