"""Opcode histogram per kernel of librnerf_b200.so (cuobjdump -sass), the evidence for "which pipe does this kernel use":
  python scripts/sass_histogram.py > profiles/<tag>_sass_histograms.txt
Lists, per kernel, the tensor-pipe / TMA / TMEM / cluster opcodes (UTCHMMA*, UTCBAR, LDTM, STTM, UBLKCP, UTMALDG, UTMASTG,
SYNCS, UCGABAR...) with counts, then the 12 most frequent opcodes."""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "samplenerfro_b200", "librnerf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = re.compile(r"^(UTC|LDTM|STTM|UBLKCP|UTMA|SYNCS|UCGABAR|HMMA|FFMA2|ELECT|R2UR|REDG|RED|ATOM|UTCBAR)")
print(f"# opcode histograms of {os.path.basename(lib)} (cuobjdump -sass, sm_100a); tensor-pipe / TMEM / TMA / barrier opcodes first")
for k, h in hist.items():
    if not k.startswith("rnerf::") and "rnerf::" not in k:
        continue
    tot = sum(h.values())
    special = {op: n for op, n in h.items() if KEY.match(op)}
    # group variants by their leading mnemonic (before the first dot) but keep .2CTA visible
    grp = collections.Counter()
    for op, n in special.items():
        parts = op.split(".")
        grp[parts[0] + (".2CTA" if "2CTA" in parts else "")] += n
    print(f"\n{k}  [{tot} instructions]")
    if grp:
        print("   " + "  ".join(f"{op}:{n}" for op, n in sorted(grp.items(), key=lambda kv: -kv[1])))
    print("   top: " + "  ".join(f"{op.split('.')[0]}:{n}" for op, n in h.most_common(12)))
