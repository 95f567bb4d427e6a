"""Summarise an ncu source-page CSV: dynamic instruction mix and the hottest SASS lines (development aid)."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hdr]
si, ei, wi = H.index("Source"), H.index("Instructions Executed"), H.index("Warp Stall Sampling (All Samples)")
mix = collections.Counter(); tot = 0; lines = []
for r in rows[hdr + 1:]:
    if len(r) <= ei or not r[ei].isdigit(): continue
    n = int(r[ei]); op = re.sub(r"^@!?U?P\d+\s+", "", r[si].strip()).split()[0].split(".")[0]
    mix[op] += n; tot += n; lines.append((int(r[wi] or 0), n, r[si].strip()))
print("total warp-inst", tot)
print("  ".join(f"{k}:{100*v/tot:.1f}%" for k, v in mix.most_common(16)))
if "--hot" in sys.argv:
    for w, n, s in sorted(lines, reverse=True)[:25]: print(f"{w:6d} {n:10d}  {s[:100]}")
