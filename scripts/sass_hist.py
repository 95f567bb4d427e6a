"""Instruction histogram of one kernel's SASS (development aid): python scripts/sass_hist.py <substr> [--dump]"""
import collections, re, subprocess, sys, os
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gigalens_b200", "libgigalens_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if sys.argv[1] not in name:
        continue
    ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", b, re.M)
    c = collections.Counter(ops)
    print(name[:90], "total", len(ops))
    print("  " + "  ".join(f"{k}:{v}" for k, v in c.most_common(18)))
    if "--dump" in sys.argv:
        open("/tmp/sass_dump.txt", "w").write(b)
