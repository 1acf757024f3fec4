"""SASS instruction histogram per kernel of libslod_b200.so (cuobjdump -sass): markers that prove the hardware paths in use
(DMMA = fp64 tensor core, UBLKCP = bulk asynchronous copy / TMA engine, SYNCS = mbarrier, LDGSTS = cp.async) and the most
frequent opcodes.  Usage: python tools/sass_histogram.py [lib] > profiles/r02_sass_histogram.txt"""
import os, re, subprocess, sys
from collections import Counter
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dealii-slod_b200", "libslod_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
print("SASS instruction histogram per kernel of dealii-slod_b200/libslod_b200.so (cuobjdump -sass, sm_100a only; static counts).")
print("Markers: DMMA.8x8x4 = fp64 tensor core (mma.sync; m16n8k8 compiles to four of them); UBLKCP = bulk asynchronous copy")
print("(cp.async.bulk, TMA engine); SYNCS = mbarrier operations; LDGSTS = cp.async.  tcgen05 (UTC*MMA) has no f64 kind and does not appear.")
print()
MARK = ("DMMA", "UBLKCP", "SYNCS", "LDGSTS", "MUFU", "UTC", "UTMA")
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    mangled = part.split("\n", 1)[0].strip()
    name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("slod::", "")
    ops = []
    for l in part.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", l)
        if m:
            ops.append(m.group(1))
    c = Counter(ops)
    fam = Counter()
    for o, n in c.items():
        if o.startswith(MARK):
            fam[".".join(o.split(".")[:2]) if o.startswith(("DMMA", "MUFU", "SYNCS")) else o.split(".")[0]] += n
    short = Counter()
    for o, n in c.items():
        short[o.split(".")[0]] += n
    print(f"{name}: {len(ops)} instructions")
    print("    markers: " + (", ".join(f"{k} {v}" for k, v in sorted(fam.items(), key=lambda kv: -kv[1])) or "-"))
    print("    top: " + ", ".join(f"{k} {v}" for k, v in short.most_common(14)))
