"""Per-kernel launch count, total time and share from an ncu launch list (--metrics gpu__time_duration.sum --csv).
Usage: python tools/launch_summary.py launches.csv [title] > summary.txt"""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[rows.index(hdr) + 1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu].strip(), 1e-6)
    name = re.sub(r"\(.*", "", r[ik])
    tot[name] += v * scale
    cnt[name] += 1
total = sum(tot.values())
print(sys.argv[2] if len(sys.argv) > 2 else "ncu launch list: per kernel name launches, total ms, share")
print()
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k:<60s} {cnt[k]:5d} launches {tot[k]:10.3f} ms {100 * tot[k] / total:6.2f} %")
