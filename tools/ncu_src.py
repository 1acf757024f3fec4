"""Summarise an ncu report per CUDA source line: python tools/ncu_src.py report.ncu-rep <kernel-regex> [top]"""
import csv, subprocess, sys, collections, io, re
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.Counter()
cur_file, cur_line, cur_src = "?", 0, ""
hdr = None
total = 0.0
for r in rows:
    if not r: continue
    if r[0] == "File Name": cur_file = r[1].split("/")[-1]; continue
    if "# Samples" in r:
        hdr = r; continue
    if hdr is None: continue
    # cuda,sass view: source lines have a line number in col0 and source text; sass lines have an address
    si = hdr.index("# Samples")
    try: v = float(r[si])
    except (ValueError, IndexError): v = None
    if re.fullmatch(r"\d+", r[0] or ""):
        cur_line, cur_src = int(r[0]), r[1]
        if v is not None:
            agg[(cur_file, cur_line, cur_src.strip()[:110])] += v; total += v
print("total samples", total)
for (f, l, s), v in agg.most_common(top):
    print("%6.2f%%  %s:%d  %s" % (100 * v / max(total, 1), f, l, s))
