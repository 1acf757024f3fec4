"""Summarise an ncu report per CUDA source line: python tools/ncu_src.py report.ncu-rep <kernel-regex> [top]"""
import csv, subprocess, sys, collections, io, re
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.Counter()
stalls = collections.defaultdict(collections.Counter)
STALL = ["stall_barrier","stall_branch_resolving","stall_dispatch","stall_lg","stall_long_sb","stall_math","stall_mio","stall_no_inst","stall_not_selected","stall_selected","stall_short_sb","stall_wait","stall_membar","stall_drain"]
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
            key = (cur_file, cur_line, cur_src.strip()[:110])
            agg[key] += v; total += v
            for nm in STALL:
                try: stalls[key][nm] += float(r[hdr.index(nm)])
                except (ValueError, IndexError): pass
print("total samples", total)
for (f, l, s), v in agg.most_common(top):
    top3 = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * c / max(v, 1)) for k, c in stalls[(f, l, s)].most_common(3))
    print("%6.2f%%  %s:%d  %-90s [%s]" % (100 * v / max(total, 1), f, l, s[:90], top3))
