"""Print the interesting numbers of a bench log (last JSON line)."""
import json, sys
l = [x for x in open(sys.argv[1]) if x.startswith('{')][-1]
d = json.loads(l)
print("step ms", round(d['ms_per_step'], 2), "patches/s", round(d['value']), "e2e ms", round(d['e2e']['ms_per_step'], 2))
for k, v in d['roofline']['kernels'].items():
    print(" ", k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
pm = d.get('parity_max') or {}
print(" parity", {k: pm.get(k) for k in ('phi_max', 'phi_median', 'phi_frac_le_1e-10', 'truncation_step_mismatches', 'K_rel_max')})
print(" roofline", d['roofline']['kernel'], round(d['roofline']['frac'], 4))
