"""Wall-clock breakdown of the host-buffer C ABI path (the `e2e` figure of bench.py) on the headline workload."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module("dealii-slod_b200")
w = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
tables = bench.make_tables(w)
ctx = pkg.SlodContext(dim=w["dim"], spacedim=w["s"], n_global_refinements=w["ref"], n_subdivisions=w["n"],
                      oversampling=w["ell"], stabilize=True, problem=0 if w["s"] == 1 else 1)
for rep in range(4):
    t = [time.perf_counter()]
    for f, tb in enumerate(tables):
        ctx.set_coefficient(f, w["r"], tb)
    t.append(time.perf_counter()); ctx.compute_basis()
    t.append(time.perf_counter()); ctx.assemble_coarse()
    t.append(time.perf_counter()); ctx.all_basis()
    t.append(time.perf_counter()); ctx.coarse_csr()
    t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print("set_coef %.1f  compute_basis %.1f  assemble_coarse %.1f  all_basis %.1f  coarse_csr %.1f  total %.1f ms" % (*d, d.sum()), flush=True)
    print("  kernel ms", [round(float(x), 2) for x in ctx.timings()[:8]], flush=True)
