"""Run the online stages (C^T f, coarse CG, C u) once on the headline workload; used under ncu for profiles/."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module("dealii-slod_b200")
name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
w = bench.WORKLOADS[name]
ctx = pkg.SlodContext(dim=w["dim"], spacedim=w["s"], n_global_refinements=w["ref"], n_subdivisions=w["n"],
                      oversampling=w["ell"], stabilize=True, problem=0 if w["s"] == 1 else 1)
for f, t in enumerate(bench.make_tables(w)):
    ctx.set_coefficient(f, w["r"], t)
ctx.compute_basis()
ctx.assemble_coarse()
G = (2 ** w["ref"]) * w["n"] + 1
w1 = np.full(G, 1.0 / (G - 1)); w1[0] = w1[-1] = 0.0
f = w1
for _ in range(w["dim"] - 1):
    f = np.multiply.outer(w1, f)
f = np.repeat(f.ravel(), w["s"])
for rep in range(2):
    t0 = time.perf_counter(); b = ctx.coarse_rhs(f)
    t1 = time.perf_counter(); u, steps, res = ctx.coarse_solve(b, max_steps=20000, tolerance=0.0, reduction=1e-10)
    t2 = time.perf_counter(); uh = ctx.prolongate(u)
    t3 = time.perf_counter()
print("rhs %.3f ms, cg %.3f ms (%d steps, res %.3e), prolongate %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, steps, res, (t3 - t2) * 1e3))
