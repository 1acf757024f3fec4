"""Histogram of the selection paths / truncation counts of a bench workload (GPU only, diagnostics of the C ABI)."""
import importlib, os, sys, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("dealii-slod_b200")
wname = sys.argv[1] if len(sys.argv) > 1 else "diffusion3d_16c_l2_n2"
w = bench.WORKLOADS[wname]
ctx = pkg.SlodContext(dim=w["dim"], spacedim=w["s"], n_global_refinements=w["ref"], n_subdivisions=w["n"],
                      oversampling=w["ell"], stabilize=True, problem=0 if w["s"] == 1 else 1)
for f, t in enumerate(bench.make_tables(w)):
    ctx.set_coefficient(f, w["r"], t)
ctx.compute_basis()
print("timings ms", np.round(ctx.timings(), 2))
paths, steps, iters = collections.Counter(), collections.Counter(), []
for p in range(ctx.n_patches):
    for d in range(w["s"]):
        dg = ctx.diagnostics(p, d)
        paths[int(dg[5])] += 1
        if dg[5] >= 2:
            steps[int(dg[1])] += 1
            iters.append(dg[6])
print("paths (0 LOD, 1 Cholesky, 2 Jacobi, 3 QL):", dict(paths))
print("truncation steps on eigen-path items:", dict(sorted(steps.items())))
if iters:
    print("QL iterations / Jacobi sweeps: mean %.1f max %d" % (np.mean(iters), max(iters)))
