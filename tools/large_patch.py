"""3-D, 4 subdivisions, oversampling 2 on 8^3 coarse cells (Ni = 6859, half band width 381, 125 coarse dofs per patch):
the configuration class SURVEY 8f row 4 names.  Runs the offline phase on the GPU (SIMT solver with its windows in global
memory) and compares a few patches with the oracle."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from parity_common import build_pair, margin_safe, selection_sensitivity

ref = int(sys.argv[1]) if len(sys.argv) > 1 else 3
t0 = time.time()
ctx, orc = build_pair(dim=3, s=1, ref=ref, n=4, ell=2)
print("create %.2f s, patches %d" % (time.time() - t0, ctx.n_patches))
t0 = time.time(); ctx.compute_basis(); print("compute_basis %.2f s" % (time.time() - t0), "kernel ms", ctx.timings()[:6])
t0 = time.time(); ctx.assemble_coarse(); print("assemble_coarse %.2f s" % (time.time() - t0))
pids = [0, ctx.n_patches // 2 + 3, ctx.n_patches - 1]
t0 = time.time(); orc.compute_basis(pids); print("oracle %d patches %.1f s" % (len(pids), time.time() - t0))
for res in orc.patches:
    phi, aphi = ctx.basis(res.pid)
    err = np.linalg.norm(phi - res.basis[0])
    print(res.pid, "Nf", len(phi), "slod", res.info["slod"], "err", err, "safe", margin_safe(res.info, 0),
          "sens", selection_sensitivity(res.info, 0) if res.info["slod"] else 0, "steps", int(ctx.diagnostics(res.pid)[1]), res.info["trunc_steps"])
