"""Run a list of unusual configurations through the offline phase on the GPU and compare a few patches each with the
oracle (same acceptance rule as tests/test_parity_gpu.py); prints one line per configuration."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from parity_common import build_pair, margin_safe, selection_sensitivity

CONFIGS = [dict(dim=2, s=1, ref=4, n=8, ell=2), dict(dim=2, s=1, ref=3, n=8, ell=3), dict(dim=2, s=2, ref=3, n=8, ell=1),
           dict(dim=2, s=2, ref=4, n=4, ell=2), dict(dim=2, s=1, ref=5, n=4, ell=3), dict(dim=3, s=1, ref=2, n=4, ell=2)]
for cfg in CONFIGS:
    t0 = time.time()
    try:
        ctx, orc = build_pair(**cfg)
        ctx.compute_basis(); ctx.assemble_coarse()
        npch = ctx.n_patches
        pids = sorted({0, npch // 3, npch // 2 + 1, npch - 1})
        orc.compute_basis(pids)
        worst, nbad, nskip = 0.0, 0, 0
        for res in orc.patches:
            for d in range(cfg["s"]):
                phi, aphi = ctx.basis(res.pid, d)
                if not margin_safe(res.info, d):
                    nskip += 1; continue
                err = np.linalg.norm(phi - res.basis[d])
                tol = 1e-10
                if err > tol and res.info["slod"]:
                    tol = max(tol, 50.0 * selection_sensitivity(res.info, d))
                worst = max(worst, err)
                nbad += err > tol
                if res.info["slod"] and int(ctx.diagnostics(res.pid, d)[1]) != res.info["trunc_steps"][d]:
                    nbad += 1
        print(cfg, "patches %d  worst err %.2e  bad %d  skipped %d  kernel ms %s  (%.1f s)" %
              (npch, worst, nbad, nskip, [round(float(x), 2) for x in ctx.timings()[:5]], time.time() - t0), flush=True)
        ctx.close()
    except Exception as e:   # noqa: BLE001
        print(cfg, "EXCEPTION", repr(e)[:200], flush=True)
