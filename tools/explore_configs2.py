"""More unusual paths against the oracle: LOD branch in 3-D, oversampling 0, coefficient finer than the sub-cells (Gauss-point
mode, 2-D), small forced chunks with the staged dense kernel, the presaved-matrix quirk."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from parity_common import build_pair, margin_safe, selection_sensitivity

def run(tag, env=None, **cfg):
    for k, v in (env or {}).items():
        os.environ[k] = v
    t0 = time.time()
    try:
        ctx, orc = build_pair(**cfg)
        ctx.compute_basis(); ctx.assemble_coarse()
        npch = ctx.n_patches
        # the presaved-matrix quirk is stateful (the FIRST full-size patch in patch order is saved): the oracle must see
        # the patches in order then, a sample would presave a different patch
        pids = list(range(npch)) if cfg.get("quirk") else sorted({0, npch // 3, npch // 2 + 1, npch - 1})
        orc.compute_basis(pids)
        worst, nbad, nskip = 0.0, 0, 0
        for res in orc.patches:
            for d in range(cfg.get("s", 1)):
                phi, aphi = ctx.basis(res.pid, d)
                if not margin_safe(res.info, d):
                    nskip += 1; continue
                err = np.linalg.norm(phi - res.basis[d])
                tol = 1e-10
                if err > tol and res.info["slod"]:
                    tol = max(tol, 50.0 * selection_sensitivity(res.info, d))
                worst = max(worst, err); nbad += err > tol
        print(tag, cfg, "patches %d worst %.2e bad %d skipped %d (%.1f s)" % (npch, worst, nbad, nskip, time.time() - t0), flush=True)
        ctx.close()
    except Exception as e:   # noqa: BLE001
        print(tag, cfg, "EXCEPTION", repr(e)[:200], flush=True)
    for k in (env or {}):
        os.environ.pop(k, None)

run("LOD-3D", dim=3, s=1, ref=3, n=2, ell=2, stabilize=False)
run("ell0-3D", dim=3, s=1, ref=2, n=2, ell=0)
run("ell0-elast", dim=2, s=2, ref=3, n=2, ell=0)
run("gauss-2D", dim=2, s=1, ref=3, n=2, ell=1, r=6)
run("gauss-2D-elast", dim=2, s=2, ref=3, n=2, ell=1, r=5)
run("chunk7-3D", env={"SLOD_CHUNK": "7"}, dim=3, s=1, ref=3, n=2, ell=2)
run("chunk3-2D", env={"SLOD_CHUNK": "3"}, dim=2, s=1, ref=4, n=2, ell=2)
run("quirk", dim=2, s=1, ref=3, n=2, ell=1, quirk=True)
run("coarse-r", dim=3, s=1, ref=3, n=2, ell=2, r=2, kind="binary1e4", seed=11)
