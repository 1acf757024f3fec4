"""End-to-end chain (basis -> K -> C^T f -> CG -> C u) against the fine FEM solution on configurations that take the
fall-back kernels (many subdivisions), plus the presaved-matrix quirk in the SLOD branch with the oracle in patch order."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from lod_vs_fem import lod_vs_fem
from parity_common import build_pair

for dim, s, ref, ell, n in [(2, 1, 4, 2, 8), (2, 1, 4, 2, 4), (3, 1, 3, 1, 4), (3, 1, 2, 2, 4), (2, 2, 3, 2, 4)]:
    try:
        en, l2, steps = lod_vs_fem(dim, s, ref, ell, n=n)
        print(f"dim {dim} s {s} ref {ref} ell {ell} n {n}: energy {en:.3e} l2 {l2:.3e} cg steps {steps}", flush=True)
    except Exception as e:   # noqa: BLE001
        print(dim, s, ref, ell, n, "EXCEPTION", repr(e)[:200], flush=True)
ctx, orc = build_pair(dim=2, s=1, ref=3, n=2, ell=1, quirk=True)
ctx.compute_basis()
orc.compute_basis()
print("quirk + SLOD, all patches: max |dphi| =", max(np.linalg.norm(ctx.basis(r.pid)[0] - r.basis[0]) for r in orc.patches))
