"""Relative energy error of the SLOD solution against the fine FEM solution, everything on the GPU (no oracle)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from parity_common import make_tables
pkg = importlib.import_module("dealii-slod_b200")


def run(dim, s, ref, ell, kind="uniform100", seed=11, n=2):
    r = min(ref + 1, 8 if dim == 2 else 6)
    ctx = pkg.SlodContext(dim=dim, spacedim=s, n_global_refinements=ref, n_subdivisions=n, oversampling=ell,
                          stabilize=True, problem=0 if s == 1 else 1)
    for f, t in enumerate(make_tables(dim, s, r, kind, seed)):
        ctx.set_coefficient(f, r, t)
    ctx.compute_basis(); ctx.assemble_coarse()
    G = 2 ** ref * n + 1
    w1 = np.full(G, 1.0 / (G - 1)); w1[0] = w1[-1] = 0.0
    F = w1
    for _ in range(dim - 1):
        F = np.multiply.outer(w1, F)
    F = (F.ravel()[:, None] * np.array([1.0] if s == 1 else [1.0, -0.5])[None, :]).ravel()
    u, steps, _ = ctx.coarse_solve(ctx.coarse_rhs(F), max_steps=50000, tolerance=0.0, reduction=1e-11)
    u_lod = ctx.prolongate(u)
    u_fem, fsteps, _ = ctx.fem_solve(F, max_steps=500000, tolerance=0.0, reduction=1e-11)
    e = ctx.fine_norms(u_lod - u_fem); nn = ctx.fine_norms(u_fem)
    ctx.close()
    return e[2] / nn[2], e[0] / nn[0], steps, fsteps


if __name__ == "__main__":
    for dim, s, ref, ell in [(2, 2, 4, 1), (2, 2, 5, 1), (2, 2, 6, 1), (2, 2, 7, 1), (2, 2, 5, 2), (2, 2, 6, 2), (2, 2, 7, 2),
                             (2, 1, 5, 1), (2, 1, 6, 1), (2, 1, 6, 2), (2, 1, 7, 2), (2, 1, 8, 2), (2, 1, 7, 3), (3, 1, 5, 2)]:
        en, l2, st, fst = run(dim, s, ref, ell)
        print(f"dim {dim} s {s} ref {ref} ell {ell}: energy {en:.3e} l2 {l2:.3e} coarse CG {st} fine CG {fst}", flush=True)
