"""SLOD solution of the GPU chain (basis -> K -> C^T f -> CG -> C u) against the fine-scale FEM solution of the same
problem (the quantity the reference reports as error_LOD_FEMh, source/LOD.cc:1252): relative energy and l2 errors.
Test infrastructure: the FEM solution comes from the oracle (sparse direct solve on the CPU)."""
import sys
import numpy as np
from parity_common import build_pair


def lod_vs_fem(dim, s, ref, ell, stabilize=True, n=2, kind="uniform100", seed=11, rhs=None):
    ctx, orc = build_pair(dim=dim, s=s, ref=ref, n=n, ell=ell, stabilize=stabilize, kind=kind, seed=seed)
    rhs = ([1.0] if s == 1 else [1.0, -0.5]) if rhs is None else rhs
    F = orc.fem_rhs(lambda p: np.tile(rhs, (len(p), 1)))
    u_fem, A = orc.fem_solve(F)
    ctx.compute_basis()
    ctx.assemble_coarse()
    u, steps, _ = ctx.coarse_solve(ctx.coarse_rhs(F), max_steps=20000, tolerance=0.0, reduction=1e-12)
    e = ctx.prolongate(u) - u_fem
    ctx.close()
    return (float(np.sqrt(e @ (A @ e)) / np.sqrt(u_fem @ (A @ u_fem))), float(np.linalg.norm(e) / np.linalg.norm(u_fem)),
            steps)


if __name__ == "__main__":
    cases = [(2, 1, 4, 1), (2, 1, 4, 2), (2, 1, 4, 3), (2, 1, 5, 2), (2, 2, 3, 1), (2, 2, 3, 2), (2, 2, 4, 2),
             (3, 1, 2, 1), (3, 1, 3, 1), (3, 1, 3, 2), (3, 1, 4, 1), (3, 1, 4, 2)]
    for dim, s, ref, ell in cases:
        for kind in ("uniform100", "binary1e4"):
            en, l2, steps = lod_vs_fem(dim, s, ref, ell, kind=kind)
            print(f"dim {dim} s {s} ref {ref} ell {ell} {kind:11s} energy {en:.3e}  l2 {l2:.3e}  cg steps {steps}", flush=True)
