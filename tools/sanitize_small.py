"""Tiny end-to-end runs (2-D diffusion, 2-D elasticity, 3-D diffusion with the cfg-4 patch shape) for compute-sanitizer."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("dealii-slod_b200")
for dim, s, ref, ell in ((2, 1, 3, 1), (2, 2, 3, 1), (3, 1, 3, 2)):
    r = ref + 1
    rs = np.random.default_rng(3)
    ctx = pkg.SlodContext(dim=dim, spacedim=s, n_global_refinements=ref, n_subdivisions=2, oversampling=ell,
                          stabilize=True, problem=0 if s == 1 else 1)
    for f in range(s):
        ctx.set_coefficient(f, r, 1.0 + 99.0 * rs.random((2 ** r) ** dim))
    ctx.compute_basis()
    ctx.assemble_coarse()
    rowptr, col, val = ctx.coarse_csr()
    phi, aphi = ctx.all_basis()
    print(dim, s, ref, ell, "ok", float(np.abs(val).max()), float(np.abs(phi).max()))
    ctx.close()
