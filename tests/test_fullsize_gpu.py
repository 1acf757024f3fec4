"""BASELINE.json configurations at FULL size through size-independent properties (the oracle would need hours there):

* K = C^T A C is symmetric -- phi_p . (A phi_q) == phi_q . (A phi_p) -- which ties together the patch solves, the
  selection, the matrix-free A phi and the coarse-matrix kernel of two DIFFERENT patches for every entry;
* every basis function has unit 2-norm and vanishes on its patch boundary; no patch reports a numerical status;
* with a constant coefficient all full-size interior patches are translates of each other: their basis functions
  must be bit-identical (exercises the closed-form geometry of 2^15 / 2^16 patches);
* scaling the coefficient by 4 leaves phi bit-identical and scales K by exactly 4."""
import importlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
from parity_common import make_tables  # noqa: E402

pkg = importlib.import_module("dealii-slod_b200")
pytestmark = pytest.mark.gpu

CONFIGS = {
    "cfg2_diffusion2d_256": dict(dim=2, s=1, ref=8, n=2, ell=2, r=8, kind="uniform100", seed=1234),
    "cfg3_elasticity2d_128": dict(dim=2, s=2, ref=7, n=2, ell=1, r=6, kind="uniform100", seed=2001),
    "cfg4_diffusion3d_32": dict(dim=3, s=1, ref=5, n=2, ell=2, r=6, kind="uniform1e4", seed=3001),
}


def _run(c, tables):
    ctx = pkg.SlodContext(dim=c["dim"], spacedim=c["s"], n_global_refinements=c["ref"], n_subdivisions=c["n"],
                          oversampling=c["ell"], stabilize=True, problem=0 if c["s"] == 1 else 1)
    for f, t in enumerate(tables):
        ctx.set_coefficient(f, c["r"], t)
    ctx.compute_basis()          # raises SLOD_ERR_NUMERIC if any patch reports a status
    ctx.assemble_coarse()
    return ctx


def _csr_to_scipy(ctx):
    import scipy.sparse as sp
    rowptr, col, val = ctx.coarse_csr()
    n = rowptr.size - 1
    return sp.csr_matrix((val.copy(), col.copy(), rowptr.copy()), shape=(n, n))


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_full_size_properties(name):
    c = CONFIGS[name]
    tables = make_tables(c["dim"], c["s"], c["r"], c["kind"], c["seed"])
    ctx = _run(c, tables)
    K = _csr_to_scipy(ctx)
    kmax = np.abs(K.data).max()
    asym = abs(K - K.T)
    assert asym.max() <= 5e-9 * kmax, (name, asym.max() / kmax)
    # the symmetric pattern is geometric: (p, q) stored iff (q, p) stored
    assert (K != 0).nnz > 0 and ((abs(K) > 0) != (abs(K.T) > 0)).nnz <= 0.001 * K.nnz
    phi, aphi = ctx.all_basis()
    nrm = np.linalg.norm(phi.reshape(phi.shape[0] * phi.shape[1], -1), axis=1)
    assert np.abs(nrm - 1.0).max() < 1e-12
    assert np.isfinite(aphi).all()
    # an interior full-size patch: phi vanishes on the whole boundary of its node box
    N = 2 ** c["ref"]
    centre = [N // 2] * c["dim"]
    pid = sum(((centre[a] >> b) & 1) << (c["dim"] * b + a) for b in range(c["ref"]) for a in range(c["dim"]))
    info = ctx.patch_info(pid)
    p = [m * c["n"] + 1 for m in info["m"]]
    f = phi[pid, 0, : c["s"] * int(np.prod(p))].reshape(p[::-1] + [c["s"]])
    for ax in range(c["dim"]):
        assert np.all(np.take(f, 0, axis=ax) == 0.0) and np.all(np.take(f, -1, axis=ax) == 0.0)


def test_constant_coefficient_translation_invariance():
    c = CONFIGS["cfg4_diffusion3d_32"]
    ones = [np.ones((2 ** c["r"]) ** c["dim"])]
    ctx = _run(c, ones)
    phi, aphi = ctx.all_basis()
    N, ell = 2 ** c["ref"], c["ell"]
    ids = np.arange(N ** 3)
    x = np.zeros_like(ids)
    y = np.zeros_like(ids)
    z = np.zeros_like(ids)
    for b in range(c["ref"]):
        x |= ((ids >> (3 * b)) & 1) << b
        y |= ((ids >> (3 * b + 1)) & 1) << b
        z |= ((ids >> (3 * b + 2)) & 1) << b
    full = np.ones(ids.shape, dtype=bool)
    for co in (x, y, z):
        full &= (co >= ell + 1) & (co <= N - 2 - ell)
    # full-size patches none of whose sides lies on the domain boundary (there the side would be a Dirichlet side,
    # id 0, instead of a patch side, id 99: a different problem) are exact translates of each other
    ref_pid = ids[full][0]
    assert np.array_equal(phi[full], np.broadcast_to(phi[ref_pid], phi[full].shape))
    assert np.array_equal(aphi[full], np.broadcast_to(aphi[ref_pid], aphi[full].shape))
    assert full.sum() == (N - 2 * ell - 2) ** 3


def test_scaling_by_four_is_exact_at_full_size():
    c = CONFIGS["cfg4_diffusion3d_32"]
    tabs = make_tables(c["dim"], c["s"], c["r"], c["kind"], c["seed"])
    ctx1 = _run(c, tabs)
    p1, a1 = (x.copy() for x in ctx1.all_basis())
    v1 = ctx1.coarse_csr()[2].copy()
    ctx1.close()
    ctx2 = _run(c, [4.0 * t for t in tabs])
    p2, a2 = ctx2.all_basis()
    assert np.array_equal(p1, p2)
    assert np.array_equal(4.0 * a1, a2)
    assert np.array_equal(4.0 * v1, ctx2.coarse_csr()[2])


@pytest.mark.parametrize("name,ell,bound", [("cfg2_diffusion2d_256", 2, 2e-2), ("cfg3_elasticity2d_128", 2, 0.2),
                                            ("cfg4_diffusion3d_32", 2, 1e-2)])
def test_full_size_solution_against_fine_fem(name, ell, bound):
    """The whole chain at BASELINE size, all on the GPU: basis -> K -> C^T f -> coarse CG -> C u against the fine FEM
    solution of the same problem (slod_fem_solve), in the energy norm (slod_fine_norms) -- the reference's
    `SLOD vs reference FEM(h)` table (source/LOD.cc:1252, 1463-1465).  For a forcing that is constant on the coarse cells
    only the localization error remains, which behaves like sigma(l) / H: measured 2.7e-3 (cfg 2), 5.7e-2 (cfg 3 shape
    with l = 2; the reference's default l = 1 gives 0.8 on a 128^2 mesh -- l must grow like log(1/H)), 3.5e-3 (cfg 4)."""
    c = dict(CONFIGS[name], ell=ell)
    tables = make_tables(c["dim"], c["s"], c["r"], c["kind"], c["seed"])
    ctx = _run(c, tables)
    G = 2 ** c["ref"] * c["n"] + 1
    w1 = np.full(G, 1.0 / (G - 1))
    w1[0] = w1[-1] = 0.0
    F = w1
    for _ in range(c["dim"] - 1):
        F = np.multiply.outer(w1, F)
    F = (F.ravel()[:, None] * np.array([1.0] if c["s"] == 1 else [1.0, -0.5])[None, :]).ravel()
    u, steps, _ = ctx.coarse_solve(ctx.coarse_rhs(F), max_steps=50000, tolerance=0.0, reduction=1e-11)
    u_lod = ctx.prolongate(u)
    u_fem, fsteps, _ = ctx.fem_solve(F, max_steps=500000, tolerance=0.0, reduction=1e-11)
    _, _, e_en = ctx.fine_norms(u_lod - u_fem)
    _, _, n_en = ctx.fine_norms(u_fem)
    print(f"{name}: coarse CG {steps} steps, fine CG {fsteps} steps, relative energy error {e_en / n_en:.3e}")
    assert e_en / n_en < bound
    ctx.close()
