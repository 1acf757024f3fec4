"""GPU parity tests of the fine-scale reference problem: slod_fem_solve against the oracle's
restatement of assemble_and_solve_fem_problem (source/LOD.cc:1004-1094; sparse direct solve of the same Q_iso_Q1
stiffness matrix) and slod_fine_norms against the exact mass / Laplace / stiffness quadratic forms.  Tolerances:
1e-9 relative for the solution (the north star's bound for fine solutions), 1e-12 for the norms."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
from parity_common import build_pair, pkg  # noqa: E402

pytestmark = pytest.mark.gpu

CASES = [
    dict(dim=2, s=1, ref=3, n=2, ell=1),
    dict(dim=2, s=1, ref=4, n=2, ell=1, kind="binary1e4", seed=1235),     # high contrast
    dict(dim=2, s=1, ref=3, n=4, ell=1),
    dict(dim=2, s=1, ref=2, n=2, ell=1, r=4),                             # coefficient finer than the sub-cells (Gauss-point table)
    dict(dim=2, s=2, ref=3, n=2, ell=1),                                  # elasticity
    dict(dim=3, s=1, ref=2, n=2, ell=1),
    dict(dim=3, s=1, ref=3, n=2, ell=1, kind="uniform1e4", seed=3001),
]


def _forcing(s, dim):
    if s == 1:
        return lambda p: 1.0 + np.sin(3.0 * p[:, [0]]) * np.cos(2.0 * p[:, [1]]) + (p[:, [2]] if dim == 3 else 0.0)
    return lambda p: np.concatenate([1.0 + p[:, [0]], np.cos(2.0 * p[:, [1]]) - 0.3], axis=1)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_fem_solve_and_norms(case):
    ctx, orc = build_pair(**case)
    s, dim = case["s"], case["dim"]
    F = orc.fem_rhs(_forcing(s, dim))
    u_ref, A = orc.fem_solve(F)
    u, steps, res = ctx.fem_solve(F, max_steps=100000, tolerance=0.0, reduction=1e-13)
    assert steps > 0 and res <= 1.0000001e-13 * np.linalg.norm(F)
    assert np.linalg.norm(u - u_ref) <= 1e-9 * np.linalg.norm(u_ref)
    # boundary rows of the right-hand side are constrained away: same answer with garbage there
    G = 2 ** case["ref"] * case["n"] + 1
    idx = np.indices((G,) * dim).reshape(dim, -1)
    on_boundary = np.repeat(((idx == 0) | (idx == G - 1)).any(axis=0), s)
    assert not u[on_boundary].any()
    F2 = F.copy()
    F2[on_boundary] = 7.0
    u2, steps2, _ = ctx.fem_solve(F2, max_steps=100000, tolerance=0.0, reduction=1e-13)
    assert steps2 == steps and np.array_equal(u, u2)
    # norms
    M, L = orc.fine_norm_matrices()
    v = np.random.default_rng(3).standard_normal(ctx.n_fine)
    l2, h1, en = ctx.fine_norms(v)
    assert abs(l2 - np.sqrt(v @ (M @ v))) <= 1e-12 * l2
    assert abs(h1 - np.sqrt(v @ (L @ v))) <= 1e-12 * h1
    assert abs(en - np.sqrt(v @ (A @ v))) <= 1e-12 * en
    # the reference's own error-table quadrature (ParsedConvergenceTable::difference, source/LOD.cc:1252): QGauss on the
    # coarse cells -- against the oracle's restatement, and close to (but not equal to) the exact norms of a smooth field
    if ctx.n_patches <= 512:
        r_l2, r_inf, r_h1 = ctx.fine_norms_reference(v)
        o_l2, o_inf, o_h1 = orc.reference_error_norms(v)
        assert abs(r_l2 - o_l2) <= 1e-12 * o_l2 and abs(r_h1 - o_h1) <= 1e-12 * o_h1 and abs(r_inf - o_inf) <= 1e-12 * o_inf
    xs = np.linspace(0.0, 1.0, G)
    smooth = np.sin(2.0 * xs)
    for _ in range(dim - 1):
        smooth = np.multiply.outer(np.cos(xs), smooth)
    smooth = np.repeat(smooth.ravel(), s)
    e_l2, e_h1s, _ = ctx.fine_norms(smooth)
    r_l2, r_inf, r_h1 = ctx.fine_norms_reference(smooth)
    assert abs(r_l2 - e_l2) <= 1e-3 * e_l2 and abs(r_h1 - np.hypot(e_l2, e_h1s)) <= 1e-2 * r_h1
    assert r_inf <= np.abs(smooth).max() * (1 + 1e-12)
    ctx.close()


def test_lod_error_table_through_the_c_abi():
    """The reference's end product (compare_lod_with_fem, source/LOD.cc:1240-1260) entirely on the GPU: SLOD solution
    vs fine FEM solution in the L2 / H1 / energy norms."""
    ctx, orc = build_pair(dim=2, s=1, ref=4, n=2, ell=2)
    F = orc.fem_rhs(lambda p: np.ones((len(p), 1)))
    ctx.compute_basis()
    ctx.assemble_coarse()
    u, _, _ = ctx.coarse_solve(ctx.coarse_rhs(F), max_steps=5000, tolerance=0.0, reduction=1e-12)
    u_lod = ctx.prolongate(u)
    u_fem, _, _ = ctx.fem_solve(F, max_steps=100000, tolerance=0.0, reduction=1e-12)
    e_l2, e_h1, e_en = ctx.fine_norms(u_lod - u_fem)
    n_l2, n_h1, n_en = ctx.fine_norms(u_fem)
    assert e_en / n_en < 6e-3 and e_l2 / n_l2 < 1e-3 and e_h1 / n_h1 < 2e-2      # measured 3.7e-3 (energy)
    u_ref, A = orc.fem_solve(F)
    e = u_lod - u_ref
    assert abs(e_en / n_en - np.sqrt(e @ (A @ e)) / np.sqrt(u_ref @ (A @ u_ref))) < 1e-6
    ctx.close()


def test_fem_solve_errors():
    ctx, _ = build_pair(dim=2, s=1, ref=3, n=2, ell=1)
    with pytest.raises(pkg.SlodError) as ei:
        ctx.fem_solve(np.ones(ctx.n_fine), max_steps=2, tolerance=0.0, reduction=1e-14)
    assert "did not converge" in str(ei.value)
    with pytest.raises(ValueError):
        ctx.fem_solve(np.ones(3))
    u, steps, _ = ctx.fem_solve(np.zeros(ctx.n_fine))
    assert steps == 0 and not u.any()
    ctx.close()
