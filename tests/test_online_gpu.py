"""GPU parity tests of the online phase: coarse right-hand side C^T f, coarse solve and
prolongation C u through the C ABI, against the oracle's restatement of LOD::solve (source/LOD.cc:975-1001, :1251).

Two kinds of checks:
* operator level: C built on the host from the basis the GPU itself returned, so that the three kernels are tested
  in isolation at rounding level (1e-12) whatever the conditioning of the basis selection is;
* end to end against the oracle's own pipeline (oracle basis -> K -> CG + SSOR(1.2) / direct solve), at the north
  star's 1e-9 relative for the coarse and fine solutions, on configurations whose selection is well conditioned
  (LOD branch and interior-dominated SLOD cases; see tests/test_parity_gpu.py for why boundary patches are not).
Both preconditioners stop on the same ReductionControl rule, so the solutions agree to the tolerance, not the step
counts."""
import os
import re
import sys

import numpy as np
import pytest
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
from parity_common import build_pair, pkg  # noqa: E402
from oracle.slod_oracle import GlibcRand, reference_random_table  # noqa: E402

pytestmark = pytest.mark.gpu


def _global_lex_nodes(info, n, G):
    lo, m = info["lo"], info["m"]
    dim = len(lo)
    p = [mm * n + 1 for mm in m]
    idx = np.indices(p[::-1]).reshape(dim, -1)[::-1]  # x fastest
    g = np.zeros(idx.shape[1], dtype=np.int64)
    mul = 1
    for a in range(dim):
        g += (idx[a] + lo[a] * n) * mul
        mul *= G
    return g


def _host_C(ctx, n, s, ref):
    """C (fine x coarse, lexicographic fine numbering) from the basis the GPU returned."""
    G = 2 ** ref * n + 1
    rows, cols, vals = [], [], []
    for pid in range(ctx.n_patches):
        g = _global_lex_nodes(ctx.patch_info(pid), n, G)
        gd = (g[:, None] * s + np.arange(s)[None, :]).ravel()
        for d in range(s):
            phi, _ = ctx.basis(pid, d)
            rows.append(gd)
            cols.append(np.full(gd.shape, s * pid + d))
            vals.append(phi[: gd.size])
    return sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(s * G ** ctx.dim, s * ctx.n_patches))


def _forcing(s):
    if s == 1:
        return lambda p: (1.0 + np.sin(3.0 * p[:, [0]]) * np.cos(2.0 * p[:, [1]]) + (p[:, [2]] if p.shape[1] == 3 else 0.0))
    return lambda p: np.concatenate([1.0 + p[:, [0]], np.cos(2.0 * p[:, [1]]) - 0.3], axis=1)


OPERATOR_CASES = [
    dict(dim=2, s=1, ref=3, n=2, ell=1),
    dict(dim=2, s=1, ref=4, n=2, ell=2),
    dict(dim=2, s=2, ref=3, n=2, ell=1),
    dict(dim=2, s=1, ref=3, n=4, ell=1),
    dict(dim=2, s=1, ref=3, n=2, ell=0),
    dict(dim=3, s=1, ref=2, n=2, ell=1),
    dict(dim=3, s=1, ref=3, n=2, ell=2, kind="uniform1e4", seed=3001),
]


@pytest.mark.parametrize("case", OPERATOR_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_online_operators(case):
    ctx, orc = build_pair(**case)
    n, s = case["n"], case["s"]
    ctx.compute_basis()
    ctx.assemble_coarse()
    C = _host_C(ctx, n, s, case["ref"])
    assert ctx.n_fine == C.shape[0]
    rng = np.random.default_rng(7)
    # b = C^T f
    f = orc.fem_rhs(_forcing(s))
    b = ctx.coarse_rhs(f)
    b_ref = C.T @ f
    assert np.abs(b - b_ref).max() <= 1e-12 * np.abs(b_ref).max()
    f2 = rng.standard_normal(ctx.n_fine)          # not zero on the boundary: the basis is
    b2 = ctx.coarse_rhs(f2)
    assert np.abs(b2 - C.T @ f2).max() <= 1e-12 * np.abs(C.T @ f2).max()
    # u_h = C u
    u = rng.standard_normal(C.shape[1])
    uh = ctx.prolongate(u)
    assert np.abs(uh - C @ u).max() <= 1e-12 * np.abs(C @ u).max()
    # K u = b against a direct solve with the matrix the GPU assembled
    rowptr, col, val = ctx.coarse_csr()
    K = sp.csr_matrix((val.copy(), col.copy(), rowptr.copy()), shape=(C.shape[1],) * 2)
    x, steps, res = ctx.coarse_solve(b, max_steps=5000, tolerance=0.0, reduction=1e-13)
    x_ref, _ = orc.solve_coarse(K, b, direct=True)
    assert steps > 0
    assert res <= 1e-13 * np.linalg.norm(b) * 1.0000001
    assert np.linalg.norm(K @ x - b) <= 1e-11 * np.linalg.norm(b)      # recurrence residual vs true residual
    assert np.linalg.norm(x - x_ref) <= 1e-9 * np.linalg.norm(x_ref)
    # bit-reproducible
    x2, steps2, _ = ctx.coarse_solve(b, max_steps=5000, tolerance=0.0, reduction=1e-13)
    assert steps2 == steps and np.array_equal(x, x2)
    # linearity of the solve
    x3, _, _ = ctx.coarse_solve(4.0 * b, max_steps=5000, tolerance=0.0, reduction=1e-13)
    assert np.array_equal(x3, 4.0 * x)
    ctx.close()


END_TO_END = [
    dict(dim=2, s=1, ref=3, n=2, ell=1, stabilize=False),     # LOD branch
    dict(dim=2, s=1, ref=4, n=2, ell=3, stabilize=False),
    dict(dim=2, s=2, ref=3, n=2, ell=1, stabilize=False),
    dict(dim=3, s=1, ref=2, n=2, ell=1, stabilize=False),
    # SLOD branch.  With l = 1 the selection is well conditioned and the coarse vector agrees at
    # 1e-9 too; with l = 2 the boundary patches' basis functions are only determined to ~1e-6 (cond(G) eps, see
    # tests/test_fullsize_parity_gpu.py) and so is the coarse vector -- its coefficients refer to slightly different
    # bases -- while the fine solution C u, the quantity the north star bounds, is insensitive: 1e-9 everywhere.
    dict(dim=2, s=1, ref=4, n=2, ell=1, stabilize=True),
    dict(dim=2, s=2, ref=3, n=2, ell=1, stabilize=True),
    dict(dim=3, s=1, ref=2, n=2, ell=1, stabilize=True),
    dict(dim=2, s=1, ref=4, n=2, ell=2, stabilize=True, coarse_tol=1e-5),
    dict(dim=3, s=1, ref=3, n=2, ell=2, stabilize=True, coarse_tol=1e-5),
]


@pytest.mark.parametrize("case", END_TO_END, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_solution_against_oracle(case):
    """Coarse and fine solution of the whole chain against the oracle's (reference preconditioner and direct solve):
    1e-9 relative, the north star's bound."""
    case = dict(case)
    coarse_tol = case.pop("coarse_tol", 1e-9)
    ctx, orc = build_pair(**case)
    s = case["s"]
    ctx.compute_basis()
    ctx.assemble_coarse()
    orc.compute_basis()
    K, C, _ = orc.assemble_global_matrix()
    f = orc.fem_rhs(_forcing(s))
    b_ref = C.T @ f
    u_ssor, it_ssor = orc.solve_coarse(K, b_ref, max_steps=2000, tolerance=0.0, reduction=1e-13)
    u_dir, _ = orc.solve_coarse(K, b_ref, direct=True)
    assert np.linalg.norm(u_ssor - u_dir) <= 1e-9 * np.linalg.norm(u_dir)
    b = ctx.coarse_rhs(f)
    assert np.linalg.norm(b - b_ref) <= max(1e-10, 0.1 * coarse_tol) * np.linalg.norm(b_ref)
    u, steps, _ = ctx.coarse_solve(b, max_steps=5000, tolerance=0.0, reduction=1e-13)
    assert np.linalg.norm(u - u_dir) <= coarse_tol * np.linalg.norm(u_dir)
    assert np.linalg.norm(u - u_ssor) <= coarse_tol * np.linalg.norm(u_dir)
    uh = ctx.prolongate(u)
    uh_ref = C @ u_dir
    assert np.linalg.norm(uh - uh_ref) <= 1e-9 * np.linalg.norm(uh_ref)
    ctx.close()


@pytest.mark.parametrize("dim,s,ref,bounds", [
    (2, 1, 4, [0.3, 6e-3, 2e-5]),      # measured on B200: 2.0e-1, 3.7e-3, 7.8e-6 (the oracle's own chain gives the same)
    (2, 2, 3, [0.25, 7e-3]),           # 1.5e-1, 4.3e-3
    (3, 1, 3, [0.15, 1e-3]),           # 9.1e-2, 6.2e-4
    (3, 1, 4, [None, 3e-3]),           # ell = 2 only: 1.6e-3 (4096 patches of the cfg 4 shape)
])
def test_slod_solution_converges_to_fine_fem(dim, s, ref, bounds):
    """The property the reference reports (error_LOD_FEMh, source/LOD.cc:1252) and the one check of the SLOD branch and of
    the 3-D extension that needs no reference run: for a forcing that is constant on the coarse cells the SLOD solution
    of the GPU chain differs from the fine-scale FEM solution only by the localization error, which decays
    super-exponentially with the oversampling (profiles/r01_lod_vs_fem.txt)."""
    from lod_vs_fem import lod_vs_fem
    prev = None
    for ell, bound in zip((1, 2, 3), bounds):
        if bound is None:
            continue
        energy, l2, steps = lod_vs_fem(dim, s, ref, ell)
        assert energy < bound and l2 < bound, (ell, energy, l2)
        if prev is not None:
            assert energy < prev / 20.0
        prev = energy


def test_poisson_lod_example_rhs_norm(golden_dir):
    """tests/Poisson_LOD_Example.output: `rhs l2 norm = 0.0808367`, `size of u 16` -- now computed by slod_coarse_rhs."""
    txt = open(os.path.join(golden_dir, "Poisson_LOD_Example.output")).read()
    rhs_norm = float(re.search(r"\n\s+rhs l2 norm = ([0-9.]+)", txt).group(1))
    size_u = int(re.search(r"size of u (\d+)", txt).group(1))
    rng = GlibcRand()
    for _ in range(12):
        rng.rand()
    tab = reference_random_table(2, 1, 100, 8, rng)
    ctx, orc = build_pair(dim=2, s=1, ref=2, n=2, ell=1, stabilize=False, r=8, quirk=True, tables=[tab])
    ctx.compute_basis()
    ctx.assemble_coarse()
    f = orc.fem_rhs_constant_one()
    assert np.abs(f - orc.fem_rhs(lambda p: np.ones((len(p), 1)))).max() < 1e-16
    b = ctx.coarse_rhs(f)
    assert float("%g" % np.linalg.norm(b)) == rhs_norm
    # deal.II's ReductionControl defaults (100 steps, tolerance 1e-10, reduction 1e-2).  With quirk B (presaved patch
    # matrix, source/LOD.cc:354-362) the reference's K is not symmetric, so CG is only good for this loose reduction --
    # in the reference too.
    u, steps, res = ctx.coarse_solve(b, max_steps=100, tolerance=1e-10, reduction=1e-2)
    assert res <= 1e-2 * np.linalg.norm(b)
    assert u.size == size_u and 0 < steps <= 100
    ctx.close()


def test_solver_control_errors():
    ctx, _ = build_pair(dim=2, s=1, ref=3, n=2, ell=1)
    with pytest.raises(pkg.SlodError):           # call order
        ctx.coarse_rhs(np.zeros(ctx.n_fine))
    ctx.compute_basis()
    with pytest.raises(pkg.SlodError):
        ctx.coarse_solve(np.ones(ctx.n_patches))
    ctx.assemble_coarse()
    b = ctx.coarse_rhs(np.ones(ctx.n_fine))
    with pytest.raises(pkg.SlodError) as ei:     # SolverControl::NoConvergence
        ctx.coarse_solve(b, max_steps=1, tolerance=0.0, reduction=1e-14)
    assert "did not converge" in str(ei.value)
    u, steps, _ = ctx.coarse_solve(np.zeros(ctx.n_patches))      # zero rhs: converged at step 0
    assert steps == 0 and not u.any()
    ctx.close()


def test_checkpoint_roundtrip(tmp_path):
    """slod_save_state / slod_load_state: a fresh handle with the same parameters serves the basis, the
    coarse matrix and the online phase from the file, bit-identically, without recomputing; wrong parameters are refused."""
    ctx, _ = build_pair(dim=2, s=1, ref=4, n=2, ell=2)
    ctx.compute_basis()
    ctx.assemble_coarse()
    f = np.random.default_rng(1).standard_normal(ctx.n_fine)
    b = ctx.coarse_rhs(f)
    u, steps, _ = ctx.coarse_solve(b, max_steps=5000, tolerance=0.0, reduction=1e-12)
    uh = ctx.prolongate(u)
    path = str(tmp_path / "offline.slod")
    ctx.save_state(path)
    phi, aphi = (x.copy() for x in ctx.all_basis())
    val = ctx.coarse_csr()[2].copy()
    ctx.close()
    ctx2 = pkg.SlodContext(dim=2, spacedim=1, n_global_refinements=4, n_subdivisions=2, oversampling=2, stabilize=True)
    with pytest.raises(pkg.SlodError):
        ctx2.coarse_rhs(f)                      # nothing computed yet
    ctx2.load_state(path)
    p2, a2 = ctx2.all_basis()
    assert np.array_equal(p2, phi) and np.array_equal(a2, aphi)
    assert np.array_equal(ctx2.coarse_csr()[2], val)
    b2 = ctx2.coarse_rhs(f)
    u2, steps2, _ = ctx2.coarse_solve(b2, max_steps=5000, tolerance=0.0, reduction=1e-12)
    assert np.array_equal(b2, b) and steps2 == steps and np.array_equal(u2, u)
    assert np.array_equal(ctx2.prolongate(u2), uh)
    launches = ctx2.launch_count
    ctx2.close()
    other = pkg.SlodContext(dim=2, spacedim=1, n_global_refinements=4, n_subdivisions=2, oversampling=1, stabilize=True)
    with pytest.raises(pkg.SlodError):
        other.load_state(path)
    other.close()
    assert launches < 400                       # CG launches only: no patch kernels ran on the second handle
