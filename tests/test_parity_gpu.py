"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances (fp64): every stage up to the Gram matrix is compared at 1e-10 relative on every case; the selected
basis and A*basis at 1e-10 wherever that is reachable, and otherwise (boundary-touching patches, where the
reference's Gram/thresholded-SVD/truncation rule, source/LOD.cc:656-725, is ill conditioned) at 50x the
change a 4-ulp perturbation of the Gram matrix causes in the ORACLE's own answer -- no fp64 implementation
can agree better than that.  Truncation step counts must agree exactly, except on
patches where a discontinuous decision of the rule is within rounding of flipping (||d||_inf within 1e-3 of 0.5, or a
singular value within two decades of the 1e-15 sigma_0 threshold): those are skipped, the reference's own answer there
depends on LAPACK's rounding.
Integer maps and the CSR pattern are bit-exact."""
import os
import re
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
from parity_common import (EPS, build_pair, cond_eff, margin_safe, patch_tolerance, pkg,  # noqa: E402
                           selection_sensitivity)
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


def test_poisson_lod_example_golden(golden_dir):
    """tests/Poisson_LOD_Example.output through the CUDA path: rhs l2 norm = 0.0808367."""
    txt = open(os.path.join(golden_dir, "Poisson_LOD_Example.output")).read()
    rhs_norm = float(re.search(r"\n\s+rhs l2 norm = ([0-9.]+)", txt).group(1))
    rng = GlibcRand()
    for _ in range(12):
        rng.rand()
    tab = reference_random_table(2, 1, 100, 8, rng)
    ctx, orc = build_pair(dim=2, s=1, ref=2, n=2, ell=1, stabilize=False, r=8, quirk=True, tables=[tab])
    ctx.compute_basis()
    orc.compute_basis()
    f = orc.fem_rhs_constant_one()
    rhs = np.zeros(ctx.n_patches)
    for res in orc.patches:
        phi, aphi = ctx.basis(res.pid)
        assert np.linalg.norm(phi - res.basis[0]) < 1e-10
        assert np.linalg.norm(aphi - res.basis_premultiplied[0]) < 1e-10 * np.linalg.norm(res.basis_premultiplied[0])
        rhs[res.pid] = phi @ f[_global_lex_nodes(ctx.patch_info(res.pid), 2, 9)]
    assert float("%g" % np.linalg.norm(rhs)) == rhs_norm


CASES = [
    dict(dim=2, s=1, ref=3, n=2, ell=1, stabilize=False),            # LOD branch
    dict(dim=2, s=1, ref=4, n=2, ell=1),                             # SLOD, small patches
    dict(dim=2, s=1, ref=4, n=2, ell=2),                             # cfg 2 shape
    dict(dim=2, s=1, ref=2, n=2, ell=3),                             # every patch is the whole domain -> LOD
    dict(dim=2, s=1, ref=3, n=4, ell=1),                             # 4 subdivisions
    dict(dim=2, s=2, ref=3, n=2, ell=1),                             # cfg 3 shape (elasticity)
    dict(dim=2, s=2, ref=4, n=2, ell=2, sample=40),
    dict(dim=2, s=1, ref=4, n=2, ell=2, kind="binary1e4", seed=1235),  # high contrast
    dict(dim=2, s=1, ref=3, n=2, ell=0),                             # oversampling 0: one-cell patches -> LOD (source/LOD.cc:563)
    dict(dim=2, s=1, ref=1, n=2, ell=1),                             # 2 x 2 cells: smallest mesh with more than one patch
    dict(dim=3, s=1, ref=1, n=2, ell=1),
    dict(dim=3, s=1, ref=2, n=2, ell=1),
    dict(dim=3, s=1, ref=3, n=2, ell=2, sample=10, kind="uniform1e4", seed=3001),   # cfg 4 shape
    dict(dim=3, s=1, ref=2, n=2, ell=2, sample=16),                  # 4^3 cells: patches of 27 .. 64 cells (tensor-core solver + SIMT dense stage)
    dict(dim=2, s=1, ref=4, n=8, ell=2, sample=8),                   # 8 subdivisions: the tensor-core plan does not fit -> SIMT solver
    dict(dim=2, s=2, ref=3, n=8, ell=1, sample=8),                   # elasticity, 8 subdivisions
    # patches too large for the shared-memory solvers (Ni = 3375, half band width 241): windows in global memory
    dict(dim=3, s=1, ref=2, n=4, ell=2, sample=6),
    # the same fall-back forced on small shapes (SIMT solver with global windows, SIMT dense with a global coefficient window)
    dict(dim=2, s=1, ref=3, n=2, ell=1, env="SLOD_FORCE_GMEM_SOLVER"),
    dict(dim=2, s=2, ref=3, n=2, ell=1, env="SLOD_FORCE_GMEM_SOLVER"),
    dict(dim=3, s=1, ref=2, n=2, ell=1, env="SLOD_FORCE_GMEM_SOLVER"),
]


def test_large_patches_global_memory_solver():
    """3-D patches with 4 subdivisions and oversampling 2 (Ni = 6859 interior dofs, half band width 381,
    125 coarse dofs) do not fit the shared-memory solvers; the SIMT solver runs them with its band / right-hand-side
    windows in global memory -- a direct banded Cholesky like the reference's direct solver (include/LODtools.h:575-580)."""
    ctx, orc = build_pair(dim=3, s=1, ref=3, n=4, ell=2)
    ctx.compute_basis()
    ctx.assemble_coarse()
    full = next(p for p in range(ctx.n_patches) if ctx.patch_info(p)["n_cells"] == 125)
    assert ctx.patch_info(full)["n_internal"] == 6859
    pids = [0, full, ctx.n_patches - 1]
    orc.compute_basis(pids)
    for res in orc.patches:
        phi, aphi = ctx.basis(res.pid)
        dg = ctx.diagnostics(res.pid)
        assert dg[7] == 0
        if not margin_safe(res.info, 0):
            continue
        assert int(dg[1]) == res.info["trunc_steps"][0]
        tol = max(1e-10, 50.0 * selection_sensitivity(res.info, 0))
        assert np.linalg.norm(phi - res.basis[0]) <= tol, (res.pid, tol)
        assert np.linalg.norm(aphi - res.basis_premultiplied[0]) <= 10 * tol * np.linalg.norm(res.basis_premultiplied[0])
    rowptr, col, val = ctx.coarse_csr()
    K = np.zeros((ctx.n_patches, ctx.n_patches))
    for i in range(ctx.n_patches):
        K[i, col[rowptr[i]:rowptr[i + 1]]] = val[rowptr[i]:rowptr[i + 1]]
    assert np.abs(K - K.T).max() <= 1e-8 * np.abs(K).max()


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_basis_and_coarse_matrix(case, monkeypatch):
    case = dict(case)
    sample = case.pop("sample", None)
    env = case.pop("env", None)
    if env:
        monkeypatch.setenv(env, "1")   # read by slod_create
    ctx, orc = build_pair(**case)
    s = case["s"]
    ctx.compute_basis()
    ctx.assemble_coarse()
    npch = ctx.n_patches
    pids = list(range(npch)) if sample is None else sorted(set(range(0, npch, max(1, npch // sample))) | {npch - 1})
    orc.compute_basis(pids)
    tol_of = {}
    n_checked_tight = 0
    for res in orc.patches:
        for d in range(s):
            phi, aphi = ctx.basis(res.pid, d)
            dg = ctx.diagnostics(res.pid, d)
            assert dg[7] == 0
            if not margin_safe(res.info, d):
                tol_of[(res.pid, d)] = None     # decision within rounding of flipping: not comparable
                continue
            err = np.linalg.norm(phi - res.basis[d])
            tol = 1e-10
            if err > tol and res.info["slod"]:
                # ill-conditioned selection: the floor is what a few-ulp perturbation of G does to the oracle
                tol = max(tol, 50.0 * selection_sensitivity(res.info, d))
            tol_of[(res.pid, d)] = tol
            n_checked_tight += tol == 1e-10
            if res.info["slod"]:
                assert int(dg[1]) == res.info["trunc_steps"][d], (res.pid, d)
            assert err <= tol, (res.pid, d, err, tol, cond_eff(res.info, d) if res.info["slod"] else 1)
            nrm = np.linalg.norm(res.basis_premultiplied[d])
            assert np.linalg.norm(aphi - res.basis_premultiplied[d]) <= 10 * tol * nrm, (res.pid, d)
    assert n_checked_tight > 0
    # stages up to the Gram matrix on a few SLOD patches: 1e-10 relative, no conditioning excuse
    for res in [r for r in orc.patches if r.info["slod"]][:: max(1, len(orc.patches) // 5)]:
        X, Minv, G = ctx.debug_stages(res.pid)
        for got, want in ((X, res.info["X"]), (Minv, res.info["Minv"]), (G, res.info["G"])):
            assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max(), res.pid
    if sample is None:
        K, _, _ = orc.assemble_global_matrix()
        rowptr, col, val = ctx.coarse_csr()
        assert np.array_equal(rowptr, K.indptr) and np.array_equal(col, K.indices)      # bit-exact pattern
        kmax = np.abs(K.data).max()
        rows = np.repeat(np.arange(K.shape[0]), np.diff(K.indptr))
        tol_row = np.array([tol_of.get((r // s, r % s)) or np.inf for r in range(K.shape[0])])
        tol_e = 20 * (tol_row[rows] + tol_row[col])
        assert (np.abs(val - K.data) <= tol_e * kmax).all()
        tight = np.isfinite(tol_e) & (tol_e <= 20 * 2e-10)
        assert tight.any()
        assert np.abs(val - K.data)[tight].max() <= 4e-9 * kmax


def test_scaling_by_power_of_two_is_exact():
    """Size-independent property: scaling alpha by 4 (exact in binary floating point) leaves phi bit-identical
    and scales A phi and K by exactly 4 -- also a determinism check (no atomics-ordered sums)."""
    import parity_common
    tabs1 = parity_common.make_tables(2, 1, 5, "uniform100", 5)
    ctx1, _ = build_pair(dim=2, s=1, ref=4, n=2, ell=2, tables=tabs1, r=5)
    ctx2, _ = build_pair(dim=2, s=1, ref=4, n=2, ell=2, tables=[4.0 * t for t in tabs1], r=5)
    for c in (ctx1, ctx2):
        c.compute_basis()
        c.assemble_coarse()
    p1, a1 = ctx1.all_basis()
    p2, a2 = ctx2.all_basis()
    assert np.array_equal(p1, p2)
    assert np.array_equal(4.0 * a1, a2)
    assert np.array_equal(4.0 * ctx1.coarse_csr()[2], ctx2.coarse_csr()[2])


def test_partition_of_patches_matches_full_run():
    """Patches are independent: computing two halves separately (the multi-GPU partition) gives the same basis."""
    import torch
    ctx, _ = build_pair(dim=2, s=1, ref=4, n=2, ell=2, seed=9)
    ctx.compute_basis()
    p_full, a_full = ctx.all_basis()
    n, s, stride = ctx.n_patches, 1, ctx.basis_stride
    phi = torch.zeros((n, s, stride), dtype=torch.float64, device="cuda")
    aphi = torch.zeros_like(phi)
    ctx.compute_basis_device(0, n // 3, phi.data_ptr(), aphi.data_ptr())
    ctx.compute_basis_device(n // 3, n, phi.data_ptr(), aphi.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(phi.cpu().numpy(), p_full)
    assert np.array_equal(aphi.cpu().numpy(), a_full)
    K = torch.zeros((n * s, ctx.ell_width), dtype=torch.float64, device="cuda")
    ctx.assemble_coarse_device(0, n // 2, phi.data_ptr(), aphi.data_ptr(), K.data_ptr())
    ctx.assemble_coarse_device(n // 2, n, phi.data_ptr(), aphi.data_ptr(), K.data_ptr())
    torch.cuda.synchronize()
    ctx.assemble_coarse()
    rowptr, col, val = ctx.ell_to_csr(K.cpu().numpy())
    r2, c2, v2 = ctx.coarse_csr()
    assert np.array_equal(rowptr, r2) and np.array_equal(col, c2) and np.array_equal(val, v2)


@pytest.mark.parametrize("env", ["SLOD_FORCE_SIMT_SOLVER", "SLOD_FORCE_SIMT_DENSE", "SLOD_SIMPLE_COARSE",
                                 "SLOD_FORCE_JACOBI", "SLOD_NO_FAST_SELECT"])
@pytest.mark.parametrize("case", [dict(dim=2, s=1, ref=4, n=2, ell=2), dict(dim=3, s=1, ref=2, n=2, ell=1)],
                         ids=["2d", "3d"])
def test_generic_fallback_kernels(env, case, monkeypatch):
    """The shape-generic kernels behind the tensor-core variants (SIMT banded solver, SIMT dense stage, per-patch
    coarse kernel, Jacobi eigen-solver, eigen pipeline without the Cholesky fast path) are product code for shapes the
    fast variants do not cover: run them on shapes both cover and demand the same answers."""
    ctx_ref, _ = build_pair(**case)
    ctx_ref.compute_basis()
    ctx_ref.assemble_coarse()
    p_ref, a_ref = (x.copy() for x in ctx_ref.all_basis())
    k_ref = ctx_ref.coarse_csr()[2].copy()
    monkeypatch.setenv(env, "1")
    ctx, _ = build_pair(**case)
    ctx.compute_basis()
    ctx.assemble_coarse()
    p, a = ctx.all_basis()
    k = ctx.coarse_csr()[2]
    steps_ref = [int(ctx_ref.diagnostics(pid)[1]) for pid in range(ctx.n_patches)]
    steps = [int(ctx.diagnostics(pid)[1]) for pid in range(ctx.n_patches)]
    assert steps == steps_ref
    # same algorithm, different summation orders / eigen-solvers: agreement at the level of the selection's
    # conditioning (cond(G) * eps reaches 1e-5 on the boundary patches of the 2-D case, see test_basis_and_coarse_matrix)
    per_patch = np.abs(p - p_ref).max(axis=(1, 2))
    assert per_patch.max() <= 1e-4 and np.median(per_patch) <= 1e-11 and np.quantile(per_patch, 0.75) <= 1e-9
    assert np.abs(k - k_ref).max() <= 1e-3 * np.abs(k_ref).max()
    assert np.abs(a - a_ref).max() <= 1e-3 * np.abs(a_ref).max()


def test_device_entry_reports_numerical_status():
    """slod_compute_basis_device must surface the per-patch status bits.  A coefficient with negative
    values makes A_ii indefinite on some patches; the failure arrives at slod_synchronize as SLOD_ERR_NUMERIC and a
    clean run afterwards reports nothing (the bits of the range are cleared in front of the kernels)."""
    import torch
    ctx, _ = build_pair(dim=2, s=1, ref=3, n=2, ell=1, seed=3)
    n, stride = ctx.n_patches, ctx.basis_stride
    phi = torch.zeros((n, 1, stride), dtype=torch.float64, device="cuda")
    aphi = torch.zeros_like(phi)
    good = 1.0 + np.random.default_rng(3).random(64)
    bad = good.copy()
    bad[::3] = -5.0
    ctx.set_coefficient(0, 3, bad)
    ctx.compute_basis_device(0, n, phi.data_ptr(), aphi.data_ptr())
    with pytest.raises(pkg.SlodError) as e:
        ctx.synchronize()
    assert e.value.code == 5 and "patch" in str(e.value)
    with pytest.raises(pkg.SlodError):
        ctx.compute_basis()          # the host-buffer path reports the same
    ctx.set_coefficient(0, 3, good)
    ctx.compute_basis_device(0, n, phi.data_ptr(), aphi.data_ptr())
    ctx.synchronize()
    assert torch.isfinite(phi).all()


@pytest.mark.parametrize("case", [dict(dim=2, s=1, ref=4, n=2, ell=2), dict(dim=3, s=1, ref=2, n=2, ell=1)],
                         ids=["2d", "3d"])
def test_multi_chunk_loop(case, monkeypatch):
    """n_patches > chunk: SLOD_CHUNK forces a workspace of 37 patches, so the range is processed
    in several chunks that reuse the same buffers back to back; results must be bit-identical to the one-chunk run."""
    ctx_ref, _ = build_pair(**case)
    ctx_ref.compute_basis()
    ctx_ref.assemble_coarse()
    p_ref, a_ref = ctx_ref.all_basis()
    k_ref = ctx_ref.coarse_csr()[2]
    monkeypatch.setenv("SLOD_CHUNK", "37")
    ctx, _ = build_pair(**case)
    ctx.compute_basis()
    ctx.assemble_coarse()
    p, a = ctx.all_basis()
    assert np.array_equal(p, p_ref) and np.array_equal(a, a_ref)
    assert np.array_equal(ctx.coarse_csr()[2], k_ref)
    assert ctx.timings()[0] > 0


def test_offline_distributed_single_rank_with_host_outputs():
    """slod_offline_distributed on a communicator-less handle (world 1) = basis + coarse rows of all patches, enqueued;
    slod_set_host_outputs makes the library copy phi / A phi / K rows to host memory on its own stream; everything is
    bit-identical to the host-buffer path."""
    import torch
    ctx, _ = build_pair(dim=2, s=1, ref=4, n=2, ell=2, seed=9)
    ctx.compute_basis()
    ctx.assemble_coarse()
    p_ref, a_ref = (x.copy() for x in ctx.all_basis())
    k_ref = ctx.coarse_csr()[2].copy()
    n, stride, ellw = ctx.n_patches, ctx.basis_stride, ctx.ell_width
    assert ctx.owned_range(0, 1) == (0, n)
    phi = torch.zeros((n, 1, stride), dtype=torch.float64, device="cuda")
    aphi = torch.zeros_like(phi)
    K = torch.zeros((n, ellw), dtype=torch.float64, device="cuda")
    h_phi = torch.empty((n, 1, stride), dtype=torch.float64, pin_memory=True)
    h_aphi = torch.empty_like(h_phi, pin_memory=True)
    h_K = torch.empty((n, ellw), dtype=torch.float64, pin_memory=True)
    ctx.set_host_outputs(h_phi.data_ptr(), h_aphi.data_ptr(), h_K.data_ptr())
    ctx.offline_distributed(phi.data_ptr(), aphi.data_ptr(), K.data_ptr(), gather_phi=True, gather_K=True)
    ctx.synchronize()
    ctx.set_host_outputs(0, 0, 0)
    assert np.array_equal(h_phi.numpy(), p_ref) and np.array_equal(h_aphi.numpy(), a_ref)
    torch.cuda.synchronize()
    assert np.array_equal(phi.cpu().numpy(), p_ref)
    assert np.array_equal(ctx.ell_to_csr(h_K.numpy())[2], k_ref)
