"""Pin the CPU oracle against the reference's own golden files (copied verbatim from
/root/reference/tests/*.output into tests/golden/)."""
import os
import re

import numpy as np
import pytest

from oracle.slod_oracle import (CoefficientTable, GlibcRand, PatchResult, SlodOracle, SlodProblem,
                                create_patches, lexicographic_to_hierarchic, morton_decode,
                                qiso_cell_matrix, reference_random_table, structured_patch_poisson)


def _read(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return f.read()


def test_glibc_rand_stream():
    r = GlibcRand()
    assert [r.rand() for _ in range(4)] == [1804289383, 846930886, 1681692777, 1714636915]
    tab = reference_random_table(2, 1, 100, 1, GlibcRand())
    assert abs(tab[0] - 84.17858887) < 1e-6


def test_create_patch_01(golden_dir):
    """tests/create_patch_01.cc: ref 5, oversampling 4 -> #cells of each of 1024 patches (Morton)."""
    txt = _read(golden_dir, "create_patch_01.output")
    vals = [int(m) for m in re.findall(r"\{(\d+)\}", txt)]
    assert len(vals) == 1024
    patches = create_patches(2, 5, 4)
    assert [len(p) for p in patches] == vals


def test_create_mesh_from_cells_01(golden_dir):
    """tests/create_mesh_from_cells_01.cc: 3x3 sub-mesh around cell (2,3) of a 7x7 mesh."""
    g = np.array(_read(golden_dir, "create_mesh_from_cells_01.output").split(), dtype=float).reshape(-1, 2)
    centres = sorted(((i + 0.5) / 7, (j + 0.5) / 7) for i in (1, 2, 3) for j in (2, 3, 4))
    assert np.allclose(np.array(sorted(map(tuple, g))), np.array(centres), atol=1e-6)


def test_fe_q_iso_q1_01(golden_dir):
    """tests/fe_q_iso_q1_01.cc: cell Laplace matrices 1-D and 2-D, n = 3, hierarchical numbering."""
    blocks = [b for b in _read(golden_dir, "fe_q_iso_q1_01.output").split("\n\n") if b.strip()]
    assert len(blocks) == 4
    for blk, dim in zip(blocks, (1, 1, 2, 2)):
        A = qiso_cell_matrix(dim, 3)
        lines = blk.rstrip("\n").split("\n")
        n = A.shape[0]
        assert len(lines) == n
        for i, line in enumerate(lines):
            # fixed-width columns of 11 characters, blank = structural zero
            for j in range(n):
                field = line[11 * j: 11 * (j + 1)].strip()
                if field:
                    assert abs(float(field) - A[i, j]) < 5.1e-4, (dim, i, j)
                else:
                    assert abs(A[i, j]) < 1e-14, (dim, i, j)


def test_hierarchic_numbering_is_permutation():
    for dim, n in ((2, 2), (2, 3), (3, 2), (3, 3)):
        h = lexicographic_to_hierarchic(dim, n)
        assert sorted(h) == list(range((n + 1) ** dim))
    # 2-D n=2: vertices 0..3, lines x=0,x=1,y=0,y=1, interior
    assert list(lexicographic_to_hierarchic(2, 2)) == [0, 6, 1, 4, 8, 5, 2, 7, 3]


def test_solve_poisson_problem_on_patch_01(golden_dir):
    g = np.array(_read(golden_dir, "solve_poisson_problem_on_patch_01.output").split(), dtype=float)
    u = structured_patch_poisson([10, 10], 7, (1, 4), 3)
    assert g.shape == u.shape == (5041,)
    # golden is printed with %.3e
    assert np.abs(u - g).max() < 5.1e-6
    assert (g != 0).sum() == (np.abs(u) > 5e-7).sum()


def test_parallel_assembly(golden_dir):
    """tests/parallel_assembly.cc: LOD<2,2>, ref 2, ell 1, n 2, basis == premultiplied == 1."""
    entries = re.findall(r"\((\d+),(\d+)\) ([-0-9.e+]+)", _read(golden_dir, "parallel_assembly.output"))
    assert len(entries) == 1024
    prob = SlodProblem(dim=2, spacedim=2, n_global_refinements=2, n_subdivisions=2, oversampling=1,
                       problem="elasticity")
    o = SlodOracle(prob)
    fake = []
    for pid in range(16):
        shape, lo = o.shape_for(morton_decode(pid, 2, 2))
        ones = np.ones((2, shape.Nf))
        fake.append(PatchResult(pid, tuple(lo), shape.m, None, ones, ones, {}))
    K, _, _ = o.assemble_global_matrix(fake)
    Kd = K.toarray()
    assert K.nnz == 1024
    for i, j, v in entries:
        assert Kd[int(i), int(j)] == float(v)


def test_poisson_lod_example(golden_dir):
    """tests/Poisson_LOD_Example.cc end to end (LOD branch).  Needs quirk A (glibc rand() advanced
    by 12 draws, float32 rounding) and quirk B (presaved patch matrix)."""
    txt = _read(golden_dir, "Poisson_LOD_Example.output")
    assert "Patches size in (4, 9)" in txt
    fem_rhs_norm = float(re.search(r"fem rhs l2 norm = ([0-9.]+)", txt).group(1))
    rhs_norm = float(re.search(r"\n\s+rhs l2 norm = ([0-9.]+)", txt).group(1))
    rng = GlibcRand()
    for _ in range(12):
        rng.rand()
    tab = reference_random_table(2, 1, 100, 8, rng)
    prob = SlodProblem(dim=2, spacedim=1, n_global_refinements=2, n_subdivisions=2, oversampling=1,
                       stabilize=False, quirk_presaved=True, coefficients=[CoefficientTable(2, 8, tab)])
    o = SlodOracle(prob)
    o.compute_basis()
    sizes = [len(p.cells) for p in o.patches]
    assert (min(sizes), max(sizes)) == (4, 9)
    K, C, AC = o.assemble_global_matrix()
    f = o.fem_rhs_constant_one()
    assert f.size == 81
    assert float("%g" % np.linalg.norm(f)) == fem_rhs_norm
    assert float("%g" % np.linalg.norm(C.T @ f)) == rhs_norm       # 0.0808367
    assert K.shape == (16, 16)


def test_fe_q_iso_q1_02_elasticity_cell_matrix(golden_dir):
    """tests/fe_q_iso_q1_02.cc: the vector-valued Q_iso_Q1(3) cell matrix of 2 eps(u):eps(v) + div u div v assembled by the
    reference's sub-cell loops (include/Elasticity.h:211-284) equals the plain FEValues double loop; the test prints the
    tolerance it passed with (1e-16).  Here: the oracle's sub-cell assembly (lambda = mu = 1) against an independent
    evaluation with global hat functions at every iterated Gauss point."""
    tol = float(_read(golden_dir, "fe_q_iso_q1_02.output").split()[0])
    assert tol == 1e-16
    n = 3
    prob = SlodProblem(dim=2, spacedim=2, n_global_refinements=0, n_subdivisions=n, oversampling=0, stabilize=False,
                       problem="elasticity", coefficients=[CoefficientTable(2, 0, np.ones(1)), CoefficientTable(2, 0, np.ones(1))])
    o = SlodOracle(prob)
    shape, lo = o.shape_for((0, 0))
    A = o.assemble_patch_stiffness(shape, lo).toarray()          # dof = 2 * node + comp, nodes x fastest
    # independent: hat functions of the (n+1)^2 node grid, 2-point Gauss rule iterated n times per axis
    G = n + 1
    h = 1.0 / n
    nodes1 = np.arange(G) * h

    def hat(x):            # values and derivatives of the G one-dimensional hats at x
        v = np.maximum(0.0, 1.0 - np.abs(x - nodes1) / h)
        d = np.where(np.abs(x - nodes1) < h, -np.sign(x - nodes1) / h, 0.0)
        return v, d

    g1 = np.array([0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0)])
    pts = np.concatenate([(c + g1) * h for c in range(n)])
    ref = np.zeros((2 * G * G, 2 * G * G))
    for y in pts:
        vy, dy = hat(y)
        for x in pts:
            vx, dx = hat(x)
            gx = np.outer(vy, dx).ravel()          # d/dx of node (ix, iy), index iy * G + ix
            gy = np.outer(dy, vx).ravel()
            w = (h / 2.0) ** 2
            # u = phi e_c: eps = sym(grad), div = d_c phi
            grads = {0: (gx, gy), 1: (gx, gy)}
            for ci in range(2):
                for cj in range(2):
                    # 2 eps_i : eps_j for unit vectors e_ci, e_cj
                    if ci == cj:
                        other = gy if ci == 0 else gx
                        same = gx if ci == 0 else gy
                        e2 = 2.0 * np.outer(same, same) + np.outer(other, other)
                    else:
                        gi_other = gy if ci == 0 else gx      # d_{cj} phi_i
                        gj_other = gy if cj == 0 else gx      # d_{ci} phi_j
                        e2 = np.outer(gi_other, gj_other)
                    di = gx if ci == 0 else gy
                    dj = gx if cj == 0 else gy
                    ref[ci::2, cj::2] += (e2 + np.outer(di, dj)) * w
    assert np.abs(A - ref).max() <= 50 * tol * max(1.0, np.abs(ref).max())
    assert np.abs(A - A.T).max() <= 50 * tol


def test_assembly_01_entries(golden_dir):
    """tests/assembly_01.cc: 1-D, 6 cells of FE_Q_iso_Q1(4), indicator basis (a shared vertex belongs to both cells),
    all-ones cell matrices: the printed entries of A_lod = C^T A C (an early prototype of the coarse assembly: the
    entries (i, i+-2) come from A coupling the two shared vertices of the cell in between)."""
    import scipy.sparse as sp
    entries = {}
    for line in _read(golden_dir, "assembly_01.output").splitlines():
        m = re.match(r"\((\d+),(\d+)\) (\S+)", line.strip())
        if m:
            entries[(int(m.group(1)), int(m.group(2)))] = float(m.group(3))
    assert len(entries) == 24
    ncell, deg = 6, 4
    nf = ncell * deg + 1
    cell_dofs = [np.arange(c * deg, c * deg + deg + 1) for c in range(ncell)]
    A = np.zeros((nf, nf))
    C = np.zeros((nf, ncell))
    for c, dofs in enumerate(cell_dofs):
        A[np.ix_(dofs, dofs)] += 1.0
        C[dofs, c] = 1.0
    Kd = C.T @ A @ C
    got = {(i, j): float(Kd[i, j]) for i in range(ncell) for j in range(ncell) if Kd[i, j] != 0.0}
    assert got == entries
    # The oracle's scatter-and-product keeps the pattern of Tmmult(C, AC) with AC stored on C's pattern (what
    # LOD::assemble_global_matrix does, source/LOD.cc:933-971): columns sharing a stored row, here |i - j| <= 1.
    Cs = sp.csc_matrix(C)
    AC = sp.csc_matrix(np.where(C != 0.0, A @ C, 0.0))
    K = SlodOracle.galerkin_product(Cs, AC, Cs)
    K.sort_indices()
    for i in range(ncell):
        cols = K.indices[K.indptr[i]:K.indptr[i + 1]]
        assert list(cols) == [j for j in range(ncell) if abs(i - j) <= 1]


def test_online_restatement_consistency():
    """fem_rhs (Gauss quadrature) reproduces the closed form for f = 1 and integrates a linear f exactly; the oracle's
    CG + SSOR(1.2) (LOD::solve, source/LOD.cc:991-998) agrees with the direct branch (source/LOD.cc:984-989)."""
    prob = SlodProblem(dim=2, spacedim=1, n_global_refinements=3, n_subdivisions=2, oversampling=1, stabilize=True,
                       coefficients=[CoefficientTable(2, 4, 1.0 + 99.0 * np.random.default_rng(5).random(256))])
    o = SlodOracle(prob)
    assert np.abs(o.fem_rhs(lambda p: np.ones((len(p), 1))) - o.fem_rhs_constant_one()).max() < 1e-16
    F = o.fem_rhs(lambda p: p[:, [0]] + 2.0 * p[:, [1]]).reshape(17, 17)     # [y][x]
    assert abs(F[5, 3] - (3 / 16 + 2 * 5 / 16) * (1 / 16) ** 2) < 1e-16        # interior node: f(node) h^2
    assert not F[0].any() and not F[:, -1].any()                              # constrained rows
    o.compute_basis()
    K, C, _ = o.assemble_global_matrix()
    b = C.T @ o.fem_rhs(lambda p: np.sin(3 * p[:, [0]]) + p[:, [1]])
    u_cg, steps = o.solve_coarse(K, b, max_steps=500, tolerance=0.0, reduction=1e-12)
    u_direct, _ = o.solve_coarse(K, b, direct=True)
    assert 0 < steps < 500
    assert np.linalg.norm(u_cg - u_direct) <= 1e-9 * np.linalg.norm(u_direct)
    with pytest.raises(RuntimeError):
        o.solve_coarse(K, b, max_steps=1, tolerance=0.0, reduction=1e-12)


def test_assembly_02_frobenius(golden_dir):
    """tests/assembly_02.cc: 1-D, 5 coarse cells x 2 sub-cells, indicator basis: C^T A C has 20 on the
    diagonal and four 10s coupling... -> Frobenius norm 48.9898 (hand-checked in SURVEY 4.2)."""
    first = _read(golden_dir, "assembly_02.output").split()[0]
    assert first == "48.9898"
    assert "%g" % np.sqrt(2400.0) == first
