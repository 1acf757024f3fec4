"""CPU oracle for the SLOD offline phase  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The product path (``dealii-slod_b200``) never does.

It is a numpy/scipy restatement of the reference's serial CPU algorithm (camillabelponer/dealii-slod,
paths relative to the reference root):

* ``LOD::create_patches``                      source/LOD.cc:122-244
* ``LOD::create_mesh_for_patch`` (face tags)   source/LOD.cc:770-858
* ``fill_dofs_indices_vector``                 include/LODtools.h:334-375
* ``projection_P1_P0`` + scatter               include/LODtools.h:7-73, source/LOD.cc:329-342,470-496
* ``DiffusionProblem::assemble_stiffness``     include/Diffusion.h:111-207
* ``ElasticityProblem::assemble_stiffness``    include/Elasticity.h:163-299
* ``problem_parameter``                        include/Diffusion.h:7-54
* ``LOD::compute_basis_function_candidates``   source/LOD.cc:296-768
* ``LOD::assemble_global_matrix``              source/LOD.cc:860-973
* ``LOD::solve`` (+ fine rhs, prolongation)    source/LOD.cc:975-1001, :1251, include/Diffusion.h:149-153,188-191

deal.II 9.6 / Trilinos / LAPACK (the third-party libraries that carry the arithmetic) are not
available in this environment, so the reference itself cannot be built; the oracle is pinned
against the reference's own golden ``tests/*.output`` files (see tests/test_oracle_golden.py).

PARITY STATUS: the LOD branch, patch construction, Q_iso_Q1 cell matrices, patch solve and the
coarse-matrix scatter/product are pinned by reference golden files.  The SLOD branch
(source/LOD.cc:596-757), 2-D elasticity end-to-end and everything 3-D are executed by no
reference test: for those the oracle is "parity unpinned" -- its fidelity rests on following
the cited lines.  The reference has no 3-D path at all; dim = 3 here is the dim-generic
extension of the same algorithm (x-outer / z-inner cell order, weights 1,2,4,8 * h^3/8).

Conventions: patch-local fine nodes are numbered lexicographically (x fastest); a fine DoF is
``spacedim * node + comp``.  The maps to deal.II's own numberings are given by
``global_fine_dof_numbering`` / ``patch_local_dof_numbering`` (restated from deal.II 9.6
semantics, not pinned by any golden file).
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

SPECIAL_NUMBER = 99  # source/LOD.cc:7


# ----------------------------------------------------------------------------------------------
# glibc rand() and the reference's random coefficient tables (quirk A)
# ----------------------------------------------------------------------------------------------
class GlibcRand:
    """glibc ``rand()`` with the default seed 1 (TYPE_3 additive feedback generator)."""

    RAND_MAX = 2147483647

    def __init__(self, seed: int = 1):
        r = [0] * 34
        r[0] = seed
        for i in range(1, 31):
            # 16807 * r[i-1] % 2147483647 computed Schrage-style on signed 32-bit ints
            hi, lo = divmod(r[i - 1], 127773)
            word = 16807 * lo - 2836 * hi
            if word < 0:
                word += 2147483647
            r[i] = word
        for i in range(31, 34):
            r[i] = r[i - 31]
        self._r = r
        for _ in range(34, 344):
            self._step()

    def _step(self) -> int:
        r = self._r
        v = (r[-31] + r[-3]) & 0xFFFFFFFF
        r.append(v)
        if len(r) > 64:
            del r[: len(r) - 34]
        return v

    def rand(self) -> int:
        return self._step() >> 1


def reference_random_table(dim: int, vmin: float, vmax: float, r: int, rng: GlibcRand) -> np.ndarray:
    """``problem_parameter`` ctor, include/Diffusion.h:19-38 (values pass through float32)."""
    n_cells = (2 ** r) ** dim
    out = np.empty(n_cells)
    denom = np.float32(np.float64(GlibcRand.RAND_MAX) / (vmax - vmin))
    for i in range(n_cells):
        out[i] = vmin + np.float64(np.float32(rng.rand()) / denom)
    return out


# ----------------------------------------------------------------------------------------------
# integer maps
# ----------------------------------------------------------------------------------------------
def morton_encode(idx, dim: int, ref: int) -> int:
    """Active-cell index of coarse cell ``idx`` after ``refine_global(ref)``: child = x + 2y (+4z)."""
    code = 0
    for b in range(ref):
        for a in range(dim):
            code |= ((idx[a] >> b) & 1) << (dim * b + a)
    return code


def morton_decode(code: int, dim: int, ref: int):
    idx = [0] * dim
    for b in range(ref):
        for a in range(dim):
            idx[a] |= ((code >> (dim * b + a)) & 1) << b
    return tuple(idx)


def patch_cells_lex(c, ell: int, N: int):
    """source/LOD.cc:151-178: centre first, then x-offset outer ... last axis innermost.

    Returns the list of cell multi-indices."""
    dim = len(c)
    out = [tuple(c)]
    ranges = [range(-ell, ell + 1)] * dim
    for off in itertools.product(*ranges):  # first axis (x) is the outer loop
        cc = tuple(c[a] + off[a] for a in range(dim))
        if all(0 <= cc[a] < N for a in range(dim)) and cc != tuple(c):
            out.append(cc)
    return out


def create_patches(dim: int, ref: int, ell: int):
    """Patch id == active cell index (Morton).  Returns list over patch id of arrays of active-cell
    ids (centre first), source/LOD.cc:184-219."""
    N = 2 ** ref
    patches = []
    for pid in range(N ** dim):
        c = morton_decode(pid, dim, ref)
        cells = patch_cells_lex(c, ell, N)
        patches.append(np.array([morton_encode(cc, dim, ref) for cc in cells], dtype=np.uint32))
    return patches


def lexicographic_to_hierarchic(dim: int, n: int) -> np.ndarray:
    """FETools::lexicographic_to_hierarchic_numbering<dim>(n) for FE_Q_iso_Q1(n) (pinned in 2-D by
    tests/fe_q_iso_q1_01.output)."""
    p = n + 1
    h = np.full(p ** dim, -1, dtype=np.int64)
    nxt = 0

    def lex(ix):
        j = 0
        for a in reversed(range(dim)):
            j = j * p + ix[a]
        return j

    ends = (0, n)
    # vertices, x fastest
    for v in itertools.product(*([ends] * dim)):
        ix = tuple(reversed(v))  # product varies last fastest -> make x fastest
        h[lex(ix)] = nxt
        nxt += 1
    inner = range(1, n)
    if dim == 2:
        lines = [((0, None)), ((n, None)), ((None, 0)), ((None, n))]  # x=0, x=1, y=0, y=1
        for ln in lines:
            for t in inner:
                ix = tuple(t if v is None else v for v in ln)
                h[lex(ix)] = nxt
                nxt += 1
    elif dim == 3:
        lines = []
        for z in (0, n):
            lines += [(0, None, z), (n, None, z), (None, 0, z), (None, n, z)]
        lines += [(0, 0, None), (n, 0, None), (0, n, None), (n, n, None)]
        for ln in lines:
            for t in inner:
                ix = tuple(t if v is None else v for v in ln)
                h[lex(ix)] = nxt
                nxt += 1
        faces = [(0, None, None), (n, None, None), (None, 0, None), (None, n, None),
                 (None, None, 0), (None, None, n)]
        for fc in faces:
            free = [a for a in range(3) if fc[a] is None]
            for t1 in inner:          # second free axis slow
                for t0 in inner:      # first free axis fast
                    ix = list(fc)
                    ix[free[0]] = t0
                    ix[free[1]] = t1
                    h[lex(tuple(ix))] = nxt
                    nxt += 1
    elif dim == 1:
        pass
    for ix in itertools.product(*([inner] * dim)):
        ixx = tuple(reversed(ix))
        h[lex(ixx)] = nxt
        nxt += 1
    assert nxt == p ** dim and (h >= 0).all()
    return h


def _cell_entity_walk(dim: int, n: int):
    """Order in which DoFHandler::distribute_dofs numbers the nodes of one cell [deal.II-internal]:
    vertices, lines, (quads,) interior -- i.e. hierarchical order.  Returns cell-local node
    multi-indices in that order."""
    l2h = lexicographic_to_hierarchic(dim, n)
    p = n + 1
    order = np.argsort(l2h)
    out = []
    for j in order:
        ix = []
        for a in range(dim):
            ix.append(j % p)
            j //= p
        out.append(tuple(ix))
    return out


def _numbering_walk(dim: int, s: int, n: int, cells, node_shape, cell_origin):
    """Walk ``cells`` (multi-indices) in order; on each cell number not-yet-numbered nodes in
    hierarchical entity order, ``s`` consecutive numbers per node (n = 2: exact; n > 2 keeps
    node-major component order)."""
    walk = _cell_entity_walk(dim, n)
    num = -np.ones(tuple(node_shape)[::-1], dtype=np.int64)  # indexed [z][y][x]
    nxt = 0
    for c in cells:
        for loc in walk:
            g = tuple(n * (c[a] - cell_origin[a]) + loc[a] for a in range(dim))
            if num[g[::-1]] < 0:
                num[g[::-1]] = nxt
                nxt += s
    return num


def global_fine_dof_numbering(dim: int, s: int, ref: int, n: int) -> np.ndarray:
    """deal.II global fine DoF index of (node, comp 0) for every global fine node, array indexed
    [z][y][x]; comp c adds c.  Cells are walked in Morton order (source/LOD.cc:90)."""
    N = 2 ** ref
    cells = [morton_decode(k, dim, ref) for k in range(N ** dim)]
    return _numbering_walk(dim, s, n, cells, [N * n + 1] * dim, [0] * dim)


def patch_local_dof_numbering(dim: int, s: int, n: int, cells, lo, m) -> np.ndarray:
    """Patch-local deal.II numbering (source/LOD.cc:365-366) of the lexicographic patch nodes."""
    return _numbering_walk(dim, s, n, cells, [m[a] * n + 1 for a in range(dim)], lo)


# ----------------------------------------------------------------------------------------------
# reference-element (sub-cell) matrices at the 2^d Gauss points of QGauss<1>(2)
# ----------------------------------------------------------------------------------------------
def subcell_gauss_matrices(dim: int, h: float, problem: str):
    """Per-Gauss-point sub-cell matrices.

    diffusion:  K[q][i][j]     = grad N_i . grad N_j * JxW                (include/Diffusion.h:181-186)
    elasticity: Kmu[q], Klam[q] = 2 eps(phi_i):eps(phi_j) * JxW,  div phi_i div phi_j * JxW
                                                                       (include/Elasticity.h:236-250)
    Local node order lexicographic (i_0 fastest); elasticity dof = 2*node + comp.
    Gauss-point order q_0 fastest.  Returns (gauss_unit_points, matrices...)."""
    g = 0.5 / math.sqrt(3.0)
    pts1 = (0.5 - g, 0.5 + g)
    nn = 2 ** dim
    qpts = [tuple(reversed(q)) for q in itertools.product(*([pts1] * dim))]
    nodes = [tuple(reversed(v)) for v in itertools.product(*([(0, 1)] * dim))]
    jxw = (h / 2.0) ** dim

    def grad(node, x):
        gvec = np.empty(dim)
        for a in range(dim):
            val = 1.0
            for b in range(dim):
                if b == a:
                    val *= (1.0 if node[b] == 1 else -1.0) / h
                else:
                    val *= x[b] if node[b] == 1 else (1.0 - x[b])
            gvec[a] = val
        return gvec

    if problem == "diffusion":
        K = np.zeros((len(qpts), nn, nn))
        for qi, x in enumerate(qpts):
            G = np.array([grad(nd, x) for nd in nodes])
            K[qi] = G @ G.T * jxw
        return qpts, K
    s = dim
    Kmu = np.zeros((len(qpts), nn * s, nn * s))
    Klam = np.zeros_like(Kmu)
    for qi, x in enumerate(qpts):
        G = [grad(nd, x) for nd in nodes]
        eps = []
        div = []
        for i in range(nn):
            for c in range(s):
                gm = np.zeros((s, dim))
                gm[c, :] = G[i]
                eps.append(0.5 * (gm + gm.T))
                div.append(G[i][c])
        for i in range(nn * s):
            for j in range(nn * s):
                Kmu[qi, i, j] = 2.0 * np.sum(eps[i] * eps[j]) * jxw
                Klam[qi, i, j] = div[i] * div[j] * jxw
    return qpts, Kmu, Klam


# ----------------------------------------------------------------------------------------------
# coefficient fields
# ----------------------------------------------------------------------------------------------
@dataclass
class CoefficientTable:
    """``problem_parameter`` (include/Diffusion.h:7-54): cell-wise values on a 2^r grid,
    lexicographic x fastest; value(p) = table[floor(x/eta) + 2^r floor(y/eta) (+ 4^r floor(z/eta))]."""
    dim: int
    r: int
    values: np.ndarray

    def at(self, pts: np.ndarray) -> np.ndarray:
        nl = 2 ** self.r
        eta = 1.0 / nl
        idx = np.zeros(pts.shape[:-1], dtype=np.int64)
        mul = 1
        for a in range(self.dim):
            ia = np.floor(pts[..., a] / eta).astype(np.int64)
            idx += mul * ia
            mul *= nl
        return self.values[idx]


# ----------------------------------------------------------------------------------------------
# the problem description and per-patch algorithm
# ----------------------------------------------------------------------------------------------
@dataclass
class SlodProblem:
    dim: int = 2
    spacedim: int = 1
    n_global_refinements: int = 2
    n_subdivisions: int = 2
    oversampling: int = 1
    stabilize: bool = False          # "Stabilize phi_LOD candidates"  include/LOD.h:140
    problem: str = "diffusion"       # or "elasticity"
    quirk_presaved: bool = False     # source/LOD.cc:354-362 (constant_coefficients = true)
    coefficients: list = field(default_factory=list)   # [alpha] or [lambda, mu]

    @property
    def N(self):
        return 2 ** self.n_global_refinements

    @property
    def H(self):
        return 0.5 ** self.n_global_refinements

    @property
    def h(self):
        return self.H / self.n_subdivisions


@dataclass
class PatchResult:
    pid: int
    lo: tuple
    m: tuple
    cells: np.ndarray                 # active-cell ids, centre first
    basis: np.ndarray                 # (spacedim, Nf) lexicographic patch numbering
    basis_premultiplied: np.ndarray   # (spacedim, Nf)
    info: dict


class PatchShape:
    """All integer structure of a patch (depends only on per-axis extents and which sides lie on
    the domain boundary)."""

    def __init__(self, dim, s, n, m, dom_lo, dom_hi, centre_rel):
        self.dim, self.s, self.n = dim, s, n
        self.m = tuple(m)
        self.p = tuple(mm * n + 1 for mm in m)
        self.n_nodes = int(np.prod(self.p))
        self.Nf = s * self.n_nodes
        p = self.p
        # node multi-indices, lexicographic x fastest
        grids = np.meshgrid(*[np.arange(pp) for pp in p], indexing="ij")   # axis order x,y,z
        # flatten with x fastest: transpose so that last axis is x
        coords = [np.transpose(g, tuple(reversed(range(dim)))).ravel() for g in grids]
        self.node_coords = np.stack(coords, axis=-1)  # (n_nodes, dim)
        on_db = np.zeros(self.n_nodes, dtype=bool)
        on_b = np.zeros(self.n_nodes, dtype=bool)
        for a in range(dim):
            lo_side = self.node_coords[:, a] == 0
            hi_side = self.node_coords[:, a] == p[a] - 1
            if dom_lo[a]:
                on_db |= lo_side
            else:
                on_b |= lo_side
            if dom_hi[a]:
                on_db |= hi_side
            else:
                on_b |= hi_side
        # include/LODtools.h:355-373: the two sets may overlap; internal = neither
        node_int = ~(on_db | on_b)
        rep = lambda mask: np.repeat(mask, s)
        self.db = np.flatnonzero(rep(on_db))
        self.b = np.flatnonzero(rep(on_b))
        self.internal = np.flatnonzero(rep(node_int))
        # cell list order: centre first then x-outer (source/LOD.cc:151-178)
        cells_rel = [tuple(centre_rel)]
        for cc in itertools.product(*[range(mm) for mm in m]):
            if cc != tuple(centre_rel):
                cells_rel.append(cc)
        self.cells_rel = cells_rel
        self.Nc = len(cells_rel)
        self.Ncd = s * self.Nc
        # sub-cell -> patch node / dof tables
        msub = tuple(mm * n for mm in m)
        self.msub = msub
        sub = [tuple(reversed(t)) for t in itertools.product(*[range(mm) for mm in reversed(msub)])]
        self.sub_coords = np.array(sub, dtype=np.int64)            # (n_sub, dim) x fastest
        corner = [tuple(reversed(v)) for v in itertools.product(*([(0, 1)] * dim))]
        strides = np.cumprod((1,) + p[:-1])
        nodes = np.zeros((len(sub), len(corner)), dtype=np.int64)
        for k, cv in enumerate(corner):
            nodes[:, k] = ((self.sub_coords + np.array(cv)) * strides).sum(axis=1)
        self.sub_nodes = nodes
        self.sub_dofs = (nodes[:, :, None] * s + np.arange(s)[None, None, :]).reshape(len(sub), -1)
        self.strides = strides

    def projection_PT(self, h):
        """Dense P^T (Nf x Ncd): source/LOD.cc:329-342 + 470-496.  Cell-local weight 1/2/4(/8) for
        vertex/line/(face/)interior nodes times h^d / 2^d, summed over patch cells; column
        s*k + comp for the k-th cell of the list."""
        dim, s, n = self.dim, self.s, self.n
        PT = np.zeros((self.Nf, self.Ncd))
        loc = [tuple(reversed(t)) for t in itertools.product(*([range(n + 1)] * dim))]
        for k, c in enumerate(self.cells_rel):
            for t in loc:
                w = 1.0
                for a in range(dim):
                    w *= 1.0 if t[a] in (0, n) else 2.0
                node = sum((c[a] * n + t[a]) * self.strides[a] for a in range(dim))
                for comp in range(s):
                    PT[s * node + comp, s * k + comp] += w * (h ** dim / 2 ** dim)
        return PT


def _patch_extent(c, ell, N):
    lo = tuple(max(ci - ell, 0) for ci in c)
    hi = tuple(min(ci + ell, N - 1) for ci in c)
    m = tuple(hi[a] - lo[a] + 1 for a in range(len(c)))
    return lo, hi, m


class SlodOracle:
    """Serial restatement of LOD::run() stages create_patches .. assemble_global_matrix."""

    def __init__(self, prob: SlodProblem):
        self.prob = prob
        self._shapes = {}
        self._presaved = None
        d = prob.dim
        if prob.problem == "diffusion":
            assert prob.spacedim == 1
            self.qpts, self.Kq = subcell_gauss_matrices(d, prob.h, "diffusion")
        else:
            assert prob.spacedim == d
            self.qpts, self.Kmu, self.Klam = subcell_gauss_matrices(d, prob.h, "elasticity")
        self.patches: list[PatchResult] = []

    # -- structure ---------------------------------------------------------------------------
    def shape_for(self, c):
        pr = self.prob
        lo, hi, m = _patch_extent(c, pr.oversampling, pr.N)
        dom_lo = tuple(l == 0 for l in lo)
        dom_hi = tuple(hh == pr.N - 1 for hh in hi)
        crel = tuple(c[a] - lo[a] for a in range(pr.dim))
        key = (m, dom_lo, dom_hi, crel)
        if key not in self._shapes:
            self._shapes[key] = PatchShape(pr.dim, pr.spacedim, pr.n_subdivisions, m, dom_lo, dom_hi, crel)
        return self._shapes[key], lo

    # -- assembly (include/Diffusion.h:143-205, include/Elasticity.h:211-295) -------------------
    def assemble_patch_stiffness(self, shape: PatchShape, lo):
        pr = self.prob
        h = pr.h
        n_sub = shape.sub_coords.shape[0]
        origin = (np.array(lo) * pr.n_subdivisions + shape.sub_coords) * h          # (n_sub, dim)
        nq = len(self.qpts)
        pts = origin[:, None, :] + np.array(self.qpts)[None, :, :] * h              # (n_sub, nq, dim)
        if pr.problem == "diffusion":
            a = pr.coefficients[0].at(pts)                                          # (n_sub, nq)
            loc = np.einsum("sq,qij->sij", a, self.Kq)
        else:
            lam = pr.coefficients[0].at(pts)
            mu = pr.coefficients[1].at(pts)
            loc = np.einsum("sq,qij->sij", mu, self.Kmu) + np.einsum("sq,qij->sij", lam, self.Klam)
        nd = shape.sub_dofs.shape[1]
        rows = np.repeat(shape.sub_dofs, nd, axis=1).ravel()
        cols = np.tile(shape.sub_dofs, (1, nd)).ravel()
        A = sp.coo_matrix((loc.ravel(), (rows, cols)), shape=(shape.Nf, shape.Nf)).tocsr()
        A.sum_duplicates()
        return A

    # -- per patch (source/LOD.cc:345-767) -------------------------------------------------------
    def compute_patch(self, pid: int) -> PatchResult:
        pr = self.prob
        dim, s, ell = pr.dim, pr.spacedim, pr.oversampling
        c = morton_decode(pid, dim, pr.n_global_refinements)
        shape, lo = self.shape_for(c)
        H, h = pr.H, pr.h
        full_size = shape.Nc == (2 * ell + 1) ** dim
        info = {}

        # source/LOD.cc:354-362, 433-451 (quirk B)
        if pr.quirk_presaved and full_size and self._presaved is not None:
            A = self._presaved.copy()
        else:
            A = self.assemble_patch_stiffness(shape, lo)
            if pr.quirk_presaved and full_size:
                self._presaved = A.copy()

        PT = shape.projection_PT(h)                       # :470-496
        b, db, it = shape.b, shape.db, shape.internal
        slod = not (not pr.stabilize or ell == 0 or pr.N ** dim == shape.Nc)   # :563-564
        if pr.stabilize and len(b) > 0:
            PT_boundary = PT[b, :].copy()                 # :498-506 (before zeroing)
        PT[b, :] = 0.0                                    # :512-518
        PT[db, :] = 0.0
        if pr.stabilize and len(b) > 0:
            S_boundary = A[b, :][:, it].toarray()         # :520-528 (unconstrained A)

        # :537-543  clear_row(j, 1): zero the row, put 1 on the diagonal
        def clear_rows(M, rows):
            M = M.tolil(copy=True)
            for j in rows:
                M.rows[j] = [int(j)]
                M.data[j] = [1.0]
            return M.tocsr()

        semi = clear_rows(A, db)
        A0 = clear_rows(semi, b)

        # :546  Gauss_elimination -> Amesos KLU; here SuperLU
        lu = spla.splu(A0.tocsc())
        Ainv_PT = lu.solve(PT)
        M = PT.T @ Ainv_PT                                # :548
        M /= H ** dim                                     # :551
        info["cond_M"] = float(np.linalg.cond(M))
        Minv = np.linalg.inv(M)                           # :553 gauss_jordan

        basis = np.zeros((s, shape.Nf))
        if not slod:
            for d in range(s):                            # :570-593
                t = Minv[:, d]
                phi = Ainv_PT @ t
                basis[d] = phi / np.linalg.norm(phi)
        else:
            Xi = Ainv_PT[it, :]                           # :609-611
            B_full = S_boundary @ Xi                      # :612
            BD = B_full @ Minv - PT_boundary @ Minv       # :616-618
            ncand = shape.Ncd - 1
            info["trunc_steps"] = []
            info["cond_G"] = []
            info["dinf"] = []
            info["sigma"] = []
            info["G"] = BD.T @ BD
            info["Minv"] = Minv
            info["X"] = Xi
            for d in range(s):
                B_d0 = BD[:, d]
                other = [k for k in range(shape.Ncd) if k != d]     # :637-640
                newBD = BD[:, other]
                G = newBD.T @ newBD                       # :660
                g = newBD.T @ B_d0                        # :662
                U, sig, Vt = np.linalg.svd(G)             # :667 compute_inverse_svd(1e-15)
                winv = np.where(sig > 1e-15 * sig[0], 1.0 / np.where(sig > 0, sig, 1.0), 0.0)
                d_i = -(Vt.T @ (winv * (U.T @ g)))        # :669-671
                info["cond_G"].append(float(sig[0] / sig[-1]) if sig[-1] > 0 else float("inf"))
                info["dinf"].append(float(np.abs(d_i).max()))
                info["sigma"].append(sig.copy())
                steps = 0
                for i in range(ncand - 1, -1, -1):        # :703-725
                    if np.abs(d_i).max() < 0.5:
                        break
                    d_i = d_i + Vt[i, :] * (U[:, i] @ g) * winv[i]
                    steps += 1
                info["trunc_steps"].append(steps)
                c_i = Minv[:, d].copy()                   # :727-731
                for idx, k in enumerate(other):           # :735-743
                    c_i += d_i[idx] * Minv[:, k]
                phi_int = Xi @ c_i                        # :745
                phi = np.zeros(shape.Nf)                  # LODtools.h:305-332
                phi[it] = phi_int
                basis[d] = phi / np.linalg.norm(phi)      # :752

        prem = np.stack([semi @ basis[d] for d in range(s)])   # :758-765
        cells = np.array([morton_encode(tuple(lo[a] + cr[a] for a in range(dim)), dim,
                                        pr.n_global_refinements) for cr in shape.cells_rel],
                         dtype=np.uint32)
        info["slod"] = slod
        return PatchResult(pid, tuple(lo), shape.m, cells, basis, prem, info)

    def compute_basis(self, pids=None):
        pr = self.prob
        n_patches = pr.N ** pr.dim
        pids = range(n_patches) if pids is None else pids
        self.patches = [self.compute_patch(p) for p in pids]
        return self.patches

    # -- global scatter + product (source/LOD.cc:860-973) -----------------------------------------
    def global_node_ids(self, res: PatchResult) -> np.ndarray:
        """Global lexicographic fine-node id of every patch node."""
        pr = self.prob
        n = pr.n_subdivisions
        G = pr.N * n + 1
        shape, _ = self.shape_for(morton_decode(res.pid, pr.dim, pr.n_global_refinements))
        g = np.zeros(shape.n_nodes, dtype=np.int64)
        mul = 1
        for a in range(pr.dim):
            g += (shape.node_coords[:, a] + res.lo[a] * n) * mul
            mul *= G
        return g

    def assemble_global_matrix(self, patches=None):
        """K = C^T (A C) with C[:, s*pid+d] = phi_{pid,d}; explicit zeros kept so that the pattern is
        the structural product pattern (patches sharing a fine node)."""
        pr = self.prob
        s = pr.spacedim
        patches = self.patches if patches is None else patches
        n_fine = s * (pr.N * pr.n_subdivisions + 1) ** pr.dim
        n_coarse = s * pr.N ** pr.dim
        rows, cols, v_phi, v_aphi = [], [], [], []
        for res in patches:
            g = self.global_node_ids(res)
            gd = (g[:, None] * s + np.arange(s)[None, :]).ravel()
            for d in range(s):
                rows.append(gd)
                cols.append(np.full(gd.shape, s * res.pid + d))
                v_phi.append(res.basis[d])
                v_aphi.append(res.basis_premultiplied[d])
        rows = np.concatenate(rows)
        cols = np.concatenate(cols)
        C = sp.csc_matrix((np.concatenate(v_phi), (rows, cols)), shape=(n_fine, n_coarse))
        AC = sp.csc_matrix((np.concatenate(v_aphi), (rows, cols)), shape=(n_fine, n_coarse))
        ones = sp.csc_matrix((np.ones(rows.shape), (rows, cols)), shape=(n_fine, n_coarse))
        return self.galerkin_product(C, AC, ones), C, AC

    @staticmethod
    def galerkin_product(C, AC, ones):
        """K = C^T (A C) on the structural pattern of the sparse product (source/LOD.cc:970-971: Tmmult of the two
        Trilinos matrices): (i, j) is an entry iff columns i and j of C store a common row -- stored zeros count.
        `ones` has the stored pattern of C with all values 1."""
        Kv = (C.T @ AC).tocsr()
        pattern = (ones.T @ ones).tocsr()   # structural product pattern (explicit zeros of C are entries)
        pattern.sort_indices()
        r, c = pattern.nonzero()
        return sp.csr_matrix((np.asarray(Kv[r, c]).ravel(), pattern.indices.copy(), pattern.indptr.copy()),
                             shape=pattern.shape)

    # -- fine right-hand side used by Poisson_LOD_Example (f = 1, zero Dirichlet rows) ------------
    def fem_rhs_constant_one(self):
        pr = self.prob
        assert pr.spacedim == 1
        G = pr.N * pr.n_subdivisions + 1
        w1 = np.full(G, pr.h)
        w1[0] = w1[-1] = 0.0    # constrained rows get rhs 0
        rhs = w1
        for _ in range(pr.dim - 1):
            rhs = np.multiply.outer(w1, rhs)
        return rhs.ravel()


    def fem_rhs(self, fn):
        """Fine right-hand side F_i = int f phi_i with QIterated(QGauss<1>(2), n) per coarse cell, i.e. 2^dim Gauss
        points per sub-cell (include/Diffusion.h:149-153, 188-191; the elasticity loop include/Elasticity.h:262-270 is
        the same per component), rows on the domain boundary constrained to 0 (source/LOD.cc:1021-1027).
        ``fn(points[npts, dim]) -> values[npts, spacedim]``.  Lexicographic numbering node * spacedim + comp."""
        pr = self.prob
        dim, s, h = pr.dim, pr.spacedim, pr.h
        M = pr.N * pr.n_subdivisions
        G = M + 1
        t = 0.5 - 0.5 / math.sqrt(3.0)
        gp = np.array([t, 1.0 - t])                 # Gauss points of the unit interval, weights 1/2 each
        F = np.zeros((G,) * dim + (s,))             # indexed [z][y][x][comp] (x fastest in the flattened order)
        cells = np.stack(np.meshgrid(*[np.arange(M)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)  # (x, y, z)
        for q in itertools.product(range(2), repeat=dim):
            pts = (cells + gp[list(q)]) * h
            val = np.asarray(fn(pts), dtype=float).reshape(len(pts), s) * (h / 2.0) ** dim
            for loc in itertools.product(range(2), repeat=dim):
                shape = np.ones(())
                for a in range(dim):
                    shape = shape * (gp[q[a]] if loc[a] == 1 else 1.0 - gp[q[a]])
                idx = tuple((cells[:, a] + loc[a]) for a in reversed(range(dim)))
                np.add.at(F, idx, val * shape)
        for a in range(dim):                         # homogeneous Dirichlet rows
            sl = [slice(None)] * (dim + 1)
            for end in (0, G - 1):
                sl[dim - 1 - a] = end
                F[tuple(sl)] = 0.0
        return F.reshape(-1)

    def fem_solve(self, F):
        """Fine-scale FEM solution of assemble_and_solve_fem_problem (source/LOD.cc:1004-1094): global Q_iso_Q1 stiffness
        of the whole domain, homogeneous Dirichlet rows, right-hand side F (from fem_rhs).  Solved directly instead of
        with the reference's CG + AMG.  Lexicographic numbering."""
        pr = self.prob
        whole = PatchShape(pr.dim, pr.spacedim, pr.n_subdivisions, (pr.N,) * pr.dim, (True,) * pr.dim, (True,) * pr.dim,
                           (0,) * pr.dim)
        A = self.assemble_patch_stiffness(whole, (0,) * pr.dim).tocsr()
        ii = whole.internal
        u = np.zeros(whole.Nf)
        u[ii] = spla.spsolve(A[ii][:, ii].tocsc(), np.asarray(F, dtype=float)[ii])
        return u, A

    def fine_norm_matrices(self):
        """Exact Q1 mass matrix and component-wise Laplace matrix of the sub-cell grid: ||v||_L2^2 = v.Mv and
        |v|_H1^2 = v.Lv are the norms compare_lod_with_fem tabulates (source/LOD.cc:1252; the reference integrates them
        with a Gauss rule on the coarse cells, which is not exact for Q_iso_Q1 functions).  Lexicographic numbering."""
        pr = self.prob
        dim, s, h = pr.dim, pr.spacedim, pr.h
        G = pr.N * pr.n_subdivisions + 1
        m1 = sp.diags([np.full(G - 1, h / 6), np.r_[h / 3, np.full(G - 2, 2 * h / 3), h / 3], np.full(G - 1, h / 6)],
                      [-1, 0, 1])
        k1 = sp.diags([np.full(G - 1, -1 / h), np.r_[1 / h, np.full(G - 2, 2 / h), 1 / h], np.full(G - 1, -1 / h)],
                      [-1, 0, 1])

        def kron_chain(mats):           # x fastest: the x factor is the innermost (last) Kronecker factor
            out = mats[-1]
            for m in reversed(mats[:-1]):
                out = sp.kron(out, m)
            return out

        M = kron_chain([m1] * dim)
        L = sum(kron_chain([k1 if a == k else m1 for a in range(dim)]) for k in range(dim))
        eye = sp.identity(s)
        return sp.kron(M, eye).tocsr(), sp.kron(L, eye).tocsr()

    def reference_error_norms(self, v):
        """(L2_norm, Linfty_norm, H1_norm) of a fine vector the way ParsedConvergenceTable::difference computes them
        (source/LOD.cc:1252, include/LOD.h:111-115): VectorTools::integrate_difference against zero on the cells of
        dof_handler_fine -- the coarse cells with FE_Q_iso_Q1(n) -- with QGauss((degree + 1) * 2) = 2 (n + 1) points per
        direction [deal.II 9.6 parsed_convergence_table: q_gauss((dh.get_fe().degree + 1) * 2)]; default norm list
        L2, Linfty, H1 (full norm).  Lexicographic numbering node * spacedim + comp."""
        pr = self.prob
        dim, s, n, N, h, H = pr.dim, pr.spacedim, pr.n_subdivisions, pr.N, pr.h, pr.H
        G = N * n + 1
        nq = 2 * (n + 1)
        gx, gw = np.polynomial.legendre.leggauss(nq)
        gx, gw = 0.5 * (gx + 1.0), 0.5 * gw
        V = np.asarray(v, dtype=float).reshape((G,) * dim + (s,))     # [z][y][x][comp]
        t = gx * n
        sub = np.minimum(t.astype(int), n - 1)
        xi = t - sub
        l2 = h1 = 0.0
        linf = 0.0
        for cell in itertools.product(range(N), repeat=dim):          # (x, y, z)
            for q in itertools.product(range(nq), repeat=dim):
                w = np.prod([gw[q[a]] * H for a in range(dim)])
                o = [cell[a] * n + sub[q[a]] for a in range(dim)]
                val = np.zeros(s)
                grad = np.zeros((dim, s))
                for l in itertools.product(range(2), repeat=dim):
                    idx = tuple(o[a] + l[a] for a in reversed(range(dim)))
                    u = V[idx]
                    sh = np.prod([xi[q[a]] if l[a] else 1.0 - xi[q[a]] for a in range(dim)])
                    val += u * sh
                    for a in range(dim):
                        g = (1.0 if l[a] else -1.0) / h
                        for a2 in range(dim):
                            if a2 != a:
                                g *= xi[q[a2]] if l[a2] else 1.0 - xi[q[a2]]
                        grad[a] += u * g
                l2 += w * float(val @ val)
                h1 += w * float((grad * grad).sum())
                linf = max(linf, float(np.abs(val).max()))
        return math.sqrt(l2), linf, math.sqrt(l2 + h1)

    # -- LOD::solve (source/LOD.cc:975-1001) and the prolongation (source/LOD.cc:1251) -----------------------------
    @staticmethod
    def solve_coarse(K, b, max_steps=100, tolerance=1e-10, reduction=1e-10, omega=1.2, direct=False):
        """CG preconditioned with one SSOR(omega = 1.2) sweep, stopped by deal.II's ReductionControl
        (||r|| <= tolerance or ||r|| <= reduction ||r_0||); ``direct=True`` is the reference's debugging branch
        (SolverDirect, source/LOD.cc:984-989).  Returns (u, steps)."""
        K = sp.csr_matrix(K)
        b = np.asarray(b, dtype=float)
        if direct:
            return spla.spsolve(K.tocsc(), b), 0
        D = K.diagonal()
        Lw = (sp.tril(K, -1) * omega + sp.diags(D)).tocsr()
        Uw = (sp.triu(K, 1) * omega + sp.diags(D)).tocsr()

        def prec(r):   # (D + w L)^-1 ... D ... (D + w U)^-1, scaled by w (2 - w)
            y = spla.spsolve_triangular(Lw, r, lower=True)
            return omega * (2.0 - omega) * spla.spsolve_triangular(Uw, D * y, lower=False)

        u = np.zeros_like(b)
        r = b.copy()
        r0 = np.linalg.norm(r)
        if r0 <= tolerance:
            return u, 0
        z = prec(r)
        p = z.copy()
        rz = r @ z
        for it in range(1, max_steps + 1):
            q = K @ p
            alpha = rz / (p @ q)
            u += alpha * p
            r -= alpha * q
            res = np.linalg.norm(r)
            if res <= tolerance or res <= reduction * r0:
                return u, it
            z = prec(r)
            rz_new = r @ z
            p = z + (rz_new / rz) * p
            rz = rz_new
        raise RuntimeError(f"SolverControl::NoConvergence after {max_steps} steps, residual {res:.3e}")


# ----------------------------------------------------------------------------------------------
# helpers used by golden tests that are not part of the LOD class
# ----------------------------------------------------------------------------------------------
def qiso_cell_matrix(dim: int, n: int, hier: bool = True) -> np.ndarray:
    """Laplace cell matrix of FE_Q_iso_Q1(n) on the unit cell (tests/fe_q_iso_q1_01.cc)."""
    h = 1.0 / n
    _, Kq = subcell_gauss_matrices(dim, h, "diffusion")
    Ks = Kq.sum(axis=0)
    p = n + 1
    A = np.zeros((p ** dim, p ** dim))
    corner = [tuple(reversed(v)) for v in itertools.product(*([(0, 1)] * dim))]
    for c in itertools.product(*([range(n)] * dim)):
        c = tuple(reversed(c))
        idx = []
        for cv in corner:
            j = 0
            for a in reversed(range(dim)):
                j = j * p + c[a] + cv[a]
            idx.append(j)
        A[np.ix_(idx, idx)] += Ks
    if hier:
        l2h = lexicographic_to_hierarchic(dim, n)
        B = np.zeros_like(A)
        B[np.ix_(l2h, l2h)] = A
        return B
    return A


def structured_patch_poisson(repetitions, n: int, centre, overlap: int):
    """tests/solve_poisson_problem_on_patch_01.cc: -Laplace u = 1 on the patch around ``centre``
    with zero values on the whole patch boundary; returns the solution scattered into the global
    lexicographic fine vector."""
    dim = len(repetitions)
    lo = [max(centre[a] - overlap, 0) for a in range(dim)]
    hi = [min(centre[a] + overlap, repetitions[a] - 1) for a in range(dim)]
    m = [hi[a] - lo[a] + 1 for a in range(dim)]
    h = 1.0 / (repetitions[0] * n)
    shape = PatchShape(dim, 1, n, m, (False,) * dim, (False,) * dim, (0,) * dim)
    _, Kq = subcell_gauss_matrices(dim, h, "diffusion")
    Ks = Kq.sum(axis=0)
    nd = shape.sub_dofs.shape[1]
    rows = np.repeat(shape.sub_dofs, nd, axis=1).ravel()
    cols = np.tile(shape.sub_dofs, (1, nd)).ravel()
    vals = np.tile(Ks.ravel(), shape.sub_dofs.shape[0])
    A = sp.coo_matrix((vals, (rows, cols)), shape=(shape.Nf, shape.Nf)).tocsr()
    rhs = np.zeros(shape.Nf)
    np.add.at(rhs, shape.sub_nodes.ravel(), h ** dim / 2 ** dim)
    it = shape.internal
    u = np.zeros(shape.Nf)
    u[it] = spla.spsolve(A[it, :][:, it].tocsc(), rhs[it])
    G = [repetitions[a] * n + 1 for a in range(dim)]
    out = np.zeros(int(np.prod(G)))
    gid = np.zeros(shape.n_nodes, dtype=np.int64)
    mul = 1
    for a in range(dim):
        gid += (shape.node_coords[:, a] + lo[a] * n) * mul
        mul *= G[a]
    out[gid] = u
    return out
