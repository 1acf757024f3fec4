"""CPU-only checks of the C-ABI library: it loads, exports every symbol the header declares, refuses to
compute without a GPU, and its integer maps agree bit-for-bit with the oracle."""
import importlib
import os
import re

import numpy as np
import pytest

from oracle.slod_oracle import (PatchResult, SlodOracle, SlodProblem, create_patches, global_fine_dof_numbering,
                                morton_decode, patch_local_dof_numbering)

pkg = importlib.import_module("dealii-slod_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    pkg.build_library()
    return pkg.load_library()


def test_exports_match_header(lib):
    hdr = open(os.path.join(ROOT, "include", "slod.h")).read()
    declared = set(re.findall(r"\b(slod_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"slod_ctx", "slod_params"}
    assert declared == set(pkg.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.SlodError) as e:
        pkg.SlodContext(dim=2)
    assert e.value.code == 3
    ctx = pkg.SlodContext(dim=2, device=-2)
    for call in (ctx.compute_basis, ctx.assemble_coarse, lambda: ctx.set_coefficient(0, 0, np.ones(1))):
        with pytest.raises(pkg.SlodError) as e:
            call()
        assert e.value.code == 3


def test_parameter_validation(lib):
    for kw in (dict(dim=1), dict(dim=2, spacedim=2), dict(dim=3, spacedim=3, problem=1), dict(dim=2, n_subdivisions=3),
               dict(dim=2, oversampling=-1), dict(dim=2, n_gpus=-1), dict(dim=2, n_gpus=2)):   # n_gpus > 1 needs devices
        with pytest.raises(pkg.SlodError):
            pkg.SlodContext(device=-2, **kw)
    # a patch without interior fine dofs (one-cell patches with one subdivision) is refused, not computed
    for kw in (dict(dim=2, n_global_refinements=2, n_subdivisions=1, oversampling=0),
               dict(dim=3, n_global_refinements=0, n_subdivisions=1, oversampling=1)):
        with pytest.raises(pkg.SlodError) as e:
            pkg.SlodContext(device=-2, **kw)
        assert e.value.code == 2 and "interior" in str(e.value)
    # fewer interior fine dofs than coarse dofs on some patch: P^T A^-1 P is singular (the reference's gauss_jordan,
    # source/LOD.cc:553, cannot invert it either) -- refused with a clear message instead of a numerical failure later
    for kw in (dict(dim=2, n_global_refinements=3, n_subdivisions=1, oversampling=1),
               dict(dim=3, n_global_refinements=2, n_subdivisions=1, oversampling=1)):
        with pytest.raises(pkg.SlodError) as e:
            pkg.SlodContext(device=-2, **kw)
        assert e.value.code == 2 and "singular" in str(e.value)


@pytest.mark.parametrize("dim,s,ref,n,ell", [(2, 1, 3, 2, 1), (2, 1, 4, 2, 2), (2, 2, 3, 2, 1), (2, 1, 2, 4, 1),
                                             (3, 1, 2, 2, 1), (3, 1, 3, 2, 2), (2, 1, 2, 2, 3)])
def test_integer_maps_bit_exact(lib, dim, s, ref, n, ell):
    ctx = pkg.SlodContext(dim=dim, spacedim=s, n_global_refinements=ref, n_subdivisions=n, oversampling=ell,
                          problem=1 if s > 1 else 0, stabilize=True, device=-2)
    prob = SlodProblem(dim=dim, spacedim=s, n_global_refinements=ref, n_subdivisions=n, oversampling=ell,
                       problem="elasticity" if s > 1 else "diffusion", stabilize=True)
    orc = SlodOracle(prob)
    patches = create_patches(dim, ref, ell)
    assert ctx.n_patches == len(patches)
    gnum = global_fine_dof_numbering(dim, s, ref, n)
    step = max(1, len(patches) // 40)
    for pid in list(range(0, len(patches), step)) + [len(patches) - 1]:
        assert np.array_equal(ctx.patch_cells(pid), patches[pid])
        c = morton_decode(pid, dim, ref)
        shape, lo = orc.shape_for(c)
        info = ctx.patch_info(pid)
        assert info["lo"] == tuple(lo) and info["m"] == shape.m
        assert (info["n_cells"], info["n_fine"], info["n_internal"], info["n_boundary"], info["n_domain_boundary"],
                info["n_coarse"]) == (shape.Nc, shape.Nf, len(shape.internal), len(shape.b), len(shape.db), shape.Ncd)
        assert np.array_equal(ctx.patch_dof_class(pid, 0), shape.internal)
        assert np.array_equal(ctx.patch_dof_class(pid, 1), shape.b)
        assert np.array_equal(ctx.patch_dof_class(pid, 2), shape.db)
        # global fine dofs
        gidx = tuple((shape.node_coords[:, a] + lo[a] * n) for a in reversed(range(dim)))
        expect = (gnum[gidx][:, None] + np.arange(s)[None, :]).ravel()
        assert np.array_equal(ctx.patch_fine_dofs(pid).astype(np.int64), expect)
        # patch-local deal.II numbering
        cells_abs = [tuple(lo[a] + cr[a] for a in range(dim)) for cr in shape.cells_rel]
        lnum = patch_local_dof_numbering(dim, s, n, cells_abs, lo, shape.m)
        lidx = tuple(shape.node_coords[:, a] for a in reversed(range(dim)))
        expect = (lnum[lidx][:, None] + np.arange(s)[None, :]).ravel()
        got = ctx.patch_local_dofs(pid).astype(np.int64)
        assert np.array_equal(got, expect)
        assert sorted(got) == list(range(shape.Nf))


@pytest.mark.parametrize("dim,s,ref,ell", [(2, 1, 3, 1), (2, 2, 2, 1), (2, 1, 4, 2), (3, 1, 2, 1)])
def test_csr_pattern_matches_sparse_product(lib, dim, s, ref, ell):
    """block-ELL -> CSR conversion is host integer logic: feed K == slot-independent values and compare
    the pattern with the structural product pattern of the oracle (source/LOD.cc:970-971)."""
    ctx = pkg.SlodContext(dim=dim, spacedim=s, n_global_refinements=ref, oversampling=ell, problem=1 if s > 1 else 0,
                          device=-2)
    prob = SlodProblem(dim=dim, spacedim=s, n_global_refinements=ref, oversampling=ell,
                       problem="elasticity" if s > 1 else "diffusion")
    orc = SlodOracle(prob)
    fake = []
    for pid in range(ctx.n_patches):
        shape, lo = orc.shape_for(morton_decode(pid, dim, ref))
        ones = np.ones((s, shape.Nf))
        fake.append(PatchResult(pid, tuple(lo), shape.m, None, ones, ones, {}))
    K, _, _ = orc.assemble_global_matrix(fake)
    hK = np.arange(ctx.n_patches * s * ctx.ell_width, dtype=np.float64).reshape(ctx.n_patches * s, -1)
    rowptr, col, val = ctx.ell_to_csr(hK)
    assert np.array_equal(rowptr, K.indptr)
    assert np.array_equal(col, K.indices)
    # values come from the right slots: slot of (p, q) = offset of centres
    w = 2 * ell + 1
    ww = 2 * w + 1
    for r in (0, K.shape[0] // 2, K.shape[0] - 1):
        p = r // s
        cp = morton_decode(p, dim, ref)
        for k in range(rowptr[r], rowptr[r + 1]):
            q, e = divmod(int(col[k]), s)
            cq = morton_decode(q, dim, ref)
            D = [cq[a] - cp[a] + w for a in range(dim)]
            slot = D[0] + ww * D[1] + (ww * ww * D[2] if dim == 3 else 0)
            assert val[k] == hK[r, slot * s + e]


def test_owned_range_is_the_reference_partition(lib):
    """slod_owned_range = create_evenly_distributed_partitioning (source/LOD.cc:116-118), as partition.py restates it."""
    ctx = pkg.SlodContext(dim=2, n_global_refinements=3, device=-2)
    part = importlib.import_module("dealii-slod_b200.partition")
    for world in (1, 2, 3, 5, 8):
        for r in range(world):
            assert ctx.owned_range(r, world) == part.owned_range(ctx.n_patches, r, world)
    with pytest.raises(pkg.SlodError):
        ctx.owned_range(3, 3)
