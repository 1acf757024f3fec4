"""Multi-GPU plumbing of the SLOD offline phase: one process per GPU, patches partitioned into contiguous ranges.

The reference declares the same ownership rule -- ``locally_owned_patches`` is
``Utilities::MPI::create_evenly_distributed_partitioning(comm, n_patches)`` (source/LOD.cc:116-118): contiguous blocks
of patch ids, the first ``n % size`` ranks holding one more -- but never finishes its MPI path
(source/LOD.cc:228, :895).  Patch ids are Morton codes, so a contiguous range is a compact box of the mesh.

Per-patch work needs no communication.  The coarse matrix ``K[(p,d),(q,e)] = phi_{p,d} . (A phi)_{q,e}``
(source/LOD.cc:970-971) couples a patch with its <= (4 l + 3)^dim neighbours, so the only exchange is an all-gather of
``A phi`` before ``k_coarse`` and an all-gather of the disjoint block-ELL row blocks after it.  ``torch.distributed``
(NCCL on GPUs, gloo in the CPU tests) carries both; no reduction is involved, results are bit-identical to a
single-rank run.
"""
from __future__ import annotations


def owned_range(n_patches: int, rank: int, world: int):
    """[begin, end) of the patch ids owned by ``rank`` (create_evenly_distributed_partitioning, source/LOD.cc:116-118)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_patches, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_ranges(n_patches: int, world: int):
    return [owned_range(n_patches, r, world) for r in range(world)]


def all_gather_rows(dist, full, ranges, rank, rows_per_patch=1):
    """In-place all-gather of row blocks of ``full`` (dim 0 = patch * rows_per_patch): rank r contributes the rows of its
    patch range.  Equal ranges use one ``all_gather_into_tensor``; ragged ranges fall back to one broadcast per rank
    (the blocks are disjoint, so there is nothing to reduce)."""
    world = len(ranges)
    if world == 1:
        return
    sizes = {e - b for b, e in ranges}
    b, e = ranges[rank]
    if len(sizes) == 1:
        flat = full.view(-1)
        mine = full[b * rows_per_patch:e * rows_per_patch].reshape(-1)
        dist.all_gather_into_tensor(flat, mine)
    else:
        for r, (rb, re) in enumerate(ranges):
            if re > rb:
                dist.broadcast(full[rb * rows_per_patch:re * rows_per_patch], src=r)


class DistributedOffline:
    """The offline phase over ``world`` ranks.  ``compute_basis(p0, p1)`` and ``assemble_coarse(p0, p1)`` are callables
    that fill the rows of the range in the shared-layout tensors ``phi``/``aphi`` ([n_patches, s, stride]) and ``K``
    ([n_patches * s, ell_width]); on a GPU they are the ``slod_*_device`` entry points, in the CPU tests an oracle."""

    def __init__(self, dist, rank, world, n_patches, spacedim, phi, aphi, K, compute_basis, assemble_coarse,
                 synchronize=None):
        self.dist, self.rank, self.world = dist, rank, world
        self.ranges = all_ranges(n_patches, world)
        self.p0, self.p1 = self.ranges[rank]
        self.s = spacedim
        self.phi, self.aphi, self.K = phi, aphi, K
        self._basis, self._coarse = compute_basis, assemble_coarse
        self._sync = synchronize    # slod_synchronize of the handle: the device entry points only enqueue

    def step(self, gather_K=True):
        self._basis(self.p0, self.p1)
        all_gather_rows(self.dist, self.aphi, self.ranges, self.rank)          # A phi of every patch, everywhere
        self._coarse(self.p0, self.p1)
        if gather_K:
            all_gather_rows(self.dist, self.K, self.ranges, self.rank, self.s)  # disjoint K row blocks
        if self._sync is not None:
            self._sync()     # numerical status of the range (SLOD_ERR_NUMERIC raises here)
