"""Multi-rank host logic on CPU: world_size-2 (and 3, ragged) `gloo` runs of the partition / all-gather plumbing that
bench.py uses with NCCL.  The per-patch compute is played by the oracle (no GPU here); the block-ELL -> CSR merge and
the integer maps come from the C library through a maps-only handle."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.slod_oracle import CoefficientTable, SlodOracle, SlodProblem, morton_decode, morton_encode

pkg = importlib.import_module("dealii-slod_b200")
part = importlib.import_module("dealii-slod_b200.partition")

CASE = dict(dim=2, s=1, ref=3, n=2, ell=1, r=4)


def test_owned_range_is_even_contiguous_partition():
    for n, world in ((64, 2), (64, 3), (5, 8), (32768, 8), (7, 1)):
        ranges = part.all_ranges(n, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        sizes = [e - b for b, e in ranges]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)   # the first n % world get one more
    with pytest.raises(ValueError):
        part.owned_range(10, 3, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle():
    c = CASE
    tab = 1.0 + 99.0 * np.random.default_rng(11).random((2 ** c["r"]) ** c["dim"])
    return SlodOracle(SlodProblem(dim=c["dim"], spacedim=c["s"], n_global_refinements=c["ref"], n_subdivisions=c["n"],
                                  oversampling=c["ell"], stabilize=True, coefficients=[CoefficientTable(c["dim"], c["r"], tab)]))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = CASE
    dim, s, ref, n, ell = c["dim"], c["s"], c["ref"], c["n"], c["ell"]
    orc = _oracle()
    ctx = pkg.SlodContext(dim=dim, spacedim=s, n_global_refinements=ref, n_subdivisions=n, oversampling=ell,
                          stabilize=True, device=-2)                  # maps only: no GPU in this test
    npch, stride, ellw = ctx.n_patches, ctx.basis_stride, ctx.ell_width
    phi = torch.zeros((npch, s, stride), dtype=torch.float64)
    aphi = torch.zeros_like(phi)
    K = torch.zeros((npch * s, ellw), dtype=torch.float64)
    N, w = 2 ** ref, 2 * ell + 1
    G = N * n + 1
    results = {}

    def compute_basis(p0, p1):
        for res in orc.compute_basis(range(p0, p1)):
            results[res.pid] = res
            nf = res.basis.shape[1]
            phi[res.pid, :, :nf] = torch.from_numpy(res.basis)
            aphi[res.pid, :, :nf] = torch.from_numpy(res.basis_premultiplied)

    def node_ids(pid):
        info = ctx.patch_info(pid)
        p = [m * n + 1 for m in info["m"]]
        ix = np.indices(p[::-1]).reshape(dim, -1)[::-1]
        return sum((ix[a] + info["lo"][a] * n) * G ** a for a in range(dim))

    def assemble_coarse(p0, p1):
        # K[(p,d),(q,e)] over the shared fine nodes, written at slot ((Dy+w)*(2w+1)+(Dx+w))*s+e (include/slod.h)
        for pid in range(p0, p1):
            cp = morton_decode(pid, dim, ref)
            gp = node_ids(pid)
            for Dy in range(-w, w + 1):
                for Dx in range(-w, w + 1):
                    cq = (cp[0] + Dx, cp[1] + Dy)
                    if not all(0 <= x < N for x in cq):
                        continue
                    qid = morton_encode(cq, dim, ref)
                    gq = node_ids(qid)
                    common, ip, iq = np.intersect1d(gp, gq, return_indices=True)
                    if common.size == 0:
                        continue
                    slot = (Dy + w) * (2 * w + 1) + (Dx + w)
                    nfp, nfq = gp.size, gq.size
                    K[pid, slot] = float(phi[pid, 0, :nfp].numpy()[ip] @ aphi[qid, 0, :nfq].numpy()[iq])

    job = part.DistributedOffline(dist, rank, world, npch, s, phi, aphi, K, compute_basis, assemble_coarse)
    assert (job.p0, job.p1) == part.owned_range(npch, rank, world)
    job.step()
    rowptr, col, val = ctx.ell_to_csr(K.numpy())
    if rank == 0:
        np.savez(out, rowptr=rowptr, col=col, val=val, aphi=aphi.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_offline_phase_matches_single_process(tmp_path, world):
    out = str(tmp_path / "k.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    orc = _oracle()
    orc.compute_basis()
    K, _, _ = orc.assemble_global_matrix()
    assert np.array_equal(got["rowptr"], K.indptr) and np.array_equal(got["col"], K.indices)
    assert np.abs(got["val"] - K.data).max() <= 1e-12 * np.abs(K.data).max()
    for res in orc.patches:                      # every rank holds A phi of every patch after the gather
        assert np.array_equal(got["aphi"][res.pid, 0, :res.basis.shape[1]], res.basis_premultiplied[0])
