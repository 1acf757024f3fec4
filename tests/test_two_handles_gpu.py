"""Two handles with DIFFERENT parameters on one device, driven concurrently from two host threads and two streams
(VERDICT r01 "what's weak" 11: the kernels read one parameter block per device).  The library binds the block per call
(capi.cu: ParamBinding), so the interleaved results must be bit-identical to the ones each handle computes alone."""
import importlib
import threading

import numpy as np
import pytest

pkg = importlib.import_module("dealii-slod_b200")
pytestmark = pytest.mark.gpu

CASES = [dict(dim=2, spacedim=1, n_global_refinements=4, n_subdivisions=2, oversampling=2),
         dict(dim=3, spacedim=1, n_global_refinements=3, n_subdivisions=2, oversampling=1),
         dict(dim=2, spacedim=2, n_global_refinements=3, n_subdivisions=2, oversampling=1, problem=1)]


def _tables(case, seed):
    r = case["n_global_refinements"] + 1
    n_fields = 1 if case["spacedim"] == 1 else 2
    rng = np.random.default_rng(seed)
    return r, [1.0 + 99.0 * rng.random((2 ** r) ** case["dim"]) for _ in range(n_fields)]


def _make(case, seed):
    ctx = pkg.SlodContext(stabilize=True, **case)   # problem=1: elasticity (two coefficient fields)
    r, tabs = _tables(case, seed)
    for f, t in enumerate(tabs):
        ctx.set_coefficient(f, r, t)
    return ctx


def _offline(ctx):
    ctx.compute_basis()
    ctx.assemble_coarse()
    phi, aphi = ctx.all_basis()
    _, _, val = ctx.coarse_csr()
    return phi.copy(), aphi.copy(), val.copy()


def test_interleaved_handles_match_sequential():
    alone = []
    for i, case in enumerate(CASES):
        ctx = _make(case, 11 + i)
        alone.append(_offline(ctx))
        ctx.close()
    # all handles alive at once, every stage of one handle followed by a stage of another
    ctxs = [_make(case, 11 + i) for i, case in enumerate(CASES)]
    for c in ctxs:
        c.compute_basis()
    for c in ctxs:
        c.assemble_coarse()
    for c, ref in zip(ctxs, alone):
        phi, aphi = c.all_basis()
        _, _, val = c.coarse_csr()
        assert np.array_equal(phi, ref[0]) and np.array_equal(aphi, ref[1]) and np.array_equal(val, ref[2])
    for c in ctxs:
        c.close()


def test_concurrent_threads_match_sequential():
    import torch
    alone = []
    for i, case in enumerate(CASES):
        ctx = _make(case, 21 + i)
        alone.append(_offline(ctx))
        ctx.close()
    results = [None] * len(CASES)
    errors = []

    def work(i):
        try:
            torch.cuda.set_device(0)
            ctx = _make(CASES[i], 21 + i)
            n, stride, s = ctx.n_patches, ctx.basis_stride, CASES[i]["spacedim"]
            st = torch.cuda.Stream()
            phi = torch.zeros((n, s, stride), dtype=torch.float64, device="cuda")
            aphi = torch.zeros_like(phi)
            for _ in range(3):   # several rounds so that the calls of the threads really interleave
                ctx.compute_basis_device(0, n, phi.data_ptr(), aphi.data_ptr(), stream=st.cuda_stream)
                ctx.synchronize()
            results[i] = (phi.cpu().numpy().reshape(-1), aphi.cpu().numpy().reshape(-1))
            ctx.close()
        except Exception as e:   # noqa: BLE001 - reported by the main thread
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(CASES))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for got, ref in zip(results, alone):
        assert np.array_equal(got[0], ref[0].reshape(-1))
        assert np.array_equal(got[1], ref[1].reshape(-1))
