"""In-library multi-GPU: NCCL inside libslod_b200.so.

* one handle, N devices (slod_params.n_gpus = N): slod_compute_basis / slod_assemble_coarse split the patches into the
  reference's contiguous even ranges (source/LOD.cc:116-118), one host thread and one NCCL rank per device;
* one handle per process (slod_comm_init + slod_offline_distributed) is exercised by `bench.py --gpus N` under torchrun,
  which also compares the all-gathered result with a single-GPU recomputation bit for bit.

Results must be BIT-identical to the single-device run: patches are independent and every reduction has a fixed order.
Needs >= 2 visible GPUs (skipped otherwise; the round-end GPU test box has one)."""
import importlib

import numpy as np
import pytest

pkg = importlib.import_module("dealii-slod_b200")
pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _run(n_gpus, case, tab, r):
    ctx = pkg.SlodContext(stabilize=True, n_gpus=n_gpus, device=0, **case)
    ctx.set_coefficient(0, r, tab)
    ctx.compute_basis()
    ctx.assemble_coarse()
    phi, aphi = ctx.all_basis()
    rowptr, col, val = ctx.coarse_csr()
    out = (phi.copy(), aphi.copy(), rowptr.copy(), col.copy(), val.copy(), ctx.timings().copy())
    ctx.close()
    return out


@pytest.mark.parametrize("case", [dict(dim=2, spacedim=1, n_global_refinements=4, n_subdivisions=2, oversampling=2),
                                  dict(dim=3, spacedim=1, n_global_refinements=3, n_subdivisions=2, oversampling=2)],
                         ids=["2d", "3d"])
def test_one_handle_drives_all_gpus(case):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    dim = case["dim"]
    r = case["n_global_refinements"] + 1          # one table cell per fine sub-cell
    tab = 1.0 + 99.0 * np.random.default_rng(5).random((2 ** r) ** dim)
    ref = _run(1, case, tab, r)
    for ng in sorted({2, n}):
        got = _run(ng, case, tab, r)
        for a, b in zip(ref[:5], got[:5]):
            assert np.array_equal(a, b)
        assert got[5][0] > 0
