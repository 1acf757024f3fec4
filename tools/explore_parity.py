"""Run on a GPU box: prints stage-wise and end-to-end parity numbers of the CUDA path against the oracle."""
import sys
import time

import numpy as np

from parity_common import EPS, build_pair, cond_eff, patch_tolerance

CASES = [
    dict(dim=2, s=1, ref=3, n=2, ell=1, stabilize=False),
    dict(dim=2, s=1, ref=4, n=2, ell=1),
    dict(dim=2, s=1, ref=4, n=2, ell=2),
    dict(dim=2, s=2, ref=3, n=2, ell=1),
    dict(dim=2, s=2, ref=4, n=2, ell=2),
    dict(dim=2, s=1, ref=3, n=4, ell=1),
    dict(dim=3, s=1, ref=2, n=2, ell=1),
    dict(dim=3, s=1, ref=3, n=2, ell=1),
    dict(dim=3, s=1, ref=3, n=2, ell=2, sample=12),
]


def main():
    for case in CASES:
        case = dict(case)
        sample = case.pop("sample", None)
        ctx, orc = build_pair(**case)
        t0 = time.time()
        try:
            ctx.compute_basis()
            ctx.assemble_coarse()
        except Exception as e:  # keep going: report the failure
            print(case, "GPU FAILED:", e)
            continue
        tg = time.time() - t0
        npch = ctx.n_patches
        pids = list(range(npch)) if sample is None else list(range(0, npch, max(1, npch // sample)))
        t0 = time.time()
        orc.compute_basis(pids)
        to = time.time() - t0
        worst = []
        stage = [0.0, 0.0, 0.0]
        for res in orc.patches:
            for d in range(case["s"]):
                phi, aphi = ctx.basis(res.pid, d)
                e = np.linalg.norm(phi - res.basis[d])
                ea = np.linalg.norm(aphi - res.basis_premultiplied[d]) / max(np.linalg.norm(res.basis_premultiplied[d]), 1e-300)
                dg = ctx.diagnostics(res.pid, d)
                info = res.info
                ce = cond_eff(info, d) if info["slod"] else 1.0
                worst.append((e / patch_tolerance(info, d), e, ea, res.pid, d, ce,
                              info["trunc_steps"][d] if info["slod"] else 0, int(dg[1]),
                              info["dinf"][d] if info["slod"] else 0, dg[0], int(dg[6])))
        # stage parity on a few patches
        for res in orc.patches[:: max(1, len(orc.patches) // 6)]:
            if not res.info["slod"]:
                continue
            X, Minv, G = ctx.debug_stages(res.pid)
            stage[0] = max(stage[0], np.abs(X - res.info["X"]).max() / np.abs(res.info["X"]).max())
            stage[1] = max(stage[1], np.abs(Minv - res.info["Minv"]).max() / np.abs(res.info["Minv"]).max())
            stage[2] = max(stage[2], np.abs(G - res.info["G"]).max() / np.abs(res.info["G"]).max())
        worst.sort(reverse=True)
        print("==", case, f"gpu {tg:.2f}s oracle {to:.2f}s  timings(ms)={np.round(ctx.timings(), 2)}")
        print("   stage rel err X/Minv/G:", ["%.1e" % v for v in stage])
        wc = [w for w in worst if w[5] < 1e6]
        print("   well-conditioned patches: %d, max dphi %.2e, max dAphi(rel) %.2e" %
              (len(wc), max([w[1] for w in wc], default=0), max([w[2] for w in wc], default=0)))
        for w in worst[:6]:
            print("   ratio %.2f dphi %.2e dAphi %.2e pid %d d %d cond_eff %.1e steps(or/gpu) %d/%d dinf %.4f/%.4f sweeps %d" % w)
        mism = [w for w in worst if w[6] != w[7]]
        print("   truncation-step mismatches:", len(mism), [(w[3], w[6], w[7]) for w in mism[:5]])
        if sample is None:
            K, _, _ = orc.assemble_global_matrix()
            rowptr, col, val = ctx.coarse_csr()
            same = np.array_equal(rowptr, K.indptr) and np.array_equal(col, K.indices)
            print("   K pattern equal:", same, " rel err max|dK|/max|K| = %.2e" % (np.abs(val - K.data).max() / np.abs(K.data).max()))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
