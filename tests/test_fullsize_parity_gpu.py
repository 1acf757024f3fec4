"""BASELINE.json configurations at FULL size, sampled against BOTH CPU checkers:

* >= 128 patches per configuration spread over the patch classes (interior, face, edge, corner by the number of axes
  on which the patch touches the domain boundary): phi, A phi and the truncation step counts of the CUDA path against
  the numpy oracle (oracle/slod_oracle.py) AND against the C++ CPU port (oracle/cpu/slod_cpu.cc);
* complete coarse-matrix rows of >= 8 patches (one Morton block of 2^dim patches in the interior and one at the domain
  corner), the CPU port computing the basis of all their <= (4 l + 3)^dim neighbours: pattern bit-exact, values at
  1e-10 * max|K| on the uniform fields;
* every patch is counted as tight (<= 1e-10 against the oracle), relaxed (ill-conditioned selection: within 50x the
  oracle's own sensitivity to a 4-ulp perturbation of the Gram matrix) or skipped (a discontinuous
  decision of source/LOD.cc:667 / :703 is within rounding of flipping); the fractions are asserted and everything is
  written to profiles/parity_r02.json when SLOD_WRITE_PARITY=1 (max, p99, per-class counts, outlier list).

Tolerances: 1e-10 (north star) wherever the selection is well conditioned.  The binary {1, 1e4} field of cfg 4b has
cond(G) up to 1e9, so most of its patches are judged against the sensitivity bound and listed."""
import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from cpu_port import CpuSlod  # noqa: E402
from parity_common import cond_eff, make_tables, margin_safe, selection_sensitivity  # noqa: E402
from oracle.slod_oracle import CoefficientTable, SlodOracle, SlodProblem  # noqa: E402

pkg = importlib.import_module("dealii-slod_b200")
pytestmark = pytest.mark.gpu

CONFIGS = {
    # min_tight: required fraction of sampled (patch, component) pairs within 1e-10 of the oracle; k_tol: bound on the
    # complete K rows of the interior block relative to max|K|.  2-D l = 2 (cfg 2) has interior patches with
    # cond(G) ~ 1e7: the two CPU implementations already differ by 3e-9 there, so neither bound can be 1e-10.
    "cfg2_diffusion2d_256": dict(dim=2, s=1, ref=8, n=2, ell=2, r=8, kind="uniform100", seed=1234, min_tight=0.6, k_tol=1e-8),
    "cfg3_elasticity2d_128": dict(dim=2, s=2, ref=7, n=2, ell=1, r=6, kind="uniform100", seed=2001, min_tight=0.9, k_tol=1e-10),
    "cfg4a_diffusion3d_32_uniform": dict(dim=3, s=1, ref=5, n=2, ell=2, r=6, kind="uniform1e4", seed=3001, min_tight=0.4, k_tol=1e-8),
    "cfg4b_diffusion3d_32_binary": dict(dim=3, s=1, ref=5, n=2, ell=2, r=6, kind="binary1e4", seed=3002, min_tight=0.0, k_tol=None),
}
N_SAMPLE = 128


def morton(c, dim, ref):
    return sum(((c[a] >> b) & 1) << (dim * b + a) for b in range(ref) for a in range(dim))


def sample_patches(c, rng):
    """>= N_SAMPLE patch ids spread over the classes (0 .. dim axes on which the patch is clipped by the domain
    boundary): an equal share per class, capped by the class population (a 2-D mesh has only (2 l)^2 corner-class
    patches), the remainder filled from the interior class; random inside a class."""
    import math
    dim, ref, ell = c["dim"], c["ref"], c["ell"]
    N = 2 ** ref
    classes = list(range(dim + 1))
    pop = [math.comb(dim, k) * (2 * ell) ** k * (N - 2 * ell) ** (dim - k) for k in classes]
    per = -(-N_SAMPLE // len(classes))
    want = [min(per, pop[k]) for k in classes]
    want[0] = min(pop[0], want[0] + N_SAMPLE - sum(want))
    out = []
    for k in classes:
        got = set()
        while len(got) < want[k]:
            clipped = rng.permutation(dim)[:k]
            cc = []
            for a in range(dim):
                if a in clipped:
                    off = int(rng.integers(0, ell))
                    cc.append(off if rng.random() < 0.5 else N - 1 - off)
                else:
                    cc.append(int(rng.integers(ell, N - ell)))
            got.add((k, morton(cc, dim, ref)))
        out += sorted(got)
    return out


def run_gpu(c, tables):
    ctx = pkg.SlodContext(dim=c["dim"], spacedim=c["s"], n_global_refinements=c["ref"], n_subdivisions=c["n"],
                          oversampling=c["ell"], stabilize=True, problem=0 if c["s"] == 1 else 1)
    for f, t in enumerate(tables):
        ctx.set_coefficient(f, c["r"], t)
    ctx.compute_basis()
    ctx.assemble_coarse()
    return ctx


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_full_size_sampled_parity(name):
    c = dict(CONFIGS[name])
    min_tight, k_tol = c.pop("min_tight"), c.pop("k_tol")
    dim, s, ref = c["dim"], c["s"], c["ref"]
    tables = make_tables(dim, s, c["r"], c["kind"], c["seed"])
    ctx = run_gpu(c, tables)
    kw = dict(dim=dim, spacedim=s, n_global_refinements=ref, n_subdivisions=c["n"], oversampling=c["ell"], stabilize=True)
    orc = SlodOracle(SlodProblem(problem="diffusion" if s == 1 else "elasticity",
                                 coefficients=[CoefficientTable(dim, c["r"], t) for t in tables], **kw))
    cpu = CpuSlod(problem=0 if s == 1 else 1, **kw)
    for f, t in enumerate(tables):
        cpu.set_coefficient(f, c["r"], t)

    sample = sample_patches(c, np.random.default_rng(17))
    pids = [p for _, p in sample]
    cpu.compute_patches(pids)
    rec = []
    for k, pid in sample:
        res = orc.compute_patch(pid)
        for d in range(s):
            phi, aphi = ctx.basis(pid, d)
            pc, ac = cpu.basis(pid, d)
            dg = ctx.diagnostics(pid, d)
            assert dg[7] == 0
            err = float(np.linalg.norm(phi - res.basis[d]))
            err_cpu = float(np.linalg.norm(phi - pc))
            err_cc = float(np.linalg.norm(pc - res.basis[d]))     # the two CPU implementations against each other
            aerr = float(np.linalg.norm(aphi - res.basis_premultiplied[d]) / np.linalg.norm(res.basis_premultiplied[d]))
            r = dict(pid=int(pid), comp=d, cls=int(k), err_oracle=err, err_cpu_port=err_cpu, cpu_port_vs_oracle=err_cc,
                     aphi_rel=aerr,
                     steps_gpu=int(dg[1]), steps_oracle=int(res.info["trunc_steps"][d]) if res.info["slod"] else 0,
                     steps_cpu_port=int(cpu.diagnostics(pid, d)[1]),
                     cond_eff=cond_eff(res.info, d) if res.info["slod"] else 1.0,
                     dinf=float(res.info["dinf"][d]) if res.info["slod"] else 0.0)
            if not margin_safe(res.info, d):
                r["verdict"] = "skipped"
            elif err <= 1e-10:
                r["verdict"] = "tight"
            else:
                r["tol"] = 50.0 * selection_sensitivity(res.info, d)
                r["verdict"] = "relaxed" if err <= r["tol"] else "FAIL"
            rec.append(r)
    verdicts = [r["verdict"] for r in rec]
    frac = {v: verdicts.count(v) / len(rec) for v in ("tight", "relaxed", "skipped", "FAIL")}
    # ---- complete K rows: an interior Morton block and the block at the domain corner ----
    N = 2 ** ref
    blocks = [morton([N // 2] * dim, dim, ref), 0]
    rowptr, col, val = ctx.coarse_csr()
    kmax = float(np.abs(val).max())
    krec = []
    phi_all, aphi_all = ctx.all_basis()
    n_sub = c["n"]

    def box(pid):
        info = ctx.patch_info(pid)
        lo = [x * n_sub for x in info["lo"]]
        p = [m * n_sub + 1 for m in info["m"]]
        return lo, p

    def self_row(pid, d):
        """Row (pid, d) of C^T (A C) from the GPU's OWN phi / A phi by plain numpy dot products over the overlap boxes:
        isolates the coarse-matrix kernel from the basis error."""
        row = pid * s + d
        lo_p, p_p = box(pid)
        fp = phi_all[pid, d, : s * int(np.prod(p_p))].reshape(p_p[::-1] + [s])
        out = []
        for q_ in col[rowptr[row]:rowptr[row + 1]]:
            qid, e = int(q_) // s, int(q_) % s
            lo_q, p_q = box(qid)
            fq = aphi_all[qid, e, : s * int(np.prod(p_q))].reshape(p_q[::-1] + [s])
            sl_p, sl_q = [], []
            for a in reversed(range(dim)):          # array axes are z, y, x
                b0 = max(lo_p[a], lo_q[a])
                b1 = min(lo_p[a] + p_p[a], lo_q[a] + p_q[a])
                sl_p.append(slice(b0 - lo_p[a], b1 - lo_p[a]))
                sl_q.append(slice(b0 - lo_q[a], b1 - lo_q[a]))
            out.append(float(np.sum(fp[tuple(sl_p)] * fq[tuple(sl_q)])))
        return np.array(out)

    for base in blocks:
        members = list(range(base, base + 2 ** dim))
        need = set()
        for pid in members:
            for r_ in range(pid * s, pid * s + s):
                need.update(int(q) // s for q in col[rowptr[r_]:rowptr[r_ + 1]])
        cpu.compute_patches(sorted(need))
        for pid in members:
            for d in range(s):
                row = pid * s + d
                cc, vv = cpu.coarse_row(row)
                lo, hi = rowptr[row], rowptr[row + 1]
                assert np.array_equal(cc, col[lo:hi]), (name, row)          # pattern bit-exact
                krec.append(dict(row=int(row), block=int(base), n=int(cc.size),
                                 rel=float(np.abs(vv - val[lo:hi]).max() / kmax),
                                 rel_self=float(np.abs(self_row(pid, d) - val[lo:hi]).max() / kmax)))
    summary = dict(config=name, n_patch_components=len(rec), fractions=frac,
                   phi_err_oracle_max=max(r["err_oracle"] for r in rec),
                   phi_err_oracle_p99=float(np.quantile([r["err_oracle"] for r in rec], 0.99)),
                   phi_err_oracle_median=float(np.median([r["err_oracle"] for r in rec])),
                   phi_err_cpu_port_max=max(r["err_cpu_port"] for r in rec),
                   cpu_port_vs_oracle_max=max(r["cpu_port_vs_oracle"] for r in rec),
                   cpu_port_vs_oracle_frac_le_1e10=float(np.mean([r["cpu_port_vs_oracle"] <= 1e-10 for r in rec])),
                   phi_err_tight_class_max=max([r["err_oracle"] for r in rec if r["verdict"] == "tight"] or [0.0]),
                   aphi_rel_max=max(r["aphi_rel"] for r in rec),
                   max_err_over_eps_cond=max(r["err_oracle"] / (2.220446049250313e-16 * r["cond_eff"]) for r in rec),
                   steps_mismatch_oracle=sum(r["steps_gpu"] != r["steps_oracle"] for r in rec if r["verdict"] != "skipped"),
                   steps_mismatch_cpu_port=sum(r["steps_gpu"] != r["steps_cpu_port"] for r in rec if r["verdict"] != "skipped"),
                   per_class={str(k): dict(n=sum(r["cls"] == k for r in rec),
                                           tight=sum(r["cls"] == k and r["verdict"] == "tight" for r in rec),
                                           max_err=max([r["err_oracle"] for r in rec if r["cls"] == k] or [0.0]))
                              for k in range(dim + 1)},
                   K_rows=len(krec), K_rel_max=max(k_["rel"] for k_ in krec),
                   K_rel_max_from_gpu_basis=max(k_["rel_self"] for k_ in krec),
                   K_rel_max_interior_block=max(k_["rel"] for k_ in krec if k_["block"] == blocks[0]),
                   outliers=sorted([r for r in rec if r["verdict"] != "tight"], key=lambda r: -r["err_oracle"])[:24])
    print(json.dumps({k: v for k, v in summary.items() if k != "outliers"}))
    if os.environ.get("SLOD_WRITE_PARITY"):
        path = os.path.join(ROOT, "gpurun_out", "parity_r02.json")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        allrec = json.load(open(path)) if os.path.exists(path) else {}
        allrec[name] = summary
        json.dump(allrec, open(path, "w"), indent=1)
    assert frac["FAIL"] == 0.0, [r for r in rec if r["verdict"] == "FAIL"][:5]
    assert summary["steps_mismatch_oracle"] == 0 and summary["steps_mismatch_cpu_port"] == 0
    assert frac["tight"] >= min_tight, frac
    assert frac["skipped"] <= 0.10, frac
    assert summary["K_rel_max_from_gpu_basis"] <= 1e-12, summary["K_rel_max_from_gpu_basis"]   # the coarse kernel itself
    if k_tol is not None:
        assert summary["K_rel_max_interior_block"] <= k_tol, summary["K_rel_max_interior_block"]
    # A phi follows phi
    for r in rec:
        if r["verdict"] == "tight":
            assert r["aphi_rel"] <= 1e-9, r
    ctx.close()
    cpu.close()
