"""An anchor for the oracle's SLOD branch that does not depend on any reference run (the branch is executed by no
reference test: "parity unpinned").  Super-localization is a mathematical property: for a right-hand side that is
constant on the coarse cells the SLOD solution differs from the fine-scale FEM solution of the same problem
(error_LOD_FEMh of the reference, source/LOD.cc:1252) only by the localization error, which decays
super-exponentially in the oversampling parameter.  A wrong boundary-flux matrix, Gram matrix, pseudo-inverse or
truncation rule (source/LOD.cc:596-757) destroys that decay.  The unstabilised branch (source/LOD.cc:563-595), which the
reference's golden test pins, does not have it."""
import numpy as np
import pytest

from oracle.slod_oracle import CoefficientTable, SlodOracle, SlodProblem


def _errors(dim, s, ref, ell, stabilize, seed=11):
    r = ref + 1
    rs = np.random.default_rng(seed)
    tabs = [1.0 + 99.0 * rs.random((2 ** r) ** dim) for _ in range(1 if s == 1 else 2)]
    prob = SlodProblem(dim=dim, spacedim=s, n_global_refinements=ref, n_subdivisions=2, oversampling=ell,
                       stabilize=stabilize, problem="diffusion" if s == 1 else "elasticity",
                       coefficients=[CoefficientTable(dim, r, t) for t in tabs])
    o = SlodOracle(prob)
    o.compute_basis()
    K, C, _ = o.assemble_global_matrix()
    rhs = [1.0] if s == 1 else [1.0, -0.5]
    F = o.fem_rhs(lambda p: np.tile(rhs, (len(p), 1)))
    u_fem, A = o.fem_solve(F)
    u, _ = o.solve_coarse(K, C.T @ F, direct=True)
    e = C @ u - u_fem
    return np.sqrt(e @ (A @ e)) / np.sqrt(u_fem @ (A @ u_fem)), np.linalg.norm(e) / np.linalg.norm(u_fem)


@pytest.mark.parametrize("dim,s,ref,bounds", [
    (2, 1, 3, [(0.2, 0.05), (2e-3, 4e-4), (5e-6, 1e-6)]),      # measured: 1.2e-1, 8.9e-4, 1.2e-6 (energy)
    (2, 2, 3, [(0.3, 0.1), (1e-2, 2e-3), (3e-5, 5e-6)]),       # measured: 1.5e-1, 4.3e-3, 8.9e-6
])
def test_slod_error_decays_superexponentially(dim, s, ref, bounds):
    prev = None
    for ell, (b_energy, b_l2) in zip((1, 2, 3), bounds):
        en, l2 = _errors(dim, s, ref, ell, True)
        assert en < b_energy and l2 < b_l2, (ell, en, l2)
        if prev is not None:
            assert en < prev / 20.0        # each layer gains more than a factor 20 (measured: 130 and 760)
        prev = en


def test_unstabilised_branch_has_no_such_decay():
    en1, _ = _errors(2, 1, 3, 1, False)
    en2, _ = _errors(2, 1, 3, 2, False)
    assert en1 > 0.5 and en2 > 0.2      # measured 0.90, 0.46: the branch the golden file pins is a poor approximation
