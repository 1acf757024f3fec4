"""C++ host mirror of the reference driver (dealii-slod_b200/host): builds against the C ABI, speaks the reference's
.prm dialect, fails loudly without a GPU (CPU tests) and reproduces the oracle's coarse matrix (GPU test)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "dealii-slod_b200", "host")

PRM = """subsection Problem
  set Output directory = .
  set Output name = t
  set Oversampling = {ell}
  set Number of subdivisions = 2
  set Number of global refinements = {ref}
  set Stabilize phi_LOD candidates = true   # SLOD
  set Compare with fine global solution = true
  subsection Coefficients
    set Constant problem coefficients = false
    set Refinement for random coefficients = {r}
    set Random seed = {seed}
  end
  subsection Right hand side
    set Function expression = {rhs}
  end
  subsection Solver
    subsection Coarse solver control
      set Max steps = 2000
      set Tolerance = 0
      set Reduction = 1e-12
    end
    subsection Fine solver control
      set Max steps = 100000
      set Tolerance = 0
      set Reduction = 1e-12
    end
  end
end
"""


@pytest.fixture(scope="module")
def apps():
    res = subprocess.run(["make", "-C", HOST], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    return {n: os.path.join(HOST, n) for n in ("main_Diffusion", "main_Elasticity", "main_Diffusion3D")}


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_template_prm_and_no_cpu_fallback(apps, tmp_path):
    if _has_gpu():
        pytest.skip("a GPU is present")
    res = subprocess.run([apps["main_Diffusion"]], cwd=tmp_path, capture_output=True, text=True)
    # like the reference: exception -> message -> exit code 1 (app/main_Diffusion.cc:23-47)
    assert res.returncode == 1
    assert "Exception on processing" in res.stderr and "no CPU fallback" in res.stderr
    assert "Running LOD Diffusion problem in 2D" in res.stdout
    prm = (tmp_path / "parameters.prm").read_text()   # ParameterAcceptor writes a template when the file is missing
    for key in ("Output directory", "Output name", "Oversampling", "Number of subdivisions",
                "Number of global refinements", "Compare with fine global solution", "Stabilize phi_LOD candidates",
                "Constant problem coefficients", "Function expression", "Max steps", "Tolerance", "Reduction"):
        assert f"set {key} = " in prm                   # include/LOD.h:133-143
    assert "subsection Problem" in prm and "subsection Coefficients" in prm
    assert "subsection Right hand side" in prm and "subsection Coarse solver control" in prm   # include/LOD.h:123,127
    used = (tmp_path / "used_parameters_2.prm").read_text()   # source/LOD.cc:60-62
    assert "set Oversampling = 1" in used


def test_prm_errors(apps, tmp_path):
    (tmp_path / "bad.prm").write_text("subsection Problem\n  set No such key = 1\nend\n")
    res = subprocess.run([apps["main_Elasticity"], "bad.prm"], cwd=tmp_path, capture_output=True, text=True)
    assert res.returncode == 1 and "no such parameter" in res.stderr
    (tmp_path / "bad2.prm").write_text("subsection Problem\n  set Oversampling = two\nend\n")
    res = subprocess.run([apps["main_Diffusion3D"], "bad2.prm"], cwd=tmp_path, capture_output=True, text=True)
    assert res.returncode == 1 and "not an integer" in res.stderr
    (tmp_path / "bad4.prm").write_text("subsection Problem\n  subsection Right hand side\n    set Function expression = sin(x)\n  end\nend\n")
    res = subprocess.run([apps["main_Diffusion"], "bad4.prm"], cwd=tmp_path, capture_output=True, text=True)
    assert res.returncode == 1 and "only constant expressions" in res.stderr
    (tmp_path / "bad5.prm").write_text("subsection Problem\n  subsection Right hand side\n    set Function expression = 1\n  end\nend\n")
    res = subprocess.run([apps["main_Elasticity"], "bad5.prm"], cwd=tmp_path, capture_output=True, text=True)
    assert res.returncode == 1 and "needs 2 components" in res.stderr
    (tmp_path / "bad3.prm").write_text("subsection Problem\n  set Oversampling = 1\n")
    res = subprocess.run([apps["main_Diffusion"], "bad3.prm"], cwd=tmp_path, capture_output=True, text=True)
    assert res.returncode == 1 and "unterminated subsection" in res.stderr


def _read_matrix(path):
    raw = open(path, "rb").read()
    n_rows, nnz = np.frombuffer(raw[:16], dtype=np.int64)
    off = 16
    rowptr = np.frombuffer(raw, dtype=np.int64, count=n_rows + 1, offset=off)
    off += 8 * (n_rows + 1)
    col = np.frombuffer(raw, dtype=np.int64, count=nnz, offset=off)
    off += 8 * nnz
    val = np.frombuffer(raw, dtype=np.float64, count=nnz, offset=off)
    return rowptr, col, val


@pytest.mark.gpu
@pytest.mark.parametrize("app,dim,s,ref,ell,r", [("main_Diffusion", 2, 1, 3, 1, 4), ("main_Elasticity", 2, 2, 3, 1, 4),
                                                 ("main_Diffusion3D", 3, 1, 2, 1, 3)])
def test_apps_match_oracle(apps, tmp_path, app, dim, s, ref, ell, r):
    from oracle.slod_oracle import CoefficientTable, GlibcRand, SlodOracle, SlodProblem, reference_random_table
    seed = 5
    rhs = [1.0] if s == 1 else [1.0, -0.5]
    (tmp_path / "p.prm").write_text(PRM.format(ell=ell, ref=ref, r=r, seed=seed, rhs="; ".join(str(v) for v in rhs)))
    res = subprocess.run([apps[app], "p.prm"], cwd=tmp_path, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert f"Number of patches = {(2 ** ref) ** dim}" in res.stdout
    rowptr, col, val = _read_matrix(tmp_path / "t_coarse_matrix.bin")
    rng = GlibcRand(seed)                       # srand(seed); the host draws Lambda then Mu like the reference
    tabs = [reference_random_table(dim, 1, 100, r, rng) for _ in range(s)]
    orc = SlodOracle(SlodProblem(dim=dim, spacedim=s, n_global_refinements=ref, n_subdivisions=2, oversampling=ell,
                                 stabilize=True, problem="diffusion" if s == 1 else "elasticity",
                                 coefficients=[CoefficientTable(dim, r, t) for t in tabs]))
    orc.compute_basis()
    K, _, _ = orc.assemble_global_matrix()
    assert np.array_equal(rowptr, K.indptr) and np.array_equal(col, K.indices)
    assert np.abs(val - K.data).max() <= 1e-8 * np.abs(K.data).max()
    # online stages (LOD::solve, source/LOD.cc:975-1001; prolongation :1251) as printed by the host
    import re
    _, C, _ = orc.assemble_global_matrix()
    f = orc.fem_rhs(lambda p: np.tile(rhs, (len(p), 1)))
    out = res.stdout
    assert abs(float(re.search(r"fem rhs l2 norm = (\S+)", out).group(1)) - np.linalg.norm(f)) <= 1e-5 * np.linalg.norm(f)
    b = C.T @ f
    assert abs(float(re.search(r"\n\s+rhs l2 norm = (\S+)", out).group(1)) - np.linalg.norm(b)) <= 1e-5 * np.linalg.norm(b)
    assert f"size of u {s * (2 ** ref) ** dim}" in out
    u, _ = orc.solve_coarse(K, b, direct=True)
    nrm = float(re.search(r"lod solution l2 norm = (\S+)", out).group(1))
    assert abs(nrm - np.linalg.norm(C @ u)) <= 1e-7 * np.linalg.norm(C @ u)
    # compare_lod_with_fem (source/LOD.cc:1240-1260): the SLOD vs FEM(h) table
    assert f"size of fem u {s * (2 ** ref * 2 + 1) ** dim}" in out
    u_fem, A = orc.fem_solve(f)
    e = u_fem - C @ u
    row = out.split("SLOD vs reference FEM(h)")[1].split("\n")[2].split()
    l2, linf, h1, en = float(row[2]), float(row[3]), float(row[4]), float(row[5])
    # the table's L2 / Linfty / H1 columns follow ParsedConvergenceTable::difference (the reference's quadrature)
    o_l2, o_inf, o_h1 = orc.reference_error_norms(e)
    assert abs(l2 - o_l2) <= 1e-4 * l2 and abs(linf - o_inf) <= 1e-4 * linf and abs(h1 - o_h1) <= 1e-4 * h1
    assert abs(en - np.sqrt(e @ (A @ e))) <= 1e-4 * en
    # the three .vtu files of the reference (include/Diffusion.h:102-105, source/LOD.cc:284-286, :1370-1372)
    nsub = 2 ** ref * 2
    for name, npts, ncells, fields in (("t_coefficients.vtu", (nsub + 1) ** dim, nsub ** dim, ["alpha" if s == 1 else "lambda"]),
                                       ("t_coarse.vtu", (2 ** ref + 1) ** dim, (2 ** ref) ** dim, ["LOD_solution", "exact_solution"]),
                                       ("t_fine.vtu", (nsub + 1) ** dim, nsub ** dim, ["fem_reference", "exact_rhs", "lod_solution"])):
        txt = (tmp_path / name).read_text()
        assert f'NumberOfPoints="{npts}" NumberOfCells="{ncells}"' in txt
        for fld in fields:
            assert f'Name="{fld}"' in txt
    # the coarse VTU carries the coarse solution in lexicographic cell order
    import xml.etree.ElementTree as ET
    arr = [a for a in ET.parse(tmp_path / "t_coarse.vtu").getroot().iter("DataArray") if a.get("Name") == "LOD_solution"][0]
    vals = np.array(arr.text.split(), dtype=float).reshape((2 ** ref) ** dim, -1)[:, :s]
    N = 2 ** ref
    lex = np.arange(N ** dim)
    code = np.zeros_like(lex)
    for a in range(dim):
        ia = (lex // N ** a) % N
        for b in range(ref):
            code |= ((ia >> b) & 1) << (dim * b + a)
    assert np.abs(vals.ravel() - u.reshape(-1, s)[code].ravel()).max() <= 1e-6 * np.abs(u).max()
