"""Quick GPU check of one configuration against the C++ CPU port: stage parity (X, Minv, G) on a few patches and
kernel times.  Usage: python tools/quick_gpu.py [ref] [n_check]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from cpu_port import CpuSlod  # noqa: E402

pkg = importlib.import_module("dealii-slod_b200")
ref = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ncheck = int(sys.argv[2]) if len(sys.argv) > 2 else 6
r = min(ref + 1, 6)
tab = 1.0 + (1e4 - 1.0) * np.random.default_rng(3001).random((2 ** r) ** 3)
kw = dict(dim=3, spacedim=1, n_global_refinements=ref, n_subdivisions=2, oversampling=2, stabilize=True)
ctx = pkg.SlodContext(**kw)
ctx.set_coefficient(0, r, tab)
cpu = CpuSlod(**kw)
cpu.set_coefficient(0, r, tab)
n = ctx.n_patches
pids = sorted(set(int(x) for x in np.linspace(0, n - 1, ncheck)) | {n // 2 + 3})
for pid in pids:
    X, Minv, G = ctx.debug_stages(pid)
    Xc, Mc, Gc = cpu.debug_stages(pid)
    print(f"patch {pid}: X {np.abs(X - Xc).max() / np.abs(Xc).max():.2e}  Minv {np.abs(Minv - Mc).max() / np.abs(Mc).max():.2e}"
          f"  G {np.abs(G - Gc).max() / np.abs(Gc).max():.2e}")
t0 = time.time()
ctx.compute_basis()
ctx.assemble_coarse()
print("first run %.3f s" % (time.time() - t0))
for _ in range(3):
    ctx.compute_basis()
    ctx.assemble_coarse()
    tm = ctx.timings()
print("kernel ms: solve %.2f (factor %.2f) dense %.2f select %.2f finish %.2f coarse %.2f" % (tm[0], tm[5], tm[1], tm[2], tm[3], tm[4]))
cpu.compute_patches(pids)
for pid in pids:
    phi, _ = ctx.basis(pid)
    pc, _ = cpu.basis(pid)
    print(f"patch {pid}: phi err {np.linalg.norm(phi - pc):.2e} steps {int(ctx.diagnostics(pid)[1])} / {int(cpu.diagnostics(pid)[1])}")
