"""Workload for ncu captures: the 3-D l = 2 problem on a 2^ref mesh, basis + coarse matrix, `reps` times.
Usage: python tools/prof_run.py [ref=4] [reps=2]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("dealii-slod_b200")
ref = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
r = min(ref + 1, 6)
tab = 1.0 + (1e4 - 1.0) * np.random.default_rng(3001).random((2 ** r) ** 3)
ctx = pkg.SlodContext(dim=3, spacedim=1, n_global_refinements=ref, n_subdivisions=2, oversampling=2, stabilize=True)
ctx.set_coefficient(0, r, tab)
for _ in range(reps):
    ctx.compute_basis()
    ctx.assemble_coarse()
print("kernel ms", ctx.timings()[:6])
