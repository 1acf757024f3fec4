"""Run one small 3-D problem with the instrumented library (SLOD_LIB=tools/libslod_prof.so) so that CTA 0 prints its
per-phase clock counts."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module("dealii-slod_b200")
w = bench.WORKLOADS["diffusion3d_16c_l2_n2"]
ctx = pkg.SlodContext(dim=3, spacedim=1, n_global_refinements=4, n_subdivisions=2, oversampling=2, stabilize=True)
ctx.set_coefficient(0, w["r"], bench.make_tables(w)[0])
import torch
n, stride = ctx.n_patches, ctx.basis_stride
phi = torch.zeros((n, 1, stride), dtype=torch.float64, device="cuda"); aphi = torch.zeros_like(phi)
ctx.compute_basis_device(0, 148, phi.data_ptr(), aphi.data_ptr())
torch.cuda.synchronize()
