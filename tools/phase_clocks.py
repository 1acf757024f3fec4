"""Run 296 interior-heavy patches of the 16^3 mesh with the instrumented library (SLOD_LIB=tools/libslod_prof.so, built
with -DSLOD_PHASE_CLOCKS) so that CTA 0 prints its per-phase clock counts; also prints the achieved occupancy."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SLOD_LIB", os.path.join(ROOT, "tools", "libslod_prof.so"))
import bench
pkg = importlib.import_module("dealii-slod_b200")
w = bench.WORKLOADS["diffusion3d_16c_l2_n2"]
ctx = pkg.SlodContext(dim=3, spacedim=1, n_global_refinements=4, n_subdivisions=2, oversampling=2, stabilize=True)
ctx.set_coefficient(0, w["r"], bench.make_tables(w)[0])
import torch
n, stride = ctx.n_patches, ctx.basis_stride
phi = torch.zeros((n, 1, stride), dtype=torch.float64, device="cuda"); aphi = torch.zeros_like(phi)
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 1792
ctx.compute_basis_device(lo, lo + 296, phi.data_ptr(), aphi.data_ptr())
ctx.synchronize()
print("kernel ms", ctx.timings()[:6])
