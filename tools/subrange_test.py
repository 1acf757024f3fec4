import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("dealii-slod_b200")
w = bench.WORKLOADS["diffusion3d_32c_l2_n2"]
ctx = pkg.SlodContext(dim=3, spacedim=1, n_global_refinements=5, n_subdivisions=2, oversampling=2, stabilize=True)
ctx.set_coefficient(0, w["r"], bench.make_tables(w)[0])
n, stride, ellw = ctx.n_patches, ctx.basis_stride, ctx.ell_width
dev = torch.device("cuda")
phi = torch.zeros((n, 1, stride), dtype=torch.float64, device=dev); aphi = torch.zeros_like(phi)
K = torch.zeros((n, ellw), dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
ctx.offline_distributed(phi.data_ptr(), aphi.data_ptr(), K.data_ptr(), False, True, st); ctx.synchronize(); torch.cuda.synchronize()
print("full ok", flush=True)
ctx.compute_basis_device(0, 16384, phi.data_ptr(), aphi.data_ptr(), st); ctx.synchronize(); print("half ok", flush=True)
phi2 = torch.zeros_like(phi); aphi2 = torch.zeros_like(phi)
q0, q1 = 16384 - 64, 16384 + 64
ctx.compute_basis_device(q0, q1, phi2.data_ptr(), aphi2.data_ptr(), st); print("enq", flush=True)
ctx.synchronize(); print("sub ok", bool(torch.equal(phi2[q0:q1], phi[q0:q1])), flush=True)
