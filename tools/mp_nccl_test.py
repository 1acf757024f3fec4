"""Two-process check of slod_comm_init + slod_offline_distributed (launched with torchrun, small problems)."""
import importlib, os, sys, time
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("dealii-slod_b200")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
cases = [(2, 4, 5), (3, 3, 4), (3, 4, 5), (3, 5, 6)] if len(sys.argv) < 2 else [(3, int(sys.argv[1]), int(sys.argv[1]) + 1)]
for dim, ref, r in cases:
    tab = 1.0 + 99.0 * np.random.default_rng(5).random((2 ** r) ** dim)
    ctx = pkg.SlodContext(dim=dim, spacedim=1, n_global_refinements=ref, n_subdivisions=2, oversampling=2, stabilize=True, device=local)
    ctx.set_coefficient(0, r, tab)
    uid = torch.tensor(list(ctx.comm_unique_id() if rank == 0 else bytes(128)), dtype=torch.uint8, device=dev)
    dist.broadcast(uid, 0)
    print(rank, "uid ok", flush=True)
    ctx.comm_init(rank, world, uid.cpu().numpy().tobytes())
    print(rank, "comm ok", flush=True)
    n, s, stride, ellw = ctx.n_patches, 1, ctx.basis_stride, ctx.ell_width
    phi = torch.zeros((n, s, stride), dtype=torch.float64, device=dev); aphi = torch.zeros_like(phi)
    K = torch.zeros((n * s, ellw), dtype=torch.float64, device=dev)
    ctx.offline_distributed(phi.data_ptr(), aphi.data_ptr(), K.data_ptr(), gather_phi=True, gather_K=True, stream=torch.cuda.current_stream().cuda_stream)
    print(rank, "enqueued", flush=True)
    ctx.synchronize()
    torch.cuda.synchronize()
    print(rank, dim, ref, "done", float(K.abs().sum()), float(phi.abs().sum()), flush=True)
    for it in range(3):
        ctx.offline_distributed(phi.data_ptr(), aphi.data_ptr(), K.data_ptr(), gather_phi=False, gather_K=True, stream=torch.cuda.current_stream().cuda_stream)
        ctx.synchronize()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    print(rank, dim, ref, "repeat ok", ctx.timings()[:5], flush=True)
    ctx.close()
dist.barrier()
dist.destroy_process_group()
