"""mfb_harela3d_sweep over real NCCL: torchrun --nproc-per-node N tools/sweep_check.py   (every rank gets all solutions; compared with per-frequency solves)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from multifebe_b200 import capi
from multifebe_b200.host import Model, Material, cube_mesh, cube_bcs, shape
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
md = Model(cube_mesh(8, shape.TRI3), cube_bcs()); mat = Material(1, 1, 0.25, 0.03)
ctx = capi.Context(local); pr = capi.Problem(ctx, md)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.frombuffer(bytearray(capi.dist_unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(uid, 0)
oms = np.linspace(0.5, 6.0, 11)
X, info = pr.sweep(oms, mat, rank, world, uid.cpu().numpy().tobytes())
ref = np.array([pr.solve_frequency(float(o), mat) for o in oms])
err = float(np.abs(X - ref).max() / np.abs(ref).max())
print("rank %d of %d: mfb_harela3d_sweep over NCCL, %d frequencies, n_dof %d: max rel diff vs per-frequency solves %.2e, info %s" % (rank, world, len(oms), md.n_dof, err, info.tolist()), flush=True)
assert err < 1e-12 and (info == 0).all()
pr.close(); ctx.close(); dist.barrier(); dist.destroy_process_group()
