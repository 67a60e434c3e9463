#!/usr/bin/env python3
"""Multi-GPU check of the single-frequency distributed path (mfb_dist_*), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py [m_small] [m_large]

Every rank builds the same S-cube model; rank 0 creates the NCCL id, torch.distributed (gloo) carries it to the others.
Checks: (1) distributed LU of a random matrix against numpy, (2) distributed frequency against the single-GPU path on a
small mesh, (3) timing of one frequency at the large mesh, single-GPU vs distributed.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from multifebe_b200 import capi
    from multifebe_b200.host import Model, Material, cube_mesh, cube_bcs, shape
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    m_small = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    m_large = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = capi.Context(local)
    mat = Material(1.0, 1.0, 0.25, 0.03)
    out = {"world": world}

    def join(pr, nb=0):
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.frombuffer(bytearray(capi.dist_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(uid, 0)
        pr.dist_init(rank, world, bytes(uid.numpy().tobytes()), nb)

    # (1) + (2) small mesh
    md = Model(cube_mesh(m_small, shape.TRI3), cube_bcs())
    n = md.n_dof
    pr = capi.Problem(ctx, md)
    x1 = pr.solve_frequency(3.0, mat)
    join(pr, 64)
    rng = np.random.default_rng(7)
    A = np.asfortranarray(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x = pr.dist_zsolve(A, b)
    xo = np.linalg.solve(A, b)
    out["lu_random_relerr"] = float(np.abs(x - xo).max() / np.abs(xo).max())
    x2 = pr.dist_solve_frequency(3.0, mat)
    out["frequency_small_relerr_vs_single_gpu"] = float(np.abs(x2 - x1).max() / np.abs(x1).max())
    out["n_small"] = n
    out["row_bounds_small"] = [int(v) for v in pr.dist_info()["row_bounds"]]
    pr.close()
    ok = out["lu_random_relerr"] < 1e-9 and out["frequency_small_relerr_vs_single_gpu"] < 1e-11

    # (3) large mesh timing
    if m_large > 0:
        md = Model(cube_mesh(m_large, shape.TRI3), cube_bcs())
        pr = capi.Problem(ctx, md)
        omega = 2.0
        for _ in range(2):
            x1 = pr.solve_frequency(omega, mat)
        s1 = pr.stats()
        dist.barrier()
        join(pr, 256)
        ts = []
        for it in range(3):
            dist.barrier()
            t0 = time.time()
            x2 = pr.dist_solve_frequency(omega, mat)
            ts.append(time.time() - t0)
        s2 = pr.stats()
        err = float(np.abs(x2 - x1).max() / np.abs(x1).max())
        ok = ok and err < 1e-9
        out.update({"n_large": md.n_dof, "frequency_large_relerr_vs_single_gpu": err,
                    "single_gpu_ms": {"assemble": s1["MS_ASSEMBLE"], "lu": s1["MS_LU"], "solve": s1["MS_SOLVE"]},
                    "dist_ms": {"assemble": s2["MS_ASSEMBLE"], "redistribute": s2["MS_REDIST"], "lu": s2["MS_DIST_LU"],
                                "solve": s2["MS_DIST_SOLVE"], "total": s2["MS_DIST_TOTAL"], "wall_best": 1e3 * min(ts)},
                    "dist_lu_tflops_per_rank": s2["GEMM_FLOPS"] / (s2["MS_DIST_LU"] * 1e-3) / 1e12})
        pr.close()
    out["ok"] = bool(ok)
    oks = [None] * world
    dist.all_gather_object(oks, bool(ok))
    out["ok_all_ranks"] = bool(all(oks))
    if rank == 0:
        print(json.dumps(out))
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if all(oks) else 1)


if __name__ == "__main__":
    main()
