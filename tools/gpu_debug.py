"""First-contact GPU script (development aid): GEMM, assembly parity vs the oracle, LU, peaks."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multifebe_b200 import capi
from multifebe_b200.host import *
from oracle import oracle as orc

ctx = capi.Context(0)
print("peaks", ctx.measure_peaks(), flush=True)
rng = np.random.default_rng(1)
for (m, n, k) in [(128, 64, 16), (200, 130, 36), (512, 512, 128), (37, 5, 2)]:
    A = rng.standard_normal((m, k)) + 1j * rng.standard_normal((m, k))
    B = rng.standard_normal((k, n)) + 1j * rng.standard_normal((k, n))
    Cm = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
    R, ms = ctx.zgemm_minus(Cm, A, B)
    ref = Cm - A @ B
    print("gemm", m, n, k, "err", np.abs(R - ref).max() / np.abs(ref).max(), "ms", ms, flush=True)

for et, msize in [(shape.TRI3, 3), (shape.QUAD4, 2), (shape.TRI6, 2), (shape.QUAD8, 2), (shape.QUAD9, 2)]:
    mesh = cube_mesh(msize, et)
    md = Model(mesh, cube_bcs())
    mat = Material(1, 1, 0.25, 0.03)
    o = orc.Oracle(md)
    t0 = time.time(); Ao, bo, st = o.assemble(3.0, mat); t1 = time.time()
    pr = capi.Problem(ctx, md)
    Ag, bg = pr.build_lse_mechanics_bem_harela(3.0, mat)
    s = pr.stats()
    print("etype", et, "ndof", md.n_dof, "oracle s", round(t1 - t0, 2), "stats", {k: v for k, v in s.items() if v}, flush=True)
    print("  oracle stats", st)
    print("  A err", np.abs(Ag - Ao).max() / np.abs(Ao).max(), "b err", np.abs(bg - bo).max() / max(np.abs(bo).max(), 1e-300), flush=True)
    # where is the worst entry
    i, j = np.unravel_index(np.argmax(np.abs(Ag - Ao)), Ao.shape)
    print("  worst", i, j, Ag[i, j], Ao[i, j])
    xo, _, _ = orc.lu_solve(Ao, bo)
    xg = pr.solve_lse_c(Ag.copy(order="F"), bg)
    print("  x err (gpu LU on gpu A)", np.abs(xg - xo).max() / np.abs(xo).max(), flush=True)
    xf = pr.solve_frequency(3.0, mat)
    print("  x err (solve_frequency)", np.abs(xf - xo).max() / np.abs(xo).max(), flush=True)
    pr.close()

# LU on a random matrix larger than one block
for n in (300, 1000):
    pass
print("done")
