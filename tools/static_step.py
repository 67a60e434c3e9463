"""Static elasticity at full size: per-phase device times.  usage: static_step.py [etype m reps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multifebe_b200 import capi
from multifebe_b200.host import *
et = int(sys.argv[1]) if len(sys.argv) > 1 else shape.QUAD9
m = int(sys.argv[2]) if len(sys.argv) > 2 else 11
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = capi.Context(0); md = Model(cube_mesh(m, et), cube_bcs()); pr = capi.Problem(ctx, md)
mat = Material(1.0, 1.0, 0.25, 0.0)
lam2mu = 2.0 * mat.mu_r * mat.nu_r / (1.0 - 2.0 * mat.nu_r) + 2.0 * mat.mu_r
for r in range(reps):
    t0 = time.time(); x = pr.solve_static(mat); wall = (time.time() - t0) * 1e3
    s = pr.stats()
    u, t = md.nodal_solution(x)
    err = np.abs(u[:, 0].real - md.node_x[:, 0] / lam2mu).max() * lam2mu
    print("static n_dof=%d" % md.n_dof, r, {k: round(s[k], 2) for k in ("MS_ZERO", "MS_REGULAR", "MS_ADAPTIVE", "MS_SINGULAR", "MS_ASSEMBLE", "MS_LU", "MS_GEMM", "MS_PANEL", "MS_SOLVE")},
          "wall %.1f ms" % wall, "K1 TFLOP/s(harmonic count) %.2f" % (s["FLOPS_REGULAR"] / s["MS_REGULAR"] / 1e9), "LU TFLOP/s %.2f" % (2.0 / 3.0 * md.n_dof ** 3 / s["MS_LU"] / 1e9),
          "exact-solution err %.2e" % err, flush=True)
