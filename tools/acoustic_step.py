"""Acoustic (inviscid fluid) region at size: per-phase device times of one frequency.  usage: acoustic_step.py [etype m reps f_hz]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multifebe_b200 import capi
from multifebe_b200.host import *
et = int(sys.argv[1]) if len(sys.argv) > 1 else shape.QUAD9
m = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
f_hz = float(sys.argv[4]) if len(sys.argv) > 4 else 40.0
L = 3.0
air = Fluid(rho=1.25, c=343.0)
t0 = time.time(); md = FluidModel(cube_mesh(m, et, L=L), room_bcs(1.0)); t_model = time.time() - t0
ctx = capi.Context(0)
t0 = time.time(); pr = capi.Problem(ctx, md); t_setup = time.time() - t0
print("acoustic etype=%d m=%d n_dof=%d n_elem=%d n_colloc=%d  host model %.2f s, set-up %.2f s" % (et, m, md.n_dof, md.n_elem, md.n_colloc, t_model, t_setup), flush=True)
omega = 2 * np.pi * f_hz
for r in range(reps):
    t0 = time.time(); x = pr.solve_frequency_fluid(omega, air); wall = (time.time() - t0) * 1e3
    s = pr.stats()
    p, un = md.nodal_solution(x)
    p_ex, ux_ex = room_analytic(md.node_x[:, 0], omega, air, L=L, P=1.0)
    err = np.abs(p - p_ex).max() / np.abs(p_ex).max()
    print("rep", r, {k: round(s[k], 3) for k in ("MS_ZERO", "MS_REGULAR", "MS_ADAPTIVE", "MS_SINGULAR", "MS_FREETERM", "MS_ASSEMBLE", "MS_LU", "MS_GEMM", "MS_PANEL", "MS_SOLVE")},
          "wall %.1f ms" % wall, "P1 GFLOP/s(algorithmic) %.1f" % (s["FLOPS_REGULAR"] / s["MS_REGULAR"] / 1e6),
          "pairs/s %.3g" % (s["PAIRS_REGULAR"] / s["MS_REGULAR"] * 1e3), "entries/s %.3g" % (md.n_dof ** 2 / s["MS_ASSEMBLE"] * 1e3),
          "LU TFLOP/s %.2f" % (8.0 / 3.0 * md.n_dof ** 3 / s["MS_LU"] / 1e9), "analytic err %.2e" % err, flush=True)
s = pr.stats()
print("plan", {k: int(s[k]) for k in ("PAIRS_REGULAR", "POINTS_REGULAR", "PAIRS_ADAPTIVE", "LEAVES", "POINTS_ADAPTIVE", "PAIRS_SINGULAR", "POINTS_SINGULAR")}, "launches", int(s["LAUNCHES"]))
# backward error of the solution against a fresh assembly
pr.build_lse_mechanics_bem_harpot(omega, air, want_host=False)
print("berr, rel_resid", pr.residual(x))
