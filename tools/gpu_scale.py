"""Scale probe (development aid): set-up, assembly and LU timings of S-cube meshes of growing size."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multifebe_b200 import capi
from multifebe_b200.host import *

os.environ["MFB_LU_TIMING"] = "1"
ctx = capi.Context(0)
cases = [(shape.TRI3, 10), (shape.TRI3, 20), (shape.QUAD9, 10), (shape.TRI3, 40), (shape.QUAD9, 20)]
if len(sys.argv) > 1:
    cases = [(int(a.split(":")[0]), int(a.split(":")[1])) for a in sys.argv[1:]]
for et, m in cases:
    t0 = time.time(); mesh = cube_mesh(m, et); md = Model(mesh, cube_bcs()); t1 = time.time()
    mat = Material(1, 1, 0.25, 0.03)
    pr = capi.Problem(ctx, md); t2 = time.time()
    om = 2.0
    for rep in range(2):
        t3 = time.time(); x = pr.solve_frequency(om, mat); t4 = time.time()
    s = pr.stats()
    n = md.n_dof
    print("etype", et, "m", m, "ndof", n, "model s", round(t1 - t0, 2), "setup s", round(t2 - t1, 2), "solve_frequency wall s", round(t4 - t3, 3), flush=True)
    print("   ", {k: (round(v, 3) if v < 1e6 else v) for k, v in s.items() if v}, flush=True)
    if s["MS_REGULAR"] > 0:
        print("    regular TFLOP/s", s["FLOPS_REGULAR"] / s["MS_REGULAR"] / 1e9, " LU TFLOP/s", 8 / 3 * n ** 3 / s["MS_LU"] / 1e9,
              " entries/s", n * n / s["MS_ASSEMBLE"] * 1e3, flush=True)
    u, t = md.nodal_solution(x)
    ua = column_analytic_u(md.node_x[:, 0], om, mat)
    print("    u1 vs analytic", np.abs(u[:, 0] - ua).max() / np.abs(ua).max(), flush=True)
    pr.close()
print("done")
