"""One poroelastic assembly of a quad9 cube (for ncu): python tools/por_step.py [m]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from multifebe_b200 import capi
from multifebe_b200.host import Poro, PoroModel, cube_mesh, shape
sys.path.insert(0, "tests")
from test_oracle_poroelastic import column_bcs

m = int(sys.argv[1]) if len(sys.argv) > 1 else 10
po = Poro(rhof=1000.0, rhos=2600.0, lam=1.0e8, mu=1.0e8, xi=0.02, phi=0.3, rhoa=150.0, R=3.0e8, Q=6.0e8, b=1.0e6)
md = PoroModel(cube_mesh(m, shape.QUAD9, L=10.0), column_bcs())
ctx = capi.Context(0)
pr = capi.Problem(ctx, md)
for _ in range(2):
    pr.build_lse_mechanics_bem_harpor(2 * np.pi * 50.0, po, want_host=False)
st = pr.stats()
print("n_dof", md.n_dof, "ms_regular", st["MS_REGULAR"], "pairs", st["PAIRS_REGULAR"], "points", st["POINTS_REGULAR"])
pr.close(); ctx.close()
