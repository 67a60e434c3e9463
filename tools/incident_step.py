"""K1 with an incident field set (every element runs as the general class) on the headline mesh: python tools/incident_step.py [m]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multifebe_b200 import capi
from multifebe_b200.host import *
m = int(sys.argv[1]) if len(sys.argv) > 1 else 40
mat = Material(1, 1, 0.25, 0.03)
ctx = capi.Context(0); md = Model(cube_mesh(m, shape.TRI3), cube_bcs(), reversed_parts=(1, 2, 3, 4, 5, 6)); pr = capi.Problem(ctx, md)
pr.build_lse_mechanics_bem_harela(9.0, mat, want_host=False); pr.build_lse_mechanics_bem_harela(9.0, mat, want_host=False)
print("plain   ", {k: round(v, 2) for k, v in pr.stats().items() if k in ("MS_REGULAR", "MS_ADAPTIVE", "MS_SINGULAR", "MS_ASSEMBLE")})
u, t = element_incident(md, plane_wave("P", [0.2, 0.1, 1.0], mat, 9.0))
pr.set_incident(u, t)
pr.build_lse_mechanics_bem_harela(9.0, mat, want_host=False); pr.build_lse_mechanics_bem_harela(9.0, mat, want_host=False)
print("incident", {k: round(v, 2) for k, v in pr.stats().items() if k in ("MS_REGULAR", "MS_ADAPTIVE", "MS_SINGULAR", "MS_ASSEMBLE")})
