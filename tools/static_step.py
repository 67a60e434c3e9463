"""One static solve of the BASELINE config 2 cube (profiling target): python tools/static_step.py [m]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multifebe_b200 import capi
from multifebe_b200.host import *
m = int(sys.argv[1]) if len(sys.argv) > 1 else 11
ctx = capi.Context(0); md = Model(cube_mesh(m, shape.QUAD9), cube_bcs()); pr = capi.Problem(ctx, md)
sm = Material(1, 1, 0.25, 0.0)
pr.solve_static(sm); pr.solve_static(sm)
print({k: round(v, 2) for k, v in pr.stats().items() if k.startswith("MS_")})
