"""One frequency at full size (profiling target for ncu): usage one_step.py [etype m]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multifebe_b200 import capi
from multifebe_b200.host import *
et = int(sys.argv[1]) if len(sys.argv) > 1 else shape.TRI3
m = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ctx = capi.Context(0); md = Model(cube_mesh(m, et), cube_bcs()); pr = capi.Problem(ctx, md)
pr.solve_frequency(9.0, Material(1, 1, 0.25, 0.03))
print({k: round(v, 2) for k, v in pr.stats().items() if k.startswith("MS_")})
