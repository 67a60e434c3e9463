"""Assembly only, several times, at full size: prints the per-kernel-phase device times of each repetition.
usage: asm_only.py [etype m reps omega]   (MFB_LIB selects a variant build of libmfb.so)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multifebe_b200 import capi
from multifebe_b200.host import *
et = int(sys.argv[1]) if len(sys.argv) > 1 else shape.TRI3
m = int(sys.argv[2]) if len(sys.argv) > 2 else 40
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
omega = float(sys.argv[4]) if len(sys.argv) > 4 else 9.0
ctx = capi.Context(0); md = Model(cube_mesh(m, et), cube_bcs()); pr = capi.Problem(ctx, md)
for r in range(reps):
    pr.build_lse_mechanics_bem_harela(omega, Material(1, 1, 0.25, 0.03), want_host=False)
    s = pr.stats()
    print(os.environ.get("MFB_LIB", "HEAD"), r, {k: round(s[k], 2) for k in ("MS_ZERO", "MS_REGULAR", "MS_ADAPTIVE", "MS_SINGULAR", "MS_ASSEMBLE")},
          "K1 TFLOP/s %.2f" % (s["FLOPS_REGULAR"] / s["MS_REGULAR"] / 1e9), flush=True)
