"""ME-TH-EL-002 on the device against the oracle WITHOUT the result file in between: is the 1e-8 miss at 0.32 Hz (tests/test_reference_examples.py) the
file's rounding or the conditioning at the resonance?  Prints, per frequency, |x_gpu - x_oracle| / max|x_oracle| on the solution vector and cond_2(A)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
from multifebe_b200 import capi
from multifebe_b200.host.casefile import CaseFile
from oracle import oracle as orc

case = CaseFile(os.path.join(ROOT, "tests", "golden", "ME-TH-EL-002", "INPUT_DATA_FILE.txt"))
md = case.build_model()
ctx = capi.Context(0); pr = capi.Problem(ctx, md); o = orc.Oracle(md)
for kf in (5, 9, 10, 11, 60):
    om = case.omega[kf]
    x = pr.solve_frequency(om, case.material)
    A, b, _ = o.assemble(om, case.material)
    x0 = np.linalg.solve(A, b)
    u, _ = md.nodal_solution(x0)
    print("kf %3d f %.3f Hz  max|u| %9.3f  |x - x0|/max|x0| %.2e  |du|/max|u| %.2e  cond2(A) %.2e" % (
        kf, om / (2 * np.pi), np.abs(u).max(), np.abs(x - x0).max() / np.abs(x0).max(),
        np.abs(md.nodal_solution(x)[0] - u).max() / np.abs(u).max(), np.linalg.cond(A)), flush=True)
pr.close(); ctx.close()
