"""ME-TH-EL-002, CPU only: is the GPU-oracle gap of profiles/r02_el002_resonance_check.log a matter of the SI-unit column scaling (then LAPACK-style
equilibration, the reference's lse_scaling, would cure it) or of the resonance itself?  Prints cond_2 of A and of the column-equilibrated A at the five
frequencies of that log, the measured gap over the latter, and the effect of a random perturbation of A of 1e-13 of the column scale (order of magnitude only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from multifebe_b200.host.casefile import CaseFile
from multifebe_b200.host import Material
from oracle import oracle as orc

c = CaseFile(os.path.join(ROOT, "tests", "golden", "ME-TH-EL-002", "INPUT_DATA_FILE.txt")); m = c.build_model()
o = orc.Oracle(m)
gap = {5: 7.67e-9, 9: 1.69e-8, 10: 2.73e-8, 11: 7.31e-8, 60: 6.24e-10}          # |x_gpu - x_oracle| / max|x|, measured on the B200
rng = np.random.default_rng(0)
for kf in (5, 9, 10, 11, 60):
    A, b, _ = o.assemble(c.omega[kf], c.material)
    sc = np.abs(A).max(axis=0)
    ce = np.linalg.cond(A / sc)
    x = np.linalg.solve(A, b)
    dA = (rng.normal(size=A.shape) + 1j * rng.normal(size=A.shape)) * 1e-13 * sc
    xp = np.linalg.solve(A + dA, b)
    print("kf %2d  cond(A) %.1e  cond(column-equilibrated A) %.1e  measured gap %.1e  gap/cond_eq %.1e  |dx|/|x| under a random 1e-13 column-relative perturbation %.1e" % (
        kf, np.linalg.cond(A), ce, gap[kf], gap[kf] / ce, np.abs(xp - x).max() / np.abs(x).max()), flush=True)
# the same problem in units where mu = 1 (stiffness and density divided by mu): the oracle's own solution must not care
s = c.material.mu_r
A, b, _ = o.assemble(c.omega[10], c.material); u, t = m.nodal_solution(np.linalg.solve(A, b))
A2, b2, _ = o.assemble(c.omega[10], Material(c.material.rho / s, 1.0, c.material.nu_r, c.material.xi)); u2, t2 = m.nodal_solution(np.linalg.solve(A2, b2))
print("kf 10 in units with mu = 1: cond %.1e -> %.1e; oracle against itself: u %.1e, t %.1e" % (np.linalg.cond(A), np.linalg.cond(A2),
      np.abs(u - u2).max() / np.abs(u).max(), np.abs(t - t2 * s).max() / np.abs(t).max()))
