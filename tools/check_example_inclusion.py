"""examples/inclusion_p_wave: the result file written on the GPU (python -m multifebe_b200 -i ... -o OUT) against the multi-region oracle run here on the
CPU at the same three frequencies.  usage: python tools/check_example_inclusion.py OUT.nso   (about two minutes per frequency: the oracle's Python loops)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from multifebe_b200.host.casefile import CaseFile
from multifebe_b200.host.export import read_nso
from oracle.multiregion import MultiRegionOracle

c = CaseFile(os.path.join(ROOT, "examples", "inclusion_p_wave", "inclusion_p_wave.dat")); md = c.build_model()
rows = read_nso(sys.argv[1])
o = MultiRegionOracle(md)
for kf, om in enumerate(c.omega):
    t0 = time.time()
    for kr, (u, t) in c.incident_arrays(md, om).items():
        md.set_incident(kr, u, t)
    A, b = o.assemble(om); x = np.linalg.solve(A, b)
    r = rows[rows[:, 0] == kf + 1]
    worst = 0.0
    for kr in (0, 1):
        prim, sec = md.nodal_solution(x, kr)
        rr = r[r[:, 2] == c.regions[kr][0]]
        idx = [list(md.mesh.node_ids).index(int(i)) for i in rr[:, 8]]
        pu = rr[:, 12:18:2] + 1j * rr[:, 13:18:2]; pt = rr[:, 18:24:2] + 1j * rr[:, 19:24:2]
        worst = max(worst, np.abs(pu - prim[idx]).max() / np.abs(prim[idx]).max(), np.abs(pt - sec[idx]).max() / np.abs(sec[idx]).max())
    print("omega %.1f  worst relative difference (u and t families, both regions) GPU file vs oracle %.2e  cond_2(A) %.1e  (%.0f s)" % (om, worst, np.linalg.cond(A), time.time() - t0), flush=True)
