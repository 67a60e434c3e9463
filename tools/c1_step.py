"""Per-frequency cost of BASELINE config C1 (t3.msh, 1386 DOF) on one GPU, one frequency at a time (round-1 path): library events + wall."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multifebe_b200 import capi
from multifebe_b200.host.casefile import CaseFile
case = CaseFile(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ME-TH-EL-001", "t3.dat"))
md = case.build_model(); ctx = capi.Context(0); pr = capi.Problem(ctx, md)
fr = case.omega
for w in range(5): pr.solve_frequency(fr[w], case.material)
acc = {}; t0 = time.time()
for kf in range(0, 300, 3):
    pr.solve_frequency(fr[kf], case.material); st = pr.stats()
    for k, v in st.items():
        if k.startswith("MS_") or k.endswith("LAUNCHES"): acc[k] = acc.get(k, 0.0) + v / 100
wall = (time.time() - t0) / 100
print("C1 per frequency: wall %.3f ms; %s" % (wall * 1e3, {k: round(v, 3) for k, v in acc.items()}))
