"""Column-scaled deviation of the coupled matrix from the multi-region oracle with and without an incident field (fluid-poroelastic two-box model):
tells a tolerance-level difference of the general poroelastic pair path from a defect.  Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from multifebe_b200 import capi
from multifebe_b200.host import MultiRegionModel, Region, FLUID, two_box_mesh, shape
from multifebe_b200.host.multiregion import PORO
from oracle.multiregion import MultiRegionOracle
from test_oracle_multiregion import BPART, LAT1, LAT2, PO
from test_coupled_from_single_region import FL, bcs_for, _random_incident

ctx = capi.Context(0)
for et in (shape.QUAD8, shape.QUAD9):
    bcs = bcs_for(FLUID, LAT1, 1, True); bcs.update(bcs_for(PORO, LAT2, 2, False))
    mrm = MultiRegionModel(two_box_mesh(2, et), [Region(FLUID, FL, [1, 3, 4, 5, 6, 7]), Region(PORO, PO, [-7, 2, 13, 14, 15, 16])], BPART, bcs)
    A0, b0 = MultiRegionOracle(mrm).assemble(1.7)
    sc = np.abs(A0).max(axis=0)
    cp = capi.CoupledProblem(ctx, mrm)
    for inc in (False, True):
        if inc:
            for kr in (0, 1):
                cp.set_incident(kr, *_random_incident(mrm, kr, 20 + kr))
            _, b0 = MultiRegionOracle(mrm).assemble(1.7)
        A, b = cp.assemble(1.7)
        d = np.abs(A - A0).max(axis=0) / sc
        j = int(np.argmax(d)); i = int(np.argmax(np.abs(A - A0)[:, j]))
        print("etype", et, "incident", inc, "max col-scaled dA %.3e at (%d,%d) |A0|=%.3e colmax=%.3e  db %.3e" % (d.max(), i, j, abs(A0[i, j]), sc[j], np.abs(b - b0).max() / np.abs(b0).max()), flush=True)
    cp.close()
