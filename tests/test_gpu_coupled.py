"""Coupled BE regions on the GPU (capi.CoupledProblem: single-region kernels + the column combination of host/coupled.py) against the
multi-region oracle (first hardware run green in round 2: profiles/r02_first_contact.log)."""
import os
import numpy as np
import pytest
from multifebe_b200.host import MultiRegionModel, Region, SOLID, FLUID, two_box_mesh, shape
from multifebe_b200.host.multiregion import PORO
from test_oracle_multiregion import BPART, LAT1, LAT2, PO
from test_coupled_from_single_region import MS, FL, bcs_for

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("kinds,ict", [((SOLID, SOLID), 0), ((FLUID, FLUID), 0), ((SOLID, FLUID), 0), ((FLUID, SOLID), 0), ((FLUID, PORO), 0), ((PORO, FLUID), 1),
                                       ((SOLID, PORO), 0), ((PORO, PORO), 0)])
def test_coupled_system_and_solution(gpu_ctx, kinds, ict):
    from multifebe_b200 import capi
    from oracle.multiregion import MultiRegionOracle
    mats = {SOLID: MS, FLUID: FL, PORO: PO}
    bcs = bcs_for(kinds[0], LAT1, 1, True); bcs.update(bcs_for(kinds[1], LAT2, 2, False))
    mrm = MultiRegionModel(two_box_mesh(2, shape.QUAD9), [Region(kinds[0], mats[kinds[0]], [1, 3, 4, 5, 6, 7]), Region(kinds[1], mats[kinds[1]], [-7, 2, 13, 14, 15, 16])],
                           BPART, bcs, interface_ctype={7: ict})
    omega = 1.7
    cp = capi.CoupledProblem(gpu_ctx, mrm)
    A, b = cp.assemble(omega)
    A0, b0 = MultiRegionOracle(mrm).assemble(omega)
    sc = np.abs(A0).max(axis=0)
    assert (np.abs(A - A0).max(axis=0) <= 1e-11 * sc).all(), (np.abs(A - A0).max(axis=0) / sc).max()
    assert np.abs(b - b0).max() <= 1e-11 * np.abs(b0).max()
    x = cp.solve_frequency(omega)
    xr = cp.solve_frequency_resident(omega)            # the combination done on the device (mfb_combine_columns / mfb_add_entries)
    x0 = np.linalg.solve(A0, b0)
    assert np.abs((xr - x0) * sc).max() <= 1e-8 * np.abs(x0 * sc).max()
    # variables of different families live on different scales: compare every unknown against the largest of its kind (by column scale)
    assert np.abs((x - x0) * sc).max() <= 1e-8 * np.abs(x0 * sc).max()
    cp.close()


@pytest.mark.parametrize("kinds,ict,planes", [((SOLID, SOLID), 0, [("y", "symmetry"), ("z", "symmetry")]), ((SOLID, FLUID), 0, [("y", "symmetry"), ("z", "antisymmetry")]),
                                              ((FLUID, PORO), 0, [("y", "symmetry"), ("z", "symmetry")]), ((PORO, PORO), 0, [("y", "antisymmetry")])])
def test_coupled_regions_with_symmetry_planes(gpu_ctx, kinds, ict, planes):
    """[symmetry planes] on a coupled model: every local assembly integrates the mirror images of its elements (quarter / half of the two-box model);
    the oracle side is pinned by the layered columns as quarter models (tests/test_oracle_multiregion_symmetry.py)."""
    from multifebe_b200 import capi
    from multifebe_b200.host import without_parts
    from oracle.multiregion import MultiRegionOracle
    mats = {SOLID: MS, FLUID: FL, PORO: PO}
    drop = {"y": (3, 13), "z": (5, 15)}
    gone = set(p for a, _ in planes for p in drop[a])
    lat1, lat2 = tuple(p for p in LAT1 if p not in gone), tuple(p for p in LAT2 if p not in gone)
    bcs = bcs_for(kinds[0], LAT1, 1, True); bcs.update(bcs_for(kinds[1], LAT2, 2, False))
    bcs = {k: v for k, v in bcs.items() if k not in gone}
    bpart = {b: b for b in BPART if b not in gone}
    mesh = without_parts(two_box_mesh(2, shape.QUAD9), gone)
    mrm = MultiRegionModel(mesh, [Region(kinds[0], mats[kinds[0]], [1] + list(lat1) + [7]), Region(kinds[1], mats[kinds[1]], [-7, 2] + list(lat2))],
                           bpart, bcs, interface_ctype={7: ict}, symmetry=planes)
    omega = 1.7
    cp = capi.CoupledProblem(gpu_ctx, mrm)
    A, b = cp.assemble(omega)
    A0, b0 = MultiRegionOracle(mrm).assemble(omega)
    sc = np.abs(A0).max(axis=0)
    assert (np.abs(A - A0).max(axis=0) <= 1e-11 * sc).all(), (np.abs(A - A0).max(axis=0) / sc).max()
    assert np.abs(b - b0).max() <= 1e-11 * max(np.abs(b0).max(), 1e-300)
    xr = cp.solve_frequency_resident(omega)
    x0 = np.linalg.solve(A0, b0)
    assert np.abs((xr - x0) * sc).max() <= 1e-8 * np.abs(x0 * sc).max()
    cp.close()


@pytest.mark.parametrize("kinds,where", [((SOLID, FLUID), (0,)), ((FLUID, PORO), (0, 1)), ((PORO, SOLID), (1,))])
def test_coupled_regions_with_incident_fields(gpu_ctx, kinds, where):
    """An incident field in one or both regions of a coupled model (CoupledProblem.set_incident): the H problem of the region runs with the field set on the
    device, its right-hand side and the free-term part join b -- host combination and resident combination against the multi-region oracle.
    (quad9 like the other coupled cases.  On the quad8 two-box mesh this test failed its 1e-11 bar on A at one diagonal entry, 1.27e-11 of the column scale, with or
    without a field (profiles/r02_coupled_incident_check.log).  That entry belongs to a corner node of the poroelastic box, a rim node with an MCA collocation point: its
    free term is phi/2 on both sides, so the difference is in the singular integral of the point's own quad8 element in the poroelastic kernel -- a miss of the bar on
    quad8 that is open, not a property of the incident field.  An earlier version of this note blamed the oracle's acos free term; that was wrong.)"""
    from multifebe_b200 import capi
    from oracle.multiregion import MultiRegionOracle
    from test_coupled_from_single_region import _random_incident
    mats = {SOLID: MS, FLUID: FL, PORO: PO}
    bcs = bcs_for(kinds[0], LAT1, 1, True); bcs.update(bcs_for(kinds[1], LAT2, 2, False))
    mrm = MultiRegionModel(two_box_mesh(2, shape.QUAD9), [Region(kinds[0], mats[kinds[0]], [1, 3, 4, 5, 6, 7]), Region(kinds[1], mats[kinds[1]], [-7, 2, 13, 14, 15, 16])],
                           BPART, bcs)
    omega = 1.7
    cp = capi.CoupledProblem(gpu_ctx, mrm)
    for kr in where:
        cp.set_incident(kr, *_random_incident(mrm, kr, 20 + kr))
    A0, b0 = MultiRegionOracle(mrm).assemble(omega)
    A, b = cp.assemble(omega)
    sc = np.abs(A0).max(axis=0)
    assert (np.abs(A - A0).max(axis=0) <= 1e-11 * sc).all(), (np.abs(A - A0).max(axis=0) / sc).max()
    assert np.abs(b - b0).max() <= 1e-11 * np.abs(b0).max()
    x0 = np.linalg.solve(A0, b0)
    xr = cp.solve_frequency_resident(omega)
    assert np.abs((xr - x0) * sc).max() <= 1e-8 * np.abs(x0 * sc).max()
    # cleared again: the plain system comes back (the H problems drop the field)
    for kr in where:
        cp.set_incident(kr)
    _, b1 = cp.assemble(omega)
    _, b10 = MultiRegionOracle(mrm).assemble(omega)
    assert np.abs(b1 - b10).max() <= 1e-11 * np.abs(b10).max()
    cp.close()
