"""Symmetry planes on coupled (multi-region) models, oracle side (CPU): the layered columns of tests/test_oracle_multiregion.py as QUARTER models.
Their sliding / rigid side walls at y = 0 and z = 0 are symmetry planes, so removing those faces and declaring the planes must leave the exact 1D
two-layer solutions in place: solid | solid, fluid | solid (both orders), fluid | fluid."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from multifebe_b200.host import Material, Fluid, MultiRegionModel, Region, SOLID, FLUID, two_box_mesh, without_parts, shape  # noqa: E402
from oracle.multiregion import MultiRegionOracle  # noqa: E402
from test_oracle_multiregion import layered_1d, solid_bcs, fluid_bcs  # noqa: E402

KEEP = (1, 2, 4, 6, 7, 14, 16)
BPART_Q = {b: b for b in KEEP}
PLANES = [("y", "symmetry"), ("z", "symmetry")]


def quarter(m, et, xs):
    return without_parts(two_box_mesh(m, et, xs=xs), {3, 5, 13, 15})


def quarter_two_layer_solid(m=2, et=shape.QUAD9):
    m1, m2 = Material(1.0, 1.0, 0.25, 0.02), Material(2.0, 3.0, 0.3, 0.05)
    bcs = {k: v for k, v in solid_bcs().items() if k in KEEP}
    mrm = MultiRegionModel(quarter(m, et, 0.4), [Region(SOLID, m1, [1, 4, 6, 7]), Region(SOLID, m2, [-7, 2, 14, 16])], BPART_Q, bcs, symmetry=PLANES)
    return mrm, m1, m2


def test_two_layer_solid_column_as_a_quarter_model():
    omega = 2.5
    mrm, m1, m2 = quarter_two_layer_solid()
    A, b = MultiRegionOracle(mrm).assemble(omega)
    x = np.linalg.solve(A, b)
    field = layered_1d(omega, [(omega / m1.c1, m1.lam + 2 * m1.mu), (omega / m2.c1, m2.lam + 2 * m2.mu), 0.4], ("f", 0.0), ("s", 1.0))
    for kr in (0, 1):
        u, t = mrm.nodal_solution(x, kr)
        ok = ~np.isnan(u[:, 0])
        ua, _ = field(np.clip(mrm.node_x[ok, 0], 0, 0.4) if kr == 0 else np.clip(mrm.node_x[ok, 0], 0.4 + 1e-12, 1))
        assert np.abs(u[ok, 0] - ua).max() < 3e-3 * np.abs(ua).max()
        assert np.abs(u[ok, 1:]).max() < 3e-3 * np.abs(ua).max()          # the planes hold the column: no lateral motion


def quarter_fluid_solid(solid_first):
    ms, fl = Material(2.0, 1.5, 0.25, 0.03), Fluid(1.0, 1.2, 0.01)
    if solid_first:
        regs = [Region(SOLID, ms, [1, 4, 6, 7]), Region(FLUID, fl, [-7, 2, 14, 16])]
        bcs = {k: v for k, v in solid_bcs().items() if k in (1, 4, 6)}; bcs.update({k: v for k, v in fluid_bcs(1.0).items() if k in (2, 14, 16)})
    else:
        regs = [Region(FLUID, fl, [1, 4, 6, 7]), Region(SOLID, ms, [-7, 2, 14, 16])]
        bcs = {k: v for k, v in fluid_bcs().items() if k in (1, 4, 6)}; bcs[1] = (0, 1.0)
        sb = solid_bcs(); bcs.update({k: sb[k - 10] for k in (14, 16)}); bcs[2] = ([0, 0, 0], [0, 0, 0])
    return MultiRegionModel(quarter(2, shape.QUAD9, 0.5), regs, BPART_Q, bcs, symmetry=PLANES), ms, fl


@pytest.mark.parametrize("solid_first", [True, False])
def test_fluid_solid_column_as_a_quarter_model(solid_first):
    omega, xs = 3.0, 0.5
    mrm, ms, fl = quarter_fluid_solid(solid_first)
    A, b = MultiRegionOracle(mrm).assemble(omega)
    x = np.linalg.solve(A, b)
    Zs, Kf = ms.lam + 2 * ms.mu, fl.rho * fl.c ** 2
    ks, kf = omega / ms.c1, omega / fl.c
    field = layered_1d(omega, [(ks, Zs), (kf, Kf), xs], ("f", 0.0), ("s", -1.0)) if solid_first else layered_1d(omega, [(kf, Kf), (ks, Zs), xs], ("s", -1.0), ("f", 0.0))
    ks_, kfl = (0, 1) if solid_first else (1, 0)
    u, _ = mrm.nodal_solution(x, ks_)
    ok = ~np.isnan(u[:, 0])
    lo, hi = (0.0, xs) if solid_first else (xs + 1e-12, 1.0)
    wa, _ = field(np.clip(mrm.node_x[ok, 0], lo, hi))
    assert np.abs(u[ok, 0] - wa).max() < 5e-3 * np.abs(wa).max()
    p, _ = mrm.nodal_solution(x, kfl)
    ok = ~np.isnan(p)
    lo, hi = (xs + 1e-12, 1.0) if solid_first else (0.0, xs)
    _, sa = field(np.clip(mrm.node_x[ok, 0], lo, hi))
    assert np.abs(p[ok] + sa).max() < 5e-3 * np.abs(sa).max()


def test_antisymmetric_plane_changes_the_coupled_system():
    """Teeth: declaring one plane antisymmetric gives another system (the symmetric one is the column)."""
    mrm, _, _ = quarter_two_layer_solid()
    A, b = MultiRegionOracle(mrm).assemble(2.5)
    bcs = {k: v for k, v in solid_bcs().items() if k in KEEP}
    m1, m2 = Material(1.0, 1.0, 0.25, 0.02), Material(2.0, 3.0, 0.3, 0.05)
    other = MultiRegionModel(quarter(2, shape.QUAD9, 0.4), [Region(SOLID, m1, [1, 4, 6, 7]), Region(SOLID, m2, [-7, 2, 14, 16])], BPART_Q, bcs,
                             symmetry=[("y", "antisymmetry"), ("z", "symmetry")])
    A2, _ = MultiRegionOracle(other).assemble(2.5)
    assert np.abs(A - A2).max() > 1e-3 * np.abs(A).max()
