"""GPU parity tests of the Biot poroelastic BE region through the C ABI (mfb_harpor3d_*) against the CPU oracle.
First hardware run (compute-sanitizer clean, all element types green): profiles/r02_first_contact.log."""
import os
import numpy as np
import pytest
from multifebe_b200.host import Poro, PoroModel, Model, Material, cube_mesh, cube_bcs, shape

pytestmark = [pytest.mark.gpu]
TOL_A, TOL_X = 1e-11, 1e-8
PO = Poro(rhof=1.0, rhos=2.2, lam=1.2, mu=1.0, xi=0.02, phi=0.35, rhoa=0.15, R=0.8, Q=0.5, b=0.4)


def column_bcs(P=1.0, q=0.0):
    bcs = {1: ([1, 0, 0, 0], [q, 0, 0, 0]), 2: ([0, 1, 1, 1], [0.2 * P, P, 0, 0])}
    for p_, free in ((3, 2), (4, 2), (5, 3), (6, 3)):
        ct = [1, 1, 1, 1]; ct[free] = 0
        bcs[p_] = (ct, [0, 0, 0, 0])
    return bcs


@pytest.mark.parametrize("et,m", [(shape.TRI3, 3), (shape.TRI6, 2), (shape.QUAD4, 3), (shape.QUAD8, 2), (shape.QUAD9, 2)])
@pytest.mark.parametrize("omega", [0.3, 4.0])
def test_assembly_and_solution_parity(gpu_ctx, oracle_lib, et, m, omega):
    from multifebe_b200 import capi
    md = PoroModel(cube_mesh(m, et), column_bcs(1.0, 0.05 + 0.02j))
    pr = capi.Problem(gpu_ctx, md)
    A, b = pr.build_lse_mechanics_bem_harpor(omega, PO)
    Ao, bo, st = oracle_lib.PorOracle(md).assemble(omega, PO)
    # columns of tau, Un, u_k, t_k live on different scales: compare every column against its own maximum
    sc = np.abs(Ao).max(axis=0)
    assert (np.abs(A - Ao).max(axis=0) <= TOL_A * sc).all(), (np.abs(A - Ao).max(axis=0) / sc).max()
    assert np.abs(b - bo).max() <= TOL_A * np.abs(bo).max()
    s = pr.stats()
    assert s["PAIRS_REGULAR"] == sum(st["pairs_regular"].values()) and s["POINTS_REGULAR"] == st["pts_regular"]
    assert s["PAIRS_ADAPTIVE"] == st["pairs_adaptive"] and s["LEAVES"] == st["leaves"] and s["PAIRS_SINGULAR"] == st["pairs_singular"]
    xo = np.linalg.solve(Ao, bo)
    x = pr.solve_frequency_poro(omega, PO)
    # one scale per family (tau, Un, the displacement vector, the traction vector): a component that vanishes by symmetry has no scale of its own
    for fam in (md.col_u[:, :1], md.col_t[:, :1], md.col_u[:, 1:], md.col_t[:, 1:]):
        cols = fam[fam >= 0]
        if len(cols):
            assert np.abs(x[cols] - xo[cols]).max() <= TOL_X * np.abs(xo[cols]).max()
    pr.close()


def test_wrong_family_is_refused(gpu_ctx):
    from multifebe_b200 import capi
    pm = PoroModel(cube_mesh(1, shape.QUAD4), column_bcs())
    pr = capi.Problem(gpu_ctx, pm)
    with pytest.raises(capi.MfbError):
        pr.build_lse_mechanics_bem_harela(1.0, Material())
    pr.close()
    pe = capi.Problem(gpu_ctx, Model(cube_mesh(1, shape.QUAD4), cube_bcs()))
    with pytest.raises(capi.MfbError):
        pe.build_lse_mechanics_bem_harpor(1.0, PO)
    pe.close()
