"""GPU parity tests of the acoustic (inviscid fluid) BE region (first brick), through the C ABI
(mfb_harpot3d_*) against the CPU oracle on the same inputs.  Tolerances as for the elastic path: assembled entries within 1e-11
relative (max-norm), solutions within 1e-8."""
import os
import numpy as np
import pytest
from multifebe_b200.host import Fluid, FluidModel, Model, Material, cube_mesh, cube_bcs, room_bcs, room_analytic, shape

pytestmark = pytest.mark.gpu
TOL_A, TOL_X = 1e-11, 1e-8
AIR = Fluid(rho=1.25, c=343.0)


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def mixed_bcs():
    """Nonzero prescribed values of both kinds, so that h and g both reach A and b."""
    return {1: (0, 0.3 - 0.1j), 2: (0, 1.0), 3: (1, 0.0), 4: (1, 2e-6 + 1e-6j), 5: (1, 0.0), 6: (0, -0.5)}


@pytest.mark.parametrize("et,m", [(shape.TRI3, 3), (shape.TRI6, 2), (shape.QUAD4, 3), (shape.QUAD8, 2), (shape.QUAD9, 2)])
@pytest.mark.parametrize("omega,fl", [(2 * np.pi * 20.0, AIR), (2 * np.pi * 300.0, Fluid(rho=1.25, c=343.0, xi=0.02)), (0.05, Fluid(rho=1.0, c=1.0))])
def test_assembly_and_solution_parity(gpu_ctx, oracle_lib, et, m, omega, fl):
    from multifebe_b200 import capi
    md = FluidModel(cube_mesh(m, et), mixed_bcs())
    pr = capi.Problem(gpu_ctx, md)
    A, b = pr.build_lse_mechanics_bem_harpot(omega, fl)
    Ao, bo, st = oracle_lib.PotOracle(md).assemble(omega, fl)
    # columns of p (entries ~ h) and of Un (entries ~ g rho omega^2) differ by orders of magnitude: compare per column scale
    sc = np.abs(Ao).max(axis=0)
    assert (np.abs(A - Ao).max(axis=0) <= TOL_A * sc).all(), (np.abs(A - Ao).max(axis=0) / sc).max()
    assert relerr(b, bo) < TOL_A
    s = pr.stats()
    assert s["PAIRS_REGULAR"] == sum(st["pairs_regular"].values()) and s["POINTS_REGULAR"] == st["pts_regular"]
    assert s["PAIRS_ADAPTIVE"] == st["pairs_adaptive"] and s["LEAVES"] == st["leaves"] and s["POINTS_ADAPTIVE"] == st["pts_adaptive"]
    assert s["PAIRS_SINGULAR"] == st["pairs_singular"] and s["POINTS_SINGULAR"] == st["pts_singular"]
    xo = np.linalg.solve(Ao, bo)
    x1 = pr.solve_lse_c(A.copy(order="F"), b)
    x2 = pr.solve_frequency_fluid(omega, fl)
    # p and Un live on different scales: compare each family against its own maximum
    for cols in (md.col_u[md.col_u >= 0], md.col_t[md.col_t >= 0]):
        assert relerr(x1[cols], xo[cols]) < TOL_X and relerr(x2[cols], xo[cols]) < TOL_X
    assert s["LAUNCHES"] >= 3
    pr.close()


def test_room_tutorial_on_the_gpu(gpu_ctx):
    """ME-TH-AC-001 (docs/examples/ME-TH-AC-001): p = P sin kx / sin kL on the rigid walls, 8 x 8 quad9 cells per wall as in cube.geo."""
    from multifebe_b200 import capi
    md = FluidModel(cube_mesh(4, shape.QUAD9, L=3.0), room_bcs(1.0))
    pr = capi.Problem(gpu_ctx, md)
    # discretisation error of the 4 x 4 quad9 mesh: kL = 0.55, 2.2 (below f_1 = 57.2 Hz) and 4.9 (between f_1 and f_2 = 114.3 Hz; measured 3.9e-3)
    for f_hz, tol in ((10.0, 2e-3), (40.0, 2e-3), (90.0, 1e-2)):
        omega = 2 * np.pi * f_hz
        p, un = md.nodal_solution(pr.solve_frequency_fluid(omega, AIR))
        p_ex, ux_ex = room_analytic(md.node_x[:, 0], omega, AIR, L=3.0, P=1.0)
        assert np.abs(p - p_ex).max() <= tol * np.abs(p_ex).max()
        sign = np.where(md.node_part == 1, -1.0, np.where(md.node_part == 2, 1.0, 0.0))
        assert np.abs(un - sign * ux_ex).max() <= 10 * tol * np.abs(ux_ex).max()
    pr.close()


def test_plan_modes_follow_the_scalar_estimator(gpu_ctx, oracle_lib):
    from multifebe_b200 import capi
    md = FluidModel(cube_mesh(4, shape.QUAD8), room_bcs())
    pr = capi.Problem(gpu_ctx, md)
    orc = oracle_lib.PotOracle(md)
    rng = np.random.default_rng(3)
    cs = rng.integers(0, md.n_colloc, 300).astype(np.int32); es = rng.integers(0, md.n_elem, 300).astype(np.int32)
    got = pr.plan_modes(cs, es)
    want = [orc.pair(int(e), md.colloc_x[int(c)], 1.0, AIR)[2] for c, e in zip(cs, es)]
    assert list(got) == want
    pr.close()


def test_wrong_family_is_refused(gpu_ctx):
    from multifebe_b200 import capi
    fm = FluidModel(cube_mesh(2, shape.QUAD4), room_bcs())
    pr = capi.Problem(gpu_ctx, fm)
    with pytest.raises(capi.MfbError):
        pr.build_lse_mechanics_bem_harela(1.0, Material())
    pr.close()
    em = Model(cube_mesh(2, shape.QUAD4), cube_bcs())
    pe = capi.Problem(gpu_ctx, em)
    with pytest.raises(capi.MfbError):
        pe.build_lse_mechanics_bem_harpot(1.0, AIR)
    pe.close()


def test_interior_pressures_match_the_oracle_composition(gpu_ctx, oracle_lib):
    from multifebe_b200 import capi
    md = FluidModel(cube_mesh(3, shape.QUAD9), room_bcs(1.0))
    omega = 2 * np.pi * 30.0
    pts = np.array([[0.5, 0.5, 0.5], [0.2, 0.7, 0.4], [0.93, 0.5, 0.5], [0.31, 0.08, 0.77], [0.5, 0.5, 0.985]])
    pr = capi.Problem(gpu_ctx, md)
    x = pr.solve_frequency_fluid(omega, AIR)
    ip = capi.InternalPoints(gpu_ctx, md, pts)
    pin = ip.pressures(omega, AIR, x)
    o = oracle_lib.PotOracle(md)
    p, un = md.nodal_solution(x)
    ref = np.zeros(len(pts), dtype=np.complex128)
    for k, xp in enumerate(pts):
        for e in range(md.n_elem):
            h, g, _ = o.pair(e, xp, omega, AIR)
            nodes = md.mesh.conn[e]
            ref[k] += (g * AIR.rho * omega ** 2) @ un[nodes] - h @ p[nodes]
    assert relerr(pin, ref) < 1e-10
    p_ex, _ = room_analytic(pts[:, 0], omega, AIR)
    assert np.abs(pin - p_ex).max() < 2e-4 * np.abs(p_ex).max()
    ip.close(); pr.close()
