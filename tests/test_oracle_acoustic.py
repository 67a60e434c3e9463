"""CPU pins of the acoustic (inviscid fluid, scalar Helmholtz) oracle: SURVEY.md section 8f rank 3, first brick.

The reference holds no numeric golden vectors for this path either; the restatement of fbem_bem_harpot3d_* /
build_lse_mechanics_bem_harpot / assemble_bem_harpot_equation is pinned by
  * the closed form of the fundamental solution (p* = e^{-ikr}/(4 pi r), q* = dp*/dn),
  * the identity c + int q* dS = O(k^2) over a closed surface (ties free term, singular, quasi-singular and regular parts),
  * the analytic solution of the reference's tutorial ME-TH-AC-001 (room with p = 0 / p = P on two opposite walls, rigid
    walls elsewhere: p = P sin kx / sin kL, U_x = P k cos kx / (rho omega^2 sin kL)), all five element types.
"""
import numpy as np
import pytest

from multifebe_b200.host import Fluid, FluidModel, cube_mesh, room_bcs, room_analytic, shape
from oracle import oracle as orc

ETYPES = [shape.TRI3, shape.TRI6, shape.QUAD4, shape.QUAD8, shape.QUAD9]


@pytest.mark.parametrize("omega", [0.05, 3.0, 400.0])
def test_fundamental_solution_closed_form(omega):
    fl = Fluid(rho=1.25, c=343.0 if omega > 100 else 1.0, xi=0.01)
    rng = np.random.default_rng(7)
    for _ in range(20):
        x_i = rng.normal(size=3); x = x_i + rng.normal(size=3) * rng.choice([0.01, 0.3, 2.0])
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        po, qo = orc.fundamental_solutions_pot(x, n, x_i, omega, fl)
        rv = x - x_i; r = np.linalg.norm(rv); k = omega / fl.c
        p_ref = np.exp(-1j * k * r) / (4 * np.pi * r)
        q_ref = -(1.0 + 1j * k * r) * np.exp(-1j * k * r) / (4 * np.pi * r * r) * (rv @ n) / r
        assert abs(po - p_ref) <= 2e-13 * abs(p_ref)
        assert abs(qo - q_ref) <= 1e-11 * max(abs(q_ref), abs(p_ref) / r)   # q* cancels when dr/dn -> 0


def test_decomposed_zexp_matches_the_seven_term_version():
    L = orc.lib()
    for z in [0.0 + 0.0j, 1e-7 - 3e-7j, 0.02 - 0.3j, -0.1 - 0.99j, 0.3 - 1.4j, -2.0 - 7.0j]:
        z_ri = np.array([z.real, z.imag]); E5 = np.zeros(10)
        L.orc_decomposed_zexp(orc._p(z_ri), orc._p(E5))
        E5 = E5[0::2] + 1j * E5[1::2]
        E7 = orc.zexp_decomposed(z)
        assert np.abs(E5 - E7[:5]).max() <= 4e-16 * max(1.0, abs(np.exp(z)))


@pytest.mark.parametrize("et", ETYPES)
def test_constant_pressure_identity(et):
    """p = 1, Un = 0 solves the Laplace limit: every row of (H + C) sums to O((kL)^2)."""
    m = {shape.TRI3: 3, shape.QUAD4: 3}.get(et, 2)
    bcs = {p: (1, 0.0) for p in range(1, 7)}          # Un known everywhere -> every column is a pressure, A = H + C
    md = FluidModel(cube_mesh(m, et), bcs)
    fl = Fluid(rho=1.0, c=1.0)
    omega = 1e-4
    A, b, st = orc.PotOracle(md).assemble(omega, fl)
    assert st["pairs_singular"] > 0 and st["pairs_adaptive"] > 0 and st["pts_regular"] > 0
    rs = A.sum(axis=1)
    assert np.abs(rs).max() < 5e-6, np.abs(rs).max()   # quadrature error qsi_relative_error = 1e-6, plus O(k^2) = 1e-8
    assert np.abs(b).max() == 0.0


@pytest.mark.parametrize("et", ETYPES)
def test_room_tutorial_analytic_solution(et):
    """ME-TH-AC-001 on the S-cube: pressure on the rigid walls and normal displacement on the two driven walls."""
    m = {shape.TRI3: 4, shape.QUAD4: 4}.get(et, 2)
    md = FluidModel(cube_mesh(m, et), room_bcs(1.0))
    fl = Fluid(rho=1.25, c=343.0)
    omega = 2 * np.pi * 20.0                     # kL = 0.37 with L = 1: below the first natural frequency (171.5 Hz for L = 1)
    A, b, _ = orc.PotOracle(md).assemble(omega, fl)
    x = np.linalg.solve(A, b)
    p, un = md.nodal_solution(x)
    p_ex, ux_ex = room_analytic(md.node_x[:, 0], omega, fl, L=1.0, P=1.0)
    tol = 2e-3 if et in (shape.TRI3, shape.QUAD4) else 1e-4   # discretisation error of the 4x4 linear / 2x2 quadratic meshes
    assert np.abs(p - p_ex).max() <= tol * np.abs(p_ex).max()
    # Un: outward normal is -x on part 1 (x=0) and +x on part 2 (x=L); zero on the rigid walls (prescribed)
    sign = np.where(md.node_part == 1, -1.0, np.where(md.node_part == 2, 1.0, 0.0))
    assert np.abs(un - sign * ux_ex).max() <= 10 * tol * np.abs(ux_ex).max()


def test_plan_uses_the_scalar_estimator_order():
    """f = 3 (bem_harpot3d.f90:1009) needs fewer points than the elastic f = 5 at the same distance."""
    L = orc.lib()
    for d in [2.5, 4.0, 8.0]:
        n3 = orc.qs_n(False, shape.QUAD9, 3, 1e-6, d, [0.0, 0.0]); n5 = orc.qs_n(False, shape.QUAD9, 5, 1e-6, d, [0.0, 0.0])
        assert 2 <= n3 <= n5
    md = FluidModel(cube_mesh(2, shape.QUAD4), room_bcs())
    from multifebe_b200.host import Model, Material, cube_bcs
    me = Model(cube_mesh(2, shape.QUAD4), cube_bcs())
    fl = Fluid(rho=1.0, c=1.0); mat = Material()
    x_i = np.array([0.3, 0.3, 0.1])
    modes_p = [orc.PotOracle(md).pair(e, x_i, 1.0, fl)[2] for e in range(md.n_elem)]
    modes_e = [orc.Oracle(me).pair(e, x_i, 1.0, mat)[2] for e in range(me.n_elem)]
    assert all(a <= b for a, b in zip(modes_p, modes_e)) and any(a < b for a, b in zip(modes_p, modes_e))


def test_interior_pressure_follows_the_room_solution():
    """Interior points of a fluid region (src/calculate_internal_points_mechanics_bem_harpot.f90): p(x) = sum_e (g rho omega^2 Un - h p) with the
    integrators of the boundary equations; ME-TH-AC-001 computes the field inside the room this way."""
    from multifebe_b200.host import InternalPointsModel
    md = FluidModel(cube_mesh(3, shape.QUAD9), room_bcs(1.0))
    fl = Fluid(rho=1.25, c=343.0)
    omega = 2 * np.pi * 30.0
    o = orc.PotOracle(md)
    A, b, _ = o.assemble(omega, fl)
    x = np.linalg.solve(A, b)
    p, un = md.nodal_solution(x)
    pts = np.array([[0.5, 0.5, 0.5], [0.2, 0.7, 0.4], [0.93, 0.5, 0.5], [0.31, 0.08, 0.77]])
    d1J = fl.rho * omega ** 2
    pin = np.zeros(len(pts), dtype=np.complex128)
    for ip, xp in enumerate(pts):
        for e in range(md.n_elem):
            h, g, _ = o.pair(e, xp, omega, fl)
            nodes = md.mesh.conn[e]
            pin[ip] += (g * d1J) @ un[nodes] - h @ p[nodes]
    p_ex, _ = room_analytic(pts[:, 0], omega, fl)
    assert np.abs(pin - p_ex).max() < 1e-4 * np.abs(p_ex).max()
    # the interior-point problem the library assembles: same elements and columns, one extra row per point, no free term
    ipm = InternalPointsModel(md, pts)
    assert ipm.ndof == 1 and ipm.n_dof == md.n_dof + 4 and ipm.row.shape == (md.n_node + 4, 1) and np.all(ipm.colloc_elem == -1)
    Ai, bi, _ = orc.PotOracle(ipm).assemble(omega, fl)
    xa = np.zeros(ipm.n_dof, dtype=np.complex128); xa[:md.n_dof] = x
    assert np.abs(-(Ai @ xa - bi)[md.n_dof:] - pin).max() < 1e-12 * np.abs(pin).max()
