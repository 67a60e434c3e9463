"""Incident wave field on the GPU path (mfb_harela3d_set_incident) against the oracle, whose term tests/test_oracle_incident.py pins by
the integral identities of a plane wave: b within 1e-11, A untouched, the solution within 1e-8."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from multifebe_b200.host import Material, Model, cube_mesh, cube_bcs, without_parts, halfspace_patch, plane_wave, element_incident, shape  # noqa: E402

pytestmark = pytest.mark.gpu
MAT = Material(rho=1.0, mu=1.0, nu=0.25, xi=0.02)
TOL_A, TOL_X = 1e-11, 1e-8
T_KNOWN = {p: ([1, 1, 1], [0, 0, 0]) for p in range(1, 7)}


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("etype,m", [(shape.TRI3, 4), (shape.QUAD4, 3), (shape.TRI6, 2), (shape.QUAD8, 2), (shape.QUAD9, 2)], ids=["tri3", "quad4", "tri6", "quad8", "quad9"])
def test_incident_field_term_matches_the_oracle(gpu_ctx, oracle_lib, etype, m):
    """Mixed kinds of condition, nonzero prescribed values (their b terms live beside the incident ones), rim nodes with non-nodal points."""
    from multifebe_b200 import capi
    bcs = {1: ([0, 0, 0], [0.1, 0, 0.2j]), 2: ([1, 1, 1], [1, 0, 0]), 3: ([1, 0, 1], [0, 0.3, 0]), 4: ([1, 0, 1], [0, 0, 0]), 5: ([1, 1, 0], [0, 0, 0]), 6: ([1, 1, 1], [0, 0, 0])}
    md = Model(cube_mesh(m, etype), bcs, reversed_parts=(1, 2, 3, 4, 5, 6))
    pr = capi.Problem(gpu_ctx, md); o = oracle_lib.Oracle(md)
    A0, b0 = pr.build_lse_mechanics_bem_harela(2.5, MAT)
    for omega, wave in [(2.5, ("P", [1.0, 0.5, 0.2], None)), (0.8, ("S", [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]))]:
        u, t = element_incident(md, plane_wave(wave[0], wave[1], MAT, omega, polarisation=wave[2], amplitude=0.5 + 0.2j))
        pr.set_incident(u, t); o.set_incident(u, t)
        A, b = pr.build_lse_mechanics_bem_harela(omega, MAT)
        Ao, bo, _ = o.assemble(omega, MAT)
        assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A, (omega, relerr(A, Ao), relerr(b, bo))
        xo, _, _ = oracle_lib.lu_solve(Ao, bo)
        assert relerr(pr.solve_frequency(omega, MAT), xo) < TOL_X
    # cleared: the system of the first call again
    pr.set_incident(None)
    A1, b1 = pr.build_lse_mechanics_bem_harela(2.5, MAT)
    assert relerr(A1, A0) < 1e-13 and relerr(b1, b0) < 1e-12
    pr.close()


def test_total_field_equal_to_the_incident_field(gpu_ctx):
    """Prescribed tractions = incident tractions on a cavity: u = u_inc, an algebraic identity of the assembled system (no oracle involved)."""
    from multifebe_b200 import capi
    omega = 3.0
    field = plane_wave("S", [0.3, -0.4, 1.0], MAT, omega, polarisation=[1.0, 1.0, 0.0], amplitude=0.7)
    for etype, m in [(shape.TRI3, 5), (shape.QUAD9, 3)]:
        md = Model(cube_mesh(m, etype), T_KNOWN, reversed_parts=(1, 2, 3, 4, 5, 6))
        u, t = element_incident(md, field)
        for e in range(md.n_elem):
            for kn, v in enumerate(md.mesh.conn[e]):
                md.cvalue[v] = t[md.elem_ptr[e] + kn]
        pr = capi.Problem(gpu_ctx, md)
        pr.set_incident(u, t)
        un, _ = md.nodal_solution(pr.solve_frequency(omega, MAT))
        want = np.array([field(xv, [1.0, 0, 0])[0] for xv in md.node_x])
        assert relerr(un, want) < 1e-9
        pr.close()


def test_incident_field_with_symmetry_planes_and_open_surfaces(gpu_ctx, oracle_lib):
    """The images of a symmetric model take the root's incident values with the sign of the plane (the reference passes the root's u_inc, t_inc
    together with the sign-multiplied hp, gp); a free-surface patch (soil-structure layout) with mixed conditions per node."""
    from multifebe_b200 import capi
    mesh = without_parts(cube_mesh(3, shape.QUAD4), {1, 3})
    bcs = {2: ([0, 0, 0], [0, 0, 0]), 4: ([1, 1, 1], [0, 0.3, 0]), 5: ([1, 1, 1], [0, 0, 0]), 6: ([1, 1, 1], [0, 0, 0.5])}
    cases = [Model(mesh, bcs, symmetry=[("x", "symmetry"), ("y", "antisymmetry")]),
             Model(halfspace_patch(5, shape.TRI6), {1: ([1, 1, 1], [0.1, 0, 0.3j]), 2: ([0, 1, 0], [1.0, 0.5, 0.2])})]
    for md in cases:
        omega = 2.0
        u, t = element_incident(md, plane_wave("P", [0.2, 0.1, 1.0], MAT, omega))
        pr = capi.Problem(gpu_ctx, md); o = oracle_lib.Oracle(md)
        pr.set_incident(u, t); o.set_incident(u, t)
        A, b = pr.build_lse_mechanics_bem_harela(omega, MAT)
        Ao, bo, _ = o.assemble(omega, MAT)
        assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A, (relerr(A, Ao), relerr(b, bo))
        pr.close()


def test_incident_field_misuse(gpu_ctx):
    from multifebe_b200 import capi
    md = Model(cube_mesh(2, shape.TRI3), cube_bcs())
    pr = capi.Problem(gpu_ctx, md)
    u, t = element_incident(md, plane_wave("P", [1.0, 0, 0], MAT, 1.0))
    with pytest.raises(ValueError):
        pr.set_incident(u[:-1], t[:-1])
    bad = u.copy(); bad[3, 1] = np.nan
    with pytest.raises(capi.MfbError):
        pr.set_incident(bad, t)
    pr.set_incident(u, t)
    with pytest.raises(capi.MfbError):
        pr.build_lse_mechanics_bem_staela(Material(rho=1.0, mu=1.0, nu=0.25, xi=0.0))      # a static assembly with an incident field set
    pr.set_incident(None)
    pr.build_lse_mechanics_bem_staela(Material(rho=1.0, mu=1.0, nu=0.25, xi=0.0))
    pr.close()


def test_fluid_region_incident_field(gpu_ctx, oracle_lib):
    """mfb_harpot3d_set_incident: hp p_inc - gp Un_inc on b (assemble_bem_harpot_equation.f90:471-481), regular / adaptive / singular pairs and free terms,
    mixed conditions with nonzero prescribed values, a symmetry plane; and the no-scattering identity p = p_inc when Un = Un_inc is prescribed on a cavity."""
    from multifebe_b200 import capi
    from multifebe_b200.host import Fluid, FluidModel, plane_wave_fluid, element_incident_fluid, room_bcs
    fl = Fluid(rho=1.2, c=1.0, xi=0.01)
    omega = 2.5
    field = plane_wave_fluid([0.3, 1.0, -0.2], fl, omega, amplitude=0.8 + 0.1j)
    cases = [FluidModel(cube_mesh(3, shape.TRI3), room_bcs(0.7), reversed_parts=(1, 2, 3, 4, 5, 6)),
             FluidModel(cube_mesh(2, shape.QUAD9), {1: (0, 0.2), 2: (1, 0.1j), 3: (1, 0.0), 4: (0, 0.0), 5: (1, 0.3), 6: (1, 0.0)}),
             FluidModel(without_parts(cube_mesh(2, shape.QUAD8), {3}), {1: (0, 0.0), 2: (0, 1.0), 4: (1, 0.0), 5: (1, 0.0), 6: (1, 0.0)}, symmetry=[("y", "antisymmetry")])]
    for md in cases:
        p_inc, un_inc = element_incident_fluid(md, field)
        pr = capi.Problem(gpu_ctx, md); o = oracle_lib.PotOracle(md)
        pr.set_incident(p_inc, un_inc); o.set_incident(p_inc, un_inc)
        A, b = pr.build_lse_mechanics_bem_harpot(omega, fl)
        Ao, bo, _ = o.assemble(omega, fl)
        assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A, (relerr(A, Ao), relerr(b, bo))
        assert relerr(pr.solve_frequency_fluid(omega, fl), np.linalg.solve(Ao, bo)) < TOL_X
        pr.set_incident(None); o.set_incident(None)
        _, b0 = pr.build_lse_mechanics_bem_harpot(omega, fl)
        assert relerr(b0, o.assemble(omega, fl)[1]) < TOL_A
        pr.close()
    md = FluidModel(cube_mesh(3, shape.QUAD9), {q: (1, 0.0) for q in range(1, 7)}, reversed_parts=(1, 2, 3, 4, 5, 6))
    p_inc, un_inc = element_incident_fluid(md, field)
    for e in range(md.n_elem):
        for kn, v in enumerate(md.mesh.conn[e]):
            md.cvalue[v] = un_inc[md.elem_ptr[e] + kn]
    pr = capi.Problem(gpu_ctx, md)
    pr.set_incident(p_inc, un_inc)
    p, _ = md.nodal_solution(pr.solve_frequency_fluid(omega, fl))
    assert relerr(p, np.array([field(x, [1.0, 0, 0])[0] for x in md.node_x])) < 1e-9
    pr.close()


def test_poroelastic_region_incident_field(gpu_ctx, oracle_lib):
    """mfb_harpor3d_set_incident against the oracle (b <= 1e-11, A untouched), mixed conditions and a symmetry plane; clearing restores the plain system."""
    from multifebe_b200 import capi
    from multifebe_b200.host import Poro, PoroModel
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_oracle_poroelastic import column_bcs
    po = Poro(rhof=1.0, rhos=2.2, lam=1.2, mu=1.0, xi=0.03, phi=0.35, rhoa=0.15, R=0.8, Q=0.5, b=0.4)
    rng = np.random.default_rng(5)
    cases = [PoroModel(cube_mesh(2, shape.QUAD9), column_bcs()),
             PoroModel(cube_mesh(3, shape.TRI3), {q: ([1, 1, 1, 1], [0.1, 0, 0.2, 0]) for q in range(1, 7)}, reversed_parts=(1, 2, 3, 4, 5, 6)),
             PoroModel(without_parts(cube_mesh(2, shape.QUAD4), {3}), {k: v for k, v in column_bcs().items() if k != 3}, symmetry=[("y", "symmetry")])]
    for md in cases:
        n_rows = int(md.elem_ptr[-1])
        u_inc = rng.normal(size=(n_rows, 4)) + 1j * rng.normal(size=(n_rows, 4)); t_inc = rng.normal(size=(n_rows, 4)) + 1j * rng.normal(size=(n_rows, 4))
        pr = capi.Problem(gpu_ctx, md); o = oracle_lib.PorOracle(md)
        A0, b0 = pr.build_lse_mechanics_bem_harpor(2.0, po)
        pr.set_incident(u_inc, t_inc); o.set_incident(u_inc, t_inc)
        A, b = pr.build_lse_mechanics_bem_harpor(2.0, po)
        Ao, bo, _ = o.assemble(2.0, po)
        assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A, (relerr(A, Ao), relerr(b, bo))
        assert relerr(pr.solve_frequency_poro(2.0, po), np.linalg.solve(Ao, bo)) < TOL_X
        pr.set_incident(None)
        A1, b1 = pr.build_lse_mechanics_bem_harpor(2.0, po)
        assert relerr(A1, A0) < 1e-13 and np.abs(b1 - b0).max() <= 1e-12 * max(np.abs(b0).max(), 1e-300)
        pr.close()
