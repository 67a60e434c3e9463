"""Incident wave field in the oracle (CPU): the term b += hp u_inc - gp t_inc of assemble_bem_harela_equation.f90:651-666, pinned by the
integral identities a plane wave satisfies -- not by a second copy of the formula.

For a field that is regular inside a closed surface, Somigliana's identity gives
  * seen from the INTERIOR region (normals outward):          c u + int t* u - int u* t = 0      ->  H u_inc - G t_inc = 0
  * seen from the EXTERIOR region (normals into the cavity):  c_e u + int t*(n) u - int u* t(n) = u  ->  H u_inc - G t_inc = u_inc(x_i)
up to the discretisation error of the mesh (interpolation of exp(-i k d.x) by the shape functions): a wrong sign of either term is O(1).
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from multifebe_b200.host import Material, Model, cube_mesh, plane_wave, element_incident, shape  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from test_oracle_symmetry import closed_cube  # noqa: E402

MAT = Material(rho=1.0, mu=1.0, nu=0.25, xi=0.02)
T_KNOWN = {p: ([1, 1, 1], [0, 0, 0]) for p in range(1, 7)}
WAVES = [("P", [1.0, 0.5, 0.2], None), ("S", [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]), ("S", [1.0, 1.0, 0.0], [0.0, 0.0, 1.0])]


def incident_rhs(md, omega, field):
    o = orc.Oracle(md)
    u, t = element_incident(md, field)
    o.set_incident(u, t)
    A, b, _ = o.assemble(omega, MAT)
    o.set_incident(None)
    A0, b0, _ = o.assemble(omega, MAT)
    assert np.abs(A - A0).max() < 1e-14 and not b0.any()      # the field touches b alone (A up to the summation order of the OpenMP threads); cleared, b is zero again
    return b


@pytest.mark.parametrize("kind,d,pol", WAVES, ids=["P", "S-z", "S-xy"])
@pytest.mark.parametrize("etype,m", [(shape.TRI3, 4), (shape.QUAD9, 2)], ids=["tri3", "quad9"])
def test_plane_wave_identities_interior_and_exterior(kind, d, pol, etype, m):
    omega = 2.0
    field = plane_wave(kind, d, MAT, omega, polarisation=pol)
    mesh = closed_cube(m, etype)
    mesh.part[:] = 1                                               # one boundary: the merged nodes of the edges belong to it
    inner = Model(mesh, T_KNOWN)                                   # the cube itself
    outer = Model(mesh, T_KNOWN, reversed_parts=(1,))              # the full space around a cubic cavity
    assert inner.n_colloc == inner.n_node                          # nodal collocation everywhere (closed surface)
    u_nodes = np.array([field(x, [1.0, 0, 0])[0] for x in inner.node_x])
    want = np.zeros(inner.n_dof, dtype=np.complex128)
    for v in range(inner.n_node):
        want[inner.row[v]] = u_nodes[v]
    tol = 4e-2 if etype == shape.TRI3 else 1e-2
    b_in = incident_rhs(inner, omega, field)
    assert np.abs(b_in).max() < tol * np.abs(want).max(), np.abs(b_in).max()
    b_out = incident_rhs(outer, omega, field)
    assert np.abs(b_out - want).max() < tol * np.abs(want).max(), np.abs(b_out - want).max()


def test_total_field_equal_to_the_incident_field_is_reproduced_exactly():
    """Exterior region, prescribed tractions equal to the incident tractions: nothing is scattered and u = u_inc -- an algebraic identity of
    the assembled system (A = H, b = G t + H u_inc - G t_inc), exact to the precision of the solver whatever the mesh."""
    omega = 3.0
    field = plane_wave("S", [0.3, -0.4, 1.0], MAT, omega, polarisation=[1.0, 1.0, 0.0], amplitude=0.7)
    md = Model(cube_mesh(3, shape.QUAD8), T_KNOWN, reversed_parts=(1, 2, 3, 4, 5, 6))     # unshared rims: one normal per node
    u, t = element_incident(md, field)
    for e in range(md.n_elem):
        for kn, v in enumerate(md.mesh.conn[e]):
            md.cvalue[v] = t[md.elem_ptr[e] + kn]
    o = orc.Oracle(md)
    o.set_incident(u, t)
    A, b, _ = o.assemble(omega, MAT)
    x, _, _ = orc.lu_solve(A, b)
    un, _ = md.nodal_solution(x)
    want = np.array([field(xv, [1.0, 0, 0])[0] for xv in md.node_x])
    assert np.abs(un - want).max() < 1e-10 * np.abs(want).max()


@pytest.mark.parametrize("etype,m", [(shape.TRI3, 4), (shape.QUAD9, 2)], ids=["tri3", "quad9"])
def test_fluid_plane_wave_identities_interior_and_exterior(etype, m):
    """The same identities for a fluid region (assemble_bem_harpot_equation.f90:471-481): H p_inc - G Un_inc = 0 seen from inside, = p_inc seen from outside."""
    from multifebe_b200.host import Fluid, FluidModel, plane_wave_fluid, element_incident_fluid
    fl = Fluid(rho=1.2, c=1.0, xi=0.01)
    omega = 2.0
    field = plane_wave_fluid([1.0, 0.5, 0.2], fl, omega, amplitude=0.8 + 0.1j)
    mesh = closed_cube(m, etype); mesh.part[:] = 1
    tol = 4e-2 if etype == shape.TRI3 else 1e-2
    for rev, expect_p in (((), False), ((1,), True)):
        md = FluidModel(mesh, {1: (1, 0.0)}, reversed_parts=rev)
        o = orc.PotOracle(md)
        p_inc, un_inc = element_incident_fluid(md, field)
        o.set_incident(p_inc, un_inc)
        A, b, _ = o.assemble(omega, fl)
        want = np.zeros(md.n_dof, dtype=np.complex128)
        pn = np.array([field(x, [1.0, 0, 0])[0] for x in md.node_x])
        if expect_p:
            want[md.row[:, 0]] = pn
        assert np.abs(b - want).max() < tol * np.abs(pn).max(), (rev, np.abs(b - want).max())
        o.set_incident(None)
        assert not o.assemble(omega, fl)[1].any()


def test_poroelastic_incident_term_no_scattering_identity():
    """assemble_bem_harpor_equation.f90:1277-1289 in the oracle: with every secondary variable (Un, t_k) prescribed equal to the incident one, the assembled
    system is A = H, b = G s + H p_inc - G s_inc = H p_inc, so the primary variables (tau, u_k) come out equal to the incident ones -- an algebraic identity
    that holds for ANY incident arrays and fails if the signs of the two terms are not those of the H and G columns."""
    from multifebe_b200.host import Poro, PoroModel
    po = Poro(rhof=1.0, rhos=2.2, lam=1.2, mu=1.0, xi=0.03, phi=0.35, rhoa=0.15, R=0.8, Q=0.5, b=0.4)
    md = PoroModel(cube_mesh(2, shape.QUAD9), {q: ([1, 1, 1, 1], [0, 0, 0, 0]) for q in range(1, 7)}, reversed_parts=(1, 2, 3, 4, 5, 6))
    rng = np.random.default_rng(3)
    prim_node = rng.normal(size=(md.n_node, 4)) + 1j * rng.normal(size=(md.n_node, 4))       # (tau, u_k)_inc per node
    sec_node = rng.normal(size=(md.n_node, 4)) + 1j * rng.normal(size=(md.n_node, 4))        # (Un, t_k)_inc per node (unshared rims: one value per node)
    n_rows = int(md.elem_ptr[-1])
    u_inc = np.zeros((n_rows, 4), dtype=np.complex128); t_inc = np.zeros((n_rows, 4), dtype=np.complex128)
    for e in range(md.n_elem):
        for kn, v in enumerate(md.mesh.conn[e]):
            u_inc[md.elem_ptr[e] + kn] = prim_node[v]; t_inc[md.elem_ptr[e] + kn] = sec_node[v]
    md.cvalue[:] = sec_node
    o = orc.PorOracle(md)
    o.set_incident(u_inc, t_inc)
    A, b, _ = o.assemble(2.0, po)
    prim, _ = md.nodal_solution(np.linalg.solve(A, b))
    assert np.abs(prim - prim_node).max() < 1e-9 * np.abs(prim_node).max()
