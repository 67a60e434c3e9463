"""Local-axes boundary conditions (ctype 2: u.l = U, ctype 3: t.l = T; assemble_bem_harela_equation.f90:107-112 + the host's condition rows of
build_lse_mechanics_harmonic.f90:204-258) in the oracle and the host model (CPU), pinned by the P-wave column on a ROTATED cube: its side walls slide
(u.n = 0, no shear), which in global axes is no longer a 0 / 1 condition per component but is exactly what ctype (2, 3, 3) says in the local axes
(n, t1, t2) of every node.  The solution must be the rotated 1D column (docs/examples/ME-TH-EL-001)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multifebe_b200.host import Material, Model, cube_mesh, column_analytic_u, shape  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def rotation(a, b):
    ca, sa, cb, sb = np.cos(a), np.sin(a), np.cos(b), np.sin(b)
    return np.array([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1]]) @ np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])


def rotated_column(m, etype, R, P=1.0, walls=([2, 3, 3], [0, 0, 0]), **kw):
    mesh = cube_mesh(m, etype)
    mesh.nodes[:] = mesh.nodes @ R.T
    bcs = {1: ([0, 0, 0], [0, 0, 0]), 2: ([1, 1, 1], list(R @ np.array([P, 0, 0])))}
    for p_ in (3, 4, 5, 6):
        bcs[p_] = walls
    return Model(mesh, bcs, **kw)


@pytest.mark.parametrize("etype,m", [(shape.QUAD9, 2), (shape.TRI6, 2), (shape.QUAD4, 4)], ids=["quad9", "tri6", "quad4"])
def test_rotated_column_with_sliding_walls(etype, m):
    R = rotation(0.5, 0.3)
    md = rotated_column(m, etype, R)
    wall = md.node_part >= 3
    assert (md.row_bc[wall] >= 0).all() and (md.col_u[wall] >= 0).all() and (md.col_t[wall] >= 0).all() and md.n_dof == 3 * md.n_node + 3 * wall.sum()
    # local axes are orthonormal, n is the face normal
    for v in np.flatnonzero(wall):
        Q = np.array([md.n_fn[v], md.t1_fn[v], md.t2_fn[v]])
        assert np.abs(Q @ Q.T - np.eye(3)).max() < 1e-12
    mat = Material(1.0, 1.0, 0.25, 0.03)
    omega = 2.0
    A, b, _ = orc.Oracle(md).assemble(omega, mat)
    md.add_condition_rows(A, b)
    u, t = md.nodal_solution(np.linalg.solve(A, b))
    ua = column_analytic_u((md.node_x @ R)[:, 0], omega, mat)
    ue = np.outer(ua, R @ np.array([1.0, 0, 0]))
    tol = 3e-2 if etype == shape.QUAD4 else 2e-3
    assert np.abs(u - ue).max() < tol * np.abs(ue).max()
    # the walls carry only a normal reaction: t.t1 = t.t2 = 0 is a row of the system, so it holds to solver precision
    for v in np.flatnonzero(wall):
        assert abs(t[v] @ md.t1_fn[v]) < 1e-10 * np.abs(t).max() and abs(t[v] @ md.t2_fn[v]) < 1e-10 * np.abs(t).max() and abs(u[v] @ md.n_fn[v]) < 1e-10 * np.abs(u).max()


def test_local_axes_on_an_unrotated_cube_equal_the_global_conditions():
    """With walls normal to the coordinate axes, ctype (2, 3, 3) is the 0 / 1 pattern of cube_bcs(): same displacements (the systems differ -- the local-axes
    one carries both variables as unknowns -- the solutions do not)."""
    from multifebe_b200.host import cube_bcs
    mat = Material(1.0, 1.0, 0.25, 0.03)
    a = rotated_column(2, shape.QUAD9, np.eye(3))
    g = Model(cube_mesh(2, shape.QUAD9), cube_bcs())
    Aa, ba, _ = orc.Oracle(a).assemble(2.5, mat); a.add_condition_rows(Aa, ba)
    Ag, bg, _ = orc.Oracle(g).assemble(2.5, mat)
    ua, ta = a.nodal_solution(np.linalg.solve(Aa, ba)); ug, tg = g.nodal_solution(np.linalg.solve(Ag, bg))
    assert np.abs(ua - ug).max() < 1e-9 * np.abs(ug).max() and np.abs(ta - tg).max() < 1e-8 * np.abs(tg).max()


def test_prescribed_local_values_and_the_reference_vector():
    """Nonzero prescribed local values (a wall pushed inwards by U along its normal) and a user reference vector for t1."""
    R = rotation(0.2, -0.4)
    md = rotated_column(2, shape.QUAD9, R, walls=([2, 3, 3], [0.01, 0.0, 0.0]), local_axes_reference=[0.0, 0.0, 1.0])
    wall = md.node_part >= 3
    for v in np.flatnonzero(wall):
        assert abs(md.t2_fn[v] @ np.array([0.0, 0.0, 1.0])) < 1e-12       # t2 = n x reference is normal to the reference vector
    mat = Material(1.0, 1.0, 0.25, 0.03)
    A, b, _ = orc.Oracle(md).assemble(1.5, mat); md.add_condition_rows(A, b)
    u, t = md.nodal_solution(np.linalg.solve(A, b))
    for v in np.flatnonzero(wall):
        assert abs(u[v] @ md.n_fn[v] - 0.01) < 1e-10
