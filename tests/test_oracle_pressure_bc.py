"""Boundary condition ctype = 10 ("normal pressure known": t_k = p n_fn(k), assemble_bem_harela_equation.f90:97-106) in the oracle (CPU), pinned by
an exact solution: a cube under the same normal pull p on all its faces is in the hydrostatic state sigma = p I, u = p (1 - 2 nu) / (2 mu (1 + nu)) x.
The octant x, y, z >= 0 with three symmetry planes holds the rigid-body modes, so the test also runs the image loop and n_fn of nodes on planes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multifebe_b200.host import Material, Model, cube_mesh, without_parts, shape  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def octant(etype, m, p=0.7):
    mesh = without_parts(cube_mesh(m, etype), {1, 3, 5})
    bcs = {q: ([10, 10, 10], [p, p, p]) for q in (2, 4, 6)}
    return Model(mesh, bcs, symmetry=[("x", "symmetry"), ("y", "symmetry"), ("z", "symmetry")])


def test_hydrostatic_octant_static():
    p = 0.7
    mat = Material(rho=1.0, mu=1.3, nu=0.2, xi=0.0)
    for etype, m in [(shape.QUAD4, 3), (shape.TRI6, 2)]:
        md = octant(etype, m, p)
        assert (md.ctype == 10).all() and (md.col_u >= 0).all()
        # the nodal normal of a face node is the face normal; on the planes' rims too (the in-plane components of the mirrored sum cancel)
        for v in range(md.n_node):
            part = int(md.node_part[v]); ax = {2: 0, 4: 1, 6: 2}[part]
            want = np.zeros(3); want[ax] = 1.0
            assert np.abs(md.n_fn[v] - want).max() < 1e-12
        A, b, _ = orc.Oracle(md).assemble_static(mat)
        x, _, _ = orc.lu_solve_real(A, b)
        u, t = md.nodal_solution(x.astype(np.complex128))
        eps = p * (1.0 - 2.0 * mat.nu_r) / (2.0 * mat.mu_r * (1.0 + mat.nu_r))
        assert np.abs(u.real - eps * md.node_x).max() < 2e-5 * eps
        assert np.abs(t.real - p * md.n_fn).max() < 1e-14


def test_pressure_condition_equals_the_traction_condition_it_stands_for():
    """ctype 10 with pressure p is ctype 1 with the traction p n_fn: same system (harmonic, reversed boundary included: an outward pressure on a cavity)."""
    mat = Material(rho=1.0, mu=1.0, nu=0.25, xi=0.02)
    for rev in ((), (1, 2, 3, 4, 5, 6)):
        mesh = cube_mesh(2, shape.QUAD9)
        a = Model(mesh, {q: ([10, 10, 10], [0.3 + 0.1j] * 3) for q in range(1, 7)}, reversed_parts=rev)
        b = Model(mesh, {q: ([1, 1, 1], [0, 0, 0]) for q in range(1, 7)}, reversed_parts=rev)
        sgn = -1.0 if rev else 1.0
        b.cvalue[:] = sgn * (0.3 + 0.1j) * a.n_fn
        Aa, ba, _ = orc.Oracle(a).assemble(2.0, mat)
        Ab, bb, _ = orc.Oracle(b).assemble(2.0, mat)
        assert np.abs(Aa - Ab).max() < 1e-14 and np.abs(ba - bb).max() < 1e-13 * np.abs(bb).max()
