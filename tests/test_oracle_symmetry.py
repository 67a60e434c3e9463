"""Symmetry planes in the oracle (CPU): a half / quarter model with a [symmetry planes] entry against the FULL model it stands for.

The restatement of the image loop (build_lse_mechanics_bem_harela.f90:1098-1107 with fbem_symmetry_multipliers,
lib/fbem/src/symmetry.f90:60-171) and of the mirrored free-term fans (:496-555) is pinned here by the physics, not by a second copy of
the same formulas: the full model is meshed explicitly (host.mirror_mesh: mirrored nodes, reversed connectivity, mirrored loads) and
solved WITHOUT any symmetry code; the reduced model must reproduce its solution on the nodes they share.  With nodal collocation on
the plane nodes the two discrete systems are the same equations, so the solutions differ only through the quadrature rules the
reference picks for an image (it tests the image against the bounding ball of the ROOT element) -- within the integration tolerance
qsi_relative_error = 1e-6, far below what a wrong sign or orientation would produce (O(1)).
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multifebe_b200.host import Material, Model, cube_mesh, mirror_mesh, without_parts, symmetry_planes  # noqa: E402
from multifebe_b200.host.shape import TRI3, TRI6, QUAD4, QUAD9  # noqa: E402
from oracle import oracle as orc  # noqa: E402

AX = {"x": 0, "y": 1, "z": 2}


def closed_cube(m, etype):
    """cube_mesh with the coincident rim nodes of its faces merged and all faces in ONE part: a closed surface without rim nodes
    (every node collocated nodally, edges and corners through their Mantic free terms)."""
    c = cube_mesh(m, etype)
    key = np.round(c.nodes * 4096).astype(np.int64)
    _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    from multifebe_b200.host import Mesh
    return Mesh(c.nodes[first], c.etype, c.part, [[int(inv[v]) for v in cc] for cc in c.conn])


def reduced_and_full(m, etype, planes, traction):
    """planes: [(axis name, kind)]; traction(x) -> prescribed traction vector, with the parity of every plane.  The closed cube [0,1]^3
    without its faces in the planes is the reduced model (one part, open along the planes, nodal collocation everywhere); the full
    model is its successive mirror image, a closed surface.  Returns (reduced Model, full Model)."""
    cube = closed_cube(m, etype)
    drop = {"x": 1, "y": 3, "z": 5}
    red_mesh = without_parts(cube, {drop[a] for a, _ in planes})
    red_mesh.part[:] = 1
    bcs = {1: ([1, 1, 1], [0, 0, 0])}
    red = Model(red_mesh, bcs, symmetry=planes, nodal_on_symplanes=True)
    assert not red.in_boundary.any() and red.n_colloc == red.n_node
    full_mesh = red_mesh
    for a, _ in planes:
        full_mesh, _ = mirror_mesh(full_mesh, AX[a])
    full = Model(full_mesh, bcs)
    assert not full.in_boundary.any()
    for mdl in (red, full):
        mdl.cvalue = np.ascontiguousarray([traction(x) for x in mdl.node_x], dtype=np.complex128)
    return red, full


def solve(model, omega, mat):
    A, b, st = orc.Oracle(model).assemble(omega, mat)
    x, _, _ = orc.lu_solve(A, b)
    return model.nodal_solution(x), st


# Prescribed tractions everywhere (no prescribed displacement on a plane node: with nodal collocation ON the plane the equation of the
# dof that the parity annihilates there is c u = 0 -- every integral cancels against its image -- which determines u but not an unknown
# traction; the reference avoids that by default: rim nodes get non-nodal collocation points, off the plane).
# parity of a field under a plane normal to axis a: symmetry v(Mx) = M v(x), antisymmetry v(Mx) = -M v(x), M = reflection of axis a
CASES = [
    ("x-sym", [("x", "symmetry")], lambda x: (x[0], 0.3 + x[1], 0.5 * x[2])),
    ("x-anti", [("x", "antisymmetry")], lambda x: (1.0 + x[1], x[0], x[0] * x[2])),
    ("z-sym", [("z", "symmetry")], lambda x: (1.0 + x[0], x[1], x[2])),
    # two planes (quarter model); the nodes of the edge x = y = 0 lie in both: fourfold fans
    ("xy-sym-sym", [("x", "symmetry"), ("y", "symmetry")], lambda x: (x[0], x[1], 0.5 + x[2])),
    ("xy-sym-anti", [("x", "symmetry"), ("y", "antisymmetry")], lambda x: (x[0] * x[1], 1.0 + x[2], x[1])),
    ("yz-anti-anti", [("y", "antisymmetry"), ("z", "antisymmetry")], lambda x: (x[1] * x[2], x[2], x[1])),
]


@pytest.mark.parametrize("name,planes,traction", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("etype", [TRI3, QUAD9], ids=["tri3", "quad9"])
def test_reduced_model_reproduces_the_full_model(name, planes, traction, etype):
    m = 3 if etype == TRI3 else 2
    red, full = reduced_and_full(m, etype, planes, traction)
    assert full.n_dof > red.n_dof
    mat = Material(rho=1.0, mu=1.0, nu=0.25, xi=0.02)
    omega = 2.5
    (ur, _), st = solve(red, omega, mat)
    (uf, _), _ = solve(full, omega, mat)
    n = red.n_node   # the reduced model's nodes are the first nodes of the full mesh (mirror_mesh appends)
    su = np.abs(uf).max()
    assert su > 1e-3
    du = np.abs(ur - uf[:n]).max() / su
    assert du < 2e-5, (name, du)
    # every image was integrated: 2 or 4 times the pairs of the root elements
    n_pairs = sum(st["pairs_regular"].values()) + st["pairs_adaptive"] + st["pairs_singular"]
    assert n_pairs == red.n_elem * red.n_colloc * (1 << len(planes))


def test_three_planes_octant_against_the_full_cube():
    """Octant of the cube [-1,1]^3 under a symmetric load: eight images per element, the nodes of the three edges along the axes lie
    in two planes (fourfold fans); no node of the octant's three outer faces lies in all three."""
    planes = [("x", "symmetry"), ("y", "symmetry"), ("z", "symmetry")]
    red, full = reduced_and_full(2, QUAD4, planes, lambda x: (x[0], x[1] * (1 + x[2] ** 2), x[2]))
    mat = Material(rho=1.0, mu=1.0, nu=0.3, xi=0.05)
    (ur, _), st = solve(red, 1.7, mat)
    (uf, _), _ = solve(full, 1.7, mat)
    n = red.n_node
    assert np.abs(ur - uf[:n]).max() / np.abs(uf).max() < 2e-5
    assert sum(st["pairs_regular"].values()) + st["pairs_adaptive"] + st["pairs_singular"] == 8 * red.n_elem * red.n_colloc


def test_wrong_parity_is_detected():
    """The check above has teeth: the same reduced model with the plane declared antisymmetric instead of symmetric is far from the full model."""
    red, full = reduced_and_full(3, TRI3, [("x", "symmetry")], CASES[0][2])
    red.symplane_eid, red.symplane_t = symmetry_planes([("x", "antisymmetry")])
    mat = Material(rho=1.0, mu=1.0, nu=0.25, xi=0.02)
    (ur, _), _ = solve(red, 2.5, mat)
    (uf, _), _ = solve(full, 2.5, mat)
    assert np.abs(ur - uf[:red.n_node]).max() / np.abs(uf).max() > 1e-2


def test_static_reduced_model_reproduces_the_full_model():
    """Static analysis: a clamped patch away from the plane removes the rigid-body modes (prescribed displacements on the face x = 1 and its
    mirror image; mixed conditions per node)."""
    red, full = reduced_and_full(3, TRI6, [("x", "symmetry")], CASES[0][2])
    for mdl in (red, full):
        clamp = np.abs(np.abs(mdl.node_x[:, 0]) - 1.0) < 1e-9
        mdl.ctype[clamp] = 0; mdl.cvalue[clamp] = 0.0
        # the unknown of a clamped dof is its traction: same column, the other kind
        mdl.col_t[clamp] = mdl.col_u[clamp]; mdl.col_u[clamp] = -1
    mat = Material(rho=1.0, mu=1.0, nu=0.25, xi=0.0)
    out = []
    for mdl in (red, full):
        A, b, _ = orc.Oracle(mdl).assemble_static(mat)
        x, _, _ = orc.lu_solve_real(A, b)
        out.append(mdl.nodal_solution(x.astype(np.complex128)))
    (ur, tr), (uf, tf) = out
    n = red.n_node
    assert np.abs(ur - uf[:n]).max() / np.abs(uf).max() < 2e-5
    assert np.abs(tr - tf[:n]).max() / np.abs(tf).max() < 2e-5


FREE = ([1, 1, 1], [0, 0, 0])


def test_default_formulation_puts_mca_points_on_plane_nodes():
    """Without nodal_on_symplanes the nodes of the open edge in the plane are rim nodes: non-nodal collocation (the reference's default,
    assign_default_bem_formulation.f90:85-92); the symmetric solution is still reproduced to discretisation accuracy."""
    cube = cube_mesh(3, QUAD4)
    mesh = without_parts(cube, {1})
    bcs = {2: ([0, 0, 0], [0, 0, 0]), 3: FREE, 4: ([1, 1, 1], [0, 0.3, 0]), 5: FREE, 6: ([1, 1, 1], [0, 0, 0.5])}
    a = Model(mesh, bcs, symmetry=[("x", "symmetry")])
    b = Model(mesh, bcs, symmetry=[("x", "symmetry")], nodal_on_symplanes=True)
    on_plane = np.abs(mesh.nodes[:, 0]) <= 1e-9
    assert a.in_boundary[on_plane].all() and a.n_colloc > b.n_colloc
    mat = Material(rho=1.0, mu=1.0, nu=0.25, xi=0.02)
    (ua, _), _ = solve(a, 2.5, mat)
    (ub, _), _ = solve(b, 2.5, mat)
    assert np.abs(ua - ub).max() / np.abs(ub).max() < 0.05


# ---- fluid (scalar) and poroelastic regions: symconf_s on the scalar variables (build_lse_mechanics_bem_harpot.f90:947-948, _harpor.f90:971-975) ----
def _generic_reduced_and_full(m, etype, planes, make_model, values):
    """As reduced_and_full, for any region type: make_model(mesh, symmetry=..., nodal_on_symplanes=...) builds the Model with every dof of the
    secondary kind prescribed; values(x) are the prescribed values at a point (with the parity of every plane)."""
    from multifebe_b200.host import Mesh  # noqa: F401
    cube = closed_cube(m, etype)
    drop = {"x": 1, "y": 3, "z": 5}
    red_mesh = without_parts(cube, {drop[a] for a, _ in planes})
    red_mesh.part[:] = 1
    red = make_model(red_mesh, symmetry=planes, nodal_on_symplanes=True)
    full_mesh = red_mesh
    for a, _ in planes:
        full_mesh, _ = mirror_mesh(full_mesh, AX[a])
    full = make_model(full_mesh)
    for mdl in (red, full):
        assert not mdl.in_boundary.any()
        mdl.cvalue = np.ascontiguousarray([np.atleast_1d(values(x)) for x in mdl.node_x], dtype=np.complex128)
    return red, full


FLUID_CASES = [
    # parity of a scalar under a plane: symmetry q(Mx) = q(x), antisymmetry q(Mx) = -q(x)
    ("x-sym", [("x", "symmetry")], lambda x: 1.0 + x[1] + x[0] ** 2),
    ("x-anti", [("x", "antisymmetry")], lambda x: x[0] * (1.0 + x[2])),
    ("xy-sym-anti", [("x", "symmetry"), ("y", "antisymmetry")], lambda x: x[1] * (1.0 + x[0] ** 2 + x[2])),
    ("xyz-anti-sym-sym", [("x", "antisymmetry"), ("y", "symmetry"), ("z", "symmetry")], lambda x: x[0] * (1.0 + x[1] ** 2)),
]


@pytest.mark.parametrize("name,planes,flux", FLUID_CASES, ids=[c[0] for c in FLUID_CASES])
def test_fluid_region_reduced_model_reproduces_the_full_model(name, planes, flux):
    from multifebe_b200.host import Fluid, FluidModel
    fl = Fluid(rho=1.0, c=1.0, xi=0.03)
    for etype, m in ((TRI3, 3), (QUAD9, 2)):
        red, full = _generic_reduced_and_full(m, etype, planes, lambda mesh, **kw: FluidModel(mesh, {1: (1, 0.0)}, **kw), flux)
        out = []
        for mdl in (red, full):
            A, b, st = orc.PotOracle(mdl).assemble(2.5, fl)
            out.append(mdl.nodal_solution(np.linalg.solve(A, b))[0])
        pr, pf = out
        assert np.abs(pf).max() > 1e-3
        assert np.abs(pr - pf[:red.n_node]).max() / np.abs(pf).max() < 2e-5, name


def test_room_tutorial_as_a_quarter_model():
    """The rigid side walls of ME-TH-AC-001 are symmetry planes of the pressure: the quarter room (walls y = 0 and z = 0 removed, planes declared) has the
    same standing wave p = P sin(k x) / sin(k L)."""
    from multifebe_b200.host import Fluid, FluidModel, room_analytic
    fl = Fluid(rho=1.25, c=343.0)
    mesh = without_parts(cube_mesh(3, QUAD9), {3, 5})
    md = FluidModel(mesh, {1: (0, 0.0), 2: (0, 1.0), 4: (1, 0.0), 6: (1, 0.0)}, symmetry=[("y", "symmetry"), ("z", "symmetry")])
    omega = 2 * np.pi * 100.0
    A, b, _ = orc.PotOracle(md).assemble(omega, fl)
    p, _ = md.nodal_solution(np.linalg.solve(A, b))
    p_ex, _ = room_analytic(md.node_x[:, 0], omega, fl)
    assert np.abs(p - p_ex).max() < 2e-3 * np.abs(p_ex).max()


def test_biot_column_as_a_quarter_model():
    """Sliding impermeable side walls are symmetry planes of a poroelastic column: the quarter column against the exact two-wave solution."""
    from multifebe_b200.host import Poro, PoroModel
    from test_oracle_poroelastic import biot_column, column_bcs
    po = Poro(rhof=1.0, rhos=2.2, lam=1.2, mu=1.0, xi=0.02, phi=0.35, rhoa=0.15, R=0.8, Q=0.5, b=0.6)
    omega = 2.0
    mesh = without_parts(cube_mesh(2, QUAD9), {3, 5})
    bcs = {k: v for k, v in column_bcs().items() if k in (1, 2, 4, 6)}
    md = PoroModel(mesh, bcs, symmetry=[("y", "symmetry"), ("z", "symmetry")])
    A, bb, _ = orc.PorOracle(md).assemble(omega, po)
    prim, sec = md.nodal_solution(np.linalg.solve(A, bb))
    field, _ = biot_column(omega, po)
    ua, Ua, sa, ta = field(md.node_x[:, 0])
    assert np.abs(prim[:, 1] - ua).max() < 4e-3 * np.abs(ua).max() and np.abs(prim[:, 2:]).max() < 4e-3 * np.abs(ua).max()
    side = md.node_part >= 3
    assert np.abs(prim[side, 0] - ta[side]).max() < 4e-3 * np.abs(ta).max()


def test_poroelastic_region_reduced_model_reproduces_the_full_model():
    """Mixed parities on a poroelastic region: Un (scalar) and t_k (vector) prescribed everywhere with the parity of each plane."""
    from multifebe_b200.host import Poro, PoroModel
    po = Poro(rhof=1.0, rhos=2.2, lam=1.2, mu=1.0, xi=0.03, phi=0.35, rhoa=0.15, R=0.8, Q=0.5, b=0.4)
    cases = [([("x", "symmetry")], lambda x: (0.3 + x[1], x[0], 0.3 + x[1], 0.5 * x[2])),
             ([("x", "antisymmetry"), ("z", "symmetry")], lambda x: (x[0], 1.0 + x[1], x[0], x[0] * x[2]))]
    for planes, vals in cases:
        red, full = _generic_reduced_and_full(2, QUAD9, planes, lambda mesh, **kw: PoroModel(mesh, {1: ([1, 1, 1, 1], [0, 0, 0, 0])}, **kw), vals)
        out = []
        for mdl in (red, full):
            A, b, _ = orc.PorOracle(mdl).assemble(2.0, po)
            out.append(mdl.nodal_solution(np.linalg.solve(A, b))[0])
        pr, pf = out
        for k in range(4):
            sc = np.abs(pf[:, k]).max()
            assert sc > 1e-4 and np.abs(pr[:, k] - pf[:red.n_node, k]).max() / sc < 5e-5, (planes, k)


def test_plane_nodes_are_collapsed_into_the_plane():
    """collapse_nodal_pos (default T, src/read_settings.f90:168-177; fbem_transformation_collapse_nodal_positions): a mesh whose plane nodes sit 3e-8 off the
    plane (inside geometric_tolerance) is snapped, so that the images touch their root elements in the same points."""
    mesh = without_parts(cube_mesh(2, QUAD4), {1})
    on = np.abs(mesh.nodes[:, 0]) < 1e-9
    mesh.nodes[on, 0] = 3e-8
    bcs = {p: ([1, 1, 1], [0, 0, 0]) for p in set(int(q) for q in mesh.part)}
    md = Model(mesh, bcs, symmetry=[("x", "symmetry")])
    assert np.abs(md.node_x[on, 0]).max() == 0.0
    mesh2 = without_parts(cube_mesh(2, QUAD4), {1}); mesh2.nodes[on, 0] = 3e-8
    assert np.abs(Model(mesh2, bcs, symmetry=[("x", "symmetry")], collapse_nodal_pos=False).node_x[on, 0]).max() == 3e-8
