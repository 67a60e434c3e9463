"""Several BE regions coupled through be-be interfaces: host numbering (multifebe_b200.host.MultiRegionModel) and the
multi-region oracle driver (oracle/multiregion.py) around the pinned single-region integrals.  The reference holds no numeric vectors
for this path; the pins are (1) a homogeneous body split in two regions must reproduce the one-region solution, (2) exact 1D
two-layer solutions: solid-solid, fluid-fluid and fluid-solid columns."""
import numpy as np
import pytest

from multifebe_b200.host import (Material, Fluid, Model, FluidModel, MultiRegionModel, Region, SOLID, FLUID, two_box_mesh, cube_mesh, cube_bcs,
                                 room_bcs, column_analytic_u, shape)
from oracle import oracle as orc
from oracle.multiregion import MultiRegionOracle

BPART = {b: b for b in (1, 2, 3, 4, 5, 6, 7, 13, 14, 15, 16)}
LAT1, LAT2 = (3, 4, 5, 6), (13, 14, 15, 16)


def solid_bcs(P=1.0):
    """x=0 clamped, x=L normal traction P, lateral faces: zero normal displacement and zero shear."""
    bcs = {1: ([0, 0, 0], [0, 0, 0]), 2: ([1, 1, 1], [P, 0, 0])}
    for a, b in zip(LAT1, LAT2):
        ct = [1, 0, 1] if a in (3, 4) else [1, 1, 0]
        bcs[a] = (ct, [0, 0, 0]); bcs[b] = (ct, [0, 0, 0])
    return bcs


def fluid_bcs(P=1.0):
    bcs = {1: (0, 0.0), 2: (0, P)}
    for a, b in zip(LAT1, LAT2):
        bcs[a] = (1, 0.0); bcs[b] = (1, 0.0)
    return bcs


def layered_1d(omega, layers, left, right):
    """Exact 1D two-layer column on [0, xs] U [xs, 1]: field f_i = a_i e^{-i k_i x} + b_i e^{i k_i x} with flux-like quantity
    s_i = Z_i f_i' continuous at xs together with f.  layers = [(k1, Z1), (k2, Z2), xs]; left = ('f'|'s', value), right likewise."""
    (k1, Z1), (k2, Z2), xs = layers
    M = np.zeros((4, 4), dtype=complex); r = np.zeros(4, dtype=complex)

    def f(k, x): return [np.exp(-1j * k * x), np.exp(1j * k * x)]
    def s(k, Z, x): return [-1j * k * Z * np.exp(-1j * k * x), 1j * k * Z * np.exp(1j * k * x)]
    M[0, :2] = f(k1, 0.0) if left[0] == "f" else s(k1, Z1, 0.0); r[0] = left[1]
    M[1, 2:] = f(k2, 1.0) if right[0] == "f" else s(k2, Z2, 1.0); r[1] = right[1]
    M[2, :2] = f(k1, xs); M[2, 2:] = [-v for v in f(k2, xs)]
    M[3, :2] = s(k1, Z1, xs); M[3, 2:] = [-v for v in s(k2, Z2, xs)]
    c = np.linalg.solve(M, r)

    def field(x):
        x = np.asarray(x, dtype=float)
        k = np.where(x <= xs, k1, k2); a = np.where(x <= xs, c[0], c[2]); b = np.where(x <= xs, c[1], c[3]); Z = np.where(x <= xs, Z1, Z2)
        fv = a * np.exp(-1j * k * x) + b * np.exp(1j * k * x)
        sv = Z * (-1j * k * a * np.exp(-1j * k * x) + 1j * k * b * np.exp(1j * k * x))
        return fv, sv
    return field


def test_numbering_is_square_and_complete():
    mesh = two_box_mesh(2, shape.QUAD4)
    mat = Material()
    mrm = MultiRegionModel(mesh, [Region(SOLID, mat, [1, 3, 4, 5, 6, 7]), Region(SOLID, mat, [-7, 2, 13, 14, 15, 16])], BPART, solid_bcs())
    n_if = len(set(int(v) for e in mrm.elems_of_boundary[7] for v in mesh.conn[e]))
    assert mrm.n_dof == 3 * (mrm.n_node - n_if) + 6 * n_if                    # interface nodes: two sets of three equations, u1 and t1 unknown
    rows = sorted(r for lst in mrm.row.values() for r in lst)
    assert rows == list(range(mrm.n_dof)) and sorted(mrm.col.values()) == list(range(mrm.n_dof))
    v1, v2 = mrm.views
    assert not v1.elem_reversed.any() and v2.elem_reversed[:len(mrm.elems_of_boundary[7])].all() and not v2.elem_reversed[len(mrm.elems_of_boundary[7]):].any()
    # both regions collocate on the interface: equation index 1 for region 1, 2 for region 2
    if_nodes = set(int(v) for e in mrm.elems_of_boundary[7] for v in mesh.conn[e])
    assert set(v1.colloc_eq[[int(n) in if_nodes for n in v1.colloc_node]]) == {1} and set(v2.colloc_eq[[int(n) in if_nodes for n in v2.colloc_node]]) == {2}
    with pytest.raises(ValueError):
        MultiRegionModel(mesh, [Region(SOLID, mat, [1, 3, 4, 5, 6, 7]), Region(SOLID, mat, [7, 2, 13, 14, 15, 16])], BPART, solid_bcs())


@pytest.mark.parametrize("et,m", [(shape.QUAD4, 2), (shape.TRI3, 2), (shape.QUAD9, 1)])
def test_split_homogeneous_solid_equals_the_one_region_model(et, m):
    """Same material on both sides of the cut: displacements and tractions on the outer faces follow the homogeneous column, the
    interface carries u(xs) and t = sigma n; quadrature error only (both models integrate different meshes)."""
    mat = Material(1.0, 1.0, 0.25, 0.03)
    omega = 2.0
    mrm = MultiRegionModel(two_box_mesh(m, et), [Region(SOLID, mat, [1, 3, 4, 5, 6, 7]), Region(SOLID, mat, [-7, 2, 13, 14, 15, 16])], BPART, solid_bcs())
    A, b = MultiRegionOracle(mrm).assemble(omega)
    x = np.linalg.solve(A, b)
    tol = 3e-2 if et != shape.QUAD9 else 5e-3                                   # discretisation error of these very coarse meshes
    for kr in (0, 1):
        u, t = mrm.nodal_solution(x, kr)
        ok = ~np.isnan(u[:, 0])
        ua = column_analytic_u(mrm.node_x[ok, 0], omega, mat)
        assert np.abs(u[ok, 0] - ua).max() < tol * np.abs(ua).max()
        assert np.abs(u[ok, 1:]).max() < tol * np.abs(ua).max()
    # interface: traction of region 1 = sigma_xx (normal +x); region 2 sees the opposite sign
    k = omega / mat.c1
    sig = lambda x_: (np.exp(-1j * k * x_) + np.exp(1j * k * x_)) / (np.exp(-1j * k) + np.exp(1j * k))      # sigma_xx/P of the column
    u1, t1 = mrm.nodal_solution(x, 0); u2, t2 = mrm.nodal_solution(x, 1)
    ifn = sorted(set(int(v) for e in mrm.elems_of_boundary[7] for v in mrm.mesh.conn[e]))
    assert np.abs(t1[ifn, 0] - sig(0.5)).max() < tol and np.abs(t2[ifn, 0] + sig(0.5)).max() < tol
    assert np.array_equal(u1[ifn], u2[ifn])


def test_two_layer_solid_column():
    m1, m2 = Material(1.0, 1.0, 0.25, 0.02), Material(2.0, 3.0, 0.3, 0.05)
    omega = 2.5
    mrm = MultiRegionModel(two_box_mesh(2, shape.QUAD9, xs=0.4), [Region(SOLID, m1, [1, 3, 4, 5, 6, 7]), Region(SOLID, m2, [-7, 2, 13, 14, 15, 16])], BPART, solid_bcs())
    A, b = MultiRegionOracle(mrm).assemble(omega)
    x = np.linalg.solve(A, b)
    field = layered_1d(omega, [(omega / m1.c1, m1.lam + 2 * m1.mu), (omega / m2.c1, m2.lam + 2 * m2.mu), 0.4], ("f", 0.0), ("s", 1.0))
    for kr in (0, 1):
        u, t = mrm.nodal_solution(x, kr)
        ok = ~np.isnan(u[:, 0])
        ua, _ = field(np.clip(mrm.node_x[ok, 0], 0, 0.4) if kr == 0 else np.clip(mrm.node_x[ok, 0], 0.4 + 1e-12, 1))
        assert np.abs(u[ok, 0] - ua).max() < 3e-3 * np.abs(ua).max()
    ifn = sorted(set(int(v) for e in mrm.elems_of_boundary[7] for v in mrm.mesh.conn[e]))
    _, t1 = mrm.nodal_solution(x, 0)
    _, s_if = field(np.array([0.4]))
    assert np.abs(t1[ifn, 0] - s_if[0]).max() < 5e-3 * abs(s_if[0])


def test_two_layer_fluid_room():
    f1, f2 = Fluid(1.25, 343.0), Fluid(1000.0, 1480.0, 0.01)
    omega = 2 * np.pi * 120.0
    mrm = MultiRegionModel(two_box_mesh(2, shape.QUAD9, xs=0.6), [Region(FLUID, f1, [1, 3, 4, 5, 6, 7]), Region(FLUID, f2, [-7, 2, 13, 14, 15, 16])], BPART, fluid_bcs())
    A, b = MultiRegionOracle(mrm).assemble(omega)
    x = np.linalg.solve(A, b)
    # f = p, s = U_x = p'/(rho omega^2)
    field = layered_1d(omega, [(omega / f1.c, 1.0 / (f1.rho * omega ** 2)), (omega / f2.c, 1.0 / (f2.rho * omega ** 2)), 0.6], ("f", 0.0), ("f", 1.0))
    for kr in (0, 1):
        p, un = mrm.nodal_solution(x, kr)
        ok = ~np.isnan(p)
        xs_ = np.clip(mrm.node_x[ok, 0], 0, 0.6) if kr == 0 else np.clip(mrm.node_x[ok, 0], 0.6 + 1e-12, 1)
        pa, _ = field(xs_)
        assert np.abs(p[ok] - pa).max() < 2e-3 * np.abs(pa).max()
    ifn = sorted(set(int(v) for e in mrm.elems_of_boundary[7] for v in mrm.mesh.conn[e]))
    _, un1 = mrm.nodal_solution(x, 0); _, un2 = mrm.nodal_solution(x, 1)
    _, U = field(np.array([0.6]))
    assert np.abs(un1[ifn] - U[0]).max() < 5e-3 * abs(U[0]) and np.array_equal(un1[ifn], -un2[ifn])


@pytest.mark.parametrize("solid_first", [True, False])
def test_fluid_solid_column(solid_first):
    """Elastic layer against a fluid layer: sigma_xx = -p and u_x = U_x at the interface (t = -p n, Un = u.n)."""
    ms, fl = Material(2.0, 1.5, 0.25, 0.03), Fluid(1.0, 1.2, 0.01)
    omega = 3.0
    Zs, Zf = ms.lam + 2 * ms.mu, -1.0 / (fl.rho * omega ** 2)       # with f = u (solid) / U = ... see below
    if solid_first:      # solid on [0, xs] clamped at x = 0, fluid on [xs, 1] with p(1) = P
        regs = [Region(SOLID, ms, [1, 3, 4, 5, 6, 7]), Region(FLUID, fl, [-7, 2, 13, 14, 15, 16])]
        bcs = solid_bcs(); bcs.update({k: v for k, v in fluid_bcs(1.0).items() if k in (2,) + LAT2})
    else:                # fluid on [0, xs] with p(0) = 0 ... use p(0) = P instead to drive it; solid on [xs, 1] with traction-free end replaced by clamped end
        regs = [Region(FLUID, fl, [1, 3, 4, 5, 6, 7]), Region(SOLID, ms, [-7, 2, 13, 14, 15, 16])]
        bcs = fluid_bcs(); bcs[1] = (0, 1.0)
        sb = solid_bcs(); bcs.update({k: sb[k] for k in LAT2}); bcs[2] = ([0, 0, 0], [0, 0, 0])
    xs = 0.5
    mrm = MultiRegionModel(two_box_mesh(2, shape.QUAD9, xs=xs), regs, BPART, bcs)
    A, b = MultiRegionOracle(mrm).assemble(omega)
    x = np.linalg.solve(A, b)
    # common 1D unknown: displacement w(x) (u_x in the solid, U_x in the fluid); stress s = Z w' with Z_s = lambda + 2 mu, and in the fluid
    # p = -K w' with K = rho c^2, i.e. sigma = -p = K w': Z_f = rho c^2.  Continuity of w and of sigma at xs.
    Kf = fl.rho * fl.c ** 2
    ks, kf = omega / ms.c1, omega / fl.c
    if solid_first:
        field = layered_1d(omega, [(ks, Zs), (kf, Kf), xs], ("f", 0.0), ("s", -1.0))       # sigma(1) = -p(1) = -P
    else:
        field = layered_1d(omega, [(kf, Kf), (ks, Zs), xs], ("s", -1.0), ("f", 0.0))       # sigma(0) = -P, clamped at x = 1
    ks_, kfl = (0, 1) if solid_first else (1, 0)
    u, t = mrm.nodal_solution(x, ks_)
    ok = ~np.isnan(u[:, 0])
    lo, hi = (0.0, xs) if solid_first else (xs + 1e-12, 1.0)
    wa, sa = field(np.clip(mrm.node_x[ok, 0], lo, hi))
    assert np.abs(u[ok, 0] - wa).max() < 5e-3 * np.abs(wa).max()
    p, un = mrm.nodal_solution(x, kfl)
    ok = ~np.isnan(p)
    lo, hi = (xs + 1e-12, 1.0) if solid_first else (0.0, xs)
    wa, sa = field(np.clip(mrm.node_x[ok, 0], lo, hi))
    assert np.abs(p[ok] + sa).max() < 5e-3 * np.abs(sa).max()                                # p = -sigma


@pytest.mark.parametrize("kinds", [(SOLID, SOLID), (FLUID, FLUID), (SOLID, FLUID), (FLUID, SOLID)])
def test_flat_scatter_descriptors_reproduce_every_case(kinds):
    """The one-rule descriptors (col_h, coef_h, col_g[3], coef_g[3] per element node and component: the form that crosses the C ABI) give
    the same system as the case-by-case restatement of assemble_bem_har{ela,pot}_equation."""
    mats = {SOLID: Material(2.0, 1.5, 0.25, 0.03), FLUID: Fluid(1.0, 1.2, 0.01)}
    sb, fb = solid_bcs(0.7 + 0.2j), fluid_bcs(0.4 - 0.1j)
    bcs = {}
    for k, lat, ends in ((kinds[0], LAT1, (1,)), (kinds[1], LAT2, (2,))):
        src = sb if k == SOLID else fb
        bcs.update({q: src[q] for q in lat + ends})
    bcs[3 if kinds[0] == SOLID else 13] = ([0, 1, 0], [0.1, 0.2j, -0.3]) if SOLID in kinds else bcs[3]     # a nonzero prescribed displacement too
    if kinds[0] != SOLID and SOLID in kinds:
        bcs[3] = fb[3]
    mrm = MultiRegionModel(two_box_mesh(1, shape.QUAD8), [Region(kinds[0], mats[kinds[0]], [1, 3, 4, 5, 6, 7]), Region(kinds[1], mats[kinds[1]], [-7, 2, 13, 14, 15, 16])], BPART, bcs)
    o = MultiRegionOracle(mrm)
    A1, b1 = o.assemble(1.7)
    A2, b2 = o.assemble(1.7, flat=True)
    assert np.abs(A1 - A2).max() <= 1e-15 * np.abs(A1).max() and np.abs(b1 - b2).max() <= 1e-15 * max(np.abs(b1).max(), 1e-300)
    assert np.abs(b1).max() > 0
