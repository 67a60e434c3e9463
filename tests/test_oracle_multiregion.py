"""Several BE regions coupled through be-be interfaces: host numbering (multifebe_b200.host.MultiRegionModel) and the
multi-region oracle driver (oracle/multiregion.py) around the pinned single-region integrals.  The reference holds no numeric vectors
for this path; the pins are (1) a homogeneous body split in two regions must reproduce the one-region solution, (2) exact 1D
two-layer solutions: solid-solid, fluid-fluid and fluid-solid columns."""
import numpy as np
import pytest

from multifebe_b200.host import Material, Fluid, MultiRegionModel, Region, SOLID, FLUID, two_box_mesh, cube_mesh, column_analytic_u, shape
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


# ---- poroelastic regions and the fluid-poroelastic interface (BASELINE config 4: harpor + harpot coupling) -------------------------------
from multifebe_b200.host import Poro, PoroModel
from multifebe_b200.host.multiregion import PORO

PO = Poro(rhof=1.0, rhos=2.2, lam=1.2, mu=1.0, xi=0.02, phi=0.35, rhoa=0.15, R=0.8, Q=0.5, b=0.4)
FL = Fluid(1.1, 1.3, 0.01)


def poro_bcs_side(parts):
    """sliding impermeable lateral faces of a poroelastic box: Un = 0, normal skeleton displacement 0, shear tractions 0."""
    out = {}
    for p_ in parts:
        free = 2 if p_ in (3, 4, 13, 14) else 3
        ct = [1, 1, 1, 1]; ct[free] = 0
        out[p_] = (ct, [0, 0, 0, 0])
    return out


def test_one_poroelastic_region_through_the_multiregion_driver_equals_the_single_region_oracle():
    mesh = cube_mesh(1, shape.QUAD9)
    bcs = {1: ([1, 0, 0, 0], [0, 0, 0, 0]), 2: ([0, 1, 1, 1], [0, 1.0, 0, 0])}
    bcs.update(poro_bcs_side((3, 4, 5, 6)))
    md = PoroModel(mesh, bcs)
    A0, b0, _ = orc.PorOracle(md).assemble(1.9, PO)
    mrm = MultiRegionModel(mesh, [Region(PORO, PO, [1, 2, 3, 4, 5, 6])], {b: b for b in range(1, 7)}, bcs)
    assert mrm.n_dof == md.n_dof
    A1, b1 = MultiRegionOracle(mrm).assemble(1.9)
    assert np.abs(A1 - A0).max() < 1e-13 * np.abs(A0).max() and np.abs(b1 - b0).max() < 1e-13 * np.abs(b0).max()
    A2, b2 = MultiRegionOracle(mrm).assemble(1.9, flat=True)
    assert np.abs(A2 - A1).max() <= 1e-15 * np.abs(A1).max() and np.abs(b2 - b1).max() <= 1e-15 * np.abs(b1).max()


def fluid_poro_1d(omega, fl, po, xs, fluid_first, imp, P=1.0):
    """Exact 1D column: an inviscid fluid layer against a saturated poroelastic layer.  Fluid end: p = P; poroelastic end: fixed and impermeable
    (u = U = 0).  Interface: normal total stress continuous (sigma_s + tau = -p); permeable: p = -tau/phi and U_f = phi U + (1 - phi) u;
    impermeable: U = u = U_f.  Unknowns: fluid (a, b) of p = a e^{-ikx} + b e^{ikx}, poroelastic (a1, b1, a2, b2)."""
    M = np.array([[po.lam + 2 * po.mu + po.Q ** 2 / po.R, po.Q], [po.Q, po.R]])
    rh11 = po.rho1 + po.rhoa - 1j * po.b / omega; rh12 = -po.rhoa + 1j * po.b / omega; rh22 = po.rho2 + po.rhoa - 1j * po.b / omega
    k2, Y = np.linalg.eig(np.linalg.solve(M, omega ** 2 * np.array([[rh11, rh12], [rh12, rh22]])))
    ks = np.sqrt(k2); ks = np.where(ks.real < 0, -ks, ks)
    kf = omega / fl.c

    def poro_rows(x):       # u, U, sigma_s, tau as rows over (a1, b1, a2, b2)
        e = [np.exp(-1j * ks[0] * x), np.exp(1j * ks[0] * x), np.exp(-1j * ks[1] * x), np.exp(1j * ks[1] * x)]
        d = [-1j * ks[0] * e[0], 1j * ks[0] * e[1], -1j * ks[1] * e[2], 1j * ks[1] * e[3]]
        yv = [Y[:, 0], Y[:, 0], Y[:, 1], Y[:, 1]]
        u = np.array([e[q] * yv[q][0] for q in range(4)]); U = np.array([e[q] * yv[q][1] for q in range(4)])
        du = np.array([d[q] * yv[q][0] for q in range(4)]); dU = np.array([d[q] * yv[q][1] for q in range(4)])
        return u, U, M[0, 0] * du + M[0, 1] * dU, M[1, 0] * du + M[1, 1] * dU

    def fluid_rows(x):      # p, U_f = p'/(rho w^2) over (a, b)
        e = np.array([np.exp(-1j * kf * x), np.exp(1j * kf * x)])
        return e, np.array([-1j * kf, 1j * kf]) * e / (fl.rho * omega ** 2)
    x_f, x_p = (0.0, 1.0) if fluid_first else (1.0, 0.0)
    S = np.zeros((6, 6), dtype=complex); r = np.zeros(6, dtype=complex)
    S[0, :2] = fluid_rows(x_f)[0]; r[0] = P
    u, U, sg, ta = poro_rows(x_p); S[1, 2:] = u; S[2, 2:] = U
    u, U, sg, ta = poro_rows(xs); pf, Uf = fluid_rows(xs)
    S[3, :2] = pf; S[3, 2:] = sg + ta                       # sigma_s + tau + p = 0
    if imp:
        S[4, 2:] = U - u; S[5, :2] = Uf; S[5, 2:] = -u
    else:
        S[4, :2] = pf; S[4, 2:] = ta / po.phi               # p + tau/phi = 0
        S[5, :2] = Uf; S[5, 2:] = -(po.phi * U + (1 - po.phi) * u)
    c = np.linalg.solve(S, r)
    return (lambda x: [row @ c[:2] for row in fluid_rows(x)]), (lambda x: [row @ c[2:] for row in poro_rows(x)])


@pytest.mark.parametrize("fluid_first", [True, False])
@pytest.mark.parametrize("imp", [False, True])
def test_fluid_against_poroelastic_column(fluid_first, imp):
    omega, xs = 2.0, 0.45
    fb = fluid_bcs(1.0)
    pend = ([1, 0, 0, 0], [0, 0, 0, 0])                      # poroelastic end: Un = 0, u = 0
    if fluid_first:
        regs = [Region(FLUID, FL, [1, 3, 4, 5, 6, 7]), Region(PORO, PO, [-7, 2, 13, 14, 15, 16])]
        bcs = {1: (0, 1.0), 2: pend}; bcs.update({q: fb[q] for q in LAT1}); bcs.update(poro_bcs_side(LAT2))
    else:
        regs = [Region(PORO, PO, [1, 3, 4, 5, 6, 7]), Region(FLUID, FL, [-7, 2, 13, 14, 15, 16])]
        bcs = {2: (0, 1.0), 1: pend}; bcs.update({q: fb[q] for q in LAT2}); bcs.update(poro_bcs_side(LAT1))
    mrm = MultiRegionModel(two_box_mesh(2, shape.QUAD9, xs=xs), regs, BPART, bcs, interface_ctype={7: 1 if imp else 0})
    o = MultiRegionOracle(mrm)
    A, b = o.assemble(omega)
    A2, b2 = o.assemble(omega, flat=True)
    assert np.abs(A - A2).max() <= 1e-15 * np.abs(A).max() and np.abs(b - b2).max() <= 1e-15 * np.abs(b).max()
    x = np.linalg.solve(A, b)
    fluid, poro = fluid_poro_1d(omega, FL, PO, xs, fluid_first, imp)
    kf_, kp_ = (0, 1) if fluid_first else (1, 0)
    # pressure on the rigid walls of the fluid box, skeleton displacement and tau on the sliding sides of the poroelastic box
    fl_side = LAT1 if fluid_first else LAT2; po_side = LAT2 if fluid_first else LAT1
    errs = []
    for bnd in fl_side:
        for v in sorted(set(int(n) for e in mrm.elems_of_boundary[bnd] for n in mrm.mesh.conn[e])):
            errs.append(abs(x[mrm.col[(v, "p1")]] - fluid(mrm.node_x[v, 0])[0]))
    assert max(errs) < 1e-2                                   # |p| = O(1)
    eu, et_ = [], []
    for bnd in po_side:
        for v in sorted(set(int(n) for e in mrm.elems_of_boundary[bnd] for n in mrm.mesh.conn[e])):
            u, U, sg, ta = poro(mrm.node_x[v, 0])
            eu.append(abs(x[mrm.col[(v, "u10")]] - u)); et_.append(abs(x[mrm.col[(v, "tau1")]] - ta))
    uref = max(abs(poro(t)[0]) for t in np.linspace(0, 1, 11)); tref = max(abs(poro(t)[3]) for t in np.linspace(0, 1, 11))
    assert max(eu) < 1e-2 * uref and max(et_) < 1e-2 * tref
    # interface unknowns: tau and u on the poroelastic side; p (impermeable) or the relative fluid displacement w (permeable)
    side = 2 if fluid_first else 1
    u, U, sg, ta = poro(xs)
    for v in sorted(set(int(n) for e in mrm.elems_of_boundary[7] for n in mrm.mesh.conn[e])):
        assert abs(x[mrm.col[(v, "tau%d" % side)]] - ta) < 2e-2 * tref and abs(x[mrm.col[(v, "u%d0" % side)]] - u) < 2e-2 * uref
        if imp:
            assert abs(x[mrm.col[(v, "p%d" % (3 - side))]] - fluid(xs)[0]) < 2e-2
        else:
            n_out = -1.0 if fluid_first else 1.0              # outward normal of the poroelastic region at the interface, x component
            assert abs(x[mrm.col[(v, "w%d" % side)]] - n_out * U) < 3e-2 * max(abs(U), uref)


def solid_poro_1d(omega, ms, po, xs, solid_first, P=1.0):
    """Exact 1D column: an elastic layer bonded to a saturated poroelastic layer through an impervious contact (u continuous, U = u,
    total normal stress continuous: sigma_solid = sigma_s + tau).  Solid end: normal traction P on the free end; poroelastic end: fixed, impermeable."""
    M = np.array([[po.lam + 2 * po.mu + po.Q ** 2 / po.R, po.Q], [po.Q, po.R]])
    rh11 = po.rho1 + po.rhoa - 1j * po.b / omega; rh12 = -po.rhoa + 1j * po.b / omega; rh22 = po.rho2 + po.rhoa - 1j * po.b / omega
    k2, Y = np.linalg.eig(np.linalg.solve(M, omega ** 2 * np.array([[rh11, rh12], [rh12, rh22]])))
    kp = np.sqrt(k2); kp = np.where(kp.real < 0, -kp, kp)
    ks = omega / ms.c1; Zs = ms.lam + 2 * ms.mu

    def poro_rows(x):
        e = [np.exp(-1j * kp[0] * x), np.exp(1j * kp[0] * x), np.exp(-1j * kp[1] * x), np.exp(1j * kp[1] * x)]
        d = [-1j * kp[0] * e[0], 1j * kp[0] * e[1], -1j * kp[1] * e[2], 1j * kp[1] * e[3]]
        yv = [Y[:, 0], Y[:, 0], Y[:, 1], Y[:, 1]]
        u = np.array([e[q] * yv[q][0] for q in range(4)]); U = np.array([e[q] * yv[q][1] for q in range(4)])
        du = np.array([d[q] * yv[q][0] for q in range(4)]); dU = np.array([d[q] * yv[q][1] for q in range(4)])
        return u, U, M[0, 0] * du + M[0, 1] * dU, M[1, 0] * du + M[1, 1] * dU

    def solid_rows(x):
        e = np.array([np.exp(-1j * ks * x), np.exp(1j * ks * x)])
        return e, Zs * np.array([-1j * ks, 1j * ks]) * e
    x_s, x_p = (0.0, 1.0) if solid_first else (1.0, 0.0)
    S = np.zeros((6, 6), dtype=complex); r = np.zeros(6, dtype=complex)
    # traction on the solid end: t = sigma n, n = -x at x = 0 and +x at x = 1; prescribed t_x = P
    S[0, :2] = solid_rows(x_s)[1] * (-1.0 if solid_first else 1.0); r[0] = P
    u, U, sg, ta = poro_rows(x_p); S[1, 2:] = u; S[2, 2:] = U
    u, U, sg, ta = poro_rows(xs); us, ss = solid_rows(xs)
    S[3, :2] = us; S[3, 2:] = -u
    S[4, 2:] = U - u
    S[5, :2] = ss; S[5, 2:] = -(sg + ta)
    c = np.linalg.solve(S, r)
    return (lambda x: [row @ c[:2] for row in solid_rows(x)]), (lambda x: [row @ c[2:] for row in poro_rows(x)])


@pytest.mark.parametrize("solid_first", [True, False])
def test_solid_bonded_to_a_poroelastic_layer(solid_first):
    ms = Material(1.8, 1.4, 0.25, 0.03)
    omega, xs = 2.0, 0.5
    sb = solid_bcs(1.0)
    pend = ([1, 0, 0, 0], [0, 0, 0, 0])
    if solid_first:
        regs = [Region(SOLID, ms, [1, 3, 4, 5, 6, 7]), Region(PORO, PO, [-7, 2, 13, 14, 15, 16])]
        bcs = {1: ([1, 1, 1], [1.0, 0, 0]), 2: pend}; bcs.update({q: sb[q] for q in LAT1}); bcs.update(poro_bcs_side(LAT2))
    else:
        regs = [Region(PORO, PO, [1, 3, 4, 5, 6, 7]), Region(SOLID, ms, [-7, 2, 13, 14, 15, 16])]
        bcs = {2: ([1, 1, 1], [1.0, 0, 0]), 1: pend}; bcs.update({q: sb[q] for q in LAT2}); bcs.update(poro_bcs_side(LAT1))
    mrm = MultiRegionModel(two_box_mesh(2, shape.QUAD9, xs=xs), regs, BPART, bcs)
    o = MultiRegionOracle(mrm)
    A, b = o.assemble(omega)
    A2, b2 = o.assemble(omega, flat=True)
    assert np.abs(A - A2).max() <= 1e-15 * np.abs(A).max() and np.abs(b - b2).max() <= 1e-15 * np.abs(b).max()
    x = np.linalg.solve(A, b)
    solid, poro = solid_poro_1d(omega, ms, PO, xs, solid_first)
    uref = max(max(abs(solid(t)[0]) for t in np.linspace(0, 1, 11)), max(abs(poro(t)[0]) for t in np.linspace(0, 1, 11)))
    tref = max(abs(poro(t)[3]) for t in np.linspace(0, 1, 11))
    so_side = LAT1 if solid_first else LAT2; po_side = LAT2 if solid_first else LAT1
    for bnd in so_side:
        for v in sorted(set(int(n) for e in mrm.elems_of_boundary[bnd] for n in mrm.mesh.conn[e])):
            assert abs(x[mrm.col[(v, "u10")]] - solid(mrm.node_x[v, 0])[0]) < 1e-2 * uref
    for bnd in po_side:
        for v in sorted(set(int(n) for e in mrm.elems_of_boundary[bnd] for n in mrm.mesh.conn[e])):
            u, U, sg, ta = poro(mrm.node_x[v, 0])
            assert abs(x[mrm.col[(v, "u10")]] - u) < 1e-2 * uref and abs(x[mrm.col[(v, "tau1")]] - ta) < 1e-2 * tref
    side = 2 if solid_first else 1
    u, U, sg, ta = poro(xs)
    n_p = -1.0 if solid_first else 1.0                        # outward normal of the poroelastic region at the interface (x component)
    for v in sorted(set(int(n) for e in mrm.elems_of_boundary[7] for n in mrm.mesh.conn[e])):
        assert abs(x[mrm.col[(v, "u%d0" % side)]] - u) < 2e-2 * uref and abs(x[mrm.col[(v, "tau%d" % side)]] - ta) < 2e-2 * tref
        assert abs(x[mrm.col[(v, "t%d0" % side)]] - n_p * sg) < 3e-2 * max(abs(sg), tref)       # skeleton traction t = sigma_s n


def test_two_poroelastic_layers_permeable_contact():
    """Two saturated layers in perfectly permeable contact: u continuous, pore pressure continuous (tau1/phi1 = tau2/phi2), relative fluid flux
    continuous (phi1 (U1 - u) = phi2 (U2 - u)), total normal stress continuous.  Fixed impermeable base at x = 0, loaded drained top at x = 1."""
    po1 = PO
    po2 = Poro(rhof=1.0, rhos=2.6, lam=2.0, mu=1.5, xi=0.03, phi=0.2, rhoa=0.1, R=0.5, Q=0.7, b=0.8)
    omega, xs = 2.0, 0.5

    def modes(po):
        M = np.array([[po.lam + 2 * po.mu + po.Q ** 2 / po.R, po.Q], [po.Q, po.R]])
        rh11 = po.rho1 + po.rhoa - 1j * po.b / omega; rh12 = -po.rhoa + 1j * po.b / omega; rh22 = po.rho2 + po.rhoa - 1j * po.b / omega
        k2, Y = np.linalg.eig(np.linalg.solve(M, omega ** 2 * np.array([[rh11, rh12], [rh12, rh22]])))
        kp = np.sqrt(k2); kp = np.where(kp.real < 0, -kp, kp)

        def rows(x):
            e = [np.exp(-1j * kp[0] * x), np.exp(1j * kp[0] * x), np.exp(-1j * kp[1] * x), np.exp(1j * kp[1] * x)]
            d = [-1j * kp[0] * e[0], 1j * kp[0] * e[1], -1j * kp[1] * e[2], 1j * kp[1] * e[3]]
            yv = [Y[:, 0], Y[:, 0], Y[:, 1], Y[:, 1]]
            u = np.array([e[q] * yv[q][0] for q in range(4)]); U = np.array([e[q] * yv[q][1] for q in range(4)])
            du = np.array([d[q] * yv[q][0] for q in range(4)]); dU = np.array([d[q] * yv[q][1] for q in range(4)])
            return u, U, M[0, 0] * du + M[0, 1] * dU, M[1, 0] * du + M[1, 1] * dU
        return rows
    r1, r2 = modes(po1), modes(po2)
    S = np.zeros((8, 8), dtype=complex); rhs = np.zeros(8, dtype=complex)
    u, U, sg, ta = r1(0.0); S[0, :4] = u; S[1, :4] = U
    u, U, sg, ta = r2(1.0); S[2, 4:] = sg; rhs[2] = 1.0; S[3, 4:] = ta
    ua, Ua, sa, tta = r1(xs); ub, Ub, sb_, ttb = r2(xs)
    S[4, :4] = ua; S[4, 4:] = -ub
    S[5, :4] = tta / po1.phi; S[5, 4:] = -ttb / po2.phi
    S[6, :4] = po1.phi * (Ua - ua); S[6, 4:] = -po2.phi * (Ub - ub)
    S[7, :4] = sa + tta; S[7, 4:] = -(sb_ + ttb)
    c = np.linalg.solve(S, rhs)
    f1 = lambda x: [row @ c[:4] for row in r1(x)]
    f2 = lambda x: [row @ c[4:] for row in r2(x)]
    regs = [Region(PORO, po1, [1, 3, 4, 5, 6, 7]), Region(PORO, po2, [-7, 2, 13, 14, 15, 16])]
    bcs = {1: ([1, 0, 0, 0], [0, 0, 0, 0]), 2: ([0, 1, 1, 1], [0, 1.0, 0, 0])}
    bcs.update(poro_bcs_side(LAT1)); bcs.update(poro_bcs_side(LAT2))
    mrm = MultiRegionModel(two_box_mesh(2, shape.QUAD9, xs=xs), regs, BPART, bcs)
    o = MultiRegionOracle(mrm)
    A, b = o.assemble(omega)
    A2, b2 = o.assemble(omega, flat=True)
    assert np.abs(A - A2).max() <= 1e-15 * np.abs(A).max() and np.abs(b - b2).max() <= 1e-15 * np.abs(b).max()
    x = np.linalg.solve(A, b)
    uref = max(max(abs(f1(t)[0]) for t in np.linspace(0, xs, 6)), max(abs(f2(t)[0]) for t in np.linspace(xs, 1, 6)))
    tref = max(max(abs(f1(t)[3]) for t in np.linspace(0, xs, 6)), max(abs(f2(t)[3]) for t in np.linspace(xs, 1, 6)))
    for side, f in ((LAT1, f1), (LAT2, f2)):
        for bnd in side:
            for v in sorted(set(int(n) for e in mrm.elems_of_boundary[bnd] for n in mrm.mesh.conn[e])):
                u, U, sg, ta = f(mrm.node_x[v, 0])
                assert abs(x[mrm.col[(v, "u10")]] - u) < 1e-2 * uref and abs(x[mrm.col[(v, "tau1")]] - ta) < 1e-2 * tref
    u, U, sg, ta = f1(xs)
    for v in sorted(set(int(n) for e in mrm.elems_of_boundary[7] for n in mrm.mesh.conn[e])):
        assert abs(x[mrm.col[(v, "u10")]] - u) < 2e-2 * uref and abs(x[mrm.col[(v, "tau1")]] - ta) < 2e-2 * tref
        assert abs(x[mrm.col[(v, "w1")]] - U) < 3e-2 * max(abs(U), uref) and abs(x[mrm.col[(v, "t10")]] - sg) < 3e-2 * max(abs(sg), tref)


@pytest.mark.parametrize("case", ["fluid_poro_perm", "poro_fluid_imp", "solid_poro", "poro_poro"])
def test_nodal_variables_on_both_sides_of_poroelastic_interfaces(case):
    """MultiRegionModel.nodal_solution rebuilds every node variable of both sides from the active unknowns: on the interface they must satisfy
    the physical contact conditions (total normal stress, pressure, normal flux) that the reference's substitutions encode."""
    omega, xs = 2.0, 0.5
    fb, sb = fluid_bcs(1.0), solid_bcs(1.0)
    pend = ([1, 0, 0, 0], [0, 0, 0, 0])
    po2 = Poro(rhof=1.0, rhos=2.6, lam=2.0, mu=1.5, xi=0.03, phi=0.2, rhoa=0.1, R=0.5, Q=0.7, b=0.8)
    if case == "fluid_poro_perm":
        regs = [Region(FLUID, FL, [1, 3, 4, 5, 6, 7]), Region(PORO, PO, [-7, 2, 13, 14, 15, 16])]; ict = 0
        bcs = {1: (0, 1.0), 2: pend}; bcs.update({q: fb[q] for q in LAT1}); bcs.update(poro_bcs_side(LAT2))
    elif case == "poro_fluid_imp":
        regs = [Region(PORO, PO, [1, 3, 4, 5, 6, 7]), Region(FLUID, FL, [-7, 2, 13, 14, 15, 16])]; ict = 1
        bcs = {2: (0, 1.0), 1: pend}; bcs.update({q: fb[q] for q in LAT2}); bcs.update(poro_bcs_side(LAT1))
    elif case == "solid_poro":
        regs = [Region(SOLID, Material(1.8, 1.4, 0.25, 0.03), [1, 3, 4, 5, 6, 7]), Region(PORO, PO, [-7, 2, 13, 14, 15, 16])]; ict = 0
        bcs = {1: ([1, 1, 1], [1.0, 0, 0]), 2: pend}; bcs.update({q: sb[q] for q in LAT1}); bcs.update(poro_bcs_side(LAT2))
    else:
        regs = [Region(PORO, PO, [1, 3, 4, 5, 6, 7]), Region(PORO, po2, [-7, 2, 13, 14, 15, 16])]; ict = 0
        bcs = {1: pend, 2: ([0, 1, 1, 1], [0, 1.0, 0, 0])}; bcs.update(poro_bcs_side(LAT1)); bcs.update(poro_bcs_side(LAT2))
    mrm = MultiRegionModel(two_box_mesh(1, shape.QUAD9, xs=xs), regs, BPART, bcs, interface_ctype={7: ict})
    A, b = MultiRegionOracle(mrm).assemble(omega)
    x = np.linalg.solve(A, b)
    (P1, S1), (P2, S2) = mrm.nodal_solution(x, 0), mrm.nodal_solution(x, 1)
    ifn = sorted(set(int(n) for e in mrm.elems_of_boundary[7] for n in mrm.mesh.conn[e]))

    def total_normal_stress(kind, P, S, v, nx):      # sigma_xx seen from the region whose outward normal at the interface is nx * e_x
        if kind == FLUID:
            return -P[v]
        if kind == SOLID:
            return S[v, 0] * nx
        return S[v, 1] * nx + P[v, 0]                # skeleton traction t_x = sigma_s n_x, plus tau

    def normal_flux(kind, reg, P, S, v, nx):         # fluid displacement relative to ... : absolute normal displacement of the pore fluid / fluid
        if kind == FLUID:
            return S[v] * nx
        if kind == SOLID:
            return P[v, 0]
        return S[v, 0] * nx                           # U_x
    k1, k2 = regs[0].kind, regs[1].kind
    for v in ifn:
        s1 = total_normal_stress(k1, P1, S1, v, 1.0); s2 = total_normal_stress(k2, P2, S2, v, -1.0)
        assert abs(s1 - s2) < 1e-10 * max(abs(s1), 1e-3)
        # skeleton displacement continuous where both sides have one
        if k1 != FLUID and k2 != FLUID:
            u1 = P1[v, 1:] if k1 == PORO else P1[v]; u2 = P2[v, 1:] if k2 == PORO else P2[v]
            assert np.abs(u1 - u2).max() == 0
        if case == "fluid_poro_perm":                # p = -tau/phi; fluid flux U_f = phi U + (1 - phi) u
            assert abs(P1[v] + P2[v, 0] / PO.phi) < 1e-12 and abs(S1[v] - (PO.phi * (-S2[v, 0]) + (1 - PO.phi) * P2[v, 1])) < 1e-12 * max(abs(S1[v]), 1e-3)
        if case == "poro_fluid_imp":                 # U = u = U_f
            assert abs(S1[v, 0] - P1[v, 1]) < 1e-12 and abs(-S2[v] - P1[v, 1]) < 1e-12
        if case == "poro_poro":                      # pore pressure and relative flux continuous
            assert abs(P1[v, 0] / PO.phi - P2[v, 0] / po2.phi) < 1e-12
            q1 = PO.phi * (S1[v, 0] - P1[v, 1]); q2 = po2.phi * (-S2[v, 0] - P2[v, 1])
            assert abs(q1 - q2) < 1e-12 * max(abs(q1), 1e-3)


def test_impedance_and_radiation_conditions_of_a_fluid_boundary():
    """Condition 2 (Un = -i/(rho c omega) p, the rho c impedance): a duct driven at x = 0 and terminated by it at x = L carries the travelling wave
    p = P e^{-ikx} alone.  Condition 3 adds the spherical-spreading term 1/(2 R rho omega^2); here only its assembly is compared (flat descriptors
    against the case-by-case branch).  assemble_bem_harpot_equation.f90:97-110."""
    fl = Fluid(1.2, 1.5)
    omega = 4.0
    bcs = {1: (0, 1.0), 2: (2, 0.0), 3: (1, 0.0), 4: (1, 0.0), 5: (1, 0.0), 6: (1, 0.0)}
    mrm = MultiRegionModel(cube_mesh(3, shape.QUAD9), [Region(FLUID, fl, [1, 2, 3, 4, 5, 6])], {b: b for b in range(1, 7)}, bcs)
    o = MultiRegionOracle(mrm)
    A, b = o.assemble(omega)
    A2, b2 = o.assemble(omega, flat=True)
    assert np.abs(A - A2).max() <= 1e-15 * np.abs(A).max() and np.abs(b - b2).max() <= 1e-15 * np.abs(b).max()
    x = np.linalg.solve(A, b)
    p, un = mrm.nodal_solution(x, 0, omega)
    k = omega / fl.c
    ex = np.exp(-1j * k * mrm.node_x[:, 0])
    assert np.abs(p - ex).max() < 2e-3
    end = mrm.node_boundary == 2
    assert np.abs(un[end] + 1j / (fl.rho * fl.c * omega) * p[end]).max() == 0          # Un = -i/(rho c omega) p; the exact value is -i k/(rho omega^2) e^{-ikL}
    assert np.abs(un[end] - (-1j * k) * ex[end] / (fl.rho * omega ** 2)).max() < 2e-3 * k / (fl.rho * omega ** 2)
    bcs[2] = (3, 2.5)
    mrm3 = MultiRegionModel(cube_mesh(1, shape.QUAD8), [Region(FLUID, fl, [1, 2, 3, 4, 5, 6])], {b: b for b in range(1, 7)}, bcs)
    o3 = MultiRegionOracle(mrm3)
    A, b = o3.assemble(omega); A2, b2 = o3.assemble(omega, flat=True)
    assert np.abs(A - A2).max() <= 1e-15 * np.abs(A).max() and np.abs(b - b2).max() <= 1e-15 * np.abs(b).max()
