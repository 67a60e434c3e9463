"""Independent pins of the oracle's ELEMENT INTEGRALS: the pair integrals  g = int u* phi_j dS,  h = int t* phi_j dS
that the oracle produces with the reference's machinery (rule estimator, Telles transformation, subdivision, polar transformation with line integrals) are
compared with brute-force quadrature that shares NOTHING with it:

  * kernels: the closed-form Green's function of the elastodynamic full space, G = [k2^2 I g2 + grad grad (g2 - g1)] / (4 pi rho w^2), evaluated in extended
    precision (numpy longdouble); tractions from G by 4th-order central differences with a step of 3e-4 r (also in extended precision, ~1e-13);
  * geometry: shape functions written here from the element definitions;
  * quadrature: a quadtree / recursive triangle split of the reference element graded towards the collocation point, 14 x 14 Gauss-Legendre per leaf
    (collapsed square on triangles); for a point ON a flat element, polar coordinates in the element plane about the point, with the 1/r^2 part of the
    traction kernel treated analytically (Cauchy principal value: F0(theta)/r integrates to F0 ln rho_max, the ln(eps) terms cancel over the full circle).

The oracle is run with the rule estimator's tolerance tightened (qsi_relative_error 1e-12) for the near pairs: with the default 1e-6 its quadrature error,
not its correctness, would be what the comparison sees.  Analytic columns see the oracle to 1e-3; this test sees every piece of the integration path to 1e-9."""
import numpy as np
import pytest
from multifebe_b200.host import Model, Material, cube_mesh, cube_bcs, shape

LD = np.longdouble
CLD = np.clongdouble
MAT = Material(1.3, 2.0, 0.3, 0.04)
OMEGA = 3.0

# ---- shape functions from the element definitions (node order of the mesh generator = the reference's) ----
QS = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1], [0, -1], [1, 0], [0, 1], [-1, 0], [0, 0]], dtype=float)


def _lag(s, t):
    return 1.0 - t * t if s == 0 else 0.5 * t * (t + s)


def phi_of(et, xi):
    """xi [P,2] -> phi [P,nn]"""
    x, y = xi[:, 0], xi[:, 1]
    if et in (shape.TRI3, shape.TRI6):
        L = [x, y, 1.0 - x - y]
        if et == shape.TRI3:
            return np.stack(L, 1)
        return np.stack([L[0] * (2 * L[0] - 1), L[1] * (2 * L[1] - 1), L[2] * (2 * L[2] - 1), 4 * L[0] * L[1], 4 * L[1] * L[2], 4 * L[2] * L[0]], 1)
    out = []
    nn = {shape.QUAD4: 4, shape.QUAD8: 8, shape.QUAD9: 9}[et]
    for k in range(nn):
        s1, s2 = QS[k]
        if et == shape.QUAD9:
            out.append(_lag(s1, x) * _lag(s2, y))
        elif et == shape.QUAD4:
            out.append(0.25 * (1 + s1 * x) * (1 + s2 * y))
        elif k < 4:
            out.append(0.25 * (1 + s1 * x) * (1 + s2 * y) * (s1 * x + s2 * y - 1))
        elif s1 == 0:
            out.append(0.5 * (1 - x * x) * (1 + s2 * y))
        else:
            out.append(0.5 * (1 + s1 * x) * (1 - y * y))
    return np.stack(out, 1)


def surface(et, xn, xi):
    """x [P,3], unit normal [P,3], jacobian [P] by central differences of the shape functions (polynomials of degree <= 2: exact up to rounding)"""
    h = 1e-5
    e1 = np.array([h, 0.0]); e2 = np.array([0.0, h])
    phi = phi_of(et, xi)
    a1 = (phi_of(et, xi + e1) - phi_of(et, xi - e1)) @ xn / (2 * h)
    a2 = (phi_of(et, xi + e2) - phi_of(et, xi - e2)) @ xn / (2 * h)
    nv = np.cross(a1, a2); J = np.linalg.norm(nv, axis=1)
    return phi, phi @ xn, nv / J[:, None], J


# ---- extended-precision kernels ----
def U_closed(X, xi, omega, mat):
    rv = X - xi[None, :]
    r = np.sqrt((rv * rv).sum(1)); dr = rv / r[:, None]
    k1, k2 = CLD(omega) / CLD(mat.c1), CLD(omega) / CLD(mat.c2)

    def d2(k):
        g = np.exp(-1j * k * r) / r
        f1 = (-1j * k - 1 / r) * g
        f2 = g * ((-1j * k - 1 / r) ** 2 + 1 / r ** 2)
        return g, f1, f2
    g1, a1, b1 = d2(k1); g2, a2, b2 = d2(k2)
    eye = np.eye(3, dtype=LD)
    rr = dr[:, :, None] * dr[:, None, :]
    hess2 = b2[:, None, None] * rr + (a2 / r)[:, None, None] * (eye[None] - rr)
    hess1 = b1[:, None, None] * rr + (a1 / r)[:, None, None] * (eye[None] - rr)
    return (k2 * k2 * g2[:, None, None] * eye[None] + hess2 - hess1) / (4 * LD(np.pi) * LD(mat.rho) * LD(omega) ** 2)


def T_closed(X, N, xi, omega, mat):
    rv = X - xi[None, :]; r = np.sqrt((rv * rv).sum(1)); h = LD(3e-4) * r
    dG = np.zeros(X.shape[:1] + (3, 3, 3), dtype=CLD)
    for j in range(3):
        e = np.zeros_like(X); e[:, j] = h
        dG[:, :, :, j] = (-U_closed(X + 2 * e, xi, omega, mat) + 8 * U_closed(X + e, xi, omega, mat) - 8 * U_closed(X - e, xi, omega, mat)
                          + U_closed(X - 2 * e, xi, omega, mat)) / (12 * h)[:, None, None]
    lam, mu = CLD(mat.lam), CLD(mat.mu)
    div = dG[:, :, 0, 0] + dG[:, :, 1, 1] + dG[:, :, 2, 2]
    return lam * div[:, :, None] * N[:, None, :] + mu * (np.einsum('plkj,pj->plk', dG, N) + np.einsum('pljk,pj->plk', dG, N))


def test_traction_kernel_to_1e11(oracle_lib):
    """tightens the t* pin of test_oracle_kernels.py (finite differences, 1e-7) to 1e-11, also close to the source point"""
    rng = np.random.default_rng(5)
    X = rng.uniform(-1, 1, (12, 3)); X[:4] *= 0.03
    nrm = rng.standard_normal((12, 3)); nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    T = T_closed(X.astype(LD), nrm.astype(LD), np.zeros(3, dtype=LD), OMEGA, MAT)
    U = U_closed(X.astype(LD), np.zeros(3, dtype=LD), OMEGA, MAT)
    for p in range(12):
        u, t = oracle_lib.fundamental_solutions(X[p], nrm[p], np.zeros(3), OMEGA, MAT)
        assert np.abs(u - U[p].astype(complex)).max() <= 1e-13 * np.abs(u).max()
        assert np.abs(t - T[p].astype(complex)).max() <= 1e-10 * np.abs(t).max()


# ---- brute-force quadrature over the reference element, graded towards the collocation point ----
GLX, GLW = np.polynomial.legendre.leggauss(14)


def _leaves_quad(xn, et, x_i, cell=(-1.0, 1.0, -1.0, 1.0), depth=0, out=None):
    out = [] if out is None else out
    a, b, c, d = cell
    corners = np.array([[a, c], [b, c], [b, d], [a, d], [0.5 * (a + b), 0.5 * (c + d)]])
    _, xc, _, _ = surface(et, xn, corners)
    size = max(np.linalg.norm(xc[2] - xc[0]), np.linalg.norm(xc[3] - xc[1]))
    dist = np.linalg.norm(xc - x_i[None], axis=1).min()
    if depth < 12 and dist < 1.2 * size:
        m1, m2 = 0.5 * (a + b), 0.5 * (c + d)
        for sub in ((a, m1, c, m2), (m1, b, c, m2), (a, m1, m2, d), (m1, b, m2, d)):
            _leaves_quad(xn, et, x_i, sub, depth + 1, out)
    else:
        out.append(cell)
    return out


def _leaves_tri(xn, et, x_i, tri=None, depth=0, out=None):
    out = [] if out is None else out
    tri = np.array([[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]]) if tri is None else tri
    pts = np.vstack([tri, tri.mean(0)[None]])
    _, xc, _, _ = surface(et, xn, pts)
    size = max(np.linalg.norm(xc[0] - xc[1]), np.linalg.norm(xc[1] - xc[2]), np.linalg.norm(xc[2] - xc[0]))
    dist = np.linalg.norm(xc - x_i[None], axis=1).min()
    if depth < 12 and dist < 1.2 * size:
        m01, m12, m20 = 0.5 * (tri[0] + tri[1]), 0.5 * (tri[1] + tri[2]), 0.5 * (tri[2] + tri[0])
        for sub in ((tri[0], m01, m20), (m01, tri[1], m12), (m20, m12, tri[2]), (m01, m12, m20)):
            _leaves_tri(xn, et, x_i, np.array(sub), depth + 1, out)
    else:
        out.append(tri)
    return out


def brute_pair(et, xn, x_i, omega, mat, reversed_=False):
    """(h, g) [nn,3,3] of the pair by brute force (x_i off the element)"""
    pts, wts = [], []
    if et in (shape.TRI3, shape.TRI6):
        for tri in _leaves_tri(xn, et, x_i):
            # collapsed square: p = v2 + s (1 - t) (v0 - v2) + t (v1 - v2), s, t in [0, 1], weight (1 - t) * 2 * area
            s, t = np.meshgrid(0.5 * (GLX + 1), 0.5 * (GLX + 1), indexing="ij")
            w = np.outer(0.5 * GLW, 0.5 * GLW) * (1 - t)
            e0, e1 = tri[0] - tri[2], tri[1] - tri[2]
            p = tri[2][None, None] + (s * (1 - t))[..., None] * e0 + t[..., None] * e1
            pts.append(p.reshape(-1, 2)); wts.append((w * abs(e0[0] * e1[1] - e0[1] * e1[0])).ravel())
    else:
        for a, b, c, d in _leaves_quad(xn, et, x_i):
            s, t = np.meshgrid(0.5 * (a + b) + 0.5 * (b - a) * GLX, 0.5 * (c + d) + 0.5 * (d - c) * GLX, indexing="ij")
            w = np.outer(0.5 * (b - a) * GLW, 0.5 * (d - c) * GLW)
            pts.append(np.stack([s.ravel(), t.ravel()], 1)); wts.append(w.ravel())
    xi = np.vstack(pts); w = np.concatenate(wts)
    phi, x, n, J = surface(et, xn, xi)
    if reversed_:
        n = -n
    U = U_closed(x.astype(LD), x_i.astype(LD), omega, mat)
    T = T_closed(x.astype(LD), n.astype(LD), x_i.astype(LD), omega, mat)
    wj = (w * J).astype(LD)
    g = np.einsum('p,pj,plk->jlk', wj, phi.astype(LD), U)
    h = np.einsum('p,pj,plk->jlk', wj, phi.astype(LD), T)
    return h.astype(complex), g.astype(complex)


@pytest.mark.parametrize("et,m", [(shape.TRI3, 2), (shape.QUAD4, 2), (shape.TRI6, 1), (shape.QUAD8, 1), (shape.QUAD9, 1)])
def test_near_and_regular_pair_integrals_against_brute_force(oracle_lib, et, m):
    """>= 10 pairs per element type: collocation points at 0.06 ... 3 characteristic lengths from the element, over its interior, an edge and a vertex"""
    md = Model(cube_mesh(m, et), cube_bcs(), qsi_relative_error=1e-12)
    o = oracle_lib.Oracle(md)
    rng = np.random.default_rng(int(et))
    n_checked, worst = 0, 0.0
    for e in (0, md.n_elem // 2):
        nodes = md.elem_node[md.elem_ptr[e]:md.elem_ptr[e + 1]]
        xn = md.node_x[nodes]
        nn = len(nodes)
        cl = np.linalg.norm(xn[0] - xn[1]) * (2.0 if nn > 4 else 1.0)
        ctr = np.array([[1 / 3, 1 / 3]]) if et in (shape.TRI3, shape.TRI6) else np.array([[0.1, -0.2]])
        edge = np.array([[0.5, 0.5]]) if et in (shape.TRI3, shape.TRI6) else np.array([[1.0, 0.3]])
        vert = np.array([[1.0, 0.0]]) if et in (shape.TRI3, shape.TRI6) else np.array([[1.0, 1.0]])
        for base_xi, dists in ((ctr, (0.06, 0.3, 1.2)), (edge, (0.1, 0.8)), (vert, (0.15,))):
            _, xb, nb, _ = surface(et, xn, base_xi)
            for d in dists:
                side = 1.0 if rng.uniform() < 0.5 else -1.0
                tang = np.cross(nb[0], rng.standard_normal(3)); tang /= np.linalg.norm(tang)
                x_i = xb[0] + d * cl * (side * nb[0] + 0.3 * tang)
                h0, g0, mode, _ = o.pair(e, x_i, OMEGA, MAT)
                h1, g1 = brute_pair(et, xn, x_i, OMEGA, MAT)
                eh = np.abs(h0 - h1).max() / np.abs(h1).max(); eg = np.abs(g0 - g1).max() / np.abs(g1).max()
                worst = max(worst, eh, eg)
                assert eh < 1e-9 and eg < 1e-9, (int(et), e, d, mode, eh, eg)
                n_checked += 1
    assert n_checked >= 10


# ---- singular pairs: collocation point inside a FLAT element, polar coordinates in the element plane ----
def _polar_flat(et, xn, xi_i, omega, mat, ntheta=60, nrho=24):
    """h (Cauchy principal value, free term NOT included), g for x_i = x(xi_i) strictly inside a flat element whose map is affine"""
    tri = et in (shape.TRI3, shape.TRI6)
    phi_i, x0, n0, _ = surface(et, xn, xi_i[None])
    x0, n0, phi_i = x0[0], n0[0], phi_i[0]
    verts = xn[:3] if tri else xn[:4]
    # in-plane orthonormal frame and the affine map x = x0 + A (xi - xi_i)
    h_ = 1e-5
    A = np.stack([(surface(et, xn, (xi_i + np.array([h_, 0.0]))[None])[1][0] - surface(et, xn, (xi_i - np.array([h_, 0.0]))[None])[1][0]) / (2 * h_),
                  (surface(et, xn, (xi_i + np.array([0.0, h_]))[None])[1][0] - surface(et, xn, (xi_i - np.array([0.0, h_]))[None])[1][0]) / (2 * h_)], 1)   # [3,2]
    e1 = A[:, 0] / np.linalg.norm(A[:, 0]); e2 = np.cross(n0, e1)
    P = np.stack([e1, e2], 1)                                   # [3,2]
    Ainv = np.linalg.inv(P.T @ A)                               # plane coordinates -> xi offsets
    vp = (verts - x0[None]) @ P                                 # polygon in plane coordinates
    nv = len(vp)
    nu_, mu_ = mat.nu_r, CLD(mat.mu)
    g = np.zeros((len(xn), 3, 3), dtype=CLD); h = np.zeros((len(xn), 3, 3), dtype=CLD)
    tx, tw = np.polynomial.legendre.leggauss(ntheta); rx, rw = np.polynomial.legendre.leggauss(nrho)
    for k in range(nv):                                         # one angular sector per polygon edge
        a, b = vp[k], vp[(k + 1) % nv]
        ta, tb = np.arctan2(a[1], a[0]), np.arctan2(b[1], b[0])
        if tb < ta:
            tb += 2 * np.pi
        ed = b - a; nrm = np.array([ed[1], -ed[0]]) / np.linalg.norm(ed); hd = a @ nrm       # edge line: p.nrm = hd
        for kt in range(ntheta):
            th = 0.5 * (ta + tb) + 0.5 * (tb - ta) * tx[kt]; wth = 0.5 * (tb - ta) * tw[kt]
            dirp = np.array([np.cos(th), np.sin(th)]); rho_max = hd / (dirp @ nrm)
            dir3 = P @ dirp
            # static 1/r^2 part of t* on a flat element (dr/dn = 0): t_lk = -(1 - 2 nu) (n_l r_k - n_k r_l) / (8 pi (1 - nu) r^2)
            F0 = -(1 - 2 * nu_) * (np.outer(n0, dir3) - np.outer(dir3, n0)) / (8 * np.pi * (1 - nu_))
            rr = 0.5 * rho_max * (rx + 1); wr = 0.5 * rho_max * rw
            X = x0[None] + rr[:, None] * dir3[None]
            xi = xi_i[None] + (rr[:, None] * dirp[None]) @ Ainv.T
            phi = phi_of(et, xi)
            U = U_closed(X.astype(LD), x0.astype(LD), omega, mat)
            T = T_closed(X.astype(LD), np.tile(n0, (nrho, 1)).astype(LD), x0.astype(LD), omega, mat)
            g += wth * np.einsum('p,pj,plk->jlk', (wr * rr).astype(LD), phi.astype(LD), U)
            # (T r^2 phi_j - F0 phi_j(x_i)) / r is regular; the subtracted part integrates to F0 phi_j(x_i) ln(rho_max)
            reg = (T * (rr ** 2)[:, None, None])[:, None, :, :] * phi[:, :, None, None] - F0[None, None] * phi_i[None, :, None, None]
            h += wth * (np.einsum('p,pjlk->jlk', (wr / rr).astype(LD), reg.astype(CLD)) + np.log(rho_max) * phi_i[:, None, None] * F0[None])
    return h.astype(complex), g.astype(complex)


@pytest.mark.parametrize("et,m", [(shape.TRI3, 2), (shape.QUAD4, 2), (shape.QUAD9, 1)])
def test_singular_pair_integrals_against_polar_brute_force(oracle_lib, et, m):
    md = Model(cube_mesh(m, et), cube_bcs())
    o = oracle_lib.Oracle(md)
    tri = et == shape.TRI3
    for e, xi_i in ((0, np.array([0.3, 0.25]) if tri else np.array([0.2, -0.35])), (md.n_elem - 1, np.array([0.6, 0.15]) if tri else np.array([-0.55, 0.4]))):
        nodes = md.elem_node[md.elem_ptr[e]:md.elem_ptr[e + 1]]
        xn = md.node_x[nodes]
        _, x0, _, _ = surface(et, xn, xi_i[None])
        h0, g0, mode, _ = o.pair(e, x0[0], OMEGA, MAT)
        assert mode == 200
        h1, g1 = _polar_flat(et, xn, xi_i, OMEGA, MAT)
        assert np.abs(g0 - g1).max() < 1e-8 * np.abs(g1).max(), (int(et), e, np.abs(g0 - g1).max() / np.abs(g1).max())
        assert np.abs(h0 - h1).max() < 1e-8 * np.abs(h1).max(), (int(et), e, np.abs(h0 - h1).max() / np.abs(h1).max())
