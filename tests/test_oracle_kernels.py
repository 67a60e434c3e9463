"""Pins the oracle's pointwise arithmetic: E_m(z) (fbem_zexp_decomposed), the elastodynamic fundamental solution against
an independent closed form evaluated in extended precision, the static limit (Kelvin), the survey's sanity values, the
N(d) rule estimator's documented switch points, Telles, nearest point."""
import math, os
import numpy as np
import pytest
from multifebe_b200.host import Material, shape

LD = np.longdouble
CLD = np.clongdouble


def test_zexp_decomposed_matches_definition(oracle_lib):
    rng = np.random.default_rng(7)
    for mag in (1e-7, 1e-3, 0.3, 0.999, 1.0, 1.001, 2.5, 9.0):
        for _ in range(5):
            ph = rng.uniform(0, 2 * np.pi)
            z = mag * complex(math.cos(ph), math.sin(ph))
            E = oracle_lib.zexp_decomposed(z)
            zl = CLD(z)
            # extended-precision series for E_m = sum_{j>=m} z^j/j!
            terms = [CLD(1)]
            for j in range(1, 80):
                terms.append(terms[-1] * zl / LD(j))
            for m in range(7):
                # |z| <= 1: tail series; |z| > 1: e^z minus the head, both in extended precision
                ref = sum(terms[m:][::-1], CLD(0)) if mag <= 1.0 else np.exp(zl) - sum(terms[:m], CLD(0))
                # |z|>1: the reference forms E_m by subtraction, so the error is eps*|e^z| relative to |E_m|
                tol = 4e-16 * max(1.0, abs(np.exp(zl)) / abs(ref)) * 8
                assert abs(CLD(E[m]) - ref) <= tol * abs(ref) + 1e-300, (mag, m)


def _closed_form_u(x, xi, omega, mat):
    """G_lk = 1/(4 pi rho w^2) [ k2^2 d_lk g2 + d_l d_k (g2 - g1) ], g_j = exp(-i k_j r)/r  (time factor exp(+i w t)),
    evaluated in extended precision."""
    rv = np.array(x, dtype=LD) - np.array(xi, dtype=LD)
    r = np.sqrt((rv * rv).sum())
    dr = rv / r
    k1, k2 = CLD(omega) / CLD(mat.c1), CLD(omega) / CLD(mat.c2)

    def d2(k):  # second derivatives of exp(-ikr)/r: f'' r,l r,k + f'/r (d_lk - r,l r,k)
        g = np.exp(-1j * k * r) / r
        f1 = (-1j * k - 1 / r) * g
        f2 = g * ((-1j * k - 1 / r) ** 2 + 1 / r ** 2)
        return g, f1, f2
    g1, a1, b1 = d2(k1)
    g2, a2, b2 = d2(k2)
    G = np.zeros((3, 3), dtype=CLD)
    for l in range(3):
        for k in range(3):
            dlk = 1.0 if l == k else 0.0
            hess2 = b2 * dr[l] * dr[k] + a2 / r * (dlk - dr[l] * dr[k])
            hess1 = b1 * dr[l] * dr[k] + a1 / r * (dlk - dr[l] * dr[k])
            G[l, k] = (k2 * k2 * dlk * g2 + hess2 - hess1) / (4 * LD(np.pi) * LD(mat.rho) * LD(omega) ** 2)
    return G


def test_fundamental_solution_against_closed_form(oracle_lib):
    mat = Material(1.3, 2.0, 0.3, 0.04)
    rng = np.random.default_rng(3)
    for omega in (1.5, 4.0, 11.0):
        for _ in range(6):
            xi = rng.uniform(-1, 1, 3)
            x = xi + rng.uniform(0.4, 1.2) * (lambda v: v / np.linalg.norm(v))(rng.standard_normal(3))
            n = (lambda v: v / np.linalg.norm(v))(rng.standard_normal(3))
            u, t = oracle_lib.fundamental_solutions(x, n, xi, omega, mat)
            G = _closed_form_u(x, xi, omega, mat)
            assert np.abs(u - G.astype(np.complex128)).max() <= 2e-12 * np.abs(G).max()
            # traction of the closed form by central differences (extended precision): t_lk = sigma_kj(U_l.) n_j
            h = LD(1e-6)
            dG = np.zeros((3, 3, 3), dtype=CLD)  # dG[l,k,j] = d U_lk / d x_j
            for j in range(3):
                e = np.zeros(3, dtype=LD); e[j] = h
                dG[:, :, j] = (_closed_form_u(np.array(x, dtype=LD) + e, xi, omega, mat) - _closed_form_u(np.array(x, dtype=LD) - e, xi, omega, mat)) / (2 * h)
            lam, mu = CLD(mat.lam), CLD(mat.mu)
            T = np.zeros((3, 3), dtype=CLD)
            for l in range(3):
                div = dG[l, 0, 0] + dG[l, 1, 1] + dG[l, 2, 2]
                for k in range(3):
                    T[l, k] = lam * div * n[k] + mu * sum((dG[l, k, j] + dG[l, j, k]) * n[j] for j in range(3))
            assert np.abs(t - T.astype(np.complex128)).max() <= 1e-7 * np.abs(T).max()


def test_static_limit_is_kelvin(oracle_lib):
    mat = Material(1.0, 1.0, 0.25, 0.0)
    x, xi, n = np.array([0.3, 0.7, 0.2]), np.zeros(3), np.array([0.0, 0.0, 1.0])
    u, t = oracle_lib.fundamental_solutions(x, n, xi, 1e-7, mat)
    r = np.linalg.norm(x); dr = x / r; drdn = dr @ n; nu, mu = 0.25, 1.0
    U = np.array([[((3 - 4 * nu) * (l == k) + dr[l] * dr[k]) / (16 * np.pi * mu * (1 - nu) * r) for k in range(3)] for l in range(3)])
    T = np.array([[-(drdn * ((1 - 2 * nu) * (l == k) + 3 * dr[l] * dr[k]) + (1 - 2 * nu) * (n[l] * dr[k] - n[k] * dr[l])) / (8 * np.pi * (1 - nu) * r * r)
                   for k in range(3)] for l in range(3)])
    assert np.abs(u - U).max() < 1e-6 * np.abs(U).max()      # u* deviates by O(omega)
    assert np.abs(t - T).max() < 1e-10 * np.abs(T).max()     # t* deviates by O(omega^2)


def test_survey_sanity_values(oracle_lib):
    # approximate (1e-10) values from an independent evaluation at survey time
    mat = Material(1.0, 1.0, 0.25, 0.02)
    u, t = oracle_lib.fundamental_solutions([0.3, 0.7, 0.2], [0, 0, 1.0], [0, 0, 0.0], 2.0, mat)
    assert abs(u[0, 0] - (-0.0054864429 - 0.0698912730j)) < 2e-9
    assert abs(u[0, 1] - (0.0163814693 - 0.0076295587j)) < 2e-9
    assert abs(t[0, 0] - (-0.0306590085 + 0.0203475757j)) < 2e-9
    assert abs(t[2, 0] - (-0.0353027951 + 0.0297758886j)) < 2e-9


def test_rule_estimator_switch_points(oracle_lib):
    # SURVEY.md appendix A.3 (re = 1e-6, f = 5): gln 2 for d > 10.85, 3 for 4.28 < d < 10.85, 4 for 2.23 < d < 4.28, 5 for 2 < d < 2.23
    bx = [0.0, 0.0]
    for d, n in ((50.0, 2), (11.0, 2), (10.7, 3), (4.4, 3), (4.2, 4), (2.3, 4), (2.2, 5), (2.01, 5)):
        assert oracle_lib.qs_n(False, shape.TRI3, 5, 1e-6, d, bx) == n, d
    # monotone in d, 0 ("cannot") very close to the element, never above 30
    prev = 0
    for d in np.geomspace(1e-3, 100, 400):
        n = oracle_lib.qs_n(False, shape.QUAD9, 5, 1e-6, float(d), [0.3, -0.2])
        assert 0 <= n <= 30
        if prev and n:
            assert n <= prev
        prev = n if n else prev
    assert oracle_lib.qs_n(False, shape.QUAD9, 5, 1e-6, 0.01, [0.0, 0.0]) == 0
    # Telles variant reaches closer than the standard one
    assert oracle_lib.qs_n(True, shape.QUAD9, 5, 1e-6, 0.05, [0.0, 0.0]) > 0


def test_telles_transformation_properties(oracle_lib):
    import ctypes as C
    L = oracle_lib.lib()
    c = np.zeros(4)
    for bar_xi in (-1.0, -0.3, 0.0, 0.8, 1.0):
        for bar_r in (0.0, 0.1, 0.5, 1.0):
            L.orc_telles(C.c_int(0), C.c_double(bar_xi), C.c_double(bar_r), c.ctypes.data_as(C.c_void_p))
            f = lambda g: ((c[0] * g + c[1]) * g + c[2]) * g + c[3]
            assert abs(f(-1.0) + 1.0) < 1e-12 and abs(f(1.0) - 1.0) < 1e-12      # maps [-1,1] onto itself
            if bar_r == 1.0:
                assert np.allclose(c, [0, 0, 1, 0], atol=1e-12)                    # identity far away
    for bar_xi in (0.0, 0.25, 1.0):
        for bar_r in (0.0, 0.3, 1.0):
            L.orc_telles(C.c_int(1), C.c_double(bar_xi), C.c_double(bar_r), c.ctypes.data_as(C.c_void_p))
            f = lambda g: ((c[0] * g + c[1]) * g + c[2]) * g + c[3]
            assert abs(f(0.0)) < 1e-12 and abs(f(1.0) - 1.0) < 1e-12
    assert oracle_lib.lib().orc_telles_barr(C.c_double(5.0)) == 1.0
    assert abs(oracle_lib.lib().orc_telles_barr(C.c_double(0.5)) - 0.5 / (0.89039 * 0.5 + 0.32883)) < 1e-16


def test_nearest_point(oracle_lib):
    tri = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 0]])
    # above an interior point: Newton converges to the foot of the perpendicular
    bx, rmin, d, method = oracle_lib.nearest(shape.TRI3, tri, [0.25, 0.25, 0.1])
    assert method == 2 and np.allclose(bx, [0.25, 0.25], atol=1e-13) and abs(rmin - 0.1) < 1e-15
    # outside, nearest point on an edge (restart on the edges)
    bx, rmin, d, method = oracle_lib.nearest(shape.TRI3, tri, [0.5, -0.2, 0.0])
    assert np.allclose(bx, [0.5, 0.0], atol=1e-13) and abs(rmin - 0.2) < 1e-15
    # far away: only the node guess is used
    bx, rmin, d, method = oracle_lib.nearest(shape.TRI3, tri, [10.0, 0.0, 0.0])
    assert method == 1 and np.allclose(bx, [1.0, 0.0]) and abs(rmin - 9.0) < 1e-15
    # curved quad9 (bulged centre): still converges, distance below the flat-plate distance
    q = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0], [0, -1, 0], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, 0, 0.3]], dtype=float)
    bx, rmin, d, method = oracle_lib.nearest(shape.QUAD9, q, [0.1, 0.2, 0.8])
    assert method in (2, 3) and rmin < 0.8 - 0.2


def test_mantic_free_term(oracle_lib):
    # smooth (flat) point: c = I/2 for any number of coplanar elements around the node
    m = 6
    ang = np.arange(m) * 2 * np.pi / m
    normals = np.tile([0.0, 0.0, 1.0], (m, 1))
    tang = np.stack([np.cos(ang), np.sin(ang), 0 * ang], axis=1)
    c, err = oracle_lib.freeterm(normals, tang, 0.3)
    assert err == 0 and np.allclose(c, 0.5 * np.eye(3), atol=1e-14)
    # cube corner seen from inside the solid (outward normals -x,-y,-z): solid angle 4pi/8 -> trace(c) = 3/8
    normals = np.array([[-1.0, 0, 0], [0, -1.0, 0], [0, 0, -1.0]])
    # forward boundary tangents (counter-clockwise around each outward normal) of the three faces meeting at the origin
    tang = np.array([[0, 0, 1.0], [1.0, 0, 0], [0, 1.0, 0]])
    c, err = oracle_lib.freeterm(normals, tang, 0.25)
    assert err == 0
    assert abs(np.trace(c) - 3 * 0.125) < 1e-13
    assert np.allclose(c, c.T, atol=1e-14)
