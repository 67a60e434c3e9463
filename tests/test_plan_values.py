"""The product's VALUE geometry (multifebe_b200/csrc/plan_values.cpp: independent derivations, see its header) against the oracle's restatement of
the reference, piece by piece and without a GPU: free-term geometry through the host-only ABI entry mfb_freeterm_terms.  (Point sets, rays, line integrals
and Telles coefficients are exercised through the pair integrals of tests/test_por_pair_host.py and the GPU parity suite.)"""
import numpy as np
import pytest
from multifebe_b200 import capi


def _corner(normals, start_tangents):
    return np.array(normals, dtype=float), np.array(start_tangents, dtype=float)


def _cases():
    s = 1 / np.sqrt(2.0)
    # convex cube corner at the origin of the octant x, y, z >= 0 ... the body occupies x, y, z <= 0 side: outward normals +x, +y, +z;
    # each face's boundary edge that leaves the node counter-clockwise (seen from outside) is listed with it
    cases = {
        "flat, four elements": _corner([[0, 0, 1]] * 4, [[1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, -1, 0]]),
        "flat, three unequal sectors": _corner([[0, 0, 1]] * 3, [[1, 0, 0], [-s, s, 0], [0, -1, 0]]),
        "convex edge (90 deg)": _corner([[0, 0, 1], [0, 0, 1], [1, 0, 0], [1, 0, 0]], [[0, 1, 0], [-1, 0, 0], [0, -1, 0], [0, 0, -1]]),
        "convex cube corner": _corner([[1, 0, 0], [0, 1, 0], [0, 0, 1]], [[0, -1, 0], [0, 0, -1], [-1, 0, 0]]),
        "concave cube corner": _corner([[1, 0, 0], [0, 1, 0], [0, 0, 1]], [[0, 1, 0], [0, 0, 1], [1, 0, 0]]),     # the body is everything but the octant x, y, z >= 0
    }
    return cases


def _closed_polyhedral(normals, tangents):
    """consistency: every tangent lies in its own face and in the previous face of the chain (so the input really is a closed fan)"""
    n, t = normals, tangents
    return all(abs(n[i] @ t[i]) < 1e-12 for i in range(len(n)))


@pytest.mark.parametrize("name", list(_cases().keys()))
def test_free_term_geometry_matches_the_oracle(oracle_lib, name):
    n, t = _cases()[name]
    assert _closed_polyhedral(n, t)
    for nu in (0.0, 0.25, 0.3 + 0.0j, 0.45):
        c_or, err = oracle_lib.freeterm(n, t, nu)
        if err:
            pytest.skip("the oracle rejects this fan (its input convention differs): %s" % name)
        c_pr, cp = capi.freeterm(n, t, nu)
        assert np.abs(c_pr - c_or).max() < 1e-13, (name, nu, c_pr, c_or)
        assert np.abs(c_pr - c_pr.T).max() < 1e-15          # the independent derivation symmetrises by construction
        assert abs(cp - {"flat, four elements": 0.5, "flat, three unequal sectors": 0.5, "convex edge (90 deg)": 0.25, "convex cube corner": 0.125,
                         "concave cube corner": 0.875}[name]) < 1e-14        # interior solid angle / 4 pi


def _values_lib():
    import ctypes as C
    L = capi.lib()
    return L, C


def test_edge_line_integrals_closed_form_graded_panels_and_brute_force_agree():
    """plan_values.cpp edge_integrals: straight edges in closed form, curved (3-node) edges with graded Gauss panels.  A straight quadratic edge pushed
    through the curved branch (mid node moved by 1e-9 of the chord, far above the straightness switch, far below the accuracy asked) must reproduce the
    closed form; a genuinely curved edge is compared with a dense composite Gauss rule in numpy."""
    L, C = _values_lib()
    f = getattr(L, "_ZN4mfbh14edge_integralsEiPKdS1_PKbPd")
    f.restype = None
    rng = np.random.default_rng(2)

    def run(et, xn, x_i, edges):
        hli = np.zeros(9)
        on = np.array(edges, dtype=np.bool_)
        f(C.c_int(et), xn.ctypes.data_as(C.c_void_p), x_i.ctypes.data_as(C.c_void_p), on.ctypes.data_as(C.c_void_p), hli.ctypes.data_as(C.c_void_p))
        return hli

    # quad9 in the plane z = 0.2 x (flat), nodes of the reference order
    qs = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1], [0, -1], [1, 0], [0, 1], [-1, 0], [0, 0]], dtype=float)
    xn = np.stack([0.7 * qs[:, 0] + 0.1 * qs[:, 1], 0.9 * qs[:, 1], 0.2 * qs[:, 0]], 1).copy()
    for x_i in (xn[8] + np.array([0.11, -0.23, 0.022]), xn[0].copy(), 0.5 * (xn[0] + xn[1])):
        on = [True] * 4
        if np.allclose(x_i, xn[0]):
            on = [False, True, True, False]
        elif np.allclose(x_i, 0.5 * (xn[0] + xn[1])):
            on = [False, True, True, True]
        h_straight = run(9, xn, x_i, on)
        xb = xn.copy(); xb[4:8] += 1e-9 * rng.standard_normal((4, 3))          # every edge now takes the curved branch
        h_curved = run(9, xb, x_i, on)
        assert np.abs(h_curved - h_straight).max() < 5e-8 * np.abs(h_straight).max() + 1e-8   # the perturbation itself moves the integral by ~1e-9
        assert np.abs(h_straight + h_straight.reshape(3, 3).T.ravel()).max() < 1e-15            # antisymmetric
    # a really curved edge: compare with a dense composite rule
    xc = xn.copy(); xc[4] += np.array([0.0, -0.12, 0.05])
    x_i = xc[8] + np.array([0.05, -0.1, 0.0])
    got = run(9, xc, x_i, [True, False, False, False])
    A, B, M = xc[0], xc[1], xc[4]
    gx, gw = np.polynomial.legendre.leggauss(40)
    I = np.zeros(3)
    edges = np.linspace(-1, 1, 65)
    for a, b in zip(edges[:-1], edges[1:]):
        u = 0.5 * (a + b) + 0.5 * (b - a) * gx
        x = M[None] + u[:, None] * (0.5 * (B - A))[None] + (u ** 2)[:, None] * (0.5 * (A + B - 2 * M))[None]
        dx = (0.5 * (B - A))[None] + 2 * u[:, None] * (0.5 * (A + B - 2 * M))[None]
        I += ((0.5 * (b - a) * gw)[:, None] * dx / np.linalg.norm(x - x_i[None], axis=1)[:, None]).sum(0)
    ref = np.zeros(9); ref[1] = -I[2]; ref[2] = I[1]; ref[5] = -I[0]; ref[3] = I[2]; ref[6] = -I[1]; ref[7] = I[0]
    assert np.abs(got - ref).max() < 1e-13 * np.abs(ref).max()


def test_free_term_of_a_flat_fan_in_general_position(oracle_lib):
    """A node inside a flat face has the free term c = 1/2 I whatever the orientation of the face.  The reference's formula
    (fbem_bem_harela3d_sbie_freeterm, bem_harela3d.f90:478-492: sum of acos(n_i . n_i+1) with a sign from (n_i x n_i+1) . t) is ill-conditioned exactly there:
    in general position the normals of coplanar neighbours agree only up to rounding, n . n' = 1 - O(1e-16), and acos makes O(1e-8) of it.  The oracle
    restates that formula and shows the noise; the product's geometry (csrc/plan_values.cpp: angles from atan2 of triple products) does not.  This is
    why parity of the free-term blocks on rotated meshes is asserted at 5e-8 (tests/test_gpu_local_axes.py) and at 1e-11 everywhere else."""
    from multifebe_b200 import capi
    rng = np.random.default_rng(7)
    worst_o = worst_p = 0.0
    for _ in range(40):
        Q, _r = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(Q) < 0:
            Q[:, 0] = -Q[:, 0]
        ne = int(rng.integers(3, 9))
        ang = np.sort(rng.uniform(0, 2 * np.pi, ne))
        ang = ang + np.linspace(0, 1e-3, ne)                        # distinct
        # element i spans the sector [ang_i, ang_i+1) of the plane z = 0; its boundary tangent at the node points along its first edge
        n_loc = np.tile([0.0, 0.0, 1.0], (ne, 1)); t_loc = np.stack([np.cos(ang), np.sin(ang), np.zeros(ne)], axis=1)
        n = n_loc @ Q.T; t = t_loc @ Q.T
        # rounding as in a real mesh: normals come from cross products of slightly different edge vectors
        n = n + rng.normal(scale=1e-16, size=n.shape); n /= np.linalg.norm(n, axis=1)[:, None]
        perm = rng.permutation(ne)
        cp_, _ = capi.freeterm(n[perm], t[perm], 0.3)
        co, err = oracle_lib.freeterm(n[perm], t[perm], 0.3)
        assert err == 0
        worst_p = max(worst_p, np.abs(cp_ - 0.5 * np.eye(3)).max()); worst_o = max(worst_o, np.abs(co - 0.5 * np.eye(3)).max())
    assert worst_p < 1e-13, worst_p
    assert worst_o < 1e-6, worst_o          # the reference formula: right, but only to the square root of the rounding of n . n'
