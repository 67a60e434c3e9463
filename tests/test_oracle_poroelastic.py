"""CPU pins of the Biot poroelastic oracle (fbem_bem_harpor3d_*, build_lse_mechanics_bem_harpor, the ordinary-boundary
scatter of assemble_bem_harpor_equation).  No reference output exists here; the restatement is pinned by
  * the decoupled limit Q = 0, rho_a = 0, b = 0, where the solid block must equal the (independently pinned) elastodynamic u*, t* of the
    drained skeleton and the fluid block the acoustic p*, q* of a fluid with K = R, rho = rho_2 -- this pins the wavenumbers and the
    coefficient tables eta, psi, chi, W0, T1, T2, T3 entry by entry;
  * the exact 1D solution of a saturated column (two compressional waves, drained loaded top, impermeable fixed base) solved by the BEM
    with full coupling -- this pins the coupling tables vartheta, T01, T02, W1, W2, the singular integration, the free terms, the numbering
    and the scatter."""
import numpy as np
import pytest

from multifebe_b200.host import Poro, PoroModel, Material, Fluid, cube_mesh, shape
from oracle import oracle as orc


def test_decoupled_limit_reduces_to_the_elastic_and_acoustic_kernels():
    po = Poro(rhof=1.3, rhos=2.0, lam=1.5, mu=1.0, xi=0.02, phi=0.4, rhoa=0.0, R=0.9, Q=0.0, b=0.0)
    mat = Material(rho=po.rho1, mu=1.0, nu=0.5 * 1.5 / (1.5 + 1.0), xi=0.02)
    fl = Fluid(rho=po.rho2, c=1.0); fl.c = np.sqrt(po.R / po.rho2)
    rng = np.random.default_rng(3)
    for omega in (0.05, 2.3, 40.0):
        for _ in range(10):
            x_i = rng.normal(size=3); x = x_i + rng.normal(size=3) * rng.choice([0.02, 0.4, 2.0])
            n = rng.normal(size=3); n /= np.linalg.norm(n)
            u, t, k = orc.fundamental_solutions_por(x, n, x_i, omega, po)
            kk = sorted([omega * np.sqrt(po.rho1 / (po.lam + 2 * po.mu)), omega * np.sqrt(po.rho2 / po.R)], key=lambda z: z.real)
            assert abs(k[0] - kk[0]) < 1e-13 * abs(kk[0]) and abs(k[1] - kk[1]) < 1e-13 * abs(kk[1]) and abs(k[2] - omega * np.sqrt(po.rho1 / po.mu)) < 1e-13 * abs(k[2])
            assert k[3] == 0 and abs(k[4] - 1.0 / (po.rho2 * omega ** 2)) < 1e-15 * abs(k[4])
            ue, te = orc.fundamental_solutions(x, n, x_i, omega, mat)
            assert np.abs(u[1:, 1:] - ue).max() < 1e-12 * np.abs(ue).max() and np.abs(t[1:, 1:] - te).max() < 1e-11 * np.abs(te).max()
            pa, qa = orc.fundamental_solutions_pot(x, n, x_i, omega, fl)
            r = np.linalg.norm(x - x_i)
            assert abs(u[0, 0] + pa) < 1e-12 * abs(pa) and abs(t[0, 0] - k[4] * qa) < 1e-11 * abs(k[4] * pa) * (1.0 / r + abs(k[1]))
            assert not u[0, 1:].any() and not u[1:, 0].any() and not t[0, 1:].any() and not t[1:, 0].any()


def biot_column(omega, po, L=1.0, P=1.0):
    """Exact 1D saturated column: u(0) = 0, U(0) = 0 (fixed impermeable base), sigma_xx(L) = P on the skeleton, tau(L) = 0 (drained top).
    Fields y = (u, U) = sum_j (a_j e^{-i k_j x} + b_j e^{i k_j x}) y_j with M k^2 y = w^2 rhohat y, M = [[lambda+2mu+Q^2/R, Q], [Q, R]]
    (lambda, mu are the DRAINED constants of the skeleton, so Biot's solid-phase constant is lambda + Q^2/R);
    sigma_xx = M11 u' + Q U' = (lambda+2mu) u' + (Q/R) tau (stress on the skeleton), tau = Q u' + R U' (fluid equivalent stress)."""
    M = np.array([[po.lam + 2 * po.mu + po.Q ** 2 / po.R, po.Q], [po.Q, po.R]])
    rh11 = po.rho1 + po.rhoa - 1j * po.b / omega; rh12 = -po.rhoa + 1j * po.b / omega; rh22 = po.rho2 + po.rhoa - 1j * po.b / omega
    Rh = np.array([[rh11, rh12], [rh12, rh22]])
    k2, Y = np.linalg.eig(np.linalg.solve(M, omega ** 2 * Rh))
    ks = np.sqrt(k2); ks = np.where(ks.real < 0, -ks, ks)
    # unknowns a1, b1, a2, b2
    def rowsat(x):
        e = [np.exp(-1j * ks[0] * x), np.exp(1j * ks[0] * x), np.exp(-1j * ks[1] * x), np.exp(1j * ks[1] * x)]
        d = [-1j * ks[0] * e[0], 1j * ks[0] * e[1], -1j * ks[1] * e[2], 1j * ks[1] * e[3]]
        yv = [Y[:, 0], Y[:, 0], Y[:, 1], Y[:, 1]]
        u = np.array([e[q] * yv[q][0] for q in range(4)]); U = np.array([e[q] * yv[q][1] for q in range(4)])
        du = np.array([d[q] * yv[q][0] for q in range(4)]); dU = np.array([d[q] * yv[q][1] for q in range(4)])
        return u, U, M[0, 0] * du + M[0, 1] * dU, M[1, 0] * du + M[1, 1] * dU
    u0, U0, _, _ = rowsat(0.0); _, _, sL, tL = rowsat(L)
    c = np.linalg.solve(np.array([u0, U0, sL, tL]), np.array([0.0, 0.0, P, 0.0], dtype=complex))

    def field(x):
        out = [np.array([rowsat(xx)[q] @ c for xx in np.atleast_1d(x)]) for q in range(4)]
        return out          # u, U, sigma_xx, tau
    return field, ks


def column_bcs(P=1.0):
    """cube_mesh parts: 1 x=0 fixed impermeable base, 2 x=L loaded drained top, 3..6 sliding impermeable sides."""
    bcs = {1: ([1, 0, 0, 0], [0, 0, 0, 0]), 2: ([0, 1, 1, 1], [0, P, 0, 0])}
    for p_, free in ((3, 2), (4, 2), (5, 3), (6, 3)):       # normal component fixed (u_y on y-faces, u_z on z-faces), shear free
        ct = [1, 1, 1, 1]; ct[free] = 0
        bcs[p_] = (ct, [0, 0, 0, 0])
    return bcs


@pytest.mark.parametrize("et,m,b", [(shape.QUAD9, 2, 0.0), (shape.QUAD9, 2, 0.6), (shape.TRI6, 2, 0.3)])
def test_saturated_column_against_the_exact_biot_solution(et, m, b):
    po = Poro(rhof=1.0, rhos=2.2, lam=1.2, mu=1.0, xi=0.02, phi=0.35, rhoa=0.15, R=0.8, Q=0.5, b=b)
    omega = 2.0
    md = PoroModel(cube_mesh(m, et), column_bcs())
    o = orc.PorOracle(md)
    A, bb, st = o.assemble(omega, po)
    assert st["pairs_singular"] > 0 and st["pairs_adaptive"] > 0
    x = np.linalg.solve(A, bb)
    prim, sec = md.nodal_solution(x)                       # (tau, u1, u2, u3), (Un, t1, t2, t3)
    field, ks = biot_column(omega, po)
    _, _, k_or = orc.fundamental_solutions_por([1, 0, 0], [1, 0, 0], [0, 0, 0], omega, po)
    assert np.allclose(sorted(ks, key=lambda z: z.real), k_or[:2], rtol=1e-12)       # the oracle's k1, k2 are the two compressional waves
    ua, Ua, sa, ta = field(md.node_x[:, 0])
    # discretisation error of the 2 x 2 (3 x 3) quadratic meshes; it falls as h^3: 1.2e-3, 3.4e-4, 1.4e-4 for m = 2, 3, 4 with b = 0.6, and
    # is largest without dissipation (1.1e-2 at m = 2)
    tol = 2e-2 if b == 0.0 else 4e-3
    scale_u, scale_t = np.abs(ua).max(), np.abs(ta).max()
    assert np.abs(prim[:, 1] - ua).max() < tol * scale_u and np.abs(prim[:, 2:]).max() < tol * scale_u
    side = (md.node_part >= 3)
    assert np.abs(prim[side, 0] - ta[side]).max() < tol * scale_t                     # tau on the impermeable sides
    base = md.node_part == 1
    assert np.abs(prim[base, 0] - ta[base]).max() < tol * scale_t                     # tau at the impermeable base
    top = md.node_part == 2
    assert np.abs(sec[top, 0] - Ua[top]).max() < 2 * tol * np.abs(Ua).max()           # Un = U_x on the drained top (normal +x)
    assert np.abs(sec[base, 1] + sa[base]).max() < 2 * tol * np.abs(sa).max()         # t_x = -sigma_xx at the base (normal -x)
