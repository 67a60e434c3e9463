"""Displacements at interior points (displacement part): Somigliana's identity u(x) = sum_e (g t - h u)
with the same integrators as the boundary equations (src/calculate_internal_points_mechanics_bem_harela.f90:160-178, :372-420).
CPU: the composition from oracle pair integrals against exact solutions.  GPU: the library against that composition."""
import numpy as np
import pytest
from multifebe_b200.host import Model, Material, InternalPointsModel, cube_mesh, cube_bcs, column_analytic_u, shape

MAT = Material(1.0, 1.0, 0.25, 0.03)
SMAT = Material(1.0, 1.3, 0.25, 0.0)
PTS = np.array([[0.5, 0.5, 0.5], [0.2, 0.7, 0.4], [0.93, 0.5, 0.5], [0.31, 0.08, 0.77], [0.5, 0.5, 0.985]])   # two of them close to the boundary


def oracle_interior_u(o, md, x, pts, pair):
    u, t = md.nodal_solution(x)
    out = np.zeros((len(pts), 3), dtype=np.complex128)
    for ip, xp in enumerate(pts):
        for e in range(md.n_elem):
            h, g = pair(e, xp)
            nodes = md.mesh.conn[e]
            out[ip] += np.einsum("jlk,jk->l", g, t[nodes]) - np.einsum("jlk,jk->l", h, u[nodes])
    return out


def test_interior_points_model_layout():
    md = Model(cube_mesh(2, shape.TRI3), cube_bcs())
    ipm = InternalPointsModel(md, PTS)
    assert ipm.n_dof == md.n_dof + 15 and ipm.n_node == md.n_node + 5 and ipm.n_colloc == 5
    assert np.all(ipm.colloc_elem == -1) and np.all(ipm.colloc_node >= md.n_node)
    assert np.array_equal(ipm.row[:md.n_node], md.row) and ipm.row[md.n_node:].min() == md.n_dof
    assert len(set(ipm.row.ravel().tolist())) == ipm.n_dof            # every row owned once


def test_static_interior_displacements_exact(oracle_lib):
    md = Model(cube_mesh(3, shape.QUAD4), cube_bcs())
    o = oracle_lib.Oracle(md)
    A, b, _ = o.assemble_static(SMAT)
    x, _, _ = oracle_lib.lu_solve_real(A, b)
    ui = oracle_interior_u(o, md, x, PTS, lambda e, xp: o.pair_static(e, xp, SMAT)[:2])
    lam2mu = 2.0 * SMAT.mu_r * SMAT.nu_r / (1.0 - 2.0 * SMAT.nu_r) + 2.0 * SMAT.mu_r
    assert np.abs(ui[:, 0].real - PTS[:, 0] / lam2mu).max() * lam2mu < 5e-5 and np.abs(ui[:, 1:]).max() * lam2mu < 5e-5


def test_harmonic_interior_displacements_follow_the_column_solution(oracle_lib):
    md = Model(cube_mesh(5, shape.QUAD9), cube_bcs())
    o = oracle_lib.Oracle(md)
    omega = 2.0
    A, b, _ = o.assemble(omega, MAT)
    x, _, _ = oracle_lib.lu_solve(A, b)
    ui = oracle_interior_u(o, md, x, PTS[:3], lambda e, xp: o.pair(e, xp, omega, MAT)[:2])
    ua = column_analytic_u(PTS[:3, 0], omega, MAT)
    assert np.abs(ui[:, 0] - ua).max() < 2e-3 * np.abs(ua).max()


@pytest.mark.gpu
@pytest.mark.parametrize("et,m", [(shape.TRI3, 4), (shape.QUAD9, 2), (shape.QUAD4, 3)])
def test_gpu_interior_displacements_match_the_oracle_composition(gpu_ctx, oracle_lib, et, m):
    from multifebe_b200 import capi
    md = Model(cube_mesh(m, et), cube_bcs())
    pr = capi.Problem(gpu_ctx, md)
    ip = capi.InternalPoints(gpu_ctx, md, PTS)
    o = oracle_lib.Oracle(md)
    omega = 3.0
    x = pr.solve_frequency(omega, MAT)
    ug = ip.displacements(omega, MAT, x)
    uo = oracle_interior_u(o, md, x, PTS, lambda e, xp: o.pair(e, xp, omega, MAT)[:2])
    assert np.abs(ug - uo).max() < 1e-10 * np.abs(uo).max()
    xs = pr.solve_static(SMAT)
    us = ip.displacements_static(SMAT, xs)
    uso = oracle_interior_u(o, md, xs.astype(np.complex128), PTS, lambda e, xp: o.pair_static(e, xp, SMAT)[:2])
    assert np.abs(us - uso.real).max() < 1e-10 * np.abs(uso).max()
    lam2mu = 2.0 * SMAT.mu_r * SMAT.nu_r / (1.0 - 2.0 * SMAT.nu_r) + 2.0 * SMAT.mu_r
    if et != shape.TRI3 or m >= 4:
        assert np.abs(us[:, 0] - PTS[:, 0] / lam2mu).max() * lam2mu < 1e-4
    ip.close(); pr.close()


# ---- stresses at interior points: the hypersingular kernels d*, s* of the oracle (fbem_bem_harela3d_hbie_*, exterior branches) ----
def test_hbie_kernels_are_the_traction_operator_applied_to_u_and_t(oracle_lib):
    """d*_lk = sigma-operator (at the collocation point, normal n_i) of u*_.k, s*_lk the same of t*_.k: finite differences of the
    (independently pinned) u*, t* with respect to x_i pin the d*, s* formulas and the S1..S5 coefficient tables."""
    rng = np.random.default_rng(4)
    lam, mu = MAT.lam, MAT.mu
    for omega in (0.3, 2.0, 9.0):
        for _ in range(4):
            x_i = rng.uniform(-0.5, 0.5, 3); x = x_i + rng.uniform(0.3, 1.5) * rng.standard_normal(3)
            n = rng.standard_normal(3); n /= np.linalg.norm(n); n_i = rng.standard_normal(3); n_i /= np.linalg.norm(n_i)
            d, s = oracle_lib.fundamental_solutions_hbie(x, n, x_i, n_i, omega, MAT)
            hstep = 1e-5
            du = np.zeros((3, 3, 3), dtype=complex); dt = np.zeros((3, 3, 3), dtype=complex)     # [m][l][k] = d/dx_i,m of u*_lk
            for m in range(3):
                e = np.zeros(3); e[m] = hstep
                up, tp = oracle_lib.fundamental_solutions(x, n, x_i + e, omega, MAT); um, tm = oracle_lib.fundamental_solutions(x, n, x_i - e, omega, MAT)
                du[m] = (up - um) / (2 * hstep); dt[m] = (tp - tm) / (2 * hstep)

            def sigma_op(df):
                out = np.zeros((3, 3), dtype=complex)
                for l in range(3):
                    for k in range(3):
                        out[l, k] = lam * n_i[l] * sum(df[p, p, k] for p in range(3)) + mu * sum(n_i[m] * (df[m, l, k] + df[l, m, k]) for m in range(3))
                return out
            assert np.abs(d - sigma_op(du)).max() < 2e-6 * np.abs(d).max()
            assert np.abs(s - sigma_op(dt)).max() < 2e-6 * np.abs(s).max()


def oracle_interior_stress(o, md, x, pts, omega, mat):
    u, t = md.nodal_solution(x)
    sig = np.zeros((len(pts), 3, 3), dtype=np.complex128)       # [point][l][kc]: traction component l on the plane with normal e_kc
    for ip, xp in enumerate(pts):
        for kc in range(3):
            n_i = np.zeros(3); n_i[kc] = 1.0
            for e in range(md.n_elem):
                m, l, mode = o.pair_hbie(e, xp, n_i, omega, mat)
                nodes = md.mesh.conn[e]
                sig[ip, :, kc] += np.einsum("jlk,jk->l", l, t[nodes]) - np.einsum("jlk,jk->l", m, u[nodes])
    return sig


def test_harmonic_interior_stresses_follow_the_column_solution(oracle_lib):
    md = Model(cube_mesh(5, shape.QUAD9), cube_bcs())
    o = oracle_lib.Oracle(md)
    omega = 2.0
    A, b, _ = o.assemble(omega, MAT)
    x, _, _ = oracle_lib.lu_solve(A, b)
    pts = PTS[:2]
    sig = oracle_interior_stress(o, md, x, pts, omega, MAT)
    k = omega / MAT.c1
    s11 = np.cos(k * pts[:, 0]) / np.cos(k * 1.0)                # P cos(k x)/cos(k L)
    s22 = MAT.lam / (MAT.lam + 2 * MAT.mu) * s11
    assert np.abs(sig[:, 0, 0] - s11).max() < 5e-3 and np.abs(sig[:, 1, 1] - s22).max() < 5e-3 and np.abs(sig[:, 2, 2] - s22).max() < 5e-3
    off = sig.copy(); off[:, 0, 0] = 0; off[:, 1, 1] = 0; off[:, 2, 2] = 0
    assert np.abs(off).max() < 5e-3


@pytest.mark.gpu
@pytest.mark.parametrize("et,m", [(shape.TRI3, 4), (shape.QUAD9, 2), (shape.QUAD8, 2)])
def test_gpu_interior_stresses_match_the_oracle_composition(gpu_ctx, oracle_lib, et, m):
    from multifebe_b200 import capi
    md = Model(cube_mesh(m, et), cube_bcs())
    pr = capi.Problem(gpu_ctx, md)
    ip = capi.InternalPoints(gpu_ctx, md, PTS)
    o = oracle_lib.Oracle(md)
    omega = 3.0
    x = pr.solve_frequency(omega, MAT)
    sg = ip.stresses(omega, MAT, x)
    so = oracle_interior_stress(o, md, x, PTS, omega, MAT)
    assert np.abs(sg - so).max() < 1e-10 * np.abs(so).max()
    ug = ip.displacements(omega, MAT, x)            # both problems live side by side
    uo = oracle_interior_u(o, md, x, PTS, lambda e, xp: o.pair(e, xp, omega, MAT)[:2])
    assert np.abs(ug - uo).max() < 1e-10 * np.abs(uo).max()
    ip.close(); pr.close()


def test_stress_model_layout():
    md = Model(cube_mesh(2, shape.TRI3), cube_bcs())
    ipm = InternalPointsModel(md, PTS, stress=True)
    assert ipm.n_colloc == 15 and ipm.n_dof == md.n_dof + 45 and ipm.colloc_n.shape == (15, 3)
    assert np.allclose(ipm.colloc_x[0], ipm.colloc_x[2]) and np.allclose(ipm.colloc_n[:3], np.eye(3))


def test_static_interior_stresses_exact(oracle_lib):
    """ME-ST-EL-002's exact solution: sigma_11 = P = 1, sigma_22 = sigma_33 = nu/(1-nu) = 1/3 (docs/examples/ME-ST-EL-002/doc_src/
    ME-ST-EL-002.tex:29-44), at interior points, from the static hypersingular identity."""
    md = Model(cube_mesh(3, shape.QUAD4), cube_bcs())
    o = oracle_lib.Oracle(md)
    A, b, _ = o.assemble_static(SMAT)
    x, _, _ = oracle_lib.lu_solve_real(A, b)
    u, t = md.nodal_solution(x)
    for xp in PTS[:3]:
        sig = np.zeros((3, 3))
        for kc in range(3):
            n_i = np.zeros(3); n_i[kc] = 1.0
            for e in range(md.n_elem):
                m, l, mode = o.pair_hbie_static(e, xp, n_i, SMAT)
                nodes = md.mesh.conn[e]
                sig[:, kc] += np.einsum("jlk,jk->l", l, t[nodes].real) - np.einsum("jlk,jk->l", m, u[nodes].real)
        assert np.abs(sig - np.diag([1.0, 1.0 / 3.0, 1.0 / 3.0])).max() < 2e-4


@pytest.mark.gpu
def test_gpu_static_interior_stresses(gpu_ctx, oracle_lib):
    from multifebe_b200 import capi
    md = Model(cube_mesh(3, shape.QUAD4), cube_bcs())
    pr = capi.Problem(gpu_ctx, md)
    ip = capi.InternalPoints(gpu_ctx, md, PTS)
    o = oracle_lib.Oracle(md)
    xs = pr.solve_static(SMAT)
    sg = ip.stresses_static(SMAT, xs)
    u, t = md.nodal_solution(xs)
    so = np.zeros_like(sg)
    for ipt, xp in enumerate(PTS):
        for kc in range(3):
            n_i = np.zeros(3); n_i[kc] = 1.0
            for e in range(md.n_elem):
                m, l, mode = o.pair_hbie_static(e, xp, n_i, SMAT)
                nodes = md.mesh.conn[e]
                so[ipt, :, kc] += np.einsum("jlk,jk->l", l, t[nodes].real) - np.einsum("jlk,jk->l", m, u[nodes].real)
    assert np.abs(sg - so).max() < 1e-10 * np.abs(so).max()
    assert np.abs(sg[:3] - np.diag([1.0, 1.0 / 3.0, 1.0 / 3.0])).max() < 2e-4
    ip.close(); pr.close()
