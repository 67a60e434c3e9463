"""CPU pins of the static-elasticity restatement of the oracle (fbem_bem_staela3d_sbie_*,
build_lse_mechanics_bem_staela, assemble_bem_staela_equation, solve_lse_r).  The reference ships no numeric golden vectors
for this path either; what pins it: Kelvin's closed form, the static limit of the (independently pinned) harmonic oracle, the
exact solution of the reference's tutorial ME-ST-EL-002 (docs/examples/ME-ST-EL-002/doc_src/ME-ST-EL-002.tex:29-44:
u1 = P x1 / (lambda + 2 mu), which linear elements reproduce to quadrature error) and the rigid-body identity (H + C) 1 = 0."""
import numpy as np
import pytest
from multifebe_b200.host import Model, Material, cube_mesh, cube_bcs, shape

MAT = Material(1.0, 1.3, 0.25, 0.0)


def test_kelvin_closed_form(oracle_lib):
    rng = np.random.default_rng(3)
    for _ in range(20):
        x_i = rng.uniform(-1, 1, 3); x = x_i + rng.uniform(0.2, 2.0) * rng.standard_normal(3)
        n = rng.standard_normal(3); n /= np.linalg.norm(n)
        u, t = oracle_lib.fundamental_solutions_static(x, n, x_i, MAT)
        rv = x - x_i; r = np.linalg.norm(rv); q = rv / r; mu, nu = MAT.mu_r, MAT.nu_r
        uk = ((3 - 4 * nu) * np.eye(3) + np.outer(q, q)) / (16 * np.pi * mu * (1 - nu) * r)
        drdn = q @ n
        tk = -(((1 - 2 * nu) * np.eye(3) + 3 * np.outer(q, q)) * drdn + (1 - 2 * nu) * (np.outer(n, q) - np.outer(q, n))) / (8 * np.pi * (1 - nu) * r * r)
        assert np.abs(u - uk).max() < 1e-15 * np.abs(uk).max() * 50 and np.abs(t - tk).max() < 1e-15 * np.abs(tk).max() * 50


@pytest.mark.parametrize("et,m", [(shape.TRI3, 3), (shape.QUAD9, 1), (shape.QUAD4, 2)])
def test_static_pairs_are_the_zero_frequency_limit_of_the_harmonic_oracle(oracle_lib, et, m):
    md = Model(cube_mesh(m, et), cube_bcs())
    o = oracle_lib.Oracle(md)
    om = 1e-5       # the harmonic kernel differs from Kelvin by O(omega) in u* (constant term psi(2)) and O(omega^2) in t*
    for e in (0, md.n_elem // 2, md.n_elem - 1):
        for c in (0, md.n_colloc // 3, md.n_colloc - 1):
            hs, gs, mode_s = o.pair_static(e, md.colloc_x[c], MAT)
            hh, gh, mode_h, _ = o.pair(e, md.colloc_x[c], om, MAT)
            assert mode_s == mode_h                          # same plan: the rule choice does not depend on the kernel
            assert np.abs(hh.imag).max() < 1e-6 and np.abs(hs - hh.real).max() < 1e-8 * max(np.abs(hs).max(), 1.0)
            assert np.abs(gs - gh.real).max() < 1e-8 * max(np.abs(gs).max(), 1.0)


@pytest.mark.parametrize("et,m", [(shape.TRI3, 2), (shape.QUAD4, 2), (shape.TRI6, 1), (shape.QUAD8, 1), (shape.QUAD9, 1)])
def test_static_column_exact_solution(oracle_lib, et, m):
    md = Model(cube_mesh(m, et), cube_bcs())
    A, b, st = oracle_lib.Oracle(md).assemble_static(MAT)
    x, _, _ = oracle_lib.lu_solve_real(A, b)
    u, t = md.nodal_solution(x)
    lam2mu = (2.0 * MAT.mu_r * MAT.nu_r / (1.0 - 2.0 * MAT.nu_r)) + 2.0 * MAT.mu_r
    ue = md.node_x[:, 0] / lam2mu
    assert np.abs(u[:, 0].real - ue).max() < 2e-5 * np.abs(ue).max()           # quadrature error (qsi_relative_error = 1e-6)
    assert np.abs(u[:, 1:]).max() < 2e-5 * np.abs(ue).max()
    assert st["pairs_singular"] > 0 and st["pts_regular"] > 0


def test_static_rigid_body_identity(oracle_lib):
    """All-traction-known cube: A = H + C acts on u; a rigid translation gives (H + C) 1 = 0 exactly in statics."""
    bcs = {p: ([1, 1, 1], [0, 0, 0]) for p in range(1, 7)}
    md = Model(cube_mesh(3, shape.TRI3), bcs)
    A, b, _ = oracle_lib.Oracle(md).assemble_static(MAT)
    for k in range(3):
        v = np.zeros(md.n_dof); v[md.col_u[:, k]] = 1.0
        assert np.abs(A @ v).max() < 5e-6 * np.abs(A).max()


def test_static_oracle_reproduces_the_committed_vectors(oracle_lib):
    """tests/golden/oracle_static.npz (tools/gen_golden.py static): regression vectors of the static oracle."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_static.npz"))
    for et, m in [(shape.TRI3, 2), (shape.QUAD9, 1)]:
        md = Model(cube_mesh(m, et), cube_bcs())
        o = oracle_lib.Oracle(md)
        A, b, _ = o.assemble_static(MAT)
        assert np.abs(A - gold[f"A:{et}:{m}"]).max() <= 1e-13 * np.abs(A).max() and np.abs(b - gold[f"b:{et}:{m}"]).max() <= 1e-13 * np.abs(b).max()
        for key in ("reg", "adp", "sing"):
            v = gold[f"pair:{et}:{key}"]; c, e, mode = int(v[0]), int(v[1]), int(v[2]); nn = md.elem_ptr[e + 1] - md.elem_ptr[e]
            h, g, mode2 = o.pair_static(e, md.colloc_x[c], MAT)
            assert mode2 == mode and np.allclose(h.ravel(), v[3:3 + 9 * nn], rtol=0, atol=1e-13 * np.abs(v[3:]).max())
