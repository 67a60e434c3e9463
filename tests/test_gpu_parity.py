"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI (multifebe_b200.capi -> libmfb.so),
against the CPU oracle on the same inputs.  Tolerances are BASELINE.json's: assembled H/G entries within 1e-11 relative
(max-norm), solutions within 1e-8 relative."""
import os
import numpy as np
import pytest
from multifebe_b200.host import Model, Material, cube_mesh, cube_bcs, halfspace_patch, column_analytic_u, shape

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
MAT = Material(1.0, 1.0, 0.25, 0.03)
TOL_A, TOL_X = 1e-11, 1e-8


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("et,m", [(shape.TRI3, 3), (shape.TRI6, 2), (shape.QUAD4, 3), (shape.QUAD8, 2), (shape.QUAD9, 2)])
@pytest.mark.parametrize("omega", [0.4, 6.0])
def test_assembly_and_solution_parity(gpu_ctx, oracle_lib, et, m, omega):
    from multifebe_b200 import capi
    md = Model(cube_mesh(m, et), cube_bcs())
    pr = capi.Problem(gpu_ctx, md)
    A, b = pr.build_lse_mechanics_bem_harela(omega, MAT)
    Ao, bo, st = oracle_lib.Oracle(md).assemble(omega, MAT)
    assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A
    s = pr.stats()
    assert s["PAIRS_REGULAR"] == sum(st["pairs_regular"].values()) and s["POINTS_REGULAR"] == st["pts_regular"]
    assert s["PAIRS_ADAPTIVE"] == st["pairs_adaptive"] and s["LEAVES"] == st["leaves"] and s["POINTS_ADAPTIVE"] == st["pts_adaptive"]
    assert s["PAIRS_SINGULAR"] == st["pairs_singular"] and s["POINTS_SINGULAR"] == st["pts_singular"]
    xo, _, _ = oracle_lib.lu_solve(Ao, bo)
    x1 = pr.solve_lse_c(A.copy(order="F"), b)             # seam 2 with host arrays (zgesv semantics)
    x2 = pr.solve_frequency(omega, MAT)                   # fused, device resident
    assert relerr(x1, xo) < TOL_X and relerr(x2, xo) < TOL_X
    assert s["LAUNCHES"] >= 3
    pr.close()


@pytest.mark.parametrize("et,m", [(shape.TRI3, 2), (shape.TRI6, 1), (shape.QUAD4, 2), (shape.QUAD8, 1), (shape.QUAD9, 1)])
def test_against_committed_golden_vectors(gpu_ctx, et, m):
    from multifebe_b200 import capi
    gold = np.load(os.path.join(HERE, "golden", "oracle_pairs.npz"))
    md = Model(cube_mesh(m, et), cube_bcs())
    pr = capi.Problem(gpu_ctx, md)
    A, b = pr.build_lse_mechanics_bem_harela(4.0, MAT)
    assert relerr(A, gold[f"A:{et}:{m}:4.0"]) < TOL_A and relerr(b, gold[f"b:{et}:{m}:4.0"]) < TOL_A
    for om in (0.7, 4.0):
        assert relerr(pr.solve_frequency(om, MAT), gold[f"x:{et}:{m}:{om}"]) < TOL_X
    pr.close()


def test_plan_decisions_match_the_oracle(gpu_ctx, oracle_lib):
    from multifebe_b200 import capi
    rng = np.random.default_rng(5)
    for et, m in [(shape.TRI3, 6), (shape.QUAD9, 3), (shape.QUAD4, 5)]:
        mesh = cube_mesh(m, et)
        # jitter interior nodes so that distances are generic (seed 12345, 0.1 cell)
        md0 = Model(mesh, cube_bcs())
        jit = np.random.default_rng(12345).uniform(-1, 1, mesh.nodes.shape) * (0.1 / m)
        for v in range(len(mesh.nodes)):
            if not md0.in_boundary[v] and et != shape.QUAD9:
                p = int(md0.node_part[v]); ax = (p - 1) // 2
                jit[v, ax] = 0.0
                mesh.nodes[v] += jit[v]
        md = Model(mesh, cube_bcs())
        pr = capi.Problem(gpu_ctx, md); o = oracle_lib.Oracle(md)
        cs = rng.integers(0, md.n_colloc, 1500); es = rng.integers(0, md.n_elem, 1500)
        got = pr.plan_modes(cs, es)
        exp = np.array([o.pair_mode(int(e), md.colloc_x[int(c)])[0] for c, e in zip(cs, es)])
        assert np.array_equal(got, exp)
        assert len(set(exp.tolist())) >= 4
        pr.close()


def test_nondefault_settings_reversed_boundary_and_open_patch(gpu_ctx, oracle_lib):
    from multifebe_b200 import capi
    cases = [
        Model(cube_mesh(2, shape.QUAD8), cube_bcs(), qsi_relative_error=1e-4, qsi_ns_max=3, precalset_gln=(2, 4, 6)),
        Model(cube_mesh(2, shape.TRI3), cube_bcs(), reversed_parts=(1, 2, 3, 4, 5, 6), qsi_relative_error=1e-8),
        Model(halfspace_patch(4, shape.QUAD9), {1: ([1, 1, 1], [0, 0, 0]), 2: ([0, 0, 0], [0, 0, 1.0])}),
        Model(halfspace_patch(5, shape.TRI6), {1: ([1, 1, 1], [0.1, 0, 0.3j]), 2: ([0, 1, 0], [1.0, 0.5, 0.2])}),
    ]
    for md in cases:
        pr = capi.Problem(gpu_ctx, md)
        A, b = pr.build_lse_mechanics_bem_harela(2.5, MAT)
        Ao, bo, _ = oracle_lib.Oracle(md).assemble(2.5, MAT)
        assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A
        xo, _, _ = oracle_lib.lu_solve(Ao, bo)
        assert relerr(pr.solve_frequency(2.5, MAT), xo) < TOL_X
        pr.close()


def test_mixed_element_types_in_one_region(gpu_ctx, oracle_lib):
    from multifebe_b200 import capi
    from multifebe_b200.host import Mesh
    a, b_ = cube_mesh(2, shape.TRI3), cube_mesh(2, shape.QUAD4)
    # faces 1-3 from the triangle mesh, faces 4-6 from the quad mesh (each face owns its nodes, so they combine freely)
    nodes, et, part, conn = [], [], [], []
    for src, parts in ((a, (1, 2, 3)), (b_, (4, 5, 6))):
        off = len(nodes); nodes += list(src.nodes)
        for k in range(src.n_elem):
            if int(src.part[k]) in parts:
                et.append(int(src.etype[k])); part.append(int(src.part[k])); conn.append(src.conn[k] + off)
    used = sorted(set(int(v) for c in conn for v in c)); remap = {v: i for i, v in enumerate(used)}
    mesh = Mesh(np.array(nodes)[used], et, part, [[remap[int(v)] for v in c] for c in conn])
    md = Model(mesh, cube_bcs())
    pr = capi.Problem(gpu_ctx, md)
    A, b = pr.build_lse_mechanics_bem_harela(1.7, MAT)
    Ao, bo, _ = oracle_lib.Oracle(md).assemble(1.7, MAT)
    assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A
    pr.close()


def test_zgemm_on_the_fp64_tensor_pipe(gpu_ctx):
    rng = np.random.default_rng(1)
    for (m, n, k) in [(128, 64, 16), (200, 130, 36), (512, 512, 128), (37, 5, 2), (1000, 333, 128)]:
        A = rng.standard_normal((m, k)) + 1j * rng.standard_normal((m, k))
        B = rng.standard_normal((k, n)) + 1j * rng.standard_normal((k, n))
        Cm = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
        R, ms = gpu_ctx.zgemm_minus(Cm, A, B)
        assert relerr(R, Cm - A @ B) < 1e-13


def _problem_of_size(gpu_ctx, n_target):
    """A Problem whose n_dof is used only as the order of a standalone linear system (seam 2)."""
    from multifebe_b200 import capi
    m = 1
    while 18 * (m + 1) ** 2 < n_target:
        m += 1
    md = Model(cube_mesh(m, shape.TRI3), cube_bcs())
    return capi.Problem(gpu_ctx, md), md.n_dof


@pytest.mark.parametrize("n_target", [18 * 4, 18 * 36, 18 * 100])
def test_lu_against_lapack(gpu_ctx, oracle_lib, n_target):
    from scipy.linalg import lapack
    pr, n = _problem_of_size(gpu_ctx, n_target)
    rng = np.random.default_rng(n)
    A = np.asfortranarray(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    A[:, 3] *= 1e-3; A[5, :] *= 40.0        # force non-trivial pivoting / scaling
    B = np.asfortranarray(rng.standard_normal((n, 3)) + 1j * rng.standard_normal((n, 3)))
    lu_ref, piv_ref, info = lapack.zgetrf(A)
    x_ref, info = lapack.zgetrs(lu_ref, piv_ref, B)
    Af = A.copy(order="F")
    x, ipiv = pr.solve_lse_c(Af, B, want_ipiv=True)
    assert np.array_equal(ipiv - 1, piv_ref)                      # same pivot sequence as zgetrf (izamax semantics)
    assert relerr(Af, lu_ref) < 1e-10                             # A overwritten by the same L\\U factors
    assert relerr(x, x_ref) < 1e-9
    assert np.abs(A @ x - B).max() / (np.abs(A).max() * np.abs(x).max() * n) < 1e-14
    # factorize = .false. re-uses the resident factors (src/multifebe.f90:119-120)
    b2 = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x2 = pr.solve_lse_c(None, b2, factorize=False)
    assert relerr(x2, np.linalg.solve(A, b2)) < 1e-9
    pr.close()


def test_singular_matrix_reports_lapack_info(gpu_ctx):
    from multifebe_b200 import capi
    pr, n = _problem_of_size(gpu_ctx, 72)
    A = np.asfortranarray(np.eye(n, dtype=complex)); A[:, 10] = 0.0
    with pytest.raises(capi.MfbError) as e:
        pr.solve_lse_c(A, np.ones(n, dtype=complex))
    assert e.value.code == 11 and "singular" in str(e.value)      # info = 11 (1-based column of the zero pivot)
    pr.close()


def test_measured_peaks_are_plausible(gpu_ctx):
    p = gpu_ctx.measure_peaks()
    assert 20.0 < p["dfma_tflops"] < 80.0 and 20.0 < p["dmma_tflops"] < 160.0 and 3000.0 < p["copy_gbs"] < 9000.0


def test_full_size_properties_30k_dof(gpu_ctx, oracle_lib):
    """BASELINE config 3 size (S-cube tri3 m=40: 30258 DOF, 19200 elements): spot parity of assembled entries against
    oracle pair integrals, backward error of the device LU solution, and the analytic column solution."""
    from multifebe_b200 import capi
    md = Model(cube_mesh(40, shape.TRI3), cube_bcs())
    assert md.n_dof == 30258
    mat = Material(1.0, 1.0, 0.25, 0.03)
    omega = 9.0
    pr = capi.Problem(gpu_ctx, md)
    pr.build_lse_mechanics_bem_harela(omega, mat, want_host=False)
    o = oracle_lib.Oracle(md)
    rng = np.random.default_rng(11)
    node_elems = {}
    for e, c in enumerate(md.mesh.conn):
        for kn, v in enumerate(c):
            node_elems.setdefault(int(v), []).append((e, kn))
    colloc_of_node = {}
    for c in range(md.n_colloc):
        colloc_of_node.setdefault(int(md.colloc_node[c]), []).append(c)
    rows, cols, expect = [], [], []
    sn_list = rng.integers(0, md.n_node, 24)
    for sn in sn_list:
        sn = int(sn)
        own = set(e for c in colloc_of_node[sn] for e in [int(md.colloc_elem[c])]) | set(e for e, _ in node_elems[sn])
        # a far node, a node on the same face a few cells away, and a node of a neighbouring (non-incident) element
        d = np.linalg.norm(md.node_x - md.node_x[sn], axis=1)
        cand = [int(rng.integers(0, md.n_node)), int(np.argsort(d)[12]), int(np.argsort(d)[40])]
        for j in cand:
            if any(e in own for e, _ in node_elems[j]):
                continue
            blk = np.zeros((3, 3), dtype=complex)
            for c in colloc_of_node[sn]:
                for e, kn in node_elems[j]:
                    h, g, mode, _ = o.pair(e, md.colloc_x[c], omega, mat)
                    for k in range(3):
                        blk[:, k] += (-g[kn, :, k]) if md.ctype[j, k] == 0 else h[kn, :, k]
            for l in range(3):
                for k in range(3):
                    rows.append(md.row[sn, l]); cols.append(md.col_t[j, k] if md.ctype[j, k] == 0 else md.col_u[j, k]); expect.append(blk[l, k])
    got = pr.get_entries(rows, cols)
    expect = np.array(expect)
    assert len(expect) > 300
    scale = np.abs(expect).reshape(-1, 9).max(axis=1).repeat(9)
    assert (np.abs(got - expect) / scale).max() < TOL_A
    # solve, then re-assemble and measure the backward error of the solution on the fresh system
    x = pr.solve_frequency(omega, mat)
    pr.build_lse_mechanics_bem_harela(omega, mat, want_host=False)
    berr, rel = pr.residual(x)
    assert berr < 1e-11 and rel < 1e-13
    u, t = md.nodal_solution(x)
    ua = column_analytic_u(md.node_x[:, 0], omega, mat)
    assert relerr(u[:, 0], ua) < 5e-3
    pr.close()


# ---- one frequency over several GPUs: the distributed path on VIRTUAL ranks of one GPU (collectives = device copies), so that
# ---- the block-cyclic index logic, the row-block assembly and the redistribution are covered by the single-GPU test tier
@pytest.mark.parametrize("m,nranks,nb", [(3, 2, 32), (3, 3, 32), (6, 2, 64), (6, 4, 64), (6, 3, 256), (4, 8, 32)])
def test_distributed_lu_on_virtual_ranks_matches_lapack(gpu_ctx, oracle_lib, m, nranks, nb):
    from multifebe_b200 import capi
    md = Model(cube_mesh(m, shape.TRI3), cube_bcs())
    n = md.n_dof
    pr = capi.Problem(gpu_ctx, md)
    pr.dist_init_loopback(nranks, nb)
    rng = np.random.default_rng(100 * m + nranks)
    A = np.asfortranarray(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))     # no diagonal dominance: pivoting is exercised
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x, ipiv = pr.dist_zsolve(A, b, want_ipiv=True)
    xo, _, piv_o = oracle_lib.lu_solve(A, b)
    assert np.array_equal(ipiv, np.asarray(piv_o) + 1)      # the pivot sequence of LAPACK zgetrf (scipy returns it 0-based)
    assert relerr(x, xo) < 1e-9
    r = A @ x - b
    assert np.abs(r).max() / (np.abs(A).sum(axis=1).max() * np.abs(x).max()) < 1e-13
    pr.close()


@pytest.mark.parametrize("et,m,nranks,nb", [(shape.TRI3, 5, 2, 64), (shape.TRI3, 5, 3, 32), (shape.QUAD9, 2, 4, 32), (shape.QUAD4, 4, 2, 256)])
def test_distributed_frequency_on_virtual_ranks_matches_single_gpu(gpu_ctx, oracle_lib, et, m, nranks, nb):
    from multifebe_b200 import capi
    md = Model(cube_mesh(m, et), cube_bcs())
    pr = capi.Problem(gpu_ctx, md)
    omega = 3.0
    x1 = pr.solve_frequency(omega, MAT)
    pr.dist_init_loopback(nranks, nb)
    info = pr.dist_info()
    rb = info["row_bounds"]
    assert rb[0] == 0 and rb[-1] == md.n_dof and np.all(np.diff(rb) >= 0)
    x2 = pr.dist_solve_frequency(omega, MAT)
    assert relerr(x2, x1) < 1e-11
    Ao, bo, _ = oracle_lib.Oracle(md).assemble(omega, MAT)
    xo, _, _ = oracle_lib.lu_solve(Ao, bo)
    assert relerr(x2, xo) < TOL_X
    x3 = pr.solve_frequency(omega, MAT)       # the single-GPU path still works on the same problem afterwards
    assert relerr(x3, x1) < 1e-13
    pr.close()


# ---- static 3D elasticity: Kelvin kernels in real arithmetic + real LU, against the static oracle ----
SMAT = Material(1.0, 1.3, 0.25, 0.0)


@pytest.mark.parametrize("et,m", [(shape.TRI3, 3), (shape.TRI6, 2), (shape.QUAD4, 3), (shape.QUAD8, 2), (shape.QUAD9, 2)])
def test_static_assembly_and_solution_parity(gpu_ctx, oracle_lib, et, m):
    from multifebe_b200 import capi
    md = Model(cube_mesh(m, et), cube_bcs())
    pr = capi.Problem(gpu_ctx, md)
    A, b = pr.build_lse_mechanics_bem_staela(SMAT)
    Ao, bo, st = oracle_lib.Oracle(md).assemble_static(SMAT)
    assert A.dtype == np.float64 and relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A
    xo, _, _ = oracle_lib.lu_solve_real(Ao, bo)
    x1 = pr.solve_lse_r(A.copy(order="F"), b)              # seam 2 with host arrays (dgesv semantics)
    x2 = pr.solve_static(SMAT)                             # fused, device resident
    assert relerr(x1, xo) < TOL_X and relerr(x2, xo) < TOL_X
    # the exact solution of the reference's tutorial ME-ST-EL-002: u1 = P x1 / (lambda + 2 mu) (linear elements: to quadrature error)
    u, t = md.nodal_solution(x2)
    lam2mu = 2.0 * SMAT.mu_r * SMAT.nu_r / (1.0 - 2.0 * SMAT.nu_r) + 2.0 * SMAT.mu_r
    assert np.abs(u[:, 0].real - md.node_x[:, 0] / lam2mu).max() < 2e-5 / lam2mu
    # a harmonic frequency on the same problem afterwards (the two paths share the resident matrix)
    xh = pr.solve_frequency(3.0, MAT)
    Ah, bh, _ = oracle_lib.Oracle(md).assemble(3.0, MAT)
    assert relerr(xh, oracle_lib.lu_solve(Ah, bh)[0]) < TOL_X
    pr.close()


def test_static_nondefault_settings_reversed_boundary_and_open_patch(gpu_ctx, oracle_lib):
    from multifebe_b200 import capi
    cases = [
        Model(cube_mesh(2, shape.QUAD8), cube_bcs(), qsi_relative_error=1e-4, qsi_ns_max=3, precalset_gln=(2, 4, 6)),
        Model(cube_mesh(2, shape.TRI3), cube_bcs(), reversed_parts=(1, 2, 3, 4, 5, 6), qsi_relative_error=1e-8),
        Model(halfspace_patch(4, shape.QUAD9), {1: ([1, 1, 1], [0, 0, 0]), 2: ([0, 0, 0], [0, 0, 1.0])}),
        Model(halfspace_patch(5, shape.TRI6), {1: ([1, 1, 1], [0.1, 0, 0.3]), 2: ([0, 1, 0], [1.0, 0.5, 0.2])}),
    ]
    for md in cases:
        pr = capi.Problem(gpu_ctx, md)
        A, b = pr.build_lse_mechanics_bem_staela(SMAT)
        Ao, bo, _ = oracle_lib.Oracle(md).assemble_static(SMAT)
        assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A
        xo, _, _ = oracle_lib.lu_solve_real(Ao, bo)
        assert relerr(pr.solve_static(SMAT), xo) < TOL_X
        pr.close()


@pytest.mark.parametrize("n_target", [18 * 4, 18 * 36, 18 * 100])
def test_real_lu_against_lapack(gpu_ctx, n_target):
    from scipy.linalg import lapack
    pr, n = _problem_of_size(gpu_ctx, n_target)
    rng = np.random.default_rng(n + 1)
    A = np.asfortranarray(rng.standard_normal((n, n)))
    A[:, 3] *= 1e-3; A[5, :] *= 40.0
    B = np.asfortranarray(rng.standard_normal((n, 3)))
    lu_ref, piv_ref, info = lapack.dgetrf(A)
    x_ref, info = lapack.dgetrs(lu_ref, piv_ref, B)
    Af = A.copy(order="F")
    x, ipiv = pr.solve_lse_r(Af, B, want_ipiv=True)
    assert np.array_equal(ipiv - 1, piv_ref)                      # the pivot sequence of dgetrf (idamax semantics)
    assert relerr(Af, lu_ref) < 1e-10 and relerr(x, x_ref) < 1e-9
    b2 = rng.standard_normal(n)
    assert relerr(pr.solve_lse_r(None, b2, factorize=False), np.linalg.solve(A, b2)) < 1e-9
    with pytest.raises(capi_error()):
        pr_zsolve_on_real(pr, n)
    pr.close()


def capi_error():
    from multifebe_b200 import capi
    return capi.MfbError


def pr_zsolve_on_real(pr, n):
    """mfb_zsolve must refuse to factorise a resident REAL system (A = NULL) instead of reading a stale imaginary plane."""
    import ctypes as C
    from multifebe_b200 import capi
    b = np.zeros(n, dtype=np.complex128)
    capi._check(capi.lib().mfb_zsolve(pr.h, C.c_int(n), None, C.c_int(n), None, capi._p(b), C.c_int(1), C.c_int(1)))


def test_static_full_size_properties_config2(gpu_ctx, oracle_lib):
    """BASELINE config 2 size (static cube, quad9 m=11: 9522 DOF, 726 elements): spot parity of assembled entries against
    static oracle pair integrals, residual of the real LU solution on the re-assembled system, exact column solution."""
    from multifebe_b200 import capi
    md = Model(cube_mesh(11, shape.QUAD9), cube_bcs())
    assert md.n_dof == 9522
    pr = capi.Problem(gpu_ctx, md)
    pr.build_lse_mechanics_bem_staela(SMAT, want_host=False)
    o = oracle_lib.Oracle(md)
    rng = np.random.default_rng(12)
    node_elems, colloc_of_node = {}, {}
    for e, c in enumerate(md.mesh.conn):
        for kn, v in enumerate(c):
            node_elems.setdefault(int(v), []).append((e, kn))
    for c in range(md.n_colloc):
        colloc_of_node.setdefault(int(md.colloc_node[c]), []).append(c)
    rows, cols, expect = [], [], []
    for sn in rng.integers(0, md.n_node, 16):
        sn = int(sn)
        own = set(int(md.colloc_elem[c]) for c in colloc_of_node[sn]) | set(e for e, _ in node_elems[sn])
        d = np.linalg.norm(md.node_x - md.node_x[sn], axis=1)
        for j in (int(rng.integers(0, md.n_node)), int(np.argsort(d)[30]), int(np.argsort(d)[90])):
            if any(e in own for e, _ in node_elems[j]):
                continue
            blk = np.zeros((3, 3))
            for c in colloc_of_node[sn]:
                for e, kn in node_elems[j]:
                    h, g, mode = o.pair_static(e, md.colloc_x[c], SMAT)
                    for k in range(3):
                        blk[:, k] += (-g[kn, :, k]) if md.ctype[j, k] == 0 else h[kn, :, k]
            for l in range(3):
                for k in range(3):
                    rows.append(md.row[sn, l]); cols.append(md.col_t[j, k] if md.ctype[j, k] == 0 else md.col_u[j, k]); expect.append(blk[l, k])
    got = pr.get_entries(rows, cols)
    expect = np.array(expect)
    assert len(expect) > 200 and np.abs(got.imag).max() == 0.0
    scale = np.abs(expect).reshape(-1, 9).max(axis=1).repeat(9)
    assert (np.abs(got.real - expect) / scale).max() < TOL_A
    x = pr.solve_static(SMAT)
    pr.build_lse_mechanics_bem_staela(SMAT, want_host=False)
    berr, rel = pr.residual(x.astype(np.complex128))
    assert berr < 1e-11 and rel < 1e-13
    u, t = md.nodal_solution(x)
    lam2mu = 2.0 * SMAT.mu_r * SMAT.nu_r / (1.0 - 2.0 * SMAT.nu_r) + 2.0 * SMAT.mu_r
    assert np.abs(u[:, 0].real - md.node_x[:, 0] / lam2mu).max() * lam2mu < 2e-5
    pr.close()


@pytest.mark.parametrize("et,m", [(shape.TRI3, 2), (shape.TRI6, 1), (shape.QUAD4, 2), (shape.QUAD8, 1), (shape.QUAD9, 1)])
def test_static_against_committed_golden_vectors(gpu_ctx, et, m):
    from multifebe_b200 import capi
    gold = np.load(os.path.join(HERE, "golden", "oracle_static.npz"))
    md = Model(cube_mesh(m, et), cube_bcs())
    pr = capi.Problem(gpu_ctx, md)
    A, b = pr.build_lse_mechanics_bem_staela(SMAT)
    assert relerr(A, gold[f"A:{et}:{m}"]) < TOL_A and relerr(b, gold[f"b:{et}:{m}"]) < TOL_A
    assert relerr(pr.solve_static(SMAT), gold[f"x:{et}:{m}"]) < TOL_X
    pr.close()


def test_nan_input_and_state_misuse_do_not_fault_the_device(gpu_ctx):
    """Round-1 advisor findings: a NaN column used to leave the pivot search without a row (illegal address, sticky context error); mfb_zsolve used to
    factorise whatever was resident (unassembled memory, or the factors of an earlier solve); omega = 0 produced NaN kernel parameters."""
    from multifebe_b200 import capi
    from scipy.linalg import lapack
    pr, n = _problem_of_size(gpu_ctx, 18 * 36)
    rng = np.random.default_rng(5)
    b = rng.standard_normal(n) + 0j
    with pytest.raises(capi.MfbError) as e:                      # nothing assembled yet
        pr.solve_lse_c(None, b, factorize=True)
    assert e.value.code == -1 and "no assembled system" in str(e.value)
    A = np.asfortranarray(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    An = A.copy(order="F"); An[:, 40] = np.nan; An[17, 300] = np.nan
    x, ipiv = pr.solve_lse_c(An, b, want_ipiv=True)              # NaN propagates (as in LAPACK), pivots stay in range, no device fault
    assert np.isnan(x).any() and ipiv.min() >= 1 and ipiv.max() <= n
    Af = A.copy(order="F")
    x = pr.solve_lse_c(Af, b)                                     # the context is still healthy
    lu_ref, piv_ref, _ = lapack.zgetrf(A); x_ref, _ = lapack.zgetrs(lu_ref, piv_ref, b)
    assert relerr(x, x_ref) < 1e-9
    with pytest.raises(capi.MfbError) as e:                      # the resident matrix now holds L\U: refactorising it is refused
        pr.solve_lse_c(None, b, factorize=True)
    assert e.value.code == -1
    assert relerr(pr.solve_lse_c(None, b, factorize=False), x_ref) < 1e-9
    mat = Material(1.0, 1.0, 0.25, 0.03)
    for bad in (0.0, -1.0, float("nan"), float("inf")):
        with pytest.raises(capi.MfbError) as e:
            pr.solve_frequency(bad, mat)
        assert e.value.code == -1 and "omega" in str(e.value)
    x = pr.solve_frequency(2.0, mat)
    assert np.isfinite(x).all()
    with pytest.raises(capi.MfbError):                           # real factors (static path) are not accepted by the complex solve
        pr.solve_static(Material(1.0, 1.0, 0.25, 0.0)); pr.solve_lse_c(None, b, factorize=False)
    pr.close()


def test_full_size_singular_and_near_entries_30k_dof(gpu_ctx, oracle_lib):
    """at the full 30258-DOF size the spot checks skipped every incident element.  Here the 3 x 3 blocks A[rows(sn), cols(j)] are compared
    for j ON the elements of the collocation node and on their neighbours -- singular pairs, the quasi-singular pairs of the MCA points (0.05 h from the
    neighbouring elements) and the free terms (1/2 delta at a smooth nodal point, phi_j(xi_i)/2 delta at an MCA point) -- for nodal and rim nodes."""
    from multifebe_b200 import capi
    md = Model(cube_mesh(40, shape.TRI3), cube_bcs())
    mat = Material(1.0, 1.0, 0.25, 0.03)
    omega = 9.0
    pr = capi.Problem(gpu_ctx, md)
    pr.build_lse_mechanics_bem_harela(omega, mat, want_host=False)
    o = oracle_lib.Oracle(md)
    node_elems = {}
    for e, c in enumerate(md.mesh.conn):
        for kn, v in enumerate(c):
            node_elems.setdefault(int(v), []).append((e, kn))
    colloc_of_node = {}
    for c in range(md.n_colloc):
        colloc_of_node.setdefault(int(md.colloc_node[c]), []).append(c)
    rng = np.random.default_rng(23)
    rim = np.flatnonzero(md.in_boundary); inner = np.flatnonzero(~md.in_boundary)
    picks = [int(v) for v in rng.choice(inner, 8, replace=False)] + [int(v) for v in rng.choice(rim, 8, replace=False)]
    rows, cols, expect, kinds = [], [], [], set()
    for sn in picks:
        own = sorted(set(e for e, _ in node_elems[sn]))
        ring1 = sorted(set(int(v) for e in own for v in md.mesh.conn[e]))                                   # nodes of the incident elements (sn included)
        ring2 = sorted(set(int(v) for j in ring1 for e, _ in node_elems[j] for v in md.mesh.conn[e]) - set(ring1))[:6]
        for j in ring1 + ring2:
            blk = np.zeros((3, 3), dtype=complex)
            for c in colloc_of_node[sn]:
                for e, kn in node_elems[j]:
                    h, g, mode, _ = o.pair(e, md.colloc_x[c], omega, mat)
                    kinds.add(int(mode) if mode in (100, 200) else 0)
                    if md.in_boundary[sn] and e == int(md.colloc_elem[c]):                                   # MCA point inside element e: free term phi_j(xi_i) / 2
                        h = h.copy(); h[kn] += 0.5 * shape.phi(int(md.etype[e]), md.colloc_xi[c])[kn] * np.eye(3)
                    for k in range(3):
                        blk[:, k] += (-g[kn, :, k]) if md.ctype[j, k] == 0 else h[kn, :, k]
            if j == sn and not md.in_boundary[sn]:                                                           # smooth nodal point: c = I / 2
                for k in range(3):
                    if md.ctype[j, k] == 1:
                        blk[k, k] += 0.5
            for l in range(3):
                for k in range(3):
                    rows.append(md.row[sn, l]); cols.append(md.col_t[j, k] if md.ctype[j, k] == 0 else md.col_u[j, k]); expect.append(blk[l, k])
    got = pr.get_entries(rows, cols)
    expect = np.array(expect)
    assert kinds == {0, 100, 200} and len(expect) > 1500                                                     # regular, quasi-singular and singular pairs all met
    scale = np.abs(expect).reshape(-1, 9).max(axis=1).repeat(9)
    assert (np.abs(got - expect) / scale).max() < TOL_A
    pr.close()
