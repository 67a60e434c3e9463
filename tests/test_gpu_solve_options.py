"""solve_lse_c's optional stages on the device (mfb_zsolve_ex: scaling = zgeequ + zlaqge, condition = zgecon, refine = zgerfs; src/solve_lse_c.f90:81-206)
against LAPACK's expert driver zgesvx (scipy), and the 3M trailing update on badly scaled BEM systems (SI soil: G columns ~1/mu = 1e-8 beside H columns
~1): VERDICT r01 weak #8 / next #9."""
import numpy as np
import pytest
from multifebe_b200.host import Model, Material, cube_mesh, cube_bcs, shape

pytestmark = [pytest.mark.gpu]


def relerr(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def _problem(gpu_ctx, m=4):
    from multifebe_b200 import capi
    md = Model(cube_mesh(m, shape.TRI3), cube_bcs())
    return capi.Problem(gpu_ctx, md), md


def test_scaling_condition_refine_against_zgesvx(gpu_ctx):
    from scipy.linalg import lapack
    pr, md = _problem(gpu_ctx)
    n = md.n_dof
    rng = np.random.default_rng(21)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) + 4.0 * np.eye(n)
    A = (10.0 ** rng.uniform(-6, 6, n))[:, None] * A * (10.0 ** rng.uniform(-5, 5, n))[None, :]       # rows and columns far apart in scale
    A = np.asfortranarray(A)
    B = np.asfortranarray(rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2)))
    out = lapack.zgesvx(A, B, fact="E")
    as_, lu, ipiv, equed, rs, cs, bs, x_ref, rcond, ferr, berr, info = out
    assert info == 0
    equed = equed.decode() if isinstance(equed, bytes) else equed
    x, inf = pr.solve_lse_c_ex(A.copy(order="F"), B, factorize=True, scaling=True, condition=True, refine=True)
    assert inf["equed"] == equed == "B"
    assert relerr(inf["r"], rs) < 1e-14 and relerr(inf["c"], cs) < 1e-14
    assert np.array_equal(inf["ipiv"] - 1, ipiv)                                     # same pivot sequence on the equilibrated matrix
    assert abs(inf["rcond"] - rcond) < 1e-6 * rcond                                  # same estimator, solves differ in rounding only
    assert relerr(x, x_ref) < 1e-10
    assert (inf["berr"] < 4 * np.finfo(float).eps).all() and (inf["ferr"] < 10 * ferr + 1e-15).all() and (inf["ferr"] > 0).all()
    # factorize = .false. re-uses factors and scale factors (src/multifebe.f90:119-120 with the flags of the first call)
    b2 = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x2, _ = pr.solve_lse_c_ex(None, b2, factorize=False, scaling=True, refine=True, equed=inf["equed"], r=inf["r"], c=inf["c"])
    assert relerr(x2, np.linalg.solve(A, b2)) < 1e-9
    # no equilibration needed -> equed 'N' and the plain solve
    A1 = np.asfortranarray(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) + 6.0 * np.eye(n))
    x3, inf3 = pr.solve_lse_c_ex(A1.copy(order="F"), b2, scaling=True, condition=True)
    assert inf3["equed"] == "N" and relerr(x3, np.linalg.solve(A1, b2)) < 1e-10
    assert abs(inf3["rcond"] - lapack.zgesvx(A1, b2.reshape(-1, 1), fact="E")[8]) < 1e-6 * inf3["rcond"]
    pr.close()


@pytest.mark.parametrize("et,m", [(shape.TRI3, 5), (shape.QUAD9, 2)])
def test_si_soil_solution_3m_update_vs_lapack(gpu_ctx, oracle_lib, et, m):
    """SI units: mu = 8e7 Pa, rho = 2000 kg/m3 on a 10 m cube.  Columns of unknown tractions hold G ~ 1/mu ~ 1e-8, columns of unknown displacements hold H ~ 1:
    the 3M complex product of the trailing update gives up the componentwise bound on imaginary parts, so the solution of the device LU is
    compared with LAPACK's on the same matrix, without and with equilibration."""
    from multifebe_b200 import capi
    mesh = cube_mesh(m, et, L=10.0)
    md = Model(mesh, cube_bcs())
    mat = Material(2000.0, 8.0e7, 0.3, 0.05)
    pr = capi.Problem(gpu_ctx, md)
    for omega in (2 * np.pi * 3.0, 2 * np.pi * 25.0):
        A, b = pr.build_lse_mechanics_bem_harela(omega, mat)
        Ao, bo, _ = oracle_lib.Oracle(md).assemble(omega, mat)
        assert relerr(A, Ao) < 1e-11 and relerr(b, bo) < 1e-11
        assert np.abs(A).max() / np.abs(A[np.abs(A) > 0]).min() > 1e9               # the system really is badly scaled
        xo, _, _ = oracle_lib.lu_solve(Ao, bo)
        x = pr.solve_frequency(omega, mat)
        u, t = md.nodal_solution(x); uo, to = md.nodal_solution(xo)
        assert relerr(u, uo) < 1e-8 and relerr(t, to) < 1e-8                        # displacements (1e-9 m) and tractions (1 Pa) each on its own scale
        xs, inf = pr.solve_lse_c_ex(A.copy(order="F"), b, scaling=True, condition=True, refine=True)
        us, ts = md.nodal_solution(xs)
        assert inf["equed"] in ("B", "C", "R") and relerr(us, uo) < 1e-8 and relerr(ts, to) < 1e-8 and inf["berr"][0] < 1e-15
    pr.close()
