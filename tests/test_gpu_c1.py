"""BASELINE config C1 on the reference's OWN input: docs/examples/ME-TH-EL-001/case_files/t3.dat + t3.msh (462 nodes, 744 tri3, 1386 DOF, 300 frequencies
lin in [0.01, 15] rad/s), committed as test vectors under tests/golden/ME-TH-EL-001/.  The case file is read by the library's own reader (the reference's
format), the systems are assembled and solved on the GPU through the C ABI and compared with the CPU oracle at frequencies spread over the band, one of
them next to the first resonance of the column (omega ~ 2.6: the analytic curve of doc_src/ME-TH-EL-001.tex:32-56 peaks there)."""
import os
import numpy as np
import pytest
from multifebe_b200.host import column_analytic_u
from multifebe_b200.host.casefile import CaseFile

pytestmark = [pytest.mark.gpu]
HERE = os.path.dirname(os.path.abspath(__file__))
CASE = os.path.join(HERE, "golden", "ME-TH-EL-001", "t3.dat")


def relerr(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.fixture(scope="module")
def c1():
    case = CaseFile(CASE)
    md = case.build_model()
    assert (md.n_node, md.n_elem, md.n_dof, len(case.omega)) == (462, 744, 1386, 300)
    return case, md


def test_c1_matrix_and_solution_parity_at_seven_frequencies(gpu_ctx, oracle_lib, c1):
    from multifebe_b200 import capi
    case, md = c1
    mat = case.material
    pr = capi.Problem(gpu_ctx, md)
    orc = oracle_lib.Oracle(md)
    fr = np.asarray(case.omega)
    picks = [0, int(np.argmin(np.abs(fr - 2.6))), 60, 120, 180, 240, 299]          # 0.01 ... 15 rad/s, and the resonance
    for kf in picks:
        om = float(fr[kf])
        A, b = pr.build_lse_mechanics_bem_harela(om, mat)
        x = pr.solve_frequency(om, mat)
        Ao, bo, _ = orc.assemble(om, mat, nthreads=0)
        xo, _, _ = oracle_lib.lu_solve(Ao, bo)
        assert relerr(A, Ao) < 1e-11 and relerr(b, bo) < 1e-11, (kf, om)
        assert relerr(x, xo) < 1e-8, (kf, om)
        u, t = md.nodal_solution(x)
        ua = column_analytic_u(md.node_x[:, 0], om, mat)
        if kf != picks[1]:      # away from the resonance the 744-element mesh follows the analytic column to a fraction of a percent at low frequency
            assert np.abs(u[:, 0] - ua).max() < (3e-3 if om < 3.0 else 0.2) * np.abs(ua).max(), (kf, om)
    pr.close()


def test_c1_through_the_standalone_driver(gpu_ctx, c1, tmp_path):
    """python -m multifebe_b200 -i t3.dat : the reference's command line on the reference's file; the *.nso rows are the nodal solutions of the 300
    frequencies in sweep order."""
    import io
    import shutil
    from multifebe_b200 import driver
    for f in ("t3.dat", "t3.msh"):
        shutil.copy(os.path.join(os.path.dirname(CASE), f), str(tmp_path / f))
    nso = driver.run(str(tmp_path / "t3.dat"), log=io.StringIO())
    rows = [s for s in open(nso) if s.strip() and not s.startswith("#")]
    assert len(rows) == 300 * 462


def test_c_abi_sweep_and_accumulating_assembly(gpu_ctx, c1):
    """mfb_harela3d_sweep (the frequency loop as one C-ABI call; single rank: no NCCL) returns what per-frequency calls return, and
    mfb_harela3d_assemble_acc(accumulate = 1) ADDS the region's system to the caller's arrays (the `+=` of src/build_lse_mechanics_harmonic.f90:73-95)."""
    from multifebe_b200 import capi
    case, md = c1
    mat = case.material
    pr = capi.Problem(gpu_ctx, md)
    oms = [float(case.omega[k]) for k in (3, 50, 120, 299)]
    X, info = pr.sweep(oms, mat)
    assert (info == 0).all()
    for k, om in enumerate(oms):
        assert relerr(X[k], pr.solve_frequency(om, mat)) < 1e-13
    A, b = pr.build_lse_mechanics_bem_harela(oms[1], mat)
    rng = np.random.default_rng(0)
    A0 = np.asfortranarray(rng.standard_normal(A.shape) + 1j * rng.standard_normal(A.shape)); b0 = rng.standard_normal(b.shape) + 0j
    A1 = A0.copy(order="F"); b1 = b0.copy()
    pr.build_lse_accumulate(oms[1], mat, A1, b1)
    assert relerr(A1 - A0, A) < 1e-14 and relerr(b1 - b0, b) < 1e-14
    pr.close()


def test_frequencies_in_flight_on_lanes_match_one_at_a_time(gpu_ctx, c1):
    """capi.ProblemLanes: 6 (context, problem) pairs on one GPU, one host thread each, 24 different frequencies assembled and solved side by side.
    The kernel parameters are launch arguments and every context owns its K1 launch state (round 1 kept both in process-wide symbols), so the results are
    those of the one-at-a-time path, bit for bit in the solution's leading digits."""
    from multifebe_b200 import capi
    case, md = c1
    mat = case.material
    oms = [float(case.omega[k]) for k in range(5, 300, 13)][:24]
    pr = capi.Problem(gpu_ctx, md)
    ref = np.array([pr.solve_frequency(om, mat) for om in oms])
    pr.close()
    lanes = capi.ProblemLanes(md, 0, n_lanes=6)
    for rep in range(2):
        X = lanes.run(oms, mat)
        assert relerr(X, ref) < 1e-12
    lanes.close()
