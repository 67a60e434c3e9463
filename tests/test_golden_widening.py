"""Committed regression vectors of the oracle for the paths of SURVEY.md 8f rank 3 (tests/golden/oracle_widening.npz, written by
`python tools/gen_golden.py widening`; oracle output, not reference output): acoustic and poroelastic single regions and coupled two-region
systems.  CPU: the oracle still reproduces them.  GPU (gated until their first hardware run): the device paths against the same vectors."""
import os
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
from gen_golden import widening_cases, widening_system, widening_probe, FULL_MATRICES

GOLD = np.load(os.path.join(HERE, "golden", "oracle_widening.npz"))
CASES = widening_cases()


def check(A, b, key, tol):
    n = len(b)
    Av, vA = A @ widening_probe(n), widening_probe(n)[::-1] @ A
    assert np.abs(b - GOLD["b:" + key]).max() <= tol * np.abs(GOLD["b:" + key]).max()
    # rows and columns of different variables live on different scales: compare the products entry by entry against |A| |v|
    sr = np.abs(A) @ np.abs(widening_probe(n)); sc = np.abs(widening_probe(n)[::-1]) @ np.abs(A)
    assert (np.abs(Av - GOLD["Av:" + key]) <= tol * sr).all() and (np.abs(vA - GOLD["vA:" + key]) <= tol * sc).all()
    if key in FULL_MATRICES:
        sc_col = np.abs(GOLD["A:" + key]).max(axis=0)
        assert (np.abs(A - GOLD["A:" + key]).max(axis=0) <= tol * sc_col).all()


@pytest.mark.parametrize("key", sorted(CASES))
def test_oracle_reproduces_the_committed_vectors(key):
    kind, model, mat, omega = CASES[key]
    A, b = widening_system(kind, model, mat, omega)
    check(A, b, key, 1e-12)
    x = np.linalg.solve(A, b)
    assert np.abs(x - GOLD["x:" + key]).max() <= 1e-9 * np.abs(GOLD["x:" + key]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(CASES))
def test_gpu_against_the_committed_vectors(gpu_ctx, key):
    from multifebe_b200 import capi
    kind, model, mat, omega = CASES[key]
    if kind == "coupled":
        cp = capi.CoupledProblem(gpu_ctx, model)
        A, b = cp.assemble(omega); x = cp.solve_frequency(omega); cp.close()
    else:
        pr = capi.Problem(gpu_ctx, model)
        if kind == "fluid":
            A, b = pr.build_lse_mechanics_bem_harpot(omega, mat); x = pr.solve_frequency_fluid(omega, mat)
        else:
            A, b = pr.build_lse_mechanics_bem_harpor(omega, mat); x = pr.solve_frequency_poro(omega, mat)
        pr.close()
    check(A, b, key, 1e-11)
    xg = GOLD["x:" + key]
    sc = np.abs(A).max(axis=0)                      # weight every unknown by the scale of its column
    assert np.abs((x - xg) * sc).max() <= 1e-8 * np.abs(xg * sc).max()
