"""The product's poroelastic point arithmetic (multifebe_b200/csrc/por_math.cuh: parameter tables, the twelve radial scalars, the exterior
and interior 4 x 4 blocks) compiled for the HOST and compared with the CPU oracle.  The header is the arithmetic core of the poroelastic
device path of the next round; no kernel uses it yet."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from multifebe_b200.host import Poro
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def pmh(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("pmh") / "libpmhpor.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", os.path.join(HERE, "native", "por_math_host.cpp"), "-o", so])
    L = C.CDLL(so)
    L.pmh_por_exterior.argtypes = [C.c_double] + [C.c_void_p] * 7
    L.pmh_por_interior.argtypes = [C.c_double] + [C.c_void_p] * 7
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


MEDIA = [Poro(rhof=1.0, rhos=2.2, lam=1.2, mu=1.0, xi=0.02, phi=0.35, rhoa=0.15, R=0.8, Q=0.5, b=0.4),
         Poro(rhof=1000.0, rhos=2650.0, lam=1.0e8, mu=0.8e8, xi=0.03, phi=0.3, rhoa=150.0, R=4.0e8, Q=9.0e8, b=1.0e6),     # water-saturated soil
         Poro(rhof=1.3, rhos=2.0, lam=1.5, mu=1.0, xi=0.02, phi=0.4, rhoa=0.0, R=0.9, Q=0.0, b=0.0)]                       # decoupled


@pytest.mark.parametrize("po", MEDIA)
@pytest.mark.parametrize("omega_scale", [0.02, 1.0, 12.0])
def test_exterior_and_interior_blocks_against_the_oracle(pmh, po, omega_scale):
    # frequencies such that k r spans both branches of E_m
    c_s = np.sqrt(abs(po.mu) / po.rho1)
    omega = omega_scale * c_s
    rng = np.random.default_rng(5)
    pr = po.props()
    for _ in range(40):
        xc = rng.normal(size=3); x = xc + rng.normal(size=3) * rng.choice([0.01, 0.3, 2.0])
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        u = np.zeros((4, 4), dtype=np.complex128); t = np.zeros((4, 4), dtype=np.complex128); k5 = np.zeros(5, dtype=np.complex128)
        pmh.pmh_por_exterior(omega, _p(pr), _p(x), _p(n), _p(xc), _p(u), _p(t), _p(k5))
        uo, to, ko = orc.fundamental_solutions_por(x, n, xc, omega, po)
        assert np.abs(k5 - ko).max() <= 1e-13 * np.abs(ko).max()
        # both sides evaluate the same regularised sums, whose direct branch subtracts: compare block by block against the block's own scale
        for (a, b) in ((u, uo), (t, to)):
            for blk in (np.s_[0, 0], np.s_[0, 1:], np.s_[1:, 0], np.s_[1:, 1:]):
                sc = np.abs(b[blk]).max()
                if sc > 0:
                    assert np.abs(a[blk] - b[blk]).max() <= 2e-10 * sc, (blk, np.abs(a[blk] - b[blk]).max() / sc)
        # interior form + the CPV kernel = exterior form; the CPV kernel is T2(1)/r^2 (n_l r,k - n_k r,l)
        ui = np.zeros((4, 4), dtype=np.complex128); ti = np.zeros((4, 4), dtype=np.complex128); fc = np.zeros((3, 3), dtype=np.complex128)
        pmh.pmh_por_interior(omega, _p(pr), _p(x), _p(n), _p(xc), _p(ui), _p(ti), _p(fc))
        assert np.abs(ui - u).max() <= 1e-13 * np.abs(u).max() and np.abs(ti - t).max() <= 1e-12 * np.abs(t).max()
        rv = x - xc; r = np.linalg.norm(rv); d = rv / r
        T21 = -po.mu / (po.lam + 2 * po.mu)
        assert np.abs(fc - T21 / r ** 2 * (np.outer(n, d) - np.outer(d, n))).max() <= 1e-13 * abs(T21) / r ** 2
