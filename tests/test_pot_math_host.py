"""The product's scalar-wave point arithmetic (multifebe_b200/csrc/pot_math.cuh, the code the P1/P2/P3 kernels inline) compiled
for the HOST and compared with the CPU oracle: p* = fs_P/(4 pi), q* = -fs_Q dr/dn/(4 pi), both branches of E_m(z)."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from multifebe_b200.host import Fluid
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def pmh(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("pmh") / "libpmh.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", os.path.join(HERE, "native", "pot_math_host.cpp"),
                           "-o", so])
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_E2_E3_both_branches(pmh):
    """The product uses the series only for |z| <= 0.1 and the direct subtraction above it (cheaper on the FP64 pipe): E_m must then
    be accurate to a few ulp of the static term 1 that it is added to in the kernel (1/r + E_2/r, 1/r^2 + E_3/r^2), and to a few ulp
    of the largest subtracted term when |z| > 1, as in the reference."""
    for z in [1e-9 - 1e-8j, -0.001 - 0.02j, -0.004 - 0.0999j, -0.004 - 0.1001j, -0.05 - 0.6j, -0.02 - 0.9998j, -0.03 - 1.0003j, -0.2 - 3.0j,
              -1.5 - 40.0j, -0.0 - 250.0j]:
        z_ri = np.array([z.real, z.imag]); out = np.zeros(4)
        pmh.pmh_E23(_p(z_ri), _p(out))
        E = orc.zexp_decomposed(z)
        E2, E3 = complex(out[0], out[1]), complex(out[2], out[3])
        if abs(z) <= 0.1:
            scale2, scale3 = abs(E[2]), abs(E[3])                      # series: accurate relative to itself
        else:
            scale2, scale3 = 1.0 + abs(z), 1.0 + abs(z) + abs(z) ** 2 / 2
        assert abs(E2 - E[2]) <= 4e-16 * scale2 + 1e-300 and abs(E3 - E[3]) <= 4e-16 * scale3 + 1e-300, (z, E2, E[2], E3, E[3])


@pytest.mark.parametrize("omega,fl", [(2 * np.pi * 20.0, Fluid(1.25, 343.0)), (2 * np.pi * 300.0, Fluid(1.25, 343.0, 0.02)), (0.05, Fluid(1.0, 1.0)),
                                      (30.0, Fluid(1.0, 1.0, 0.05))])
def test_point_formula_against_the_oracle(pmh, omega, fl):
    rng = np.random.default_rng(11)
    c4pi = 1.0 / (4.0 * np.pi)
    pmh.pmh_point.argtypes = [C.c_double, C.c_double] + [C.c_void_p] * 5
    for _ in range(200):
        xc = rng.normal(size=3); x = xc + rng.normal(size=3) * rng.choice([1e-3, 0.05, 0.5, 3.0])
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        out = np.zeros(4); c_ri = np.array([fl.c.real, fl.c.imag])
        pmh.pmh_point(omega, fl.rho, _p(c_ri), _p(x), _p(n), _p(xc), _p(out))
        po, qo = orc.fundamental_solutions_pot(x, n, xc, omega, fl)
        p_gpu = c4pi * complex(out[2], out[3]); q_gpu = -c4pi * complex(out[0], out[1])
        r = np.linalg.norm(x - xc)
        # Both sides evaluate the reference's regularised form, whose direct branch subtracts (E_2 = e^z - 1 - z, E_3 = E_2 - z^2/2): the
        # rounding error is relative to the largest term, (1 + |z|)/r for p* and (1 + |z| + |z|^2)/r^2 for q*, not to the result
        # (which a damped medium makes exponentially smaller).  4 ulp of that scale.
        az = abs(omega / fl.c) * r
        assert abs(p_gpu - po) <= 4 * 2.2e-16 * c4pi * (1.0 + az) / r
        assert abs(q_gpu - qo) <= 4 * 2.2e-16 * c4pi * (1.0 + az + az * az) / (r * r)
