"""Quadrature tables of the oracle (and, through data/quad_tables.h, of the product) against the reference's own data:
  * tests/golden/quad_tables.json: exactly-rounded checksums generated from the reference's .rc data statements
    (tools/gen_golden.py) -- travels to the GPU box;
  * the .rc files themselves, value for value, when /root/reference is present (build container only);
  * the defining mathematical properties of each rule family."""
import json, math, os, sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "quad_tables.json")))
FAM = {"gl11": 0, "gl01": 1, "gj01": 2}


@pytest.mark.parametrize("fam", ["gl11", "gl01", "gj01"])
def test_1d_rules_match_reference_checksums(oracle_lib, fam):
    for n in range(1, 33):
        x, w = oracle_lib.tables(FAM[fam], n)
        got = [math.fsum(v * (i + 1) for i, v in enumerate(x)).hex(), math.fsum(w).hex(), math.fsum(a * b for a, b in zip(x, w)).hex()]
        assert got == GOLD[f"{fam}:{n}"], (fam, n)


def test_wandzura_rules_match_reference_checksums(oracle_lib):
    for order in range(1, 31):
        x, w = oracle_lib.tables(3, order)
        npt = len(w)
        got = [npt, math.fsum(v * (i + 1) for i, v in enumerate(x[:npt])).hex(), math.fsum(v * (i + 1) for i, v in enumerate(x[npt:])).hex(),
               math.fsum(w).hex()]
        assert got == GOLD[f"wantri:{order}"], order


@pytest.mark.skipif(not os.path.isdir("/root/reference/lib/fbem/src/resources_quad_rules"), reason="reference tree not present")
def test_tables_value_for_value_against_reference_rc(oracle_lib):
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    from gen_quad_tables import parse_rc, REF
    for fam in ("gl11", "gl01", "gj01"):
        d = parse_rc(f"{REF}/{fam}.rc")
        for n in range(1, 33):
            x, w = oracle_lib.tables(FAM[fam], n)
            assert list(x) == d[fam + "_xi"][(n - 1) * 32:(n - 1) * 32 + n]
            assert list(w) == d[fam + "_w"][(n - 1) * 32:(n - 1) * 32 + n]
    d = parse_rc(f"{REF}/wantri.rc")
    for order in range(1, 31):
        x, w = oracle_lib.tables(3, order)
        npt = d["wantri_n"][order - 1]
        assert len(w) == npt
        assert list(x[:npt]) == d["wantri_xi1"][(order - 1) * 176:(order - 1) * 176 + npt]
        assert list(x[npt:]) == d["wantri_xi2"][(order - 1) * 176:(order - 1) * 176 + npt]
        assert list(w) == d["wantri_w"][(order - 1) * 176:(order - 1) * 176 + npt]


def test_gauss_legendre_exactness(oracle_lib):
    for n in (2, 5, 9, 15, 30):
        x, w = oracle_lib.tables(0, n)
        for p in range(0, 2 * n, 3):
            exact = 0.0 if p % 2 else 2.0 / (p + 1)
            assert abs(np.dot(w, x ** p) - exact) < 5e-14
        x, w = oracle_lib.tables(1, n)
        for p in range(0, 2 * n, 3):
            assert abs(np.dot(w, x ** p) - 1.0 / (p + 1)) < 5e-14
        # Gauss-Jacobi weight (1-x) on [0,1]: int (1-x) x^p = 1/((p+1)(p+2))
        x, w = oracle_lib.tables(2, n)
        for p in range(0, 2 * n, 3):
            assert abs(np.dot(w, x ** p) - 1.0 / ((p + 1) * (p + 2))) < 5e-14


def test_wandzura_exactness(oracle_lib):
    # the low orders are truncated in the reference (that truncation is part of parity), so only ~1e-7 is asserted there
    for order in (3, 5, 7, 9, 11, 13, 15, 17):
        x, w = oracle_lib.tables(3, order)
        npt = len(w)
        x1, x2 = x[:npt], x[npt:]
        assert abs(w.sum() - 0.5) < 1e-6
        for a in range(order + 1):
            for b in range(order + 1 - a):
                exact = math.factorial(a) * math.factorial(b) / math.factorial(a + b + 2)
                assert abs(np.dot(w, x1 ** a * x2 ** b) - exact) < 1e-6
