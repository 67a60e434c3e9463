"""The per-pair bodies of the poroelastic kernels (multifebe_b200/csrc/por_pair.cuh, shared by poro.cu and this host build) run lane-serially
over the product's own quadrature plans (plan_host.cpp) and compared with the oracle's pair integrals: regular, adaptive and singular pairs of
every element type, both orientations.  What stays untested without a GPU is the lane mapping, the warp reduction and the scatter of poro.cu."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from multifebe_b200.host import PoroModel, cube_mesh, shape
from oracle import oracle as orc
from test_oracle_multiregion import PO, poro_bcs_side

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def pph(tmp_path_factory):
    d = tmp_path_factory.mktemp("pph")
    so = str(d / "libpph.so"); objs = []
    for src in ("plan_host.cpp", "plan_values.cpp"):      # the product's host planner: decision core + value geometry
        o = str(d / src.replace(".cpp", ".o")); objs.append(o)
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fopenmp", "-c", os.path.join(ROOT, "multifebe_b200", "csrc", src), "-o", o])
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", os.path.join(HERE, "native", "por_pair_host.cpp"), "-x", "none"] + objs +
                          ["-o", so, "-lquadmath", "-fopenmp"])
    L = C.CDLL(so)
    L.pph_pair.restype = C.c_int
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("et,m", [(shape.TRI3, 2), (shape.TRI6, 1), (shape.QUAD4, 2), (shape.QUAD8, 1), (shape.QUAD9, 1)])
@pytest.mark.parametrize("reversed_parts", [(), (1, 2, 3, 4, 5, 6)])
def test_pair_integrals_of_the_kernel_bodies_against_the_oracle(pph, et, m, reversed_parts):
    bcs = {1: ([1, 0, 0, 0], [0, 0, 0, 0]), 2: ([0, 1, 1, 1], [0, 1.0, 0, 0])}
    bcs.update(poro_bcs_side((3, 4, 5, 6)))
    md = PoroModel(cube_mesh(m, et), bcs, reversed_parts=reversed_parts)
    o = orc.PorOracle(md)
    omega = 2.3
    pr = PO.props()
    gl = np.ascontiguousarray(md.precalset_gln, dtype=np.int32)
    seen = set()
    rng = np.random.default_rng(3)
    far = [np.array([3.0, -2.0, 4.0]), np.array([0.5, 0.5, 1.6])]
    for e in range(0, md.n_elem, max(1, md.n_elem // 6)):
        nodes = md.elem_node[md.elem_ptr[e]:md.elem_ptr[e + 1]]
        nn = len(nodes)
        xn = np.ascontiguousarray(md.node_x[nodes], dtype=np.float64)
        # the collocation points of the model (singular on the own element, near on its neighbours) and two far exterior points
        pts = [md.colloc_x[c] for c in rng.choice(md.n_colloc, size=min(md.n_colloc, 14), replace=False)]
        pts += [md.colloc_x[c] for c in range(md.n_colloc) if md.colloc_elem[c] == e][:3] + far
        for x_i in pts:
            x_i = np.ascontiguousarray(x_i, dtype=np.float64)
            h0, g0, mode0 = o.pair(e, x_i, omega, PO)
            h = np.zeros((nn, 4, 4), dtype=np.complex128); g = np.zeros((nn, 4, 4), dtype=np.complex128)
            mode = pph.pph_pair(C.c_int(int(md.etype[e])), _p(xn), C.c_int(int(md.elem_reversed[e])), _p(x_i), C.c_double(omega), _p(pr),
                                C.c_double(md.qsi_relative_error), C.c_int(md.qsi_ns_max), C.c_int(len(gl)), _p(gl), C.c_double(md.geometric_tolerance), _p(h), _p(g))
            assert mode == {100: 1, 200: 2}.get(mode0, 0)            # the oracle returns the rule of a regular pair, 100 (adaptive) or 200 (singular)
            seen.add(mode)
            for a, b in ((h, h0), (g, g0)):
                for blk in (np.s_[:, 0, 0], np.s_[:, 0, 1:], np.s_[:, 1:, 0], np.s_[:, 1:, 1:]):
                    sc = np.abs(b[blk]).max()            # a block may vanish (point in the element's plane): floor from the whole array.  The product's
                    # rays / line integrals / point sets are derived independently of the oracle's (csrc/plan_values.cpp), so a vanishing block is two
                    # different sets of rounding errors, not a copy of the same ones: (x - x_i).n of a flat element is rounding noise ~1e-17 that the
                    # polar quadrature divides by r^2 down to r ~ 1e-3, so both sides hold noise of ~1e-12 of the array's scale there
                    assert np.abs(a[blk] - b[blk]).max() <= 1e-10 * sc + 1e-11 * np.abs(b).max(), (mode, blk, np.abs(a[blk] - b[blk]).max() / sc)
    assert seen == {0, 1, 2}
