"""Local-axes boundary conditions (ctype 2 / 3) on the GPU path: kind 2 of the scatter (h to the column of u_k, -g to the column of t_k) and the host's
condition rows handed over with mfb_set_condition_rows, against the oracle + host rows that tests/test_oracle_local_axes.py pins on the rotated column."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from multifebe_b200.host import Material, Model, cube_mesh, column_analytic_u, without_parts, shape  # noqa: E402
from test_oracle_local_axes import rotation, rotated_column  # noqa: E402

pytestmark = pytest.mark.gpu
MAT = Material(1.0, 1.0, 0.25, 0.03)


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def assert_matrix_parity(md, A, Ao):
    """1e-11 everywhere except the free-term blocks (row of node v, column of u of node v) of a mesh in GENERAL position, where the oracle -- like the
    reference formula it restates, fbem_bem_harela3d_sbie_freeterm: sum of acos(n_i . n_i+1) over the elements around the node, bem_harela3d.f90:478-492 --
    carries its own noise: on a flat face the normals of neighbouring elements are equal up to rounding, n.n' = 1 - O(1e-16), and acos turns that into
    O(1e-8) with a sign taken from a cross product that is pure rounding.  (Measured: oracle 0.4999999976, device 0.4999999999999983 for a node whose
    free term is exactly 1/2; the device forms the solid angle from atan2 of triple products, csrc/plan_values.cpp.)  On axis-aligned meshes the normals
    are bit-identical and this does not arise."""
    D = np.abs(A - Ao); sc = np.abs(Ao).max()
    own = np.zeros(A.shape, dtype=bool)
    for v in range(md.n_node):
        r = md.row[v]; c = md.col_u[v]
        if (r >= 0).all() and (c >= 0).all():
            own[np.ix_(r, c)] = True
    assert D[~own].max() < 1e-11 * sc, D[~own].max() / sc
    assert D[own].max() < 5e-8 * sc, D[own].max() / sc


@pytest.mark.parametrize("etype,m", [(shape.TRI3, 3), (shape.QUAD4, 3), (shape.TRI6, 2), (shape.QUAD8, 2), (shape.QUAD9, 2)], ids=["tri3", "quad4", "tri6", "quad8", "quad9"])
def test_local_axes_parity_and_rotated_column(gpu_ctx, oracle_lib, etype, m):
    from multifebe_b200 import capi
    R = rotation(0.5, 0.3)
    md = rotated_column(m, etype, R, walls=([2, 3, 3], [0.002, 0.0, 0.01j]))
    pr = capi.Problem(gpu_ctx, md); o = oracle_lib.Oracle(md)
    for omega in (0.8, 2.0):
        A, b = pr.build_lse_mechanics_bem_harela(omega, MAT)
        Ao, bo, _ = o.assemble(omega, MAT); md.add_condition_rows(Ao, bo)
        assert_matrix_parity(md, A, Ao)
        assert relerr(b, bo) < 1e-11, (omega, relerr(b, bo))
    x = pr.solve_frequency(2.0, MAT)
    xo = np.linalg.solve(A, b)                       # LAPACK on the device-assembled system (the oracle's free terms carry 1e-9 of their own noise here)
    sc = np.abs(Ao).max(axis=0)                      # u and t unknowns live on different scales
    assert np.abs((x - xo) * sc).max() <= 1e-8 * np.abs(xo * sc).max()
    assert np.abs((x - np.linalg.solve(Ao, bo)) * sc).max() <= 1e-6 * np.abs(xo * sc).max()
    pr.close()
    # the physics, straight from the GPU: homogeneous sliding walls -> the rotated 1D column
    md = rotated_column(m, etype, R)
    pr = capi.Problem(gpu_ctx, md)
    u, _ = md.nodal_solution(pr.solve_frequency(2.0, MAT))
    ue = np.outer(column_analytic_u((md.node_x @ R)[:, 0], 2.0, MAT), R @ np.array([1.0, 0, 0]))
    tol = {shape.TRI3: 6e-2, shape.QUAD4: 4e-2}.get(etype, 2e-3)
    assert relerr(u, ue) < tol
    pr.close()


def test_local_axes_static_symmetry_and_the_two_seam_path(gpu_ctx, oracle_lib):
    """Static analysis; a half model (symmetry plane) whose remaining walls slide in local axes; and the seam where the HOST adds the rows itself: without
    mfb_set_condition_rows the library returns the BEM rows alone, exactly the oracle's."""
    from multifebe_b200 import capi
    R = rotation(0.0, 0.0)
    smat = Material(1.0, 1.0, 0.25, 0.0)
    md = rotated_column(2, shape.QUAD9, rotation(0.4, 0.2))
    pr = capi.Problem(gpu_ctx, md); o = oracle_lib.Oracle(md)
    A, b = pr.build_lse_mechanics_bem_staela(smat)
    Ao, bo, _ = o.assemble_static(smat); Ao = Ao.astype(np.complex128); bo = bo.astype(np.complex128); md.add_condition_rows(Ao, bo)
    assert_matrix_parity(md, A, Ao.real)
    assert relerr(b, bo.real) < 1e-11
    pr.close()
    # half model: the wall y = 0 is a symmetry plane, the others slide in local axes
    mesh = without_parts(cube_mesh(2, shape.QUAD9), {3})
    bcs = {1: ([0, 0, 0], [0, 0, 0]), 2: ([1, 1, 1], [1.0, 0, 0]), 4: ([2, 3, 3], [0, 0, 0]), 5: ([2, 3, 3], [0, 0, 0]), 6: ([2, 3, 3], [0, 0, 0])}
    md = Model(mesh, bcs, symmetry=[("y", "symmetry")])
    pr = capi.Problem(gpu_ctx, md); o = oracle_lib.Oracle(md)
    A, b = pr.build_lse_mechanics_bem_harela(2.0, MAT)
    Ao, bo, _ = o.assemble(2.0, MAT)
    Abem = Ao.copy(); md.add_condition_rows(Ao, bo)
    assert relerr(A, Ao) < 1e-11 and relerr(b, bo) < 1e-11
    u, _ = md.nodal_solution(pr.solve_frequency(2.0, MAT))
    assert relerr(u[:, 0], column_analytic_u(md.node_x[:, 0], 2.0, MAT)) < 2e-3
    # two-seam path: the host keeps the condition rows to itself
    from multifebe_b200.capi import lib, _check
    import ctypes as C
    _check(lib().mfb_set_condition_rows(pr.h, C.c_int(0), None, None, None))
    A2, _ = pr.build_lse_mechanics_bem_harela(2.0, MAT)
    assert relerr(A2, Abem) < 1e-11
    pr.close()


def test_single_frequency_multi_gpu_mode_with_the_round_2_features(gpu_ctx):
    """One frequency over several (virtual) ranks -- row-block assembly, block-cyclic distributed LU -- on a model that uses what round 2 added: a symmetry
    plane, local-axes walls with nonzero prescribed values (their condition rows live in the last rank's row range and must be added once, not once per
    rank), an incident field.  Same solution as the single-GPU path."""
    from multifebe_b200 import capi
    from multifebe_b200.host import plane_wave, element_incident
    mesh = without_parts(cube_mesh(3, shape.QUAD4), {3})
    bcs = {1: ([0, 0, 0], [0, 0, 0]), 2: ([1, 1, 1], [1.0, 0, 0]), 4: ([2, 3, 3], [0.003, 0.0, 0.02j]), 5: ([2, 3, 3], [0, 0.01, 0]), 6: ([10, 10, 10], [0.2, 0.2, 0.2])}
    md = Model(mesh, bcs, symmetry=[("y", "symmetry")])
    u_inc, t_inc = element_incident(md, plane_wave("P", [0.2, 0.0, 1.0], MAT, 2.0, amplitude=0.1))
    pr = capi.Problem(gpu_ctx, md)
    pr.set_incident(u_inc, t_inc)
    x1 = pr.solve_frequency(2.0, MAT)
    for nranks, nb in ((2, 64), (3, 32)):
        pr.dist_init_loopback(nranks, nb)
        x2 = pr.dist_solve_frequency(2.0, MAT)
        sc = np.abs(x1).max()
        assert np.abs(x2 - x1).max() < 1e-9 * sc, (nranks, np.abs(x2 - x1).max() / sc)
    pr.close()
