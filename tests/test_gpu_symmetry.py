"""Symmetry planes on the GPU path against the oracle (which tests/test_oracle_symmetry.py pins against explicitly mirrored full models).

mfb_harela3d_setup_sym == the reference's [symmetry planes] section + the image loop of build_lse_mechanics_bem_harela.f90:1098-1107:
bit-exact quadrature decisions for every (collocation point, element image) pair, A and b within 1e-11, the solution within 1e-8.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from multifebe_b200.host import Material, Model, cube_mesh, without_parts, shape  # noqa: E402
from test_oracle_symmetry import CASES, reduced_and_full  # noqa: E402

pytestmark = pytest.mark.gpu
MAT = Material(rho=1.0, mu=1.0, nu=0.25, xi=0.02)
TOL_A, TOL_X = 1e-11, 1e-8
FREE = ([1, 1, 1], [0, 0, 0])


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name,planes,traction", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("etype,m", [(shape.TRI3, 3), (shape.QUAD9, 2), (shape.TRI6, 2), (shape.QUAD4, 3), (shape.QUAD8, 2)], ids=["tri3", "quad9", "tri6", "quad4", "quad8"])
def test_reduced_models_nodal_collocation_on_the_planes(gpu_ctx, oracle_lib, name, planes, traction, etype, m):
    """Half / quarter models whose plane nodes are collocated nodally: singular integration over the images that touch the collocation node,
    free terms from the mirrored fans (two- and fourfold), signs of symmetric and antisymmetric planes on A (h part) and b (g part)."""
    from multifebe_b200 import capi
    red, _ = reduced_and_full(m, etype, planes, traction)
    pr = capi.Problem(gpu_ctx, red)
    o = oracle_lib.Oracle(red)
    omega = 2.5
    A, b = pr.build_lse_mechanics_bem_harela(omega, MAT)
    Ao, bo, st = o.assemble(omega, MAT)
    assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A, (name, relerr(A, Ao), relerr(b, bo))
    xo, _, _ = oracle_lib.lu_solve(Ao, bo)
    assert relerr(pr.solve_frequency(omega, MAT), xo) < TOL_X
    # discrete decisions of every pair, images included (image ks of element r is element ks * n_elem + r on both sides)
    n_img = red.n_elem << len(planes)
    cs, es = np.meshgrid(np.arange(red.n_colloc), np.arange(n_img), indexing="ij")
    cs, es = cs.ravel(), es.ravel()
    got = pr.plan_modes(cs, es)
    exp = np.array([o.pair_mode(int(e), red.colloc_x[int(c)])[0] for c, e in zip(cs, es)])
    assert np.array_equal(got, exp)
    assert (exp == 200).sum() > (exp[es < red.n_elem] == 200).sum()      # some image touches a collocation node on a plane
    s = pr.stats()
    assert s["PAIRS_SINGULAR"] == st["pairs_singular"] and s["PAIRS_ADAPTIVE"] == st["pairs_adaptive"]
    pr.close()


def test_default_formulation_and_mixed_conditions(gpu_ctx, oracle_lib):
    """The reference's default: the nodes of the open edge in the plane are rim nodes with non-nodal collocation points; several parts with
    different kinds of boundary condition (clamped, loaded, free, mixed per dof), prescribed values that are not zero (b terms of the images)."""
    from multifebe_b200 import capi
    for et, m, planes, drop in [(shape.TRI3, 3, [("x", "symmetry")], {1}), (shape.QUAD9, 2, [("x", "antisymmetry"), ("z", "symmetry")], {1, 5}),
                                (shape.QUAD4, 3, [("x", "symmetry"), ("y", "symmetry"), ("z", "antisymmetry")], {1, 3, 5})]:
        mesh = without_parts(cube_mesh(m, et), drop)
        allb = {2: ([0, 0, 0], [0.1, 0, 0.2j]), 3: FREE, 4: ([1, 0, 1], [0, 0.3, 0.5]), 5: ([0, 1, 1], [0, 0.2, 0]), 6: ([1, 1, 1], [0.4, 0, 0.5])}
        bcs = {p: allb[p] for p in set(int(q) for q in mesh.part)}
        md = Model(mesh, bcs, symmetry=planes)
        assert md.in_boundary.any()
        pr = capi.Problem(gpu_ctx, md); o = oracle_lib.Oracle(md)
        for omega in (0.7, 6.0):
            A, b = pr.build_lse_mechanics_bem_harela(omega, MAT)
            Ao, bo, _ = o.assemble(omega, MAT)
            assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A, (et, omega, relerr(A, Ao), relerr(b, bo))
        xo, _, _ = oracle_lib.lu_solve(Ao, bo)
        assert relerr(pr.solve_frequency(6.0, MAT), xo) < TOL_X
        pr.close()


def test_static_analysis_with_symmetry(gpu_ctx, oracle_lib):
    from multifebe_b200 import capi
    mesh = without_parts(cube_mesh(3, shape.TRI6), {1, 3})
    bcs = {2: ([0, 0, 0], [0, 0, 0]), 4: ([1, 1, 1], [0, 0.3, 0]), 5: FREE, 6: ([1, 1, 1], [0, 0, 0.5])}
    md = Model(mesh, bcs, symmetry=[("x", "symmetry"), ("y", "symmetry")])
    smat = Material(rho=1.0, mu=1.0, nu=0.3, xi=0.0)
    pr = capi.Problem(gpu_ctx, md); o = oracle_lib.Oracle(md)
    A, b = pr.build_lse_mechanics_bem_staela(smat)
    Ao, bo, _ = o.assemble_static(smat)
    assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A
    xo, _, _ = oracle_lib.lu_solve_real(Ao, bo)
    assert relerr(pr.solve_static(smat), xo) < TOL_X
    pr.close()


def test_interior_points_with_symmetry(gpu_ctx, oracle_lib):
    """Displacements and stresses at interior points of a half model: the images contribute to Somigliana's identity and to its
    hypersingular form with the same multipliers (calculate_internal_points_mechanics_bem_harela.f90 loops the same images)."""
    from multifebe_b200 import capi
    name, planes, traction = CASES[3]     # quarter model, two symmetric planes
    red, full = reduced_and_full(3, shape.TRI3, planes, traction)
    pts = np.array([[0.3, 0.4, 0.5], [0.05, 0.6, 0.2], [0.5, 0.02, 0.9]])
    pr = capi.Problem(gpu_ctx, red); o = oracle_lib.Oracle(red)
    omega = 2.5
    x = pr.solve_frequency(omega, MAT)
    ip = capi.InternalPoints(gpu_ctx, red, pts)
    ug = ip.displacements(omega, MAT, x)
    sg = ip.stresses(omega, MAT, x)
    u, t = red.nodal_solution(x)
    n_img = red.n_elem << len(planes)
    uo = np.zeros((len(pts), 3), dtype=np.complex128); so = np.zeros((len(pts), 3, 3), dtype=np.complex128)
    for ipt, xp in enumerate(pts):
        for e in range(n_img):
            nodes = red.mesh.conn[e % red.n_elem]
            h, g = o.pair(e, xp, omega, MAT)[:2]
            uo[ipt] += np.einsum("jlk,jk->l", g, t[nodes]) - np.einsum("jlk,jk->l", h, u[nodes])
            for kc in range(3):
                n_i = np.zeros(3); n_i[kc] = 1.0
                mm, ll, _ = o.pair_hbie(e, xp, n_i, omega, MAT)
                so[ipt, :, kc] += np.einsum("jlk,jk->l", ll, t[nodes]) - np.einsum("jlk,jk->l", mm, u[nodes])
    assert relerr(ug, uo) < 1e-10 and relerr(sg, so) < 1e-10
    # and the physics: the same points inside the explicitly mirrored full model
    prf = capi.Problem(gpu_ctx, full)
    xf = prf.solve_frequency(omega, MAT)
    ipf = capi.InternalPoints(gpu_ctx, full, pts)
    assert relerr(ug, ipf.displacements(omega, MAT, xf)) < 2e-5
    assert relerr(sg, ipf.stresses(omega, MAT, xf)) < 2e-5
    ipf.close(); prf.close(); ip.close(); pr.close()


def test_setup_sym_rejects_bad_planes(gpu_ctx):
    from multifebe_b200 import capi
    mesh = without_parts(cube_mesh(2, shape.TRI3), {1})
    bcs = {p: FREE for p in set(int(q) for q in mesh.part)}
    md = Model(mesh, bcs, symmetry=[("x", "symmetry"), ("y", "symmetry")])
    md.symplane_eid = np.array([2, 1], dtype=np.int32)        # not ascending
    with pytest.raises(capi.MfbError):
        capi.Problem(gpu_ctx, md)
    md.symplane_eid = np.array([1, 2], dtype=np.int32); md.symplane_t = np.array([[-1, 1, 1], [1, 0.5, 1]], dtype=np.float64)
    with pytest.raises(capi.MfbError):
        capi.Problem(gpu_ctx, md)


def test_fluid_and_poroelastic_regions_with_symmetry(gpu_ctx, oracle_lib):
    """symconf_s on the scalar variables: image loops of build_lse_mechanics_bem_harpot / _harpor, against the oracles that
    tests/test_oracle_symmetry.py pins on mirrored full models, the quarter room and the quarter Biot column."""
    from multifebe_b200 import capi
    from multifebe_b200.host import Fluid, FluidModel, Poro, PoroModel
    from test_oracle_symmetry import FLUID_CASES, _generic_reduced_and_full
    fl = Fluid(rho=1.0, c=1.0, xi=0.03)
    for name, planes, flux in FLUID_CASES:
        for etype, m in ((shape.TRI3, 3), (shape.QUAD9, 2)):
            red, _ = _generic_reduced_and_full(m, etype, planes, lambda mesh, **kw: FluidModel(mesh, {1: (1, 0.0)}, **kw), flux)
            pr = capi.Problem(gpu_ctx, red); o = oracle_lib.PotOracle(red)
            A, b = pr.build_lse_mechanics_bem_harpot(2.5, fl)
            Ao, bo, _ = o.assemble(2.5, fl)
            assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A, (name, relerr(A, Ao), relerr(b, bo))
            assert relerr(pr.solve_frequency_fluid(2.5, fl), np.linalg.solve(Ao, bo)) < TOL_X
            pr.close()
    # default formulation (rim nodes with non-nodal points), mixed conditions, nonzero prescribed pressure
    mesh = without_parts(cube_mesh(3, shape.QUAD8), {3, 5})
    md = FluidModel(mesh, {1: (0, 0.0), 2: (0, 1.0 + 0.5j), 4: (1, 0.2), 6: (1, 0.0)}, symmetry=[("y", "symmetry"), ("z", "antisymmetry")])
    pr = capi.Problem(gpu_ctx, md)
    A, b = pr.build_lse_mechanics_bem_harpot(3.0, fl)
    Ao, bo, _ = oracle_lib.PotOracle(md).assemble(3.0, fl)
    assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A
    pr.close()
    po = Poro(rhof=1.0, rhos=2.2, lam=1.2, mu=1.0, xi=0.03, phi=0.35, rhoa=0.15, R=0.8, Q=0.5, b=0.4)
    cases = [([("x", "symmetry")], lambda x: (0.3 + x[1], x[0], 0.3 + x[1], 0.5 * x[2]), shape.QUAD9, 2),
             ([("x", "antisymmetry"), ("z", "symmetry")], lambda x: (x[0], 1.0 + x[1], x[0], x[0] * x[2]), shape.TRI3, 3),
             ([("x", "antisymmetry"), ("y", "antisymmetry"), ("z", "symmetry")], lambda x: (x[0] * x[1], x[1], x[0], x[0] * x[1] * x[2]), shape.QUAD4, 2)]
    for planes, vals, etype, m in cases:
        red, _ = _generic_reduced_and_full(m, etype, planes, lambda mesh, **kw: PoroModel(mesh, {1: ([1, 1, 1, 1], [0, 0, 0, 0])}, **kw), vals)
        pr = capi.Problem(gpu_ctx, red); o = oracle_lib.PorOracle(red)
        A, b = pr.build_lse_mechanics_bem_harpor(2.0, po)
        Ao, bo, _ = o.assemble(2.0, po)
        assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A, (planes, relerr(A, Ao), relerr(b, bo))
        assert relerr(pr.solve_frequency_poro(2.0, po), np.linalg.solve(Ao, bo)) < TOL_X
        pr.close()
    # quarter Biot column: mixed kinds per part (general path of the kernels), MCA points on the planes' rims
    from test_oracle_poroelastic import column_bcs
    mesh = without_parts(cube_mesh(2, shape.QUAD9), {3, 5})
    md = PoroModel(mesh, {k: v for k, v in column_bcs().items() if k in (1, 2, 4, 6)}, symmetry=[("y", "symmetry"), ("z", "symmetry")])
    pr = capi.Problem(gpu_ctx, md)
    A, b = pr.build_lse_mechanics_bem_harpor(2.0, po)
    Ao, bo, _ = oracle_lib.PorOracle(md).assemble(2.0, po)
    assert relerr(A, Ao) < TOL_A and relerr(b, bo) < TOL_A
    pr.close()
