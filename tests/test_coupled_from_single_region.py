"""Coupled regions from single-region assemblies (multifebe_b200/host/coupled.py): the H / G auxiliary problems, the private interface nodes
and the column combination through the flat descriptors must reproduce the multi-region oracle.  Here the single-region assembler is the CPU
oracle; on the GPU it is Problem.build_lse_mechanics_bem_* (tests/test_gpu_coupled.py, pending its first hardware run)."""
import ctypes as C
import numpy as np
import pytest

from multifebe_b200.host import Material, Fluid, Poro, MultiRegionModel, Region, SOLID, FLUID, two_box_mesh, shape
from multifebe_b200.host.multiregion import PORO
from multifebe_b200.host.coupled import assemble_coupled, local_models
from oracle import oracle as orc
from oracle.multiregion import MultiRegionOracle
from test_oracle_multiregion import BPART, LAT1, LAT2, solid_bcs, fluid_bcs, poro_bcs_side, PO

MS, FL = Material(2.0, 1.5, 0.25, 0.03), Fluid(1.0, 1.2, 0.01)


def oracle_local_assemble(model, region, omega):
    if region.kind == SOLID:
        return orc.Oracle(model).assemble(omega, region.material)[0]
    if region.kind == FLUID:
        return orc.PotOracle(model).assemble(omega, region.material)[0]
    return orc.PorOracle(model).assemble(omega, region.material)[0]


def oracle_freeterm(normals, tangents, nu, tol):
    cm, err = orc.freeterm(normals, tangents, nu, tol)
    cp = C.c_double(0.0)
    n_ = np.ascontiguousarray(normals, dtype=np.float64); t_ = np.ascontiguousarray(tangents, dtype=np.float64)
    err = err or orc.lib().orc_freeterm_pot(C.c_int(len(n_)), orc._p(n_), orc._p(t_), C.c_double(tol), C.byref(cp))
    assert not err
    return cm, cp.value


def bcs_for(kind, lat, end, end_value):
    if kind == SOLID:
        sb = solid_bcs(1.0); out = {q: sb[q] for q in lat}; out[end] = ([0, 1, 0], [0.1, 0.2 - 0.1j, 0.0]) if end_value else ([0, 0, 0], [0, 0, 0])
    elif kind == FLUID:
        fb = fluid_bcs(1.0); out = {q: fb[q] for q in lat}; out[end] = (0, 0.7 + 0.1j) if end_value else (1, 0.0)
    else:
        out = poro_bcs_side(lat); out[end] = ([0, 1, 1, 1], [0.3, 1.0, 0.0, 0.2j]) if end_value else ([1, 0, 0, 0], [0, 0, 0, 0])
    return out


@pytest.mark.parametrize("kinds,ict", [((SOLID, SOLID), 0), ((FLUID, FLUID), 0), ((SOLID, FLUID), 0), ((FLUID, SOLID), 0), ((FLUID, PORO), 0), ((PORO, FLUID), 1),
                                       ((SOLID, PORO), 0), ((PORO, PORO), 0)])
@pytest.mark.parametrize("et", [shape.QUAD8, shape.TRI3])
def test_coupled_system_from_single_region_assemblies(kinds, ict, et):
    mats = {SOLID: MS, FLUID: FL, PORO: PO}
    bcs = bcs_for(kinds[0], LAT1, 1, True); bcs.update(bcs_for(kinds[1], LAT2, 2, False))
    mrm = MultiRegionModel(two_box_mesh(1, et), [Region(kinds[0], mats[kinds[0]], [1, 3, 4, 5, 6, 7]), Region(kinds[1], mats[kinds[1]], [-7, 2, 13, 14, 15, 16])],
                           BPART, bcs, interface_ctype={7: ict})
    omega = 1.7
    A0, b0 = MultiRegionOracle(mrm).assemble(omega)
    A1, b1 = assemble_coupled(mrm, omega, oracle_local_assemble, oracle_freeterm)
    # columns of different variables live on different scales
    sc = np.abs(A0).max(axis=0)
    assert (np.abs(A1 - A0).max(axis=0) <= 1e-12 * sc).all(), (np.abs(A1 - A0).max(axis=0) / sc).max()
    assert np.abs(b1 - b0).max() <= 1e-12 * np.abs(b0).max() and np.abs(b0).max() > 0


def test_local_models_are_valid_single_region_inputs():
    mrm = MultiRegionModel(two_box_mesh(1, shape.QUAD9), [Region(SOLID, MS, [1, 3, 4, 5, 6, 7]), Region(FLUID, FL, [-7, 2, 13, 14, 15, 16])], BPART,
                           {**bcs_for(SOLID, LAT1, 1, True), **bcs_for(FLUID, LAT2, 2, False)})
    for kr in (0, 1):
        mH, mG, mp = local_models(mrm, kr)
        nd = mrm.regions[kr].ndof
        for m in (mH, mG):
            assert m.row.shape == (m.n_node, nd) and (m.colloc_elem == -1).all() and m.n_dof >= mp["n_rows"]
            used = m.row[m.row >= 0]
            assert len(set(used.tolist())) == len(used) == mp["n_rows"]                      # every local row owned once
            cols = (m.col_t if m is mG else m.col_u)
            assert (cols[np.unique(m.elem_node)] >= 0).all() and cols.max() < m.n_dof        # every element node has its column
        # private copies: the G problem has one node per (interface element, local node) more than the mesh
        n_if = sum(len(mrm.mesh.conn[e]) for e in mrm.elems_of_boundary[7])
        assert mG.n_node == mrm.n_node + n_if and mH.n_node == mrm.n_node


def test_product_free_term_helper_gives_the_same_system():
    """The same combination with the PRODUCT's free-term helper (mfb_freeterm_terms through capi.freeterm; host only) instead of the oracle's."""
    from multifebe_b200 import capi
    bcs = bcs_for(SOLID, LAT1, 1, True); bcs.update(bcs_for(PORO, LAT2, 2, False))
    mrm = MultiRegionModel(two_box_mesh(2, shape.TRI3), [Region(SOLID, MS, [1, 3, 4, 5, 6, 7]), Region(PORO, PO, [-7, 2, 13, 14, 15, 16])], BPART, bcs)
    A0, b0 = MultiRegionOracle(mrm).assemble(1.3)
    A1, b1 = assemble_coupled(mrm, 1.3, oracle_local_assemble, capi.freeterm)
    sc = np.abs(A0).max(axis=0)
    assert (np.abs(A1 - A0).max(axis=0) <= 1e-12 * sc).all() and np.abs(b1 - b0).max() <= 1e-12 * np.abs(b0).max()


def test_impedance_condition_through_the_single_region_route():
    """A rho c-terminated duct (fluid condition 2, which the single-region device entry points refuse) assembled through the H / G route."""
    from multifebe_b200.host import cube_mesh
    bcs = {1: (0, 1.0), 2: (2, 0.0), 3: (1, 0.0), 4: (1, 0.0), 5: (1, 0.0), 6: (3, 2.5)}
    mrm = MultiRegionModel(cube_mesh(2, shape.QUAD4), [Region(FLUID, FL, [1, 2, 3, 4, 5, 6])], {b: b for b in range(1, 7)}, bcs)
    A0, b0 = MultiRegionOracle(mrm).assemble(3.0)
    A1, b1 = assemble_coupled(mrm, 3.0, oracle_local_assemble, oracle_freeterm)
    sc = np.abs(A0).max(axis=0)
    assert (np.abs(A1 - A0).max(axis=0) <= 1e-12 * sc).all() and np.abs(b1 - b0).max() <= 1e-12 * np.abs(b0).max()


def _same_terms(ta, ref):
    row_map, tH, tG, en = ref
    assert np.array_equal(ta.row_map, row_map)
    for arr, lst in ((ta.H, tH), (ta.G, tG)):
        assert np.array_equal(arr[0], [t[0] for t in lst]) and np.array_equal(arr[1], [t[1] for t in lst])
        assert np.allclose(arr[2], [t[2] for t in lst], rtol=1e-14, atol=0)
    assert np.array_equal(ta.E[0], [e[0] for e in en]) and np.array_equal(ta.E[1], [e[1] for e in en])
    assert np.allclose(ta.E[2], [e[2] for e in en], rtol=1e-13, atol=0)


def test_term_arrays_kept_between_frequencies_equal_fresh_term_lists():
    """TermArrays (what CoupledProblem.solve_frequency_resident hands to the device): the second frequency reuses the structure, rescales J of a
    poroelastic region, and rebuilds the lists when an impedance condition makes the coefficients frequency dependent."""
    from multifebe_b200.host.coupled import TermArrays, combination_terms, local_models
    from multifebe_b200.host import cube_mesh
    bcs = bcs_for(FLUID, LAT1, 1, True); bcs.update(bcs_for(PORO, LAT2, 2, False))
    mrm = MultiRegionModel(two_box_mesh(1, shape.QUAD8), [Region(FLUID, FL, [1, 3, 4, 5, 6, 7]), Region(PORO, PO, [-7, 2, 13, 14, 15, 16])], BPART, bcs,
                           interface_ctype={7: 0})
    for kr in range(2):
        mp = local_models(mrm, kr)[2]
        ta = TermArrays(mrm, kr, mp, oracle_freeterm)
        for omega in (1.7, 0.9, 0.9, 3.1):
            _same_terms(ta.at(omega), combination_terms(mrm, kr, mp, omega, oracle_freeterm))
        assert ta.omega_dependent is False
    duct = MultiRegionModel(cube_mesh(1, shape.QUAD8), [Region(FLUID, FL, [1, 2, 3, 4, 5, 6])], {b: b for b in range(1, 7)},
                            {1: (0, 1.0), 2: (2, 0.0), 3: (1, 0.0), 4: (1, 0.0), 5: (1, 0.0), 6: (1, 0.0)})
    mp = local_models(duct, 0)[2]
    ta = TermArrays(duct, 0, mp, oracle_freeterm)
    for omega in (4.0, 2.5):
        _same_terms(ta.at(omega), combination_terms(duct, 0, mp, omega, oracle_freeterm))
    assert ta.omega_dependent is True


def _random_incident(mrm, kr, seed):
    v = mrm.views[kr]
    rng = np.random.default_rng(seed)
    n = int(v.elem_ptr[-1])
    return (rng.normal(size=(n, v.ndof)) + 1j * rng.normal(size=(n, v.ndof)), rng.normal(size=(n, v.ndof)) + 1j * rng.normal(size=(n, v.ndof)))


@pytest.mark.parametrize("kinds,where", [((SOLID, FLUID), (0,)), ((FLUID, PORO), (0, 1)), ((PORO, SOLID), (1,)), ((SOLID, SOLID), (0, 1))])
def test_incident_field_of_a_coupled_region_through_the_single_region_route(kinds, where):
    """region%n_incidentfields > 0 in a coupled model: every pair of the region, interface elements included, adds hp u_inc - gp t_inc to b
    (assemble_bem_har{ela,pot,por}_equation.f90, the block after the coupling `select case`).  The H problem of the region run with the field set
    must give the same right-hand side as the multi-region oracle, for any arrays."""
    mats = {SOLID: MS, FLUID: FL, PORO: PO}
    bcs = bcs_for(kinds[0], LAT1, 1, True); bcs.update(bcs_for(kinds[1], LAT2, 2, False))
    mrm = MultiRegionModel(two_box_mesh(1, shape.QUAD8), [Region(kinds[0], mats[kinds[0]], [1, 3, 4, 5, 6, 7]), Region(kinds[1], mats[kinds[1]], [-7, 2, 13, 14, 15, 16])],
                           BPART, bcs)
    omega = 1.7
    A_plain, b_plain = MultiRegionOracle(mrm).assemble(omega)
    for kr in where:
        mrm.set_incident(kr, *_random_incident(mrm, kr, 10 + kr))

    def local(model, region, om, incident=None):
        o = {SOLID: orc.Oracle, FLUID: orc.PotOracle}.get(region.kind, orc.PorOracle)(model)
        if incident is not None:
            o.set_incident(*incident)
        res = o.assemble(om, region.material)
        return res[0] if incident is None else (res[0], res[1])
    A0, b0 = MultiRegionOracle(mrm).assemble(omega)
    A1, b1 = assemble_coupled(mrm, omega, local, oracle_freeterm)
    sc = np.abs(A0).max(axis=0)
    assert np.array_equal(A0, A_plain) and np.abs(b0 - b_plain).max() > 0.1 * np.abs(b_plain).max()      # the field only changes b, and does change it
    assert (np.abs(A1 - A0).max(axis=0) <= 1e-12 * sc).all()
    assert np.abs(b1 - b0).max() <= 1e-12 * np.abs(b0).max()
    mrm.set_incident(where[0])                                                                          # cleared again
    assert where[0] not in mrm.incident


def test_coupled_incident_no_scattering_identity():
    """Two solid regions of the same material, the same incident field in both: with u_inc per node, t_inc = +-t per interface node (the outward normals
    of the two regions are opposite) and zero on the traction-free outer faces, the total field IS the incident one -- an algebraic identity of
    H (u - u_inc) = G (t - t_inc) per region with u1 = u2, t1 = -t2 on the interface, which pins the sign the reversed interface elements take."""
    free = ([1, 1, 1], [0, 0, 0])
    bcs = {q: free for q in (1, 2) + LAT1 + LAT2}
    mrm = MultiRegionModel(two_box_mesh(2, shape.TRI6), [Region(SOLID, MS, [1, 3, 4, 5, 6, 7]), Region(SOLID, MS, [-7, 2, 13, 14, 15, 16])], BPART, bcs)
    rng = np.random.default_rng(5)
    u_node = rng.normal(size=(mrm.n_node, 3)) + 1j * rng.normal(size=(mrm.n_node, 3))
    t_node = np.zeros((mrm.n_node, 3), dtype=np.complex128)
    if_nodes = np.unique(np.concatenate([mrm.mesh.conn[e] for e in mrm.elems_of_boundary[7]]))
    t_node[if_nodes] = rng.normal(size=(len(if_nodes), 3)) + 1j * rng.normal(size=(len(if_nodes), 3))   # traction of region 1 on the interface
    for kr, sg in ((0, 1.0), (1, -1.0)):
        v = mrm.views[kr]
        mrm.set_incident(kr, u_node[v.elem_node], sg * t_node[v.elem_node])
    A, b = MultiRegionOracle(mrm).assemble(1.3)
    x = np.linalg.solve(A, b)
    u1, t1 = mrm.nodal_solution(x, 0)
    u2, t2 = mrm.nodal_solution(x, 1)
    for u, t, sg in ((u1, t1, 1.0), (u2, t2, -1.0)):
        ok = ~np.isnan(u[:, 0])
        assert np.abs(u[ok] - u_node[ok]).max() < 1e-9 * np.abs(u_node).max()
        assert np.abs(t[ok] - sg * t_node[ok]).max() < 1e-9 * np.abs(t_node).max()


def test_coupled_model_with_a_bem_formulation_per_boundary():
    """[bem formulation over boundaries] on a coupled model: MCA points at every node of the interface (sbie_mca, default displacement of the element order) and a
    rim displacement of 0.02 on another boundary; the views, the auxiliary single-region models and the multi-region oracle must all see the same points."""
    bcs = bcs_for(SOLID, LAT1, 1, True); bcs.update(bcs_for(FLUID, LAT2, 2, False))
    regs = [Region(SOLID, MS, [1, 3, 4, 5, 6, 7]), Region(FLUID, FL, [-7, 2, 13, 14, 15, 16])]
    plain = MultiRegionModel(two_box_mesh(1, shape.QUAD8), regs, BPART, bcs)
    mrm = MultiRegionModel(two_box_mesh(1, shape.QUAD8), regs, BPART, bcs, formulation={7: ("sbie_mca", 0.0), 1: ("sbie_boundary_mca", 0.02)})
    on7 = mrm.node_boundary == 7
    assert set(mrm.mca_delta[on7]) == {-1.0} and set(mrm.mca_delta[(mrm.node_boundary == 1) & mrm.in_boundary]) == {0.02}
    assert mrm.n_dof == plain.n_dof and mrm.views[0].n_colloc == plain.views[0].n_colloc          # one quad8 per face: every node is a rim node already ...
    k = int(np.flatnonzero(mrm.views[0].colloc_node == np.flatnonzero(on7)[0])[0])
    assert not np.allclose(mrm.views[0].colloc_x[k], plain.views[0].colloc_x[k])                   # ... but the points sit at 0.2254 instead of 0.05 from the rim
    A0, b0 = MultiRegionOracle(mrm).assemble(1.7)
    A1, b1 = assemble_coupled(mrm, 1.7, oracle_local_assemble, oracle_freeterm)
    sc = np.abs(A0).max(axis=0)
    assert (np.abs(A1 - A0).max(axis=0) <= 1e-12 * sc).all() and np.abs(b1 - b0).max() <= 1e-12 * np.abs(b0).max()
    # the physics does not depend on where the points are: the solutions of the two collocation schemes agree to the discretisation error
    Ap, bp = MultiRegionOracle(plain).assemble(1.7)
    u0, _ = mrm.nodal_solution(np.linalg.solve(A0, b0), 0); up, _ = plain.nodal_solution(np.linalg.solve(Ap, bp), 0)
    ok = ~np.isnan(u0[:, 0])
    assert np.abs(u0[ok] - up[ok]).max() < 0.1 * np.abs(up[ok]).max()
