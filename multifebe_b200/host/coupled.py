"""Coupled BE regions assembled from SINGLE-REGION assemblies: the route by which several regions reach the device
with the validated single-region kernels, without a new scatter in any kernel.

For every region two auxiliary single-region problems are assembled with the region's own integrator (collocation points flagged
`colloc_elem = -1`, i.e. without free terms):
    H problem   every boundary condition "secondary variable known, value 0"  ->  A_H[r, cH(node, k)] =  sum_e h_e(j, l, k)
    G problem   every boundary condition "primary variable known,  value 0"   ->  A_G[r, cG(node', k)] = -sum_e g_e(j, l, k)
In the G problem the nodes of INTERFACE elements are private copies (one per element and local node), because the coupling coefficient
of g can depend on the element through its unit normal n_fn at the node (t = -p n, Un = u.n ...); everywhere else columns are per node.
The coupled system is then a column combination of those two matrices with the flat scatter descriptors of
MultiRegionModel.scatter_descriptors (one rule for every branch of assemble_bem_har{ela,pot,por}_equation.f90), plus the free terms
(1/2 phi_j at MCA points; c, J c_pot and Mantic's matrix at nodal points) routed through the same descriptors:
    A[row, hcol] += hcoef * H        A[row, gcol_t] += gcoef_t * G        (column -1 = right-hand side)
`assemble_coupled(mrm, omega, local_assemble, freeterm)` is backend-agnostic: `local_assemble(model, region, omega) -> A_loc` is
Problem.build_lse_mechanics_bem_* on the GPU and the single-region oracle in the CPU tests, where the result must equal the multi-region
oracle (tests/test_coupled_from_single_region.py).  The combination itself is O(n_dof^2) column updates; this first version does it with
numpy on host copies of the local matrices -- on the device it is one small kernel over resident matrices.
"""
import numpy as np
from . import shape as sh
from .multiregion import SOLID, FLUID, PORO
from .model import NODAL_XI_MARK


class LocalModel:
    """Duck-typed single-region model (the attributes capi.Problem / the oracles read)."""
    pass


def local_models(mrm, kr):
    """(H model, G model, maps) of region kr.  maps: row_of_node[(node, eq)] = first local row of a collocation node; cH[(node, k)],
    cG[(le, j, k)] = local columns; g_owner[column] = the (le, j, k) that combines it."""
    v, r = mrm.views[kr], mrm.regions[kr]
    nd = r.ndof
    n_g = mrm.n_node
    # local rows: one block of nd rows per collocation NODE of the region (MCA points of a node share the rows), in first-visit order
    row_of_node = {}
    for c in range(v.n_colloc):
        key = (int(v.colloc_node[c]), int(v.colloc_eq[c]))
        if key not in row_of_node:
            row_of_node[key] = len(row_of_node) * nd
    n_rows = len(row_of_node) * nd
    # H columns: per node of the region's elements
    cH = {}
    for le in range(v.n_elem):
        for kn in range(v.elem_ptr[le], v.elem_ptr[le + 1]):
            nde = int(v.elem_node[kn])
            for k in range(nd):
                cH.setdefault((nde, k), len(cH))
    # G nodes: private copies for interface elements
    interface = [mrm.boundary_regions[int(b)][1] is not None for b in v.elem_boundary]
    g_node = np.array(v.elem_node, dtype=np.int32).copy()
    extra_x = []
    for le in range(v.n_elem):
        if interface[le]:
            for kn in range(v.elem_ptr[le], v.elem_ptr[le + 1]):
                extra_x.append(mrm.node_x[int(v.elem_node[kn])]); g_node[kn] = n_g + len(extra_x) - 1
    gcols, cG, g_owner = {}, {}, {}                     # (G node, k) -> column; (le, j, k) -> column; column -> first (le, j, k) that uses it
    for le in range(v.n_elem):
        for j, kn in enumerate(range(v.elem_ptr[le], v.elem_ptr[le + 1])):
            for k in range(nd):
                key = (int(g_node[kn]), k)
                if key not in gcols:
                    gcols[key] = len(gcols); g_owner[gcols[key]] = (le, j, k)
                cG[(le, j, k)] = gcols[key]
    n_gcols = len(gcols)

    def make(elem_node, node_x, cols, n_cols, secondary):
        m = LocalModel()
        nn = len(node_x)
        m.ndof, m.n_node, m.n_elem, m.n_colloc = nd, nn, v.n_elem, v.n_colloc
        m.node_x = np.ascontiguousarray(node_x, dtype=np.float64)
        m.mesh = None
        m.etype, m.elem_ptr, m.elem_reversed = v.etype, v.elem_ptr, v.elem_reversed
        m.elem_node = np.ascontiguousarray(elem_node, dtype=np.int32)
        m.qsi_relative_error, m.qsi_ns_max = mrm.qsi_relative_error, mrm.qsi_ns_max
        m.precalset_gln, m.geometric_tolerance = mrm.precalset_gln, mrm.geometric_tolerance
        m.symplane_eid, m.symplane_t, m.symplane_s = mrm.symplane_eid, mrm.symplane_t, mrm.symplane_s     # the local assemblies integrate the mirror images
        m.colloc_x = v.colloc_x
        m.colloc_node = v.colloc_node
        m.colloc_elem = -np.ones(v.n_colloc, dtype=np.int32)              # no free term: added by assemble_coupled
        m.colloc_kn = np.zeros(v.n_colloc, dtype=np.int32)
        m.colloc_xi = np.full((v.n_colloc, 2), NODAL_XI_MARK, dtype=np.float64)
        m.n_dof = max(n_rows, n_cols, 1)
        m.row = -np.ones((nn, nd), dtype=np.int32)
        for (nde, eq), r0 in row_of_node.items():
            if m.row[nde, 0] >= 0 and m.row[nde, 0] != r0:
                raise ValueError("a node collocates for both sides of an interface inside one region")
            m.row[nde] = r0 + np.arange(nd)
        m.col_u = -np.ones((nn, nd), dtype=np.int32); m.col_t = -np.ones((nn, nd), dtype=np.int32)
        tgt = m.col_t if secondary else m.col_u
        for (nde, k), cidx in cols.items():
            tgt[nde, k] = cidx
        m.ctype = np.full((nn, nd), 0 if secondary else 1, dtype=np.int32)
        m.cvalue = np.zeros((nn, nd), dtype=np.complex128)
        return m
    mH = make(v.elem_node, mrm.node_x, cH, len(cH), False)
    gx = np.vstack([mrm.node_x, np.array(extra_x).reshape(-1, 3)]) if extra_x else mrm.node_x
    mG = make(g_node, gx, gcols, n_gcols, True)
    return mH, mG, dict(row_of_node=row_of_node, cH=cH, cG=cG, g_owner=g_owner, n_rows=n_rows)


def free_term_block(mrm, kr, c, omega, freeterm):
    """(own element le, block (nn, nd, nd)) added to h of the collocation point's own element.  freeterm(normals, tangents, nu, tol) -> (c 3x3, cp)
    is Mantic's matrix and the scalar (solid-angle) free term of the node's element fan."""
    v, r = mrm.views[kr], mrm.regions[kr]
    nd = r.ndof
    le, kn, sn = int(v.colloc_elem[c]), int(v.colloc_kn[c]), int(v.colloc_node[c])
    et = int(v.etype[le]); nodes = v.elem_node[v.elem_ptr[le]:v.elem_ptr[le + 1]]
    nn = len(nodes)
    mat = r.material
    if r.kind == PORO:
        J = poro_J(mat, omega)
        unit = np.diag([J, 1.0, 1.0, 1.0]).astype(np.complex128)
    else:
        unit = np.eye(nd, dtype=np.complex128)
    blk = np.zeros((nn, nd, nd), dtype=np.complex128)
    if v.colloc_xi[c, 0] != NODAL_XI_MARK:
        phi = sh.phi(et, v.colloc_xi[c])
        for j in range(nn):
            blk[j] = 0.5 * phi[j] * unit
        return le, blk
    if et == sh.QUAD9 and kn == 8:
        blk[kn] = 0.5 * unit
        return le, blk
    cache = mrm.__dict__.setdefault("_free_term_cache", {})              # geometry + Poisson's ratio only: kept between frequencies
    if (kr, c) not in cache:
        fans = mrm.__dict__.setdefault("_node_fans", {})
        if kr not in fans:                                               # node -> [(element, local node)] of the region's boundary
            fan = {}
            for le2 in range(v.n_elem):
                for kn2, nde in enumerate(v.elem_node[v.elem_ptr[le2]:v.elem_ptr[le2 + 1]]):
                    fan.setdefault(int(nde), []).append((le2, kn2))
            fans[kr] = fan
        rev = bool(v.elem_reversed[le])
        ns, ts = [], []
        for le2, kn2 in fans[kr][sn]:
            nodes2 = v.elem_node[v.elem_ptr[le2]:v.elem_ptr[le2 + 1]]
            n, tbp, tbm = sh.node_normal_tangents(int(v.etype[le2]), mrm.node_x[nodes2], kn2)
            ns.append(-n if rev else n); ts.append(tbm if rev else tbp)
        nu = mat.nu if r.kind != FLUID else 0.0
        cache[(kr, c)] = freeterm(np.array(ns), np.array(ts), nu, mrm.geometric_tolerance)
    cm, cp = cache[(kr, c)]
    if r.kind == SOLID:
        blk[kn] = cm
    elif r.kind == FLUID:
        blk[kn] = cp
    else:
        blk[kn, 0, 0] = unit[0, 0] * cp; blk[kn, 1:, 1:] = cm
    return le, blk


def combination_terms(mrm, kr, mp, omega, freeterm):
    """The coupled system as a list of column operations on the two local matrices of region kr, shared by the host and the device
    combination: (row_map, terms_H, terms_G, entries) with row_map[r] = global row of local row r; terms_* = lists of (local column,
    global column or -1 for the right-hand side, coefficient) -- `A[row_map, gcol] += coef * A_loc[:n_rows, lcol]`; entries = list of
    (global row, global column or -1, value) for the free terms."""
    v, r = mrm.views[kr], mrm.regions[kr]
    nd = r.ndof
    D = mrm.scatter_descriptors(kr, omega)
    row_map = np.zeros(mp["n_rows"], dtype=np.int32)
    for (nde, eq), r0 in mp["row_of_node"].items():
        row_map[r0:r0 + nd] = mrm.row[(nde, eq)]
    terms_H, terms_G, entries = [], [], []
    done_h = set()
    for le in range(v.n_elem):
        for j, kn in enumerate(range(v.elem_ptr[le], v.elem_ptr[le + 1])):
            nde = int(v.elem_node[kn])
            for k in range(nd):
                q = kn * nd + k
                if (nde, k) not in done_h:                             # the H column of a node is shared by its elements: combine it once
                    done_h.add((nde, k))
                    if int(D["hcol"][q]) >= -1:
                        terms_H.append((mp["cH"][(nde, k)], int(D["hcol"][q]), complex(D["hcoef"][q])))
                cg = mp["cG"][(le, j, k)]
                if mp["g_owner"][cg] == (le, j, k):                    # a per-node G column (ordinary boundary) is combined once
                    for t in range(4):
                        if int(D["gcol"][q, t]) >= -1:
                            terms_G.append((cg, int(D["gcol"][q, t]), -complex(D["gcoef"][q, t])))     # the G problem holds -g
    for c in range(v.n_colloc):                                        # free terms through the h descriptors of the own element
        le, blk = free_term_block(mrm, kr, c, omega, freeterm)
        rows_g = mrm.row[(int(v.colloc_node[c]), int(v.colloc_eq[c]))]
        for j in range(blk.shape[0]):
            for k in range(nd):
                q = (int(v.elem_ptr[le]) + j) * nd + k
                if int(D["hcol"][q]) < -1:
                    continue
                for l in range(nd):
                    if blk[j, l, k] != 0:
                        entries.append((int(rows_g[l]), int(D["hcol"][q]), complex(D["hcoef"][q] * blk[j, l, k]), r.kind == PORO and l == 0 and k == 0))
    return row_map, terms_H, terms_G, entries


def poro_J(mat, omega):
    """J = 1 / ((rho2 + rhoa - i b / omega) omega^2): the only frequency dependence of the free-term entries of a poroelastic region."""
    return 1.0 / ((mat.rho2 + mat.rhoa - 1j * mat.b / omega) * omega ** 2)


class TermArrays:
    """combination_terms as flat arrays, kept between frequencies: the structure (columns, rows) does not depend on the frequency, and the
    coefficients depend on it only through impedance conditions (then the lists are rebuilt) and through J in the fluid-phase free term of a
    poroelastic region (rescaled in place)."""

    def __init__(self, mrm, kr, mp, freeterm):
        self.mrm, self.kr, self.mp, self.freeterm = mrm, kr, mp, freeterm
        self.omega = None
        self.omega_dependent = None

    def at(self, omega):
        r = self.mrm.regions[self.kr]
        if self.omega_dependent is None:
            D1, D2 = self.mrm.scatter_descriptors(self.kr, omega), self.mrm.scatter_descriptors(self.kr, 2.0 * omega)
            self.omega_dependent = not (np.array_equal(D1["hcoef"], D2["hcoef"]) and np.array_equal(D1["gcoef"], D2["gcoef"]))
        if self.omega is None or (self.omega_dependent and omega != self.omega):
            row_map, tH, tG, en = combination_terms(self.mrm, self.kr, self.mp, omega, self.freeterm)
            self.row_map = np.ascontiguousarray(row_map, dtype=np.int32)
            self.H = tuple(np.array([t[i] for t in tH], dtype=dt) for i, dt in ((0, np.int32), (1, np.int32), (2, np.complex128)))
            self.G = tuple(np.array([t[i] for t in tG], dtype=dt) for i, dt in ((0, np.int32), (1, np.int32), (2, np.complex128)))
            self.E = [np.array([e[i] for e in en], dtype=dt) for i, dt in ((0, np.int32), (1, np.int32), (2, np.complex128), (3, bool))]
        elif omega != self.omega and r.kind == PORO and len(self.E[2]):
            self.E[2][self.E[3]] *= poro_J(r.material, omega) / poro_J(r.material, self.omega)
        self.omega = omega
        return self


def assemble_coupled(mrm, omega, local_assemble, freeterm, locals_=None):
    """-> A (n_dof x n_dof), b of the coupled system at frequency omega from two single-region assemblies per region (host combination).
    locals_[kr] = local_models(mrm, kr) when the caller keeps the auxiliary models (and the problems set up from them) between frequencies."""
    n = mrm.n_dof
    A = np.zeros((n, n), dtype=np.complex128); b = np.zeros(n, dtype=np.complex128)
    for kr in range(len(mrm.views)):
        r = mrm.regions[kr]
        mH, mG, mp = locals_[kr] if locals_ is not None else local_models(mrm, kr)
        row_map, terms_H, terms_G, entries = combination_terms(mrm, kr, mp, omega, freeterm)
        inc = mrm.incident.get(kr)
        for model, terms in ((mH, terms_H), (mG, terms_G)):
            if inc is not None and model is mH:                          # incident field: the H problem (all conditions "secondary known, 0") run with
                Aloc, bloc = local_assemble(model, r, omega, incident=inc)     # the field set has b_loc = sum_e (hp u_inc - gp t_inc), free terms apart
                b[row_map] += bloc[:mp["n_rows"]]
            else:
                Aloc = local_assemble(model, r, omega)
            for lcol, gcol, coef in terms:
                if gcol >= 0:
                    A[row_map, gcol] += coef * Aloc[:mp["n_rows"], lcol]
                else:
                    b[row_map] += coef * Aloc[:mp["n_rows"], lcol]
        for row, gcol, val, _ in entries:
            if gcol >= 0:
                A[row, gcol] += val
            else:
                b[row] += val
        for row, val in incident_free_terms(mrm, kr, omega, freeterm):
            b[row] += val
    return A, b


def incident_free_terms(mrm, kr, omega, freeterm):
    """[(global row, value)]: the free term of every collocation point of region kr times the incident field at its own element, the part of
    hp u_inc the auxiliary problems (set up without free terms) leave out.  Empty when the region has no incident field."""
    inc = mrm.incident.get(kr)
    if inc is None:
        return []
    v, nd, out = mrm.views[kr], mrm.regions[kr].ndof, []
    for c in range(v.n_colloc):
        le, blk = free_term_block(mrm, kr, c, omega, freeterm)
        rows_g = mrm.row[(int(v.colloc_node[c]), int(v.colloc_eq[c]))]
        ui = inc[0][int(v.elem_ptr[le]):int(v.elem_ptr[le + 1])]
        for l in range(nd):
            val = np.sum(blk[:, l, :] * ui)
            if val != 0:
                out.append((int(rows_g[l]), complex(val)))
    return out
