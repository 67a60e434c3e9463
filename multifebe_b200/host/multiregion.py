"""Several BE regions coupled through be-be interface boundaries: host-side numbering and the
per-region views, built the way the reference's host does.

  * regions           src/read_regions.f90: every region lists its boundaries; a NEGATIVE id means the boundary is seen with reversed
                      orientation from this region (it is region 2 of that boundary), so a boundary listed by two regions is a be-be
                      interface whose mesh normal points out of its region 1.
  * DOF numbering     src/build_auxiliary_variables_mechanics_harmonic.f90:68-900: regions -> boundaries -> elements -> nodes in
                      first-visit order.  Ordinary boundary: as in host/model.py.  Interface (perfect bonding / perfect contact):
                        fluid(1)-fluid(2)   rows (1,1), (1,2); columns p1, Un1          (p2 = p1, Un2 = -Un1)                     :593-622
                        fluid(1)-solid(2)   row (1,1), column p1; rows (k,2), columns u2_k  (Un1 = u2.n1, t2 = -p1 n2)            :626-661
                        solid(1)-fluid(2)   rows (k,1), columns u1_k; row (1,2), column p2  (t1 = -p2 n1, Un2 = u1.n2)            :751-792
                        solid(1)-solid(2)   per k: rows (k,1), (k,2); columns u1_k, t1_k    (u2 = u1, t2 = -t1; ctype 0)          :800-836
  * collocation       every region collocates at the nodes of ITS boundaries (nodal SBIE, or MCA on part rims as in host/model.py) and
                      writes the equation into rows (., eq) with eq = 1 when the region is region 1 of the collocation boundary, else 2
                      (src/build_lse_mechanics_bem_harela.f90:1118-1136, src/build_lse_mechanics_bem_harpot.f90).
Only the numbering, the views and the nodal-solution map live here; the integrals are the single-region ones.  The device path for
coupled regions is capi.CoupledProblem over host/coupled.py: this model feeds it and the multi-region oracle.
Poroelastic regions (four components per node) couple to fluids (perfectly permeable or impermeable, `interface_ctype`), to solids (bonded,
impervious) and to each other (perfectly permeable): numbering build_auxiliary_variables_mechanics_harmonic.f90:626-700, :895-1230.
"""
import numpy as np
from . import shape as sh
from .model import MCA_BOUNDARY_DELTA, NODAL_XI_MARK, mca_deltas, mca_delta_of

SOLID, FLUID, PORO = "solid", "fluid", "poro"


class Region:
    def __init__(self, kind, material, boundaries):
        """kind SOLID | FLUID; material host.Material | host.Fluid; boundaries: signed boundary ids (negative = reversed)."""
        if kind not in (SOLID, FLUID, PORO):
            raise ValueError("region kind must be 'solid', 'fluid' or 'poro'")
        self.kind, self.material, self.boundaries = kind, material, [int(b) for b in boundaries]
        self.ndof = {SOLID: 3, FLUID: 1, PORO: 4}[kind]


def _var_name(kind, k, secondary, side):
    """Key of the column of component k of a node variable: solid u/t, fluid p/un, poroelastic tau/w (k = 0) and u/t (k = 1..3)."""
    if kind == FLUID:
        return ("un%d" if secondary else "p%d") % side
    if kind == SOLID:
        return ("t%d%d" if secondary else "u%d%d") % (side, k)
    if k == 0:
        return ("w%d" if secondary else "tau%d") % side
    return ("t%d%d" if secondary else "u%d%d") % (side, k - 1)


class RegionView:
    """What one region's integrator needs (the flat arrays of host.Model, on the GLOBAL node set): its elements with the orientation
    seen from the region, and its collocation points."""
    pass


class MultiRegionModel:
    def __init__(self, mesh, regions, boundary_part, bcs, qsi_relative_error=1e-6, qsi_ns_max=16, precalset_gln=(2, 3, 4, 5, 6, 7, 8, 9),
                 geometric_tolerance=1e-6, interface_ctype=None, symmetry=None, formulation=None):
        """boundary_part {boundary id: part id of the mesh}; bcs {boundary id: (ctypes, values)} for the ORDINARY boundaries (solid: three
        components, fluid: scalars, poroelastic: tau / Un then the three skeleton components), 0 = primary variable known, 1 = secondary
        variable known.  interface_ctype {boundary id: 0 | 1}: condition of a fluid-poroelastic interface, 0 perfectly permeable (default),
        1 perfectly impermeable (node%ctype(1,1) of such a boundary)."""
        self.interface_ctype = dict(interface_ctype or {})
        self.mesh, self.regions = mesh, list(regions)
        # [symmetry planes] of the model (every BE region integrates the mirror images of its elements; the nodes of the open edges in the planes are rim
        # nodes with non-nodal collocation points, the reference's default)
        from .model import symmetry_planes, symmetry_scalars
        self.symplane_eid, self.symplane_t = symmetry_planes(symmetry)
        self.symplane_s = symmetry_scalars(symmetry, self.symplane_eid, self.symplane_t)
        nn, ne = len(mesh.nodes), mesh.n_elem
        self.n_node, self.n_elem = nn, ne
        self.node_x = mesh.nodes
        self.qsi_relative_error, self.qsi_ns_max = float(qsi_relative_error), int(qsi_ns_max)
        self.precalset_gln = np.array(precalset_gln, dtype=np.int32)
        self.geometric_tolerance = float(geometric_tolerance)
        # --- boundaries: which regions use them
        self.boundary_part = dict(boundary_part)
        self.boundary_regions = {b: [None, None] for b in boundary_part}          # [region 1 (positive), region 2 (reversed)]
        for kr, r in enumerate(self.regions):
            for b in r.boundaries:
                slot = 0 if b > 0 else 1
                if abs(b) not in self.boundary_regions or self.boundary_regions[abs(b)][slot] is not None:
                    raise ValueError("boundary %d: unknown, or listed twice with the same orientation" % abs(b))
                self.boundary_regions[abs(b)][slot] = kr
        for b, (r1, r2) in self.boundary_regions.items():
            if r1 is None:
                raise ValueError("boundary %d has no region 1 (it must be listed with a positive id by one region)" % b)
        part_boundary = {p: b for b, p in self.boundary_part.items()}
        # --- part rims (nodes of edges owned by a single element of the part) and node -> boundary
        node_b = -np.ones(nn, dtype=np.int64)
        edge_count = {}
        for e in range(ne):
            c, p = mesh.conn[e], int(mesh.part[e])
            b = part_boundary[p]
            for v in c:
                if node_b[v] not in (-1, b):
                    raise ValueError("node %d is shared by two boundaries; each boundary must own its nodes" % v)
                node_b[v] = b
            for ed in sh.edges_of(int(mesh.etype[e])):
                key = (p,) + tuple(sorted((int(c[ed[0]]), int(c[ed[1]]))))
                edge_count.setdefault(key, []).append([int(c[k]) for k in ed])
        in_boundary = np.zeros(nn, dtype=bool)
        for lst in edge_count.values():
            if len(lst) == 1:
                in_boundary[lst[0]] = True
        self.node_boundary, self.in_boundary = node_b, in_boundary
        self.mca_delta = mca_deltas(in_boundary, node_b, formulation)         # formulation {boundary id: (kind, delta)}: [bem formulation over boundaries]
        self.elems_of_boundary = {b: [e for e in range(ne) if int(mesh.part[e]) == p] for b, p in self.boundary_part.items()}
        # --- boundary conditions of the ordinary boundaries
        self.ctype = {}
        self.cvalue = {}
        for b, (r1, r2) in self.boundary_regions.items():
            if r2 is None:
                nd = self.regions[r1].ndof
                ct, cv = bcs[b]
                ct = np.atleast_1d(np.asarray(ct, dtype=np.int32)); cv = np.atleast_1d(np.asarray(cv, dtype=np.complex128))
                allowed = {0, 1, 2, 3} if self.regions[r1].kind == FLUID else {0, 1}     # fluid: 2 = rho c impedance, 3 = spherical radiation with radius cvalue
                if len(ct) != nd or len(cv) != nd or not set(ct.tolist()) <= allowed:
                    raise ValueError("boundary %d: %d condition(s) of type %s expected" % (b, nd, sorted(allowed)))
                self.ctype[b], self.cvalue[b] = ct, cv
        # --- numbering.  row[(node, eq)] -> list of rows (1 or 3); col[(node, name)] -> column, names 'p1','un1','p2','u1k','t1k','u2k'
        self.row, self.col = {}, {}
        row = col = 0
        used = np.zeros(nn, dtype=bool)
        for kr, r in enumerate(self.regions):
            for sb in r.boundaries:
                b = abs(sb)
                r1, r2 = self.boundary_regions[b]
                for e in self.elems_of_boundary[b]:
                    for v in mesh.conn[e]:
                        v = int(v)
                        if used[v]:
                            continue
                        used[v] = True
                        if r2 is None:                                   # ordinary boundary of region r1
                            nd = self.regions[r1].ndof
                            self.row[(v, 1)] = list(range(row, row + nd)); row += nd
                            for k in range(nd):
                                self.col[(v, _var_name(self.regions[r1].kind, k, self.ctype[b][k] == 0, 1))] = col; col += 1
                            continue
                        k1, k2 = self.regions[r1].kind, self.regions[r2].kind
                        if {k1, k2} == {SOLID, PORO}:
                            # perfect bonding, impervious contact (ctype 0): tau, u_k, t_k of the poroelastic side are active; its fluid-phase row first,
                            # then the rows of both regions interleaved per component (:895-930 solid(1)-poro(2), :1065-1098 poro(1)-solid(2))
                            pe = 1 if k1 == PORO else 2
                            se_ = 3 - pe
                            prow = [row]; row += 1
                            self.col[(v, "tau%d" % pe)] = col; col += 1
                            r1rows, r2rows = [], []
                            for k in range(3):
                                r1rows.append(row); r2rows.append(row + 1); row += 2
                                self.col[(v, "u%d%d" % (pe, k))] = col; self.col[(v, "t%d%d" % (pe, k))] = col + 1; col += 2
                            self.row[(v, pe)] = prow + (r1rows if pe == 1 else r2rows)
                            self.row[(v, se_)] = r2rows if pe == 1 else r1rows
                        elif k1 == PORO and k2 == PORO:
                            # perfectly permeable contact (ctype 0): tau, Un, u_k, t_k of region 1 are active; rows of both regions interleaved per
                            # component (build_auxiliary_variables_mechanics_harmonic.f90:1130-1230)
                            if self.interface_ctype.get(b, 0) != 0:
                                raise ValueError("boundary %d: of the poroelastic-poroelastic contacts only the perfectly permeable one is built" % b)
                            r1rows, r2rows = [], []
                            for k in range(4):
                                r1rows.append(row); r2rows.append(row + 1); row += 2
                                self.col[(v, _var_name(PORO, k, False, 1))] = col; self.col[(v, _var_name(PORO, k, True, 1))] = col + 1; col += 2
                            self.row[(v, 1)] = r1rows; self.row[(v, 2)] = r2rows
                        elif PORO in (k1, k2):
                            imp = self.interface_ctype.get(b, 0) == 1
                            fe, pe = (1, 2) if k1 == FLUID else (2, 1)           # equation index / variable suffix of the fluid and of the poroelastic side
                            def number_poro():
                                nonlocal row, col
                                self.row[(v, pe)] = list(range(row, row + 4)); row += 4
                                self.col[(v, "tau%d" % pe)] = col; col += 1
                                if not imp:
                                    self.col[(v, "w%d" % pe)] = col; col += 1
                                for k in range(3):
                                    self.col[(v, "u%d%d" % (pe, k))] = col; col += 1
                            def number_fluid():
                                nonlocal row, col
                                self.row[(v, fe)] = [row]; row += 1
                                if imp:
                                    self.col[(v, "p%d" % fe)] = col; col += 1
                            # the fluid equation comes first in both orders (:626-700 fluid(1)-poro(2), :990-1037 poro(1)-fluid(2))
                            number_fluid(); number_poro()
                        elif k1 == FLUID and k2 == FLUID:
                            self.row[(v, 1)] = [row]; self.row[(v, 2)] = [row + 1]; row += 2
                            self.col[(v, "p1")] = col; self.col[(v, "un1")] = col + 1; col += 2
                        elif k1 == FLUID and k2 == SOLID:
                            self.row[(v, 1)] = [row]; row += 1
                            self.col[(v, "p1")] = col; col += 1
                            self.row[(v, 2)] = list(range(row, row + 3)); row += 3
                            for k in range(3):
                                self.col[(v, "u2%d" % k)] = col; col += 1
                        elif k1 == SOLID and k2 == FLUID:
                            self.row[(v, 1)] = list(range(row, row + 3)); row += 3
                            for k in range(3):
                                self.col[(v, "u1%d" % k)] = col; col += 1
                            self.row[(v, 2)] = [row]; row += 1
                            self.col[(v, "p2")] = col; col += 1
                        else:                                            # solid - solid, perfect bonding
                            r1rows, r2rows = [], []
                            for k in range(3):
                                r1rows.append(row); r2rows.append(row + 1); row += 2
                                self.col[(v, "u1%d" % k)] = col; self.col[(v, "t1%d" % k)] = col + 1; col += 2
                            self.row[(v, 1)] = r1rows; self.row[(v, 2)] = r2rows
        if row != col:
            raise ValueError("the system is not square: %d equations, %d unknowns" % (row, col))
        self.n_dof = row
        # --- per-region views
        self.views = [self._view(kr) for kr in range(len(self.regions))]
        self.incident = {}                                  # region index -> (u_inc, t_inc) at the nodes of the region's elements, see set_incident

    def _view(self, kr):
        mesh, r = self.mesh, self.regions[kr]
        v = RegionView()
        v.kind, v.material, v.ndof = r.kind, r.material, r.ndof
        elems, rev, ebnd = [], [], []
        for sb in r.boundaries:
            for e in self.elems_of_boundary[abs(sb)]:
                elems.append(e); rev.append(1 if sb < 0 else 0); ebnd.append(abs(sb))
        v.elem_global = np.array(elems, dtype=np.int32)
        v.elem_boundary = np.array(ebnd, dtype=np.int32)
        v.n_node, v.n_elem = self.n_node, len(elems)
        v.node_x = self.node_x
        v.mesh = mesh
        v.etype = mesh.etype[elems].astype(np.int32)
        v.elem_ptr = np.zeros(len(elems) + 1, dtype=np.int32)
        v.elem_ptr[1:] = np.cumsum([len(mesh.conn[e]) for e in elems])
        v.elem_node = np.concatenate([mesh.conn[e] for e in elems]).astype(np.int32)
        v.elem_reversed = np.array(rev, dtype=np.uint8)
        v.qsi_relative_error, v.qsi_ns_max = self.qsi_relative_error, self.qsi_ns_max
        v.precalset_gln, v.geometric_tolerance = self.precalset_gln, self.geometric_tolerance
        v.symplane_eid, v.symplane_t, v.symplane_s = self.symplane_eid, self.symplane_t, self.symplane_s
        # collocation points of the region (loop order of the reference: boundaries, elements, nodes)
        cx, cnode, celem, ckn, cxi, ceq = [], [], [], [], [], []
        collocated = np.zeros(self.n_node, dtype=bool)
        for le, e in enumerate(elems):
            et = int(mesh.etype[e]); c = mesh.conn[e]
            xn = self.node_x[c]
            eq = 1 if self.boundary_regions[ebnd[le]][0] == kr else 2
            for kn, nd in enumerate(c):
                nd = int(nd)
                if self.mca_delta[nd] != 0.0:
                    xi = sh.move_xi_from_edge(et, sh.XI_NODES[et][kn], mca_delta_of(et, self.mca_delta[nd]))
                    cx.append(sh.position(et, xn, xi)); cxi.append(xi)
                elif not collocated[nd]:
                    collocated[nd] = True
                    cx.append(sh.position(et, xn, sh.XI_NODES[et][kn])); cxi.append([NODAL_XI_MARK, NODAL_XI_MARK])
                else:
                    continue
                cnode.append(nd); celem.append(le); ckn.append(kn); ceq.append(eq)
        v.colloc_x = np.ascontiguousarray(cx, dtype=np.float64)
        v.colloc_node = np.array(cnode, dtype=np.int32)
        v.colloc_elem = np.array(celem, dtype=np.int32)          # LOCAL element index of the view
        v.colloc_kn = np.array(ckn, dtype=np.int32)
        v.colloc_xi = np.ascontiguousarray(cxi, dtype=np.float64)
        v.colloc_eq = np.array(ceq, dtype=np.int32)
        v.n_colloc = len(cnode)
        # placeholders for the single-region set-up signatures (the multi-region scatter does not use them)
        z = np.zeros((self.n_node, r.ndof), dtype=np.int32)
        v.row, v.col_u, v.col_t, v.ctype = z, z, z, np.ones((self.n_node, r.ndof), dtype=np.int32)
        v.cvalue = np.zeros((self.n_node, r.ndof), dtype=np.complex128)
        v.n_dof = 1
        return v

    # ---- flat scatter descriptors of one region: the form in which the coupling crosses the C ABI
    def set_incident(self, kr, u_inc=None, t_inc=None):
        """Incident wave field of region kr (region%n_incidentfields > 0): the primary and the secondary variables of the incident field at the
        nodes of every element of the region's view, (sum nn, ndof) complex each in the view's element order, the secondary ones formed with the
        REGION's outward normal (views[kr].elem_reversed applied by the caller).  Every (collocation point, element) pair of the region, ordinary
        boundaries and interfaces alike, then adds hp u_inc - gp t_inc to the right-hand side, free term included
        (src/assemble_bem_harela_equation.f90:651-666, assemble_bem_harpot_equation.f90:471-481, assemble_bem_harpor_equation.f90:1277-1289; the
        unknowns of the coupled system are the TOTAL fields).  The field depends on the frequency: set it before every assembly.  None clears it."""
        if u_inc is None:
            self.incident.pop(kr, None); return
        v = self.views[kr]
        n = int(v.elem_ptr[-1])
        u = np.ascontiguousarray(u_inc, dtype=np.complex128).reshape(n, v.ndof); t = np.ascontiguousarray(t_inc, dtype=np.complex128).reshape(n, v.ndof)
        self.incident[kr] = (u, t)

    def impedance_coefficient(self, kr, b, omega):
        """Un = -coef * p on a fluid boundary with condition 2 (Un = -i/(rho c omega) p) or 3 (Un = -(i/(rho c omega) + 1/(2 R rho omega^2)) p,
        R = the prescribed value): assemble_bem_harpot_equation.f90:97-110."""
        mat = self.regions[kr].material
        coef = 1j / (mat.rho * mat.c * omega)
        if self.ctype[b][0] == 3:
            coef = coef + 1.0 / (2.0 * self.cvalue[b][0] * mat.rho * omega ** 2)
        return coef

    def scatter_descriptors(self, kr, omega=None):
        """Every case of assemble_bem_har{ela,pot,por}_equation.f90 reduced to one rule.  For the element instance le of the region, its node j
        and source component k (solid: k = 0..2; fluid: k = 0; poroelastic: k = 0 fluid phase, 1..3 skeleton) = column index of the node block:
            A[row_l, hcol] += hcoef * h(j, l, k)          (hcol == -1: b[row_l] += hcoef * h;  -2: nothing)
            A[row_l, gcol_t] += gcoef_t * g(j, l, k)      t = 0..3 (same conventions)
        with g of a fluid element already multiplied by rho omega^2.  Index = (elem_ptr[le] + j) * ndof + k (g targets: * 4 + t).
        Returns dict(hcol, hcoef, gcol, gcoef)."""
        v, r = self.views[kr], self.regions[kr]
        nd = r.ndof
        n = int(v.elem_ptr[-1]) * nd
        hcol = np.full(n, -2, dtype=np.int32); hcoef = np.zeros(n, dtype=np.complex128)
        gcol = np.full((n, 4), -2, dtype=np.int32); gcoef = np.zeros((n, 4), dtype=np.complex128)
        for le in range(v.n_elem):
            bnd = int(v.elem_boundary[le])
            r1, r2 = self.boundary_regions[bnd]
            first = r1 == kr
            side = 1 if first else 2
            et = int(v.etype[le])
            nodes = v.elem_node[v.elem_ptr[le]:v.elem_ptr[le + 1]]
            for j, sn in enumerate(nodes):
                sn = int(sn)
                n_fn = sh.unit_normal(et, self.node_x[nodes], sh.XI_NODES[et][j]) if r2 is not None else None
                for k in range(nd):
                    q = (int(v.elem_ptr[le]) + j) * nd + k
                    if r2 is None:
                        ct, cv = self.ctype[bnd][k], self.cvalue[bnd][k]
                        if ct in (2, 3):                              # p unknown, Un = -coef p: A(row, col_p) += h + coef g
                            if omega is None:
                                raise ValueError("boundary %d: impedance conditions need the frequency" % bnd)
                            hcol[q], hcoef[q] = self.col[(sn, "p1")], 1.0
                            gcol[q, 0], gcoef[q, 0] = self.col[(sn, "p1")], self.impedance_coefficient(kr, bnd, omega)
                        elif ct == 0:
                            hcol[q], hcoef[q] = -1, -cv; gcol[q, 0], gcoef[q, 0] = self.col[(sn, _var_name(r.kind, k, True, 1))], -1.0
                        else:
                            hcol[q], hcoef[q] = self.col[(sn, _var_name(r.kind, k, False, 1))], 1.0; gcol[q, 0], gcoef[q, 0] = -1, cv
                        continue
                    k1, k2 = self.regions[r1].kind, self.regions[r2].kind
                    other = k2 if first else k1
                    sgn = 1.0 if first else -1.0                              # n_fn is outward from region 1: the normal of THIS region is sgn * n_fn
                    if k1 == PORO and k2 == PORO:        # assemble_bem_harpor_equation.f90:696-722 (region 1) / :860-975 (region 2), perfectly permeable
                        f1, f2 = self.regions[r1].material.phi, self.regions[r2].material.phi
                        hcol[q] = self.col[(sn, _var_name(PORO, k, False, 1))]
                        if first:
                            hcoef[q] = 1.0
                            gcol[q, 0], gcoef[q, 0] = self.col[(sn, _var_name(PORO, k, True, 1))], -1.0
                        elif k == 0:                      # tau2 = phi2/phi1 tau1; Un2 = -phi1/phi2 Un1 - (1 - phi1/phi2) u1 . n1
                            hcoef[q] = f2 / f1
                            gcol[q, 0], gcoef[q, 0] = self.col[(sn, "w1")], f1 / f2
                            for t in range(3):
                                gcol[q, 1 + t], gcoef[q, 1 + t] = self.col[(sn, "u1%d" % t)], (1.0 - f1 / f2) * n_fn[t]
                        else:                             # u2 = u1; t2 = -t1 - (1 - phi2/phi1) tau1 n1
                            hcoef[q] = 1.0
                            gcol[q, 0], gcoef[q, 0] = self.col[(sn, "t1%d" % (k - 1))], 1.0
                            gcol[q, 1], gcoef[q, 1] = self.col[(sn, "tau1")], (1.0 - f2 / f1) * n_fn[k - 1]
                        continue
                    if {k1, k2} == {SOLID, PORO}:     # assemble_bem_harela_equation.f90:262-285 / :430-455; assemble_bem_harpor_equation.f90:627-660 / :807-830
                        ps = 1 if k1 == PORO else 2
                        if r.kind == SOLID:               # u = u_p; t = -t_p -/+ tau n
                            hcol[q], hcoef[q] = self.col[(sn, "u%d%d" % (ps, k))], 1.0
                            gcol[q, 0], gcoef[q, 0] = self.col[(sn, "t%d%d" % (ps, k))], 1.0
                            gcol[q, 1], gcoef[q, 1] = self.col[(sn, "tau%d" % ps)], -sgn * n_fn[k]
                        elif k == 0:                      # Un = u . n (impervious contact)
                            hcol[q], hcoef[q] = self.col[(sn, "tau%d" % ps)], 1.0
                            for t in range(3):
                                gcol[q, t], gcoef[q, t] = self.col[(sn, "u%d%d" % (ps, t))], -sgn * n_fn[t]
                        else:
                            hcol[q], hcoef[q] = self.col[(sn, "u%d%d" % (ps, k - 1))], 1.0
                            gcol[q, 0], gcoef[q, 0] = self.col[(sn, "t%d%d" % (ps, k - 1))], -1.0
                        continue
                    if PORO in (k1, k2):
                        imp = self.interface_ctype.get(bnd, 0) == 1
                        po = self.regions[r1 if k1 == PORO else r2].material
                        ps = 1 if k1 == PORO else 2                           # side index of the poroelastic variables
                        fs_ = 3 - ps
                        phi = po.phi
                        if r.kind == FLUID:                                   # assemble_bem_harpot_equation.f90:183-210 (region 1) / :236-262 (region 2)
                            if imp:                                           # p active; Un = u . n
                                hcol[q], hcoef[q] = self.col[(sn, "p%d" % fs_)], 1.0
                                for t in range(3):
                                    gcol[q, t], gcoef[q, t] = self.col[(sn, "u%d%d" % (ps, t))], -sgn * n_fn[t]
                            else:                                             # p = -tau/phi; Un = phi w + (1 - phi) u . n, w seen from the fluid = -w of the poro side
                                hcol[q], hcoef[q] = self.col[(sn, "tau%d" % ps)], -1.0 / phi
                                gcol[q, 0], gcoef[q, 0] = self.col[(sn, "w%d" % ps)], phi
                                for t in range(3):
                                    gcol[q, 1 + t], gcoef[q, 1 + t] = self.col[(sn, "u%d%d" % (ps, t))], -sgn * (1.0 - phi) * n_fn[t]
                        elif k == 0:                                          # fluid phase of the poroelastic side (assemble_bem_harpor_equation.f90:583-601 / :755-773)
                            hcol[q], hcoef[q] = self.col[(sn, "tau%d" % ps)], 1.0
                            if imp:                                           # Un = u . n
                                for t in range(3):
                                    gcol[q, t], gcoef[q, t] = self.col[(sn, "u%d%d" % (ps, t))], -sgn * n_fn[t]
                            else:
                                gcol[q, 0], gcoef[q, 0] = self.col[(sn, "w%d" % ps)], -1.0
                        else:                                                 # skeleton (:603-625 / :776-799)
                            hcol[q], hcoef[q] = self.col[(sn, "u%d%d" % (ps, k - 1))], 1.0
                            if imp:                                           # t_k = -(p + tau) n_k
                                gcol[q, 0], gcoef[q, 0] = self.col[(sn, "p%d" % fs_)], sgn * n_fn[k - 1]
                                gcol[q, 1], gcoef[q, 1] = self.col[(sn, "tau%d" % ps)], sgn * n_fn[k - 1]
                            else:                                             # t_k = (1 - phi)/phi tau n_k
                                gcol[q, 0], gcoef[q, 0] = self.col[(sn, "tau%d" % ps)], -sgn * (1.0 - phi) / phi * n_fn[k - 1]
                        continue
                    if r.kind == SOLID and other == SOLID:
                        hcol[q], hcoef[q] = self.col[(sn, "u1%d" % k)], 1.0
                        gcol[q, 0], gcoef[q, 0] = self.col[(sn, "t1%d" % k)], (-1.0 if first else 1.0)
                    elif r.kind == SOLID:
                        hcol[q], hcoef[q] = self.col[(sn, ("u1%d" if first else "u2%d") % k)], 1.0
                        gcol[q, 0], gcoef[q, 0] = self.col[(sn, "p2" if first else "p1")], (n_fn[k] if first else -n_fn[k])
                    elif other == FLUID:
                        hcol[q], hcoef[q] = self.col[(sn, "p1")], 1.0
                        gcol[q, 0], gcoef[q, 0] = self.col[(sn, "un1")], (-1.0 if first else 1.0)
                    else:
                        hcol[q], hcoef[q] = self.col[(sn, "p1" if first else "p2")], 1.0
                        for t in range(3):
                            gcol[q, t], gcoef[q, t] = self.col[(sn, ("u2%d" if first else "u1%d") % t)], (-n_fn[t] if first else n_fn[t])
        return dict(hcol=hcol, hcoef=hcoef, gcol=gcol, gcoef=gcoef)

    # ---- nodal variables from the solution vector (assign_solution_mechanics_harmonic.f90 with the interface substitutions)
    def nodal_solution(self, x, kr, omega=None):
        """Primary and secondary variables of region kr at the nodes of its boundaries: solid (u (n,3), t (n,3)), fluid (p (n), Un (n));
        nodes the region does not touch are NaN.  Interface values follow the coupling relations listed in the module docstring, with
        n = the mesh normal of the element node averaged over the node's elements (only used for the fluid-solid relations)."""
        r = self.regions[kr]
        nan = np.nan + 0j
        if r.kind == PORO:
            return self._nodal_solution_poro(x, kr)
        if r.kind == SOLID:
            P = np.full((self.n_node, 3), nan); S = np.full((self.n_node, 3), nan)
        else:
            P = np.full(self.n_node, nan); S = np.full(self.n_node, nan)
        nrm = self.node_normals()
        for sb in r.boundaries:
            b = abs(sb)
            r1, r2 = self.boundary_regions[b]
            first = r1 == kr
            for e in self.elems_of_boundary[b]:
                for v in self.mesh.conn[e]:
                    v = int(v)
                    if r2 is None:
                        for k in range(r.ndof):
                            known = self.cvalue[b][k]
                            if r.kind == SOLID:
                                if self.ctype[b][k] == 0:
                                    P[v, k] = known; S[v, k] = x[self.col[(v, "t1%d" % k)]]
                                else:
                                    S[v, k] = known; P[v, k] = x[self.col[(v, "u1%d" % k)]]
                            else:
                                if self.ctype[b][k] == 0:
                                    P[v] = known; S[v] = x[self.col[(v, "un1")]]
                                elif self.ctype[b][k] == 1:
                                    S[v] = known; P[v] = x[self.col[(v, "p1")]]
                                else:
                                    P[v] = x[self.col[(v, "p1")]]; S[v] = -self.impedance_coefficient(kr, b, omega) * P[v]
                        continue
                    k1, k2 = self.regions[r1].kind, self.regions[r2].kind
                    n1 = nrm[v]                                        # outward from region 1
                    if PORO in (k1, k2):                               # this region is the fluid or the solid side of a poroelastic interface
                        ps = 1 if k1 == PORO else 2
                        n_p = n1 if ps == 1 else -n1                   # outward from the poroelastic region
                        tau = x[self.col[(v, "tau%d" % ps)]]
                        u = np.array([x[self.col[(v, "u%d%d" % (ps, k))]] for k in range(3)])
                        if r.kind == FLUID:
                            phi = self.regions[r1 if ps == 1 else r2].material.phi
                            if self.interface_ctype.get(b, 0) == 1:    # impermeable: p active, Un = u . n_f
                                P[v] = x[self.col[(v, "p%d" % (3 - ps))]]; S[v] = -(u @ n_p)
                            else:                                      # permeable: p = -tau/phi, U_f.n = phi U.n + (1 - phi) u.n
                                P[v] = -tau / phi; S[v] = -(phi * x[self.col[(v, "w%d" % ps)]] + (1.0 - phi) * (u @ n_p))
                        else:                                          # bonded solid: u = u_p, t_s = -t_p - tau n_p
                            tp = np.array([x[self.col[(v, "t%d%d" % (ps, k))]] for k in range(3)])
                            P[v] = u; S[v] = -tp - tau * n_p
                        continue
                    if k1 == FLUID and k2 == FLUID:
                        P[v] = x[self.col[(v, "p1")]]; S[v] = x[self.col[(v, "un1")]] * (1.0 if first else -1.0)
                    elif k1 == SOLID and k2 == SOLID:
                        for k in range(3):
                            P[v, k] = x[self.col[(v, "u1%d" % k)]]; S[v, k] = x[self.col[(v, "t1%d" % k)]] * (1.0 if first else -1.0)
                    else:
                        fl_first = k1 == FLUID
                        pcol = self.col[(v, "p1" if fl_first else "p2")]
                        u = np.array([x[self.col[(v, ("u2%d" if fl_first else "u1%d") % k)]] for k in range(3)])
                        n_out = n1 if first else -n1                   # outward from THIS region
                        if r.kind == FLUID:
                            P[v] = x[pcol]; S[v] = u @ n_out
                        else:
                            P[v] = u; S[v] = -x[pcol] * n_out
        return P, S

    def _nodal_solution_poro(self, x, kr):
        """(tau, u1, u2, u3) and (Un, t1, t2, t3) of a poroelastic region at the nodes of its boundaries (NaN elsewhere), with the interface
        relations of assemble_bem_harpor_equation.f90 (fluid: permeable t_k = (1 - phi)/phi tau n_k, impermeable Un = u.n, t_k = -(p + tau) n_k;
        bonded solid: Un = u.n; second side of a permeable poroelastic contact: tau2 = phi2/phi1 tau1, u2 = u1,
        Un2 = -phi1/phi2 Un1 - (1 - phi1/phi2) u1.n1, t2 = -t1 - (1 - phi2/phi1) tau1 n1)."""
        r = self.regions[kr]
        nan = np.nan + 0j
        P = np.full((self.n_node, 4), nan); S = np.full((self.n_node, 4), nan)
        nrm = self.node_normals()
        for sb in r.boundaries:
            b = abs(sb)
            r1, r2 = self.boundary_regions[b]
            first = r1 == kr
            for e in self.elems_of_boundary[b]:
                for v in self.mesh.conn[e]:
                    v = int(v)
                    if r2 is None:
                        for k in range(4):
                            if self.ctype[b][k] == 0:
                                P[v, k] = self.cvalue[b][k]; S[v, k] = x[self.col[(v, _var_name(PORO, k, True, 1))]]
                            else:
                                S[v, k] = self.cvalue[b][k]; P[v, k] = x[self.col[(v, _var_name(PORO, k, False, 1))]]
                        continue
                    other = self.regions[r2 if first else r1]
                    n_out = nrm[v] if first else -nrm[v]               # outward from this region
                    side = 1 if first else 2
                    if other.kind == PORO:
                        f1, f2 = self.regions[r1].material.phi, self.regions[r2].material.phi
                        tau1 = x[self.col[(v, "tau1")]]; w1 = x[self.col[(v, "w1")]]
                        u1 = np.array([x[self.col[(v, "u1%d" % k)]] for k in range(3)]); t1 = np.array([x[self.col[(v, "t1%d" % k)]] for k in range(3)])
                        if first:
                            P[v] = [tau1, *u1]; S[v] = [w1, *t1]
                        else:
                            n1 = nrm[v]
                            P[v] = [f2 / f1 * tau1, *u1]
                            S[v] = [-f1 / f2 * w1 - (1.0 - f1 / f2) * (u1 @ n1), *(-t1 - (1.0 - f2 / f1) * tau1 * n1)]
                        continue
                    tau = x[self.col[(v, "tau%d" % side)]]
                    u = np.array([x[self.col[(v, "u%d%d" % (side, k))]] for k in range(3)])
                    P[v] = [tau, *u]
                    if other.kind == SOLID:
                        S[v] = [u @ n_out, *[x[self.col[(v, "t%d%d" % (side, k))]] for k in range(3)]]
                    elif self.interface_ctype.get(b, 0) == 1:
                        pf = x[self.col[(v, "p%d" % (3 - side))]]
                        S[v] = [u @ n_out, *(-(pf + tau) * n_out)]
                    else:
                        phi = r.material.phi
                        S[v] = [x[self.col[(v, "w%d" % side)]], *((1.0 - phi) / phi * tau * n_out)]
        return P, S

    def node_normals(self):
        """Unit mesh normal at every node (average of element(se)%n_fn over the node's elements, build_data_at_functional_nodes.f90:360-389)."""
        acc = np.zeros((self.n_node, 3))
        for e in range(self.n_elem):
            et = int(self.mesh.etype[e]); c = self.mesh.conn[e]
            for kn, v in enumerate(c):
                acc[int(v)] += sh.unit_normal(et, self.node_x[c], sh.XI_NODES[et][kn])
        nrm = np.linalg.norm(acc, axis=1)
        nrm[nrm == 0] = 1.0
        return acc / nrm[:, None]
