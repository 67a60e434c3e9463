"""Flat (C-ABI ready) description of one elastic BE region, built the way the reference's host does.

Mirrors, for the `be` / ordinary-boundary case used by the harmonic 3D hot path:
  * default BEM formulation   src/assign_default_bem_formulation.f90:74-93 (part-rim nodes -> SBIE-MCA, delta 0.05)
  * collocation points        src/build_data_at_collocation_points.f90:124-160,285-310
  * DOF numbering             src/build_auxiliary_variables_mechanics_harmonic.f90:151-198
  * material                  src/read_regions.f90:717-722  (mu_c = mu(1+2i xi), lambda_c = 2 mu_c nu/(1-2nu))
Everything here is O(N) host bookkeeping; the arrays it produces are exactly what crosses the C ABI
(include/mfb.h) and what the oracle consumes.
"""
import numpy as np
from . import shape as sh

MCA_BOUNDARY_DELTA = 0.05          # assign_default_bem_formulation.f90:40
NODAL_XI_MARK = -9.0               # colloc_xi marker for nodal SBIE collocation


class Material:
    def __init__(self, rho=1.0, mu=1.0, nu=0.25, xi=0.02):
        self.rho, self.mu_r, self.nu_r, self.xi = float(rho), float(mu), float(nu), float(xi)
        self.mu = self.mu_r * (1.0 + 1j * 2.0 * self.xi)
        self.nu = complex(self.nu_r)
        self.lam = 2.0 * self.mu * self.nu_r / (1.0 - 2.0 * self.nu_r)
        self.c1 = np.sqrt((self.lam + 2.0 * self.mu) / self.rho)
        self.c2 = np.sqrt(self.mu / self.rho)


MCA_ORDER_DELTA = {1: 0.42264973, 2: 0.22540333}        # default_mca_linear_delta / _quadratic_delta (read_bem_formulation_selected_nodes.f90:42-43)


def mca_deltas(in_boundary, node_group, formulation):
    """Per node: the MCA displacement of its collocation points, 0 = nodal collocation, < 0 = the default of the element's order.
    Default formulation (assign_default_bem_formulation.f90:74-93): the nodes of the rim of a boundary are SBIE-MCA with delta 0.05.
    formulation {boundary (part) id: (kind, delta)} = the records of [bem formulation over boundaries] (read_bem_formulation_selected_nodes.f90:74-160):
    `sbie` every node nodal; `sbie_boundary_mca <delta>` rim nodes MCA (delta <= 0: 0.05); `sbie_mca <delta>` every node MCA (delta <= 0: by element order).
    node_group[v] = the boundary (part) id of node v."""
    d = np.where(in_boundary, MCA_BOUNDARY_DELTA, 0.0)
    for g, (kind, delta) in (formulation or {}).items():
        sel = np.asarray(node_group) == g
        if kind == "sbie":
            if np.any(in_boundary & sel):
                # nodal collocation on the rim of an open boundary: the free term and the singular integrals then run over the elements of the coincident
                # ("common") nodes of the neighbouring boundaries (build_lse_mechanics_bem_harela.f90:452-492), which this library does not collect
                raise ValueError("boundary %s: `sbie` (nodal collocation everywhere) on a boundary with an open rim is not covered; use sbie_boundary_mca" % (g,))
            d[sel] = 0.0
        elif kind == "sbie_boundary_mca":
            d[sel] = np.where(in_boundary[sel], delta if delta > 0 else MCA_BOUNDARY_DELTA, 0.0)
        elif kind == "sbie_mca":
            d[sel] = delta if delta > 0 else -1.0
        else:
            raise ValueError("BEM formulation %r is not covered (sbie, sbie_boundary_mca, sbie_mca)" % (kind,))
    return d


def mca_delta_of(etype, d):
    return d if d > 0 else MCA_ORDER_DELTA[1 if etype in (sh.TRI3, sh.QUAD4) else 2]


def symmetry_planes(symmetry):
    """[(axis, kind), ...] -> (symplane_eid int32[n], symplane_t float64[n,3]) in the reference's internal order (x, y, z):
    symmetry: t = -1 on the normal axis, +1 on the others; antisymmetry: the opposite signs (read_symmetry_planes.f90:160-228)."""
    planes = {}
    for axis, kind in (symmetry or ()):
        ax = {"x": 1, "y": 2, "z": 3, 1: 1, 2: 2, 3: 3}[axis]
        if ax in planes:
            raise ValueError("symmetry plane %r given twice" % (axis,))
        if isinstance(kind, str):
            sgn = {"symmetry": 1.0, "antisymmetry": -1.0}[kind]
            t = np.full(3, sgn); t[ax - 1] = -sgn
        else:      # the explicit form of the case file: the three translation multipliers themselves (optionally preceded by the scalar one)
            t = np.array(kind[-3:], dtype=np.float64)
            if t.shape != (3,) or not np.all(np.abs(t) == 1.0):
                raise ValueError("symmetry multipliers must be three values +1 or -1")
        planes[ax] = t
    eid = np.array(sorted(planes), dtype=np.int32)
    t = np.ascontiguousarray([planes[a] for a in sorted(planes)], dtype=np.float64).reshape(len(eid), 3)
    return eid, t


def symmetry_scalars(symmetry, eid, t):
    """symplane_s of every plane: the multiplier of scalar variables (fluid pressure, fluid-phase stress): +1 symmetry, -1 antisymmetry
    (read_symmetry_planes.f90:160-228); with the explicit multipliers of the case file, the first of the four values."""
    given = {}
    for axis, kind in (symmetry or ()):
        ax = {"x": 1, "y": 2, "z": 3, 1: 1, 2: 2, 3: 3}[axis]
        if not isinstance(kind, str) and len(kind) == 4:
            given[ax] = float(kind[0])
    return np.array([given.get(int(a), t[i, int(a) % 3]) for i, a in enumerate(eid)], dtype=np.float64)


class Model:
    """bcs: {part_id: ([ctype_x, ctype_y, ctype_z], [value_x, value_y, value_z])}, ctype 0 = u known, 1 = t known.
    ndof = equations / unknowns per node: 3 for an elastic solid region, 1 for an inviscid fluid region (FluidModel)."""

    def __init__(self, mesh, bcs, reversed_parts=(), qsi_relative_error=1e-6, qsi_ns_max=16,
                 precalset_gln=(2, 3, 4, 5, 6, 7, 8, 9), geometric_tolerance=1e-6, ndof=3, part_order=None,
                 symmetry=None, nodal_on_symplanes=False, local_axes_reference=None, collapse_nodal_pos=True, formulation=None):
        """symmetry: the [symmetry planes] section (src/read_symmetry_planes.f90:76-283) as a list of (axis, kind), axis 'x' | 'y' | 'z'
        (plane_n1 / plane_n2 / plane_n3: the plane through the origin normal to that axis), kind 'symmetry' | 'antisymmetry'.
        nodal_on_symplanes: open edges that lie in a symmetry plane do not make their nodes boundary-of-the-boundary nodes, so those nodes
        keep the nodal SBIE (what a user of the reference selects in [bem formulation over nodes]; by default the reference gives them the
        SBIE with MCA like any other rim node, assign_default_bem_formulation.f90:85-92)."""
        self.mesh = mesh
        self.symplane_eid, self.symplane_t = symmetry_planes(symmetry)
        self.symplane_s = symmetry_scalars(symmetry, self.symplane_eid, self.symplane_t)
        if collapse_nodal_pos:   # the nodes of a symmetry plane are put exactly in it (fbem_transformation_collapse_nodal_positions, lib/fbem/src/geometry.f90:5602-5610;
            for ax in self.symplane_eid:   # the setting's default is T): an image then touches its root element in the SAME point, which the singular test needs
                on = np.abs(mesh.nodes[:, ax - 1]) <= float(geometric_tolerance)
                mesh.nodes[on, ax - 1] = 0.0
        for ax in self.symplane_eid:   # fbem_check_nodes_symplanes_configuration (lib/fbem/src/data_structures.f90:1197-1230): the mesh stays on one side
            xa = mesh.nodes[:, ax - 1]
            off = xa[np.abs(xa) > float(geometric_tolerance)]
            if len(off) and off.min() < 0.0 < off.max():
                raise ValueError("the mesh crosses the symmetry plane normal to axis %d" % ax)
        self.ndof = nd = int(ndof)
        nn = len(mesh.nodes)
        ne = mesh.n_elem
        self.n_node, self.n_elem = nn, ne
        self.node_x = mesh.nodes
        self.etype = mesh.etype.astype(np.int32)
        self.elem_ptr = np.zeros(ne + 1, dtype=np.int32)
        self.elem_ptr[1:] = np.cumsum([len(c) for c in mesh.conn])
        self.elem_node = np.concatenate(mesh.conn).astype(np.int32)
        self.elem_reversed = np.array([1 if int(p) in reversed_parts else 0 for p in mesh.part], dtype=np.uint8)
        self.qsi_relative_error, self.qsi_ns_max = float(qsi_relative_error), int(qsi_ns_max)
        self.precalset_gln = np.array(precalset_gln, dtype=np.int32)
        self.geometric_tolerance = float(geometric_tolerance)

        # --- part rims: nodes of edges that belong to a single element of the part (data_structures.f90:1617-1636)
        node_part = -np.ones(nn, dtype=np.int64)
        edge_count = {}
        for e in range(ne):
            c, p = mesh.conn[e], int(mesh.part[e])
            for v in c:
                if node_part[v] not in (-1, p):
                    raise ValueError("node %d is shared by two boundaries; each boundary must own its nodes" % v)
                node_part[v] = p
            for ed in sh.edges_of(int(mesh.etype[e])):
                key = (p,) + tuple(sorted((int(c[ed[0]]), int(c[ed[1]]))))
                edge_count.setdefault(key, []).append([int(c[k]) for k in ed])
        in_boundary = np.zeros(nn, dtype=bool)
        for key, lst in edge_count.items():
            if len(lst) == 1:
                if nodal_on_symplanes and any(np.all(np.abs(mesh.nodes[lst[0], ax - 1]) <= self.geometric_tolerance) for ax in self.symplane_eid):
                    continue
                in_boundary[lst[0]] = True
        self.in_boundary = in_boundary
        self.node_part = node_part
        self.mca_delta = mca_deltas(in_boundary, node_part, formulation)      # formulation {part id: (kind, delta)}: [bem formulation over boundaries]

        # --- boundary conditions per node
        self.ctype = np.zeros((nn, nd), dtype=np.int32)
        self.cvalue = np.zeros((nn, nd), dtype=np.complex128)
        for v in range(nn):
            ct, cv = bcs[int(node_part[v])]
            for k in range(nd):
                if int(ct[k]) == 4 and nd == 3:
                    # prescribed infinitesimal rotation field (transfer_conditions_bem_boundaries_mechanics_harmonic.f90:89-99): a ctype-0 condition with
                    # u_k = theta (axis x (x - center))_k; the value is (center, axis, theta), axis normalised by the reader
                    center, axis, theta = cv[k]
                    axis = np.asarray(axis, dtype=np.float64); axis = axis / np.linalg.norm(axis)
                    self.ctype[v, k] = 0
                    self.cvalue[v, k] = complex(theta) * np.cross(axis, mesh.nodes[v] - np.asarray(center, dtype=np.float64))[k]
                else:
                    self.ctype[v, k] = int(ct[k]); self.cvalue[v, k] = cv[k]

        # --- DOF numbering: region -> boundary (part id order) -> element -> node, first visit
        # part_order: the boundaries in the order of the region's list in the case file (default: ascending part id)
        parts = list(part_order) if part_order is not None else sorted(set(int(p) for p in mesh.part))
        if sorted(parts) != sorted(set(int(p) for p in mesh.part)):
            raise ValueError("part_order must list every part of the mesh once")
        order = [e for p in parts for e in range(ne) if int(mesh.part[e]) == p]
        self.elem_order = np.array(order, dtype=np.int32)
        self.row = -np.ones((nn, nd), dtype=np.int32)
        self.row_bc = -np.ones((nn, nd), dtype=np.int32)      # node%row(k,0): the condition row of a local-axes dof (ctype 2 / 3)
        self.col_u = -np.ones((nn, nd), dtype=np.int32)
        self.col_t = -np.ones((nn, nd), dtype=np.int32)
        row = col = 0
        seen = np.zeros(nn, dtype=bool)
        for e in order:
            for v in mesh.conn[e]:
                if seen[v]:
                    continue
                seen[v] = True
                for k in range(nd):
                    self.row[v, k] = row; row += 1
                    if self.ctype[v, k] == 0:
                        self.col_t[v, k] = col
                    elif self.ctype[v, k] == 1 or (self.ctype[v, k] == 10 and nd == 3):
                        self.col_u[v, k] = col          # 10: normal pressure known, t_k = p n_fn(k); the unknown is u_k
                    elif self.ctype[v, k] in (2, 3) and nd == 3:
                        # local axes (build_auxiliary_variables_mechanics_harmonic.f90:188-197): u_k and t_k are both unknowns, and the node gets a condition row
                        self.col_u[v, k] = col; self.col_t[v, k] = col + 1; col += 1
                        self.row_bc[v, k] = row; row += 1
                    else:
                        raise ValueError("only ctype 0 / 1 (and 2 / 3 / 10 on elastic regions) are supported on this path")
                    col += 1
        assert row == col
        self.n_dof = row

        # --- nodal unit normals node()%n_fn (src/build_data_at_functional_nodes.f90:355-390): normalised sum of the normals of the elements around the
        # node (as meshed: the orientation of the boundary in the region enters later, through the sign of the ctype-10 term), mirror images included
        self.n_fn = np.zeros((nn, 3))
        for e in range(ne):
            et = int(mesh.etype[e]); c = mesh.conn[e]
            for kn, v in enumerate(c):
                self.n_fn[v] += sh.unit_normal(et, self.node_x[c], sh.XI_NODES[et][kn])
        for ax in self.symplane_eid:
            on = np.abs(self.node_x[:, ax - 1]) <= self.geometric_tolerance
            self.n_fn[on, ax - 1] = 0.0       # n + M n, applied once per plane the node lies in: the component along the plane's axis cancels, the others double
        norm = np.linalg.norm(self.n_fn, axis=1)
        self.n_fn[norm > 0] /= norm[norm > 0, None]
        # unit tangents of the local axes (build_data_at_functional_nodes.f90:406-446): t2 = n x reference vector (the user's, else e1, e2, e3 in turn), t1 = t2 x n
        self.t1_fn = np.zeros((nn, 3)); self.t2_fn = np.zeros((nn, 3))
        refs = ([np.asarray(local_axes_reference, dtype=np.float64)] if local_axes_reference is not None else []) + [np.eye(3)[i] for i in range(3)]
        for v in range(nn):
            n = self.n_fn[v]
            if not n.any():
                continue
            for ref in refs:
                t2 = np.cross(n, ref)
                if np.linalg.norm(t2) > self.geometric_tolerance:
                    break
            t2 = t2 / np.linalg.norm(t2)
            t1 = np.cross(t2, n)
            self.t1_fn[v], self.t2_fn[v] = t1 / np.linalg.norm(t1), t2

        # --- collocation points (loop order of build_lse_mechanics_bem_harela.f90:1118-1136)
        cx, cnode, celem, ckn, cxi = [], [], [], [], []
        collocated = np.zeros(nn, dtype=bool)
        for e in order:
            et = int(mesh.etype[e]); c = mesh.conn[e]
            xn = self.node_x[c]
            for kn, v in enumerate(c):
                if self.mca_delta[v] != 0.0:
                    xi = sh.move_xi_from_edge(et, sh.XI_NODES[et][kn], mca_delta_of(et, self.mca_delta[v]))
                    cx.append(sh.position(et, xn, xi)); cxi.append(xi)
                elif not collocated[v]:
                    collocated[v] = True
                    cx.append(sh.position(et, xn, sh.XI_NODES[et][kn])); cxi.append([NODAL_XI_MARK, NODAL_XI_MARK])
                else:
                    continue
                cnode.append(v); celem.append(e); ckn.append(kn)
        self.colloc_x = np.ascontiguousarray(cx, dtype=np.float64)
        self.colloc_node = np.array(cnode, dtype=np.int32)
        self.colloc_elem = np.array(celem, dtype=np.int32)
        self.colloc_kn = np.array(ckn, dtype=np.int32)
        self.colloc_xi = np.ascontiguousarray(cxi, dtype=np.float64)
        self.n_colloc = len(cnode)

    def condition_rows(self):
        """The rows the host writes for local-axes conditions (src/build_lse_mechanics_harmonic.f90:204-258): dof k of such a node refers to the axis
        l_1 = n, l_2 = t_1, l_3 = t_2; ctype 2: u . l_k = U, ctype 3: t . l_k = T.  -> rows, cols (-1: right-hand side), values."""
        rows, cols, vals = [], [], []
        if self.ndof != 3:
            return np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.complex128)
        axes = (self.n_fn, self.t1_fn, self.t2_fn)
        for v in range(self.n_node):
            for k in range(3):
                ct = int(self.ctype[v, k])
                if ct not in (2, 3):
                    continue
                r = int(self.row_bc[v, k]); target = self.col_u if ct == 2 else self.col_t
                for kci in range(3):
                    rows.append(r); cols.append(int(target[v, kci])); vals.append(complex(axes[k][v, kci]))
                rows.append(r); cols.append(-1); vals.append(complex(self.cvalue[v, k]))
        return np.array(rows, dtype=np.int32), np.array(cols, dtype=np.int32), np.array(vals, dtype=np.complex128)

    def add_condition_rows(self, A, b):
        """A, b of the BEM assembly plus the condition rows (what the reference's host does after build_lse_mechanics_bem_harela)."""
        rows, cols, vals = self.condition_rows()
        for r, c, v in zip(rows, cols, vals):
            if c < 0:
                b[r] += v
            else:
                A[r, c] += v
        return A, b

    # --- reference's assign_solution_mechanics_harmonic.f90:192-205: nodal u,t from the solution vector
    def nodal_solution(self, x):
        u = np.zeros((self.n_node, self.ndof), dtype=np.complex128)
        t = np.zeros((self.n_node, self.ndof), dtype=np.complex128)
        for k in range(self.ndof):
            known_u = self.ctype[:, k] == 0
            u[known_u, k] = self.cvalue[known_u, k]
            t[known_u, k] = x[self.col_t[known_u, k]]
            t[~known_u, k] = self.cvalue[~known_u, k]
            u[~known_u, k] = x[self.col_u[~known_u, k]]
            if self.ndof == 3:
                p10 = self.ctype[:, k] == 10          # normal pressure known: t_k = p n_fn(k)
                t[p10, k] = self.cvalue[p10, k] * self.n_fn[p10, k]
                loc = (self.ctype[:, k] == 2) | (self.ctype[:, k] == 3)      # local axes: both are unknowns of the system
                u[loc, k] = x[self.col_u[loc, k]]; t[loc, k] = x[self.col_t[loc, k]]
        return u, t


class Fluid:
    """Inviscid fluid: `fluid c <c> rho <rho>` of the [materials] section (src/read_regions.f90:544-546, :599-604): region%property_r(1) =
    rho, region%property_c(4) = c.  xi > 0 gives the hysteretic variant c (1 + 2 i xi)^(1/2) (the stiffness K = rho c^2 is damped)."""

    def __init__(self, rho=1.25, c=343.0, xi=0.0):
        self.rho, self.c_r, self.xi = float(rho), float(c), float(xi)
        self.c = complex(self.c_r) * np.sqrt(1.0 + 2j * self.xi)


class Poro:
    """Biot poroelastic medium (`biot_poroelastic_medium` material, src/read_regions.f90:737-860): fluid and solid densities rhof, rhos, drained
    Lame constants lambda, mu of the skeleton, hysteretic damping xi (applied to lambda, mu, R, Q), porosity phi, added density rhoa, Biot's
    coupling parameters R, Q and the dissipation constant b.  rho1 = (1 - phi) rhos, rho2 = phi rhof (property_r(13:14))."""

    def __init__(self, rhof=1000.0, rhos=2600.0, lam=1.0e8, mu=1.0e8, xi=0.0, phi=0.3, rhoa=0.0, R=1.0e8, Q=1.0e8, b=0.0):
        self.rhof, self.rhos, self.phi, self.rhoa, self.b, self.xi = float(rhof), float(rhos), float(phi), float(rhoa), float(b), float(xi)
        d = 1.0 + 2j * self.xi
        self.lam, self.mu, self.R, self.Q = lam * d, mu * d, R * d, Q * d
        self.rho1, self.rho2 = (1.0 - self.phi) * self.rhos, self.phi * self.rhof
        self.nu = 0.5 * self.lam / (self.lam + self.mu)

    def props(self):
        """The argument list of fbem_bem_harpor3d_calculate_parameters (lambda, mu, rho1, rho2, rhoa, R, Q, b), flat."""
        return np.array([self.lam.real, self.lam.imag, self.mu.real, self.mu.imag, self.rho1, self.rho2, self.rhoa,
                         self.R.real, self.R.imag, self.Q.real, self.Q.imag, self.b], dtype=np.float64)


class PoroModel(Model):
    """One Biot poroelastic BE region: four equations and four unknowns per node, component 0 = fluid phase (tau known: ctype 0, Un known:
    ctype 1), components 1..3 = solid skeleton (u_k known: 0, t_k known: 1) -- the open-pore conditions of an ordinary boundary
    (assemble_bem_harpor_equation.f90:78-110, :140-170).  bcs: {part_id: ([ct_tau, ct_1, ct_2, ct_3], [values])}.  col_u holds the columns
    of (tau, u_k), col_t those of (Un, t_k).  Device path: mfb_harpor3d_* (csrc/poro.cu, tests/test_gpu_poroelastic.py)."""

    def __init__(self, mesh, bcs, **kw):
        Model.__init__(self, mesh, bcs, ndof=4, **kw)


class FluidModel(Model):
    """One inviscid-fluid BE region (scalar wave propagation): one equation and one unknown per node.
    bcs: {part_id: (ctype, value)}: ctype 0 = p known (Un unknown), 1 = Un known (p unknown), `conditions over be boundaries` of an
    inviscid fluid boundary (src/read_conditions_bem_boundaries_mechanics.f90; scatter assemble_bem_harpot_equation.f90:78-96).
    DOF numbering: build_auxiliary_variables_mechanics_harmonic.f90 (fluid boundary: row(1,1); col(1,1) = p or col(2,1) = Un).
    col_u holds the column of p, col_t the column of Un."""

    def __init__(self, mesh, bcs, **kw):
        bcs3 = {p: ([int(ct)], [complex(cv)]) for p, (ct, cv) in bcs.items()}
        Model.__init__(self, mesh, bcs3, ndof=1, **kw)

    def nodal_solution(self, x):
        p, un = Model.nodal_solution(self, x)
        return p[:, 0], un[:, 0]


# the reference's acoustic room tutorial (docs/examples/ME-TH-AC-001/case_files/room.dat: walls 1-4 rigid (Un = 0), p = 1 on one x-face,
# p = 0 on the other) on cube_mesh() part ids (1 x=0, 2 x=L, 3 y=0, 4 y=L, 5 z=0, 6 z=L)
def room_bcs(P=1.0):
    return {1: (0, 0.0), 2: (0, P), 3: (1, 0.0), 4: (1, 0.0), 5: (1, 0.0), 6: (1, 0.0)}


def room_analytic(x1, omega, fluid, L=1.0, P=1.0):
    """p(x) = P sin(kx)/sin(kL), U_x(x) = P k cos(kx)/(rho omega^2 sin(kL)) (docs/examples/ME-TH-AC-001/doc_src/ME-TH-AC-001.tex:47-50)."""
    k = omega / fluid.c
    return P * np.sin(k * x1) / np.sin(k * L), P * k * np.cos(k * x1) / (fluid.rho * omega ** 2 * np.sin(k * L))


# the reference's harmonic cube tutorial (docs/examples/ME-TH-EL-001/case_files/t3.dat:38-55), by physical name
# of t3.msh: 1 front(z=1) 2 right(x=1) 3 back(z=0) 4 left(x=0) 5 top(y=1) 6 bottom(y=0)
ME_TH_EL_001_BCS = {
    1: ([1, 1, 0], [0, 0, 0]), 2: ([1, 1, 1], [1, 0, 0]), 3: ([1, 1, 0], [0, 0, 0]),
    4: ([0, 0, 0], [0, 0, 0]), 5: ([1, 0, 1], [0, 0, 0]), 6: ([1, 0, 1], [0, 0, 0]),
}


def cube_bcs():
    """Same physical problem on cube_mesh() part ids (1 x=0, 2 x=L, 3 y=0, 4 y=L, 5 z=0, 6 z=L):
    x=0 clamped, x=L unit normal traction, lateral faces: zero normal displacement, zero shear (1D P-wave column)."""
    return {1: ([0, 0, 0], [0, 0, 0]), 2: ([1, 1, 1], [1, 0, 0]),
            3: ([1, 0, 1], [0, 0, 0]), 4: ([1, 0, 1], [0, 0, 0]),
            5: ([1, 1, 0], [0, 0, 0]), 6: ([1, 1, 0], [0, 0, 0])}


def column_analytic_u(x1, omega, mat, L=1.0, P=1.0):
    """u1(x1) of the clamped-free P-wave column (docs/examples/ME-TH-EL-001/doc_src/ME-TH-EL-001.tex:32-56)."""
    k = omega / mat.c1
    return -P * (np.exp(-1j * k * x1) - np.exp(1j * k * x1)) / ((mat.lam + 2 * mat.mu) * 1j * k * (np.exp(-1j * k * L) + np.exp(1j * k * L)))


class InternalPointsModel:
    """The flat arrays of a problem whose collocation points are points INSIDE the region of `model` (reference: [internal points]
    section, src/calculate_internal_points_mechanics_bem_harela.f90): same elements, boundary conditions and columns as the
    boundary problem; every interior point gets a dummy node (used by no element) that owns its three rows, placed after the
    boundary rows; colloc_elem = -1 tells the library that there is no free term."""

    def __init__(self, model, points, stress=False):
        """stress=True: the hypersingular problem -- every point appears three times with the unit normals e_1, e_2, e_3
        (colloc_n), so that its nine rows are the traction vectors on the three coordinate planes = the stress tensor."""
        m = model
        nd = int(getattr(m, "ndof", 3))          # 1: inviscid fluid region -- the interior rows give the pressure (calculate_internal_points_mechanics_bem_harpot.f90)
        if stress and nd != 3:
            raise ValueError("the hypersingular interior-point problem is built for elastic regions only")
        self.ndof = nd
        pts0 = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        self.n_points = len(pts0)
        self.colloc_n = None
        if stress:
            pts = np.repeat(pts0, 3, axis=0)
            self.colloc_n = np.ascontiguousarray(np.tile(np.eye(3), (len(pts0), 1)))
        else:
            pts = pts0
        nip = len(pts)
        self.base, self.points = m, pts0
        self.mesh = m.mesh
        self.n_node, self.n_elem = m.n_node + nip, m.n_elem
        self.node_x = np.ascontiguousarray(np.vstack([m.node_x, pts]))
        self.etype, self.elem_ptr, self.elem_node, self.elem_reversed = m.etype, m.elem_ptr, m.elem_node, m.elem_reversed
        self.qsi_relative_error, self.qsi_ns_max = m.qsi_relative_error, m.qsi_ns_max
        self.precalset_gln, self.geometric_tolerance = m.precalset_gln, m.geometric_tolerance
        self.symplane_eid, self.symplane_t = getattr(m, "symplane_eid", np.zeros(0, np.int32)), getattr(m, "symplane_t", np.zeros((0, 3)))
        self.symplane_s = getattr(m, "symplane_s", np.zeros(0))
        if hasattr(m, "n_fn"):
            self.n_fn = np.ascontiguousarray(np.vstack([m.n_fn, np.zeros((nip, 3))]))
        self.n_dof = m.n_dof + nd * nip
        dummy_rows = (m.n_dof + np.arange(nd * nip, dtype=np.int32)).reshape(nip, nd)
        none = -np.ones((nip, nd), dtype=np.int32)
        self.row = np.ascontiguousarray(np.vstack([m.row, dummy_rows]), dtype=np.int32)
        self.col_u = np.ascontiguousarray(np.vstack([m.col_u, none]), dtype=np.int32)
        self.col_t = np.ascontiguousarray(np.vstack([m.col_t, none]), dtype=np.int32)
        self.ctype = np.ascontiguousarray(np.vstack([m.ctype, np.ones((nip, nd), dtype=np.int32)]), dtype=np.int32)
        self.cvalue = np.ascontiguousarray(np.vstack([m.cvalue, np.zeros((nip, nd), dtype=np.complex128)]))
        self.colloc_x = pts
        self.colloc_node = (m.n_node + np.arange(nip)).astype(np.int32)
        self.colloc_elem = -np.ones(nip, dtype=np.int32)
        self.colloc_kn = np.zeros(nip, dtype=np.int32)
        self.colloc_xi = np.full((nip, 2), NODAL_XI_MARK, dtype=np.float64)
        self.n_colloc = nip
