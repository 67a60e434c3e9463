"""CPU oracle for several BE regions coupled through be-be interfaces (TEST INFRASTRUCTURE ONLY, like oracle.py).

The (collocation point, element) integrals are the pinned single-region ones of harela3d_oracle.cpp (fbem_bem_harela3d_sbie_auto /
fbem_bem_harpot3d_sbie_auto through orc_pair / orc_pair_pot, which also apply the orientation of the element as seen from the region).
This module restates the DRIVER around them for coupled regions:
  * region loop, element loop, collocation loop   src/build_lse_mechanics_harmonic.f90 (do kr), src/build_lse_mechanics_bem_harela.f90:227-238,
                                                  :972-1324, src/build_lse_mechanics_bem_harpot.f90:211-217, :722-1133
  * free terms                                    build_lse_mechanics_bem_harela.f90:273-747, build_lse_mechanics_bem_harpot.f90:243-660
  * scatter of an ordinary boundary               assemble_bem_harela_equation.f90:78-113, assemble_bem_harpot_equation.f90:78-96
  * scatter of a be-be boundary                   assemble_bem_harela_equation.f90:161-331 (region 1) / :332-460 (region 2, reversed);
                                                  assemble_bem_harpot_equation.f90:152-183 (region 1) / :213-236 (region 2)
Parity status: unpinned by reference output (no Fortran compiler here); pinned by analytic two-layer solutions and by the equivalence of
a split homogeneous body with the one-region model (tests/test_oracle_multiregion.py).
"""
import ctypes as C
import numpy as np

from . import oracle as orc
from multifebe_b200.host import shape as sh
from multifebe_b200.host.multiregion import SOLID, FLUID, PORO, _var_name


def _node_normal_tangents(et, xn, node):
    n = np.zeros(3); tp = np.zeros(3); tm = np.zeros(3)
    xn = np.ascontiguousarray(xn, dtype=np.float64)
    orc.lib().orc_node_normal_tangents(C.c_int(et), orc._p(xn), C.c_int(node), orc._p(n), orc._p(tp), orc._p(tm))
    return n, tp, tm


class MultiRegionOracle:
    def __init__(self, mrm):
        self.m = mrm
        self.h = [orc.Oracle(v) if v.kind == SOLID else (orc.PotOracle(v) if v.kind == FLUID else orc.PorOracle(v)) for v in mrm.views]
        # node -> (local element, local node) incidences per region view
        self.inc = []
        for v in mrm.views:
            d = {}
            for le in range(v.n_elem):
                for kn in range(v.elem_ptr[le], v.elem_ptr[le + 1]):
                    d.setdefault(int(v.elem_node[kn]), []).append((le, kn - int(v.elem_ptr[le])))
            self.inc.append(d)

    # ---- free term of one collocation point: added to h of its own element (h[kn] += c, or h[:] += phi/2 for an MCA point)
    def _free_term(self, kr, c, omega):
        v = self.m.views[kr]
        le, kn, sn = int(v.colloc_elem[c]), int(v.colloc_kn[c]), int(v.colloc_node[c])
        et = int(v.etype[le]); nodes = v.elem_node[v.elem_ptr[le]:v.elem_ptr[le + 1]]
        nn = len(nodes)
        solid = v.kind == SOLID
        poro = v.kind == PORO
        if poro:                                                         # c(0,0) = J c_pot, c(1:3,1:3) = Mantic's matrix (build_lse_mechanics_bem_harpor.f90:583-615)
            mat = v.material
            J = 1.0 / ((mat.rho2 + mat.rhoa - 1j * mat.b / omega) * omega ** 2)
        hp = np.zeros((nn, 3, 3), dtype=np.complex128) if solid else (np.zeros((nn, 4, 4), dtype=np.complex128) if poro else np.zeros(nn, dtype=np.complex128))
        if v.colloc_xi[c, 0] != -9.0:                                    # MCA point: 1/2 phi_j(xi_i)
            phi = sh.phi(et, v.colloc_xi[c])
            for j in range(nn):
                if solid:
                    hp[j] += 0.5 * phi[j] * np.eye(3)
                elif poro:
                    hp[j] += 0.5 * phi[j] * np.diag([J, 1.0, 1.0, 1.0])
                else:
                    hp[j] += 0.5 * phi[j]
            return le, hp
        on_edge = not (et == sh.QUAD9 and kn == 8)                       # only the centre node of quad9 lies inside its element
        if not on_edge:
            cfree = 0.5 * np.eye(3) if solid else (0.5 * np.diag([J, 1.0, 1.0, 1.0]) if poro else 0.5)
        else:
            rev = bool(v.elem_reversed[le])
            ns, ts = [], []
            for (le2, kn2) in self.inc[kr][sn]:
                nodes2 = v.elem_node[v.elem_ptr[le2]:v.elem_ptr[le2 + 1]]
                n, tbp, tbm = _node_normal_tangents(int(v.etype[le2]), v.node_x[nodes2], kn2)
                ns.append(-n if rev else n); ts.append(tbm if rev else tbp)
            if solid:
                cfree, err = orc.freeterm(np.array(ns), np.array(ts), v.material.nu, self.m.geometric_tolerance)
            elif poro:
                cela, err = orc.freeterm(np.array(ns), np.array(ts), mat.nu, self.m.geometric_tolerance)
                cp = C.c_double(0.0)
                n_ = np.ascontiguousarray(ns, dtype=np.float64); t_ = np.ascontiguousarray(ts, dtype=np.float64)
                err = err or orc.lib().orc_freeterm_pot(C.c_int(len(ns)), orc._p(n_), orc._p(t_), C.c_double(self.m.geometric_tolerance), C.byref(cp))
                cfree = np.zeros((4, 4), dtype=np.complex128); cfree[0, 0] = J * cp.value; cfree[1:, 1:] = cela
            else:
                cp = C.c_double(0.0)
                n_ = np.ascontiguousarray(ns, dtype=np.float64); t_ = np.ascontiguousarray(ts, dtype=np.float64)
                err = orc.lib().orc_freeterm_pot(C.c_int(len(ns)), orc._p(n_), orc._p(t_), C.c_double(self.m.geometric_tolerance), C.byref(cp))
                cfree = cp.value
            if err:
                raise RuntimeError("oracle: invalid normals/tangents configuration in free-term")
        hp[kn] += cfree
        return le, hp

    # ---- scatter of one (collocation node, element) block of region kr
    def _scatter(self, kr, le, sn_col, eq, h, g, A, b):
        m = self.m
        v = m.views[kr]
        bnd = int(v.elem_boundary[le])
        r1, r2 = m.boundary_regions[bnd]
        nodes = v.elem_node[v.elem_ptr[le]:v.elem_ptr[le + 1]]
        et = int(v.etype[le])
        rows = m.row[(sn_col, eq)]
        first = r1 == kr
        solid = v.kind == SOLID
        for kn, sn in enumerate(nodes):
            sn = int(sn)
            if v.kind == PORO and r2 is None:                            # assemble_bem_harpor_equation.f90:78-110, :140-170 (open-pore conditions 0 / 1)
                ct, cv = m.ctype[bnd], m.cvalue[bnd]
                for il in range(4):
                    for ik in range(4):
                        if ct[ik] == 0:
                            A[rows[il], m.col[(sn, _var_name(PORO, ik, True, 1))]] -= g[kn, il, ik]; b[rows[il]] -= h[kn, il, ik] * cv[ik]
                        else:
                            A[rows[il], m.col[(sn, _var_name(PORO, ik, False, 1))]] += h[kn, il, ik]; b[rows[il]] += g[kn, il, ik] * cv[ik]
                continue
            if r2 is not None and m.regions[r1].kind == PORO and m.regions[r2].kind == PORO:
                # POROELASTIC MEDIA (1) - POROELASTIC MEDIA (2), perfectly permeable: assemble_bem_harpor_equation.f90:696-722 (region 1: every
                # variable of region 1 unknown) and :860-975 (region 2: tau2 = phi2/phi1 tau1, Un2 = -phi1/phi2 Un1 - (1 - phi1/phi2) u1.n1, u2 = u1,
                # t2 = -t1 - (1 - phi2/phi1) tau1 n1)
                n_fn = sh.unit_normal(et, v.node_x[nodes], sh.XI_NODES[et][kn])
                f1, f2 = m.regions[r1].material.phi, m.regions[r2].material.phi
                for il in range(4):
                    row = rows[il]
                    if first:
                        A[row, m.col[(sn, "tau1")]] += h[kn, il, 0]; A[row, m.col[(sn, "w1")]] -= g[kn, il, 0]
                        for ik in range(3):
                            A[row, m.col[(sn, "u1%d" % ik)]] += h[kn, il, ik + 1]; A[row, m.col[(sn, "t1%d" % ik)]] -= g[kn, il, ik + 1]
                    else:
                        A[row, m.col[(sn, "tau1")]] += h[kn, il, 0] * (f2 / f1)
                        A[row, m.col[(sn, "w1")]] += g[kn, il, 0] * (f1 / f2)
                        for ik in range(3):
                            A[row, m.col[(sn, "u1%d" % ik)]] += g[kn, il, 0] * (1.0 - f1 / f2) * n_fn[ik]
                            A[row, m.col[(sn, "u1%d" % ik)]] += h[kn, il, ik + 1]
                            A[row, m.col[(sn, "t1%d" % ik)]] += g[kn, il, ik + 1]
                            A[row, m.col[(sn, "tau1")]] += g[kn, il, ik + 1] * (1.0 - f2 / f1) * n_fn[ik]
                continue
            if r2 is not None and {m.regions[r1].kind, m.regions[r2].kind} == {SOLID, PORO}:
                self._scatter_solid_poro(kr, kn, sn, bnd, rows, first, sh.unit_normal(et, v.node_x[nodes], sh.XI_NODES[et][kn]), h, g, A)
                continue
            if r2 is not None and PORO in (m.regions[r1].kind, m.regions[r2].kind):
                self._scatter_fluid_poro(kr, kn, sn, bnd, rows, first, sh.unit_normal(et, v.node_x[nodes], sh.XI_NODES[et][kn]), h, g, A)
                continue
            if r2 is None:                                                # ordinary boundary
                ct, cv = m.ctype[bnd], m.cvalue[bnd]
                if solid:
                    for il in range(3):
                        for ik in range(3):
                            if ct[ik] == 0:
                                A[rows[il], m.col[(sn, "t1%d" % ik)]] -= g[kn, il, ik]; b[rows[il]] -= h[kn, il, ik] * cv[ik]
                            else:
                                A[rows[il], m.col[(sn, "u1%d" % ik)]] += h[kn, il, ik]; b[rows[il]] += g[kn, il, ik] * cv[ik]
                else:
                    if ct[0] == 0:
                        A[rows[0], m.col[(sn, "un1")]] -= g[kn]; b[rows[0]] -= h[kn] * cv[0]
                    elif ct[0] == 1:
                        A[rows[0], m.col[(sn, "p1")]] += h[kn]; b[rows[0]] += g[kn] * cv[0]
                    elif ct[0] == 2:                                      # p unknown, Un = -i/(rho c omega) p (assemble_bem_harpot_equation.f90:97-102)
                        A[rows[0], m.col[(sn, "p1")]] += h[kn] + 1j / v.material.rho / v.material.c / self._omega * g[kn]
                    else:                                                 # p unknown, Un = -(i/(rho c omega) + 1/(2 R rho omega^2)) p (:103-110)
                        A[rows[0], m.col[(sn, "p1")]] += h[kn] + (1j / v.material.rho / v.material.c / self._omega
                                                                + 1.0 / (2.0 * cv[0] * v.material.rho * self._omega ** 2)) * g[kn]
                continue
            k1, k2 = m.regions[r1].kind, m.regions[r2].kind
            # element(se_int)%n_fn(:,kn): unit normal of the element at its node, mesh orientation = outward from region 1
            n_fn = sh.unit_normal(et, v.node_x[nodes], sh.XI_NODES[et][kn])
            if solid:
                other = k2 if first else k1
                for il in range(3):
                    for ik in range(3):
                        if other == SOLID:                                # u1, t1 active; region 2: t2 = -t1
                            A[rows[il], m.col[(sn, "u1%d" % ik)]] += h[kn, il, ik]
                            A[rows[il], m.col[(sn, "t1%d" % ik)]] += (-g[kn, il, ik] if first else g[kn, il, ik])
                        elif first:                                       # solid(1)-fluid(2): t1 = -p2 n1
                            A[rows[il], m.col[(sn, "u1%d" % ik)]] += h[kn, il, ik]
                            A[rows[il], m.col[(sn, "p2")]] += g[kn, il, ik] * n_fn[ik]
                        else:                                             # fluid(1)-solid(2): t2 = -p1 n2 = +p1 n1
                            A[rows[il], m.col[(sn, "u2%d" % ik)]] += h[kn, il, ik]
                            A[rows[il], m.col[(sn, "p1")]] -= g[kn, il, ik] * n_fn[ik]
            else:
                other = k2 if first else k1
                if other == FLUID:                                        # p1, Un1 active; region 2: Un2 = -Un1
                    A[rows[0], m.col[(sn, "p1")]] += h[kn]
                    A[rows[0], m.col[(sn, "un1")]] += (-g[kn] if first else g[kn])
                elif first:                                               # fluid(1)-solid(2): Un1 = u2 . n1
                    A[rows[0], m.col[(sn, "p1")]] += h[kn]
                    for ik in range(3):
                        A[rows[0], m.col[(sn, "u2%d" % ik)]] -= g[kn] * n_fn[ik]
                else:                                                     # solid(1)-fluid(2): Un2 = u1 . n2 = -u1 . n1
                    A[rows[0], m.col[(sn, "p2")]] += h[kn]
                    for ik in range(3):
                        A[rows[0], m.col[(sn, "u1%d" % ik)]] += g[kn] * n_fn[ik]

    def _scatter_solid_poro(self, kr, kn, sn, bnd, rows, first, n_fn, h, g, A):
        """be-be boundary between a viscoelastic solid and a poroelastic medium, perfect bonding with impervious contact (ctype 0), written out
        for each side and order: assemble_bem_harela_equation.f90:262-285 (solid = region 1), :430-455 (solid = region 2);
        assemble_bem_harpor_equation.f90:627-660 (poroelastic = region 1), :807-830 (poroelastic = region 2).  n_fn = normal of region 1."""
        m = self.m
        v = m.views[kr]
        r1, r2 = m.boundary_regions[bnd]
        col = m.col
        poro_first = m.regions[r1].kind == PORO
        if v.kind == SOLID:
            for il in range(3):
                for ik in range(3):
                    if first:                                             # VISCOELASTIC SOLID (1) - POROELASTIC MEDIA (2): u1 = u2, t1 = -t2 + tau2 n1
                        A[rows[il], col[(sn, "u2%d" % ik)]] += h[kn, il, ik]
                        A[rows[il], col[(sn, "t2%d" % ik)]] += g[kn, il, ik]
                        A[rows[il], col[(sn, "tau2")]] -= g[kn, il, ik] * n_fn[ik]
                    else:                                                 # POROELASTIC MEDIA (1) - VISCOELASTIC SOLID (2), seen from the solid
                        A[rows[il], col[(sn, "u1%d" % ik)]] += h[kn, il, ik]
                        A[rows[il], col[(sn, "t1%d" % ik)]] += g[kn, il, ik]
                        A[rows[il], col[(sn, "tau1")]] += g[kn, il, ik] * n_fn[ik]
            return
        for il in range(4):
            row = rows[il]
            if first:                                                     # POROELASTIC MEDIA (1) - VISCOELASTIC SOLID (2): Un1 = u1 . n1
                A[row, col[(sn, "tau1")]] += h[kn, il, 0]
                for ik in range(3):
                    A[row, col[(sn, "u1%d" % ik)]] -= g[kn, il, 0] * n_fn[ik]
                    A[row, col[(sn, "u1%d" % ik)]] += h[kn, il, ik + 1]
                    A[row, col[(sn, "t1%d" % ik)]] -= g[kn, il, ik + 1]
            else:                                                         # VISCOELASTIC SOLID (1) - POROELASTIC MEDIA (2), seen from the poroelastic medium
                A[row, col[(sn, "tau2")]] += h[kn, il, 0]
                for ik in range(3):
                    A[row, col[(sn, "u2%d" % ik)]] += g[kn, il, 0] * n_fn[ik]
                    A[row, col[(sn, "u2%d" % ik)]] += h[kn, il, ik + 1]
                    A[row, col[(sn, "t2%d" % ik)]] -= g[kn, il, ik + 1]

    def _scatter_fluid_poro(self, kr, kn, sn, bnd, rows, first, n_fn, h, g, A):
        """be-be boundary between an inviscid fluid and a poroelastic medium, perfectly permeable (ctype 0) or impermeable (1), written out as
        the reference does for each side and each order: assemble_bem_harpot_equation.f90:183-210 (fluid = region 1), :236-262 (fluid = region 2);
        assemble_bem_harpor_equation.f90:577-626 (poroelastic = region 1), :749-801 (poroelastic = region 2).  n_fn = normal of region 1."""
        m = self.m
        v = m.views[kr]
        r1, r2 = m.boundary_regions[bnd]
        imp = m.interface_ctype.get(bnd, 0) == 1
        poro_first = m.regions[r1].kind == PORO
        phi = m.regions[r1 if poro_first else r2].material.phi
        col = m.col
        if v.kind == FLUID:
            row = rows[0]
            if first:                                                     # INVISCID FLUID (1) - POROELASTIC MEDIA (2)
                if not imp:
                    A[row, col[(sn, "tau2")]] -= h[kn] / phi
                    A[row, col[(sn, "w2")]] += g[kn] * phi
                    for ik in range(3):
                        A[row, col[(sn, "u2%d" % ik)]] -= g[kn] * (1.0 - phi) * n_fn[ik]
                else:
                    A[row, col[(sn, "p1")]] += h[kn]
                    for ik in range(3):
                        A[row, col[(sn, "u2%d" % ik)]] -= g[kn] * n_fn[ik]
            else:                                                         # POROELASTIC MEDIA (1) - INVISCID FLUID (2), seen from the fluid
                if not imp:
                    A[row, col[(sn, "tau1")]] -= h[kn] / phi
                    A[row, col[(sn, "w1")]] += g[kn] * phi
                    for ik in range(3):
                        A[row, col[(sn, "u1%d" % ik)]] += g[kn] * (1.0 - phi) * n_fn[ik]
                else:
                    A[row, col[(sn, "p2")]] += h[kn]
                    for ik in range(3):
                        A[row, col[(sn, "u1%d" % ik)]] += g[kn] * n_fn[ik]
            return
        for il in range(4):
            row = rows[il]
            if first:                                                     # POROELASTIC MEDIA (1) - INVISCID FLUID (2)
                A[row, col[(sn, "tau1")]] += h[kn, il, 0]
                if not imp:
                    A[row, col[(sn, "w1")]] -= g[kn, il, 0]
                else:
                    for ik in range(3):
                        A[row, col[(sn, "u1%d" % ik)]] -= g[kn, il, 0] * n_fn[ik]
                for ik in range(3):
                    A[row, col[(sn, "u1%d" % ik)]] += h[kn, il, ik + 1]
                    if not imp:
                        A[row, col[(sn, "tau1")]] -= g[kn, il, ik + 1] * (1.0 - phi) / phi * n_fn[ik]
                    else:
                        A[row, col[(sn, "p2")]] += g[kn, il, ik + 1] * n_fn[ik]
                        A[row, col[(sn, "tau1")]] += g[kn, il, ik + 1] * n_fn[ik]
            else:                                                         # INVISCID FLUID (1) - POROELASTIC MEDIA (2), seen from the poroelastic medium
                A[row, col[(sn, "tau2")]] += h[kn, il, 0]
                if not imp:
                    A[row, col[(sn, "w2")]] -= g[kn, il, 0]
                else:
                    for ik in range(3):
                        A[row, col[(sn, "u2%d" % ik)]] += g[kn, il, 0] * n_fn[ik]
                for ik in range(3):
                    A[row, col[(sn, "u2%d" % ik)]] += h[kn, il, ik + 1]
                    if not imp:
                        A[row, col[(sn, "tau2")]] += g[kn, il, ik + 1] * (1.0 - phi) / phi * n_fn[ik]
                    else:
                        A[row, col[(sn, "p1")]] -= g[kn, il, ik + 1] * n_fn[ik]
                        A[row, col[(sn, "tau2")]] -= g[kn, il, ik + 1] * n_fn[ik]

    def _scatter_flat(self, kr, le, sn_col, eq, h, g, A, b, D):
        """The same scatter driven by the flat descriptors of MultiRegionModel.scatter_descriptors (what a device kernel consumes)."""
        v = self.m.views[kr]
        nd = v.ndof
        rows = self.m.row[(sn_col, eq)]
        nn = int(v.elem_ptr[le + 1] - v.elem_ptr[le])
        for j in range(nn):
            for k in range(nd):
                q = (int(v.elem_ptr[le]) + j) * nd + k
                for il, row in enumerate(rows):
                    hv = h[j, il, k] if nd > 1 else h[j]
                    gv = g[j, il, k] if nd > 1 else g[j]
                    if D["hcol"][q] >= 0:
                        A[row, D["hcol"][q]] += D["hcoef"][q] * hv
                    elif D["hcol"][q] == -1:
                        b[row] += D["hcoef"][q] * hv
                    for t in range(4):
                        c = D["gcol"][q, t]
                        if c >= 0:
                            A[row, c] += D["gcoef"][q, t] * gv
                        elif c == -1:
                            b[row] += D["gcoef"][q, t] * gv

    def assemble(self, omega, flat=False):
        """-> A (n_dof x n_dof), b of one frequency for the coupled system.  flat = True: scatter through the flat descriptors."""
        m = self.m
        self._omega = omega
        desc = [m.scatter_descriptors(kr, omega) for kr in range(len(m.views))] if flat else None
        n = m.n_dof
        A = np.zeros((n, n), dtype=np.complex128); b = np.zeros(n, dtype=np.complex128)
        for kr, v in enumerate(m.views):
            hd, mat = self.h[kr], v.material
            d1J = mat.rho * omega ** 2 if v.kind == FLUID else None
            for c in range(v.n_colloc):
                x_i, sn_col, eq = v.colloc_x[c], int(v.colloc_node[c]), int(v.colloc_eq[c])
                le_own, hfree = self._free_term(kr, c, omega)
                n_img = v.n_elem << len(getattr(m, "symplane_eid", ()))   # symmetry planes: element ks * n_elem + le is image ks of element le, multipliers applied
                for li in range(n_img):
                    le = li % v.n_elem
                    if v.kind == SOLID:
                        h, g, _, _ = hd.pair(li, x_i, omega, mat)
                    elif v.kind == PORO:
                        h, g, _ = hd.pair(li, x_i, omega, mat)
                    else:
                        h, g, _ = hd.pair(li, x_i, omega, mat)
                        g = g * d1J                                       # the flux unknown is Un = (dp/dn)/(rho omega^2)
                    if li == le_own:
                        h = h + hfree
                    if kr in m.incident:                                  # incident field of the region: b += hp u_inc - gp t_inc on every boundary class
                        ui, ti = m.incident[kr]                           # (an image takes the root's values; its multipliers are already in h, g)
                        sl = slice(int(v.elem_ptr[le]), int(v.elem_ptr[le + 1]))
                        rows = m.row[(sn_col, eq)]
                        if v.ndof == 1:
                            b[rows[0]] += np.dot(h, ui[sl, 0]) - np.dot(g, ti[sl, 0])
                        else:
                            for il, row in enumerate(rows):
                                b[row] += np.sum(h[:, il, :] * ui[sl]) - np.sum(g[:, il, :] * ti[sl])
                    if flat:
                        self._scatter_flat(kr, le, sn_col, eq, h, g, A, b, desc[kr])
                    else:
                        self._scatter(kr, le, sn_col, eq, h, g, A, b)
        return A, b
