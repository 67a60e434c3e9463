"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see harela3d_oracle.cpp header).

May be imported only by tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke().  Never by multifebe_b200/.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_setup.restype = C.c_void_p
        L.orc_telles_barr.restype = C.c_double
        L.orc_telles_barr.argtypes = [C.c_double]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _ri(z):
    z = complex(z)
    return np.array([z.real, z.imag], dtype=np.float64)


def _set_incident(o, u_inc, t_inc, nd):
    if u_inc is None:
        lib().orc_set_incident(o.h, None, None)
        return
    u = np.ascontiguousarray(u_inc, dtype=np.complex128).reshape(-1, nd); t = np.ascontiguousarray(t_inc, dtype=np.complex128).reshape(-1, nd)
    assert len(u) == len(t) == int(o.m.elem_ptr[-1])
    lib().orc_set_incident(o.h, _p(u), _p(t))


def _apply_symmetry(L, h, m):
    """[symmetry planes] of the model: image elements and multipliers inside the oracle (orc_set_symmetry_s)."""
    eid = np.ascontiguousarray(getattr(m, "symplane_eid", np.zeros(0)), dtype=np.int32)
    if len(eid):
        t = np.ascontiguousarray(m.symplane_t, dtype=np.float64)
        sc = np.ascontiguousarray(m.symplane_s, dtype=np.float64)
        if L.orc_set_symmetry_s(h, C.c_int(len(eid)), _p(eid), _p(t), _p(sc)):
            raise ValueError("oracle: invalid symmetry planes")


class Oracle:
    """Oracle handle for one Model (multifebe_b200.host.Model)."""

    def __init__(self, model):
        L = lib()
        m = self.m = model
        self._keep = [np.ascontiguousarray(a) for a in (
            m.node_x, m.etype, m.elem_ptr, m.elem_node, m.elem_reversed, m.colloc_x, m.colloc_node, m.colloc_elem,
            m.colloc_kn, m.colloc_xi, m.row, m.col_u, m.col_t, m.ctype, m.precalset_gln)]
        k = self._keep
        self.h = C.c_void_p(L.orc_setup(
            C.c_int(m.n_node), _p(k[0]), C.c_int(m.n_elem), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]),
            C.c_int(m.n_colloc), _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]),
            _p(k[10]), _p(k[11]), _p(k[12]), _p(k[13]), C.c_int(m.n_dof),
            C.c_double(m.qsi_relative_error), C.c_int(m.qsi_ns_max), C.c_int(len(m.precalset_gln)), _p(k[14]),
            C.c_double(m.geometric_tolerance)))
        if (np.asarray(m.ctype) == 10).any():
            nf = np.ascontiguousarray(m.n_fn, dtype=np.float64); self._keep.append(nf)
            L.orc_set_node_normals(self.h, _p(nf))
        _apply_symmetry(L, self.h, m)

    def __del__(self):
        try:
            lib().orc_free(self.h)
        except Exception:
            pass

    def set_incident(self, u_inc=None, t_inc=None):
        """Incident field at the nodes of every element, (sum nn, 3) complex each in element order (element()%incident_c); None clears."""
        _set_incident(self, u_inc, t_inc, 3)

    def assemble(self, omega, mat, nthreads=0):
        """-> A (n_dof x n_dof, Fortran order), b (n_dof), stats dict.  One frequency, A and b start at zero."""
        n = self.m.n_dof
        A = np.zeros((n, n), dtype=np.complex128, order="F")
        b = np.zeros(n, dtype=np.complex128)
        st = np.zeros(44, dtype=np.int64)
        cv = np.ascontiguousarray(self.m.cvalue)
        err = lib().orc_assemble(self.h, C.c_double(omega), _p(_ri(mat.lam)), _p(_ri(mat.mu)), C.c_double(mat.rho),
                                 _p(_ri(mat.nu)), _p(cv), _p(A), _p(b), C.c_int(nthreads), _p(st))
        if err:
            raise RuntimeError("oracle: invalid normals/tangents configuration in free-term")
        stats = {"pairs_regular": {g: int(st[g]) for g in range(33) if st[g]}, "pts_regular": int(st[33]),
                 "pairs_adaptive": int(st[34]), "leaves": int(st[35]), "pts_adaptive": int(st[36]),
                 "pairs_singular": int(st[37]), "pts_singular": int(st[38]), "li_points": int(st[39])}
        return A, b, stats

    def assemble_static(self, mat, nthreads=0):
        """Static elasticity (build_lse_mechanics_bem_staela + assemble_bem_staela_equation): real A, b and the plan stats."""
        n = self.m.n_dof
        A = np.zeros((n, n), dtype=np.float64, order="F")
        b = np.zeros(n, dtype=np.float64)
        st = np.zeros(44, dtype=np.int64)
        cv = np.ascontiguousarray(self.m.cvalue.real, dtype=np.float64)
        err = lib().orc_assemble_static(self.h, C.c_double(mat.mu_r), C.c_double(mat.nu_r), _p(cv), _p(A), _p(b), C.c_int(nthreads), _p(st))
        if err == 7:
            raise RuntimeError("oracle: the static assembly produced a nonzero imaginary part")
        if err:
            raise RuntimeError("oracle: invalid normals/tangents configuration in free-term")
        stats = {"pairs_regular": {g: int(st[g]) for g in range(33) if st[g]}, "pts_regular": int(st[33]),
                 "pairs_adaptive": int(st[34]), "leaves": int(st[35]), "pts_adaptive": int(st[36]),
                 "pairs_singular": int(st[37]), "pts_singular": int(st[38]), "li_points": int(st[39])}
        return A, b, stats

    def pair_static(self, e, x_i, mat):
        nn = int(self.m.elem_ptr[e % self.m.n_elem + 1] - self.m.elem_ptr[e % self.m.n_elem])   # e >= n_elem: a symmetry image
        h = np.zeros((nn, 3, 3)); g = np.zeros((nn, 3, 3))
        x_i = np.ascontiguousarray(x_i, dtype=np.float64)
        mode = lib().orc_pair_static(self.h, C.c_int(e), _p(x_i), C.c_double(mat.mu_r), C.c_double(mat.nu_r), _p(h), _p(g))
        return h, g, mode

    def assemble_colloc_sample(self, omega, mat, c_offset, c_stride, nthreads=0):
        """Bounded sample for the CPU baseline: all elements x every c_stride-th collocation point.
        -> (compact A_s (3*n_sample x n_dof), b_s, n_sample, quadrature points evaluated)."""
        ns = len(range(c_offset, self.m.n_colloc, c_stride))
        A = np.zeros((3 * ns, self.m.n_dof), dtype=np.complex128, order="F")
        b = np.zeros(3 * ns, dtype=np.complex128)
        pts = C.c_longlong(0)
        cv = np.ascontiguousarray(self.m.cvalue)
        n = lib().orc_assemble_colloc_sample(self.h, C.c_double(omega), _p(_ri(mat.lam)), _p(_ri(mat.mu)), C.c_double(mat.rho), _p(cv),
                                             C.c_int(c_offset), C.c_int(c_stride), _p(A), _p(b), C.c_int(nthreads), C.byref(pts))
        assert n == ns
        return A, b, ns, pts.value

    def assemble_colloc_sample_static(self, mat, c_offset, c_stride, nthreads=0):
        """The bounded CPU-baseline sample with the static kernels (complex containers with zero imaginary parts)."""
        ns = len(range(c_offset, self.m.n_colloc, c_stride))
        A = np.zeros((3 * ns, self.m.n_dof), dtype=np.complex128, order="F")
        b = np.zeros(3 * ns, dtype=np.complex128)
        pts = C.c_longlong(0)
        cv = np.ascontiguousarray(self.m.cvalue)
        n = lib().orc_assemble_colloc_sample_static(self.h, C.c_double(mat.mu_r), C.c_double(mat.nu_r), _p(cv), C.c_int(c_offset), C.c_int(c_stride),
                                                    _p(A), _p(b), C.c_int(nthreads), C.byref(pts))
        assert n == ns
        return A, b, ns, pts.value

    def pair(self, e, x_i, omega, mat):
        """h, g (n,3,3) complex of one (collocation point, element) pair and the integration mode.
        e >= n_elem addresses image e // n_elem of root element e % n_elem (symmetry planes), signs symconf_t applied."""
        nn = int(self.m.elem_ptr[e % self.m.n_elem + 1] - self.m.elem_ptr[e % self.m.n_elem])
        h = np.zeros((nn, 3, 3), dtype=np.complex128)
        g = np.zeros((nn, 3, 3), dtype=np.complex128)
        st = np.zeros(8, dtype=np.int64)
        x_i = np.ascontiguousarray(x_i, dtype=np.float64)
        mode = lib().orc_pair(self.h, C.c_int(e), _p(x_i), C.c_double(omega), _p(_ri(mat.lam)), _p(_ri(mat.mu)),
                              C.c_double(mat.rho), _p(h), _p(g), _p(st))
        return h, g, mode, st

    def pair_hbie(self, e, x_i, n_i, omega, mat):
        """m, l (n,3,3) complex of the hypersingular equation for a point off the element with unit normal n_i, and the mode
        (e >= n_elem: an image element, as in pair)."""
        nn = int(self.m.elem_ptr[e % self.m.n_elem + 1] - self.m.elem_ptr[e % self.m.n_elem])
        m = np.zeros((nn, 3, 3), dtype=np.complex128); l = np.zeros((nn, 3, 3), dtype=np.complex128)
        x_i = np.ascontiguousarray(x_i, dtype=np.float64); n_i = np.ascontiguousarray(n_i, dtype=np.float64)
        mode = lib().orc_pair_hbie(self.h, C.c_int(e), _p(x_i), _p(n_i), C.c_double(omega), _p(_ri(mat.lam)), _p(_ri(mat.mu)), C.c_double(mat.rho), _p(m), _p(l))
        if mode < 0:
            raise RuntimeError("oracle: hypersingular integration with the collocation point on the element is not restated")
        return m, l, mode

    def pair_hbie_static(self, e, x_i, n_i, mat):
        """Static (Kelvin) m, l (n,3,3) of the hypersingular equation (fbem_bem_staela3d_hbie_ext_pre / _ext_adp)."""
        nn = int(self.m.elem_ptr[e % self.m.n_elem + 1] - self.m.elem_ptr[e % self.m.n_elem])   # e >= n_elem: a symmetry image
        m = np.zeros((nn, 3, 3)); l = np.zeros((nn, 3, 3))
        x_i = np.ascontiguousarray(x_i, dtype=np.float64); n_i = np.ascontiguousarray(n_i, dtype=np.float64)
        mode = lib().orc_pair_hbie_static(self.h, C.c_int(e), _p(x_i), _p(n_i), C.c_double(mat.mu_r), C.c_double(mat.nu_r), _p(m), _p(l))
        if mode < 0:
            raise RuntimeError("oracle: hypersingular integration with the collocation point on the element is not restated")
        return m, l, mode

    def pair_mode(self, e, x_i):
        x_i = np.ascontiguousarray(x_i, dtype=np.float64)
        d = C.c_double(0.0)
        bx = np.zeros(2)
        mode = lib().orc_pair_mode(self.h, C.c_int(e), _p(x_i), C.byref(d), _p(bx))
        return mode, d.value, bx

    def element_data(self, e):
        cl, br, g = C.c_double(), C.c_double(), C.c_int()
        bc = np.zeros(3)
        lib().orc_element_data(self.h, C.c_int(e), C.byref(cl), C.byref(g), _p(bc), C.byref(br))
        return cl.value, g.value, bc, br.value


def zexp_decomposed(z):
    E = np.zeros(7, dtype=np.complex128)
    lib().orc_zexp_decomposed(_p(_ri(z)), _p(E))
    return E


def fundamental_solutions(x, n, x_i, omega, mat):
    u = np.zeros((3, 3), dtype=np.complex128)
    t = np.zeros((3, 3), dtype=np.complex128)
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, n, x_i)]
    lib().orc_fundamental_solutions(_p(a[0]), _p(a[1]), _p(a[2]), C.c_double(omega), _p(_ri(mat.lam)), _p(_ri(mat.mu)),
                                    C.c_double(mat.rho), _p(u), _p(t))
    return u, t


def qs_n(telles, etype, f, re, d, barxi):
    bx = np.ascontiguousarray(barxi, dtype=np.float64)
    return lib().orc_qs_n(C.c_int(int(telles)), C.c_int(etype), C.c_int(f), C.c_double(re), C.c_double(d), _p(bx))


def nearest(etype, x_nodes, x_i):
    x = np.ascontiguousarray(x_nodes, dtype=np.float64)
    xi = np.ascontiguousarray(x_i, dtype=np.float64)
    bx = np.zeros(2)
    rmin, d, method = C.c_double(), C.c_double(), C.c_int()
    lib().orc_nearest(C.c_int(etype), _p(x), _p(xi), _p(bx), C.byref(rmin), C.byref(d), C.byref(method))
    return bx, rmin.value, d.value, method.value


def tables(family, n):
    L = lib()
    if family == 3:
        npt = L.orc_wantri_n(C.c_int(n))
        x = np.zeros(2 * npt); w = np.zeros(npt)
    else:
        x = np.zeros(n); w = np.zeros(n)
    L.orc_tables(C.c_int(family), C.c_int(n), _p(x), _p(w))
    return x, w


def freeterm(normals, tangents, nu, tol=1e-6):
    n = np.ascontiguousarray(normals, dtype=np.float64)
    t = np.ascontiguousarray(tangents, dtype=np.float64)
    c = np.zeros((3, 3), dtype=np.complex128)
    err = lib().orc_freeterm(C.c_int(len(n)), _p(n), _p(t), C.c_double(tol), _p(_ri(nu)), _p(c))
    return c, err


def fundamental_solutions_hbie(x, n, x_i, n_i, omega, mat):
    """d*, s* (3,3) [l][k] (fbem_bem_harela3d_hbie_d / _s, lib/fbem/src/bem_harela3d.f90:2472-2568)."""
    d = np.zeros((3, 3), dtype=np.complex128); s = np.zeros((3, 3), dtype=np.complex128)
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, n, x_i, n_i)]
    lib().orc_fundamental_solutions_hbie(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), C.c_double(omega), _p(_ri(mat.lam)), _p(_ri(mat.mu)), C.c_double(mat.rho), _p(d), _p(s))
    return d, s


def fundamental_solutions_static(x, n, x_i, mat):
    """Kelvin u*, t* (3,3) [l][k] (fbem_bem_staela3d_sbie_u / _t, lib/fbem/src/bem_staela3d.f90:408-461)."""
    u = np.zeros((3, 3)); t = np.zeros((3, 3))
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, n, x_i)]
    lib().orc_fundamental_solutions_static(_p(a[0]), _p(a[1]), _p(a[2]), C.c_double(mat.mu_r), C.c_double(mat.nu_r), _p(u), _p(t))
    return u, t


def lu_solve_real(A, b):
    """The reference's solve_lse_r default path (src/solve_lse_r.f90:137,189): dgetrf + dgetrs (scipy-bundled OpenBLAS)."""
    from scipy.linalg import lapack
    lu, piv, info = lapack.dgetrf(A, overwrite_a=False)
    if info != 0:
        raise RuntimeError("dgetrf info=%d" % info)
    x, info = lapack.dgetrs(lu, piv, b)
    if info != 0:
        raise RuntimeError("dgetrs info=%d" % info)
    return x, lu, piv


def lu_solve(A, b):
    """The reference's solve_lse_c default path (src/solve_lse_c.f90:124,176): zgetrf + zgetrs from the
    OpenBLAS the interpreter ships (scipy-bundled OpenBLAS; MultiFEBE links an unpinned OpenBLAS)."""
    from scipy.linalg import lapack
    lu, piv, info = lapack.zgetrf(A, overwrite_a=False)
    if info != 0:
        raise RuntimeError("zgetrf info=%d" % info)
    x, info = lapack.zgetrs(lu, piv, b)
    if info != 0:
        raise RuntimeError("zgetrs info=%d" % info)
    return x, lu, piv


class PotOracle:
    """Oracle handle for one FluidModel (inviscid fluid BE region: fbem_bem_harpot3d_* + build_lse_mechanics_bem_harpot +
    assemble_bem_harpot_equation, one equation / unknown per node)."""

    def __init__(self, model):
        L = lib()
        L.orc_setup_pot.restype = C.c_void_p
        m = self.m = model
        assert m.ndof == 1
        self._keep = [np.ascontiguousarray(a) for a in (
            m.node_x, m.etype, m.elem_ptr, m.elem_node, m.elem_reversed, m.colloc_x, m.colloc_node, m.colloc_elem,
            m.colloc_kn, m.colloc_xi, m.row, m.col_u, m.col_t, m.ctype, m.precalset_gln)]
        k = self._keep
        self.h = C.c_void_p(L.orc_setup_pot(
            C.c_int(m.n_node), _p(k[0]), C.c_int(m.n_elem), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]),
            C.c_int(m.n_colloc), _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]),
            _p(k[10]), _p(k[11]), _p(k[12]), _p(k[13]), C.c_int(m.n_dof),
            C.c_double(m.qsi_relative_error), C.c_int(m.qsi_ns_max), C.c_int(len(m.precalset_gln)), _p(k[14]),
            C.c_double(m.geometric_tolerance)))
        _apply_symmetry(L, self.h, m)

    def __del__(self):
        try:
            lib().orc_free(self.h)
        except Exception:
            pass

    def set_incident(self, p_inc=None, un_inc=None):
        """Incident field of a fluid region at the nodes of every element: pressure and normal displacement, (sum nn,) complex each; None clears."""
        _set_incident(self, p_inc, un_inc, 1)

    def assemble(self, omega, fluid, nthreads=0):
        """-> A (n_dof x n_dof, Fortran order), b (n_dof), stats dict; one frequency, A and b start at zero."""
        n = self.m.n_dof
        A = np.zeros((n, n), dtype=np.complex128, order="F")
        b = np.zeros(n, dtype=np.complex128)
        st = np.zeros(44, dtype=np.int64)
        cv = np.ascontiguousarray(self.m.cvalue)
        err = lib().orc_assemble_pot(self.h, C.c_double(omega), C.c_double(fluid.rho), _p(_ri(fluid.c)), _p(cv), _p(A), _p(b), C.c_int(nthreads), _p(st))
        if err:
            raise RuntimeError("oracle: potential assembly failed (%d)" % err)
        stats = {"pairs_regular": {g: int(st[g]) for g in range(33) if st[g]}, "pts_regular": int(st[33]),
                 "pairs_adaptive": int(st[34]), "leaves": int(st[35]), "pts_adaptive": int(st[36]),
                 "pairs_singular": int(st[37]), "pts_singular": int(st[38])}
        return A, b, stats

    def pair(self, e, x_i, omega, fluid):
        """h, g (n) complex of one (collocation point, element) pair (g NOT yet scaled by rho omega^2) and the integration mode."""
        nn = int(self.m.elem_ptr[e % self.m.n_elem + 1] - self.m.elem_ptr[e % self.m.n_elem])
        h = np.zeros(nn, dtype=np.complex128); g = np.zeros(nn, dtype=np.complex128)
        x_i = np.ascontiguousarray(x_i, dtype=np.float64)
        mode = lib().orc_pair_pot(self.h, C.c_int(e), _p(x_i), C.c_double(omega), C.c_double(fluid.rho), _p(_ri(fluid.c)), _p(h), _p(g))
        return h, g, mode


def fundamental_solutions_pot(x, n, x_i, omega, fluid):
    """p*, q* of fbem_bem_harpot3d_sbie_p / _q."""
    x, n, x_i = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, n, x_i))
    po = np.zeros(2); qo = np.zeros(2)
    lib().orc_fundamental_solutions_pot(_p(x), _p(n), _p(x_i), C.c_double(omega), C.c_double(fluid.rho), _p(_ri(fluid.c)), _p(po), _p(qo))
    return complex(po[0], po[1]), complex(qo[0], qo[1])


class PorOracle:
    """Oracle handle for one PoroModel (Biot poroelastic BE region: fbem_bem_harpor3d_* + build_lse_mechanics_bem_harpor + the ordinary-boundary
    scatter of assemble_bem_harpor_equation, four equations / unknowns per node)."""

    def __init__(self, model):
        L = lib()
        L.orc_setup_por.restype = C.c_void_p
        m = self.m = model
        assert m.ndof == 4
        self._keep = [np.ascontiguousarray(a) for a in (
            m.node_x, m.etype, m.elem_ptr, m.elem_node, m.elem_reversed, m.colloc_x, m.colloc_node, m.colloc_elem,
            m.colloc_kn, m.colloc_xi, m.row, m.col_u, m.col_t, m.ctype, m.precalset_gln)]
        k = self._keep
        self.h = C.c_void_p(L.orc_setup_por(
            C.c_int(m.n_node), _p(k[0]), C.c_int(m.n_elem), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]),
            C.c_int(m.n_colloc), _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]),
            _p(k[10]), _p(k[11]), _p(k[12]), _p(k[13]), C.c_int(m.n_dof),
            C.c_double(m.qsi_relative_error), C.c_int(m.qsi_ns_max), C.c_int(len(m.precalset_gln)), _p(k[14]),
            C.c_double(m.geometric_tolerance)))
        _apply_symmetry(L, self.h, m)

    def __del__(self):
        try:
            lib().orc_free(self.h)
        except Exception:
            pass

    def set_incident(self, u_inc=None, t_inc=None):
        """Incident field of a poroelastic region at the nodes of every element: (tau, u_k) and (Un, t_k), (sum nn, 4) complex each; None clears."""
        _set_incident(self, u_inc, t_inc, 4)

    def assemble(self, omega, poro, nthreads=0):
        n = self.m.n_dof
        A = np.zeros((n, n), dtype=np.complex128, order="F")
        b = np.zeros(n, dtype=np.complex128)
        st = np.zeros(44, dtype=np.int64)
        cv = np.ascontiguousarray(self.m.cvalue)
        pr = poro.props()
        err = lib().orc_assemble_por(self.h, C.c_double(omega), _p(pr), _p(cv), _p(A), _p(b), C.c_int(nthreads), _p(st))
        if err:
            raise RuntimeError("oracle: poroelastic assembly failed (%d)" % err)
        stats = {"pairs_regular": {g: int(st[g]) for g in range(33) if st[g]}, "pts_regular": int(st[33]), "pairs_adaptive": int(st[34]),
                 "leaves": int(st[35]), "pts_adaptive": int(st[36]), "pairs_singular": int(st[37]), "pts_singular": int(st[38])}
        return A, b, stats

    def pair(self, e, x_i, omega, poro):
        nn = int(self.m.elem_ptr[e % self.m.n_elem + 1] - self.m.elem_ptr[e % self.m.n_elem])
        h = np.zeros((nn, 4, 4), dtype=np.complex128); g = np.zeros((nn, 4, 4), dtype=np.complex128)
        x_i = np.ascontiguousarray(x_i, dtype=np.float64)
        pr = poro.props()
        mode = lib().orc_pair_por(self.h, C.c_int(e), _p(x_i), C.c_double(omega), _p(pr), _p(h), _p(g))
        return h, g, mode


def fundamental_solutions_por(x, n, x_i, omega, poro):
    """u*, t* (4 x 4, [l][k]) of the poroelastic fundamental solution and (k1, k2, k3, Z, J)."""
    x, n, x_i = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, n, x_i))
    u = np.zeros((4, 4), dtype=np.complex128); t = np.zeros((4, 4), dtype=np.complex128); k = np.zeros(5, dtype=np.complex128)
    pr = poro.props()
    lib().orc_fundamental_solutions_por(_p(x), _p(n), _p(x_i), C.c_double(omega), _p(pr), _p(u), _p(t), _p(k))
    return u, t, k
