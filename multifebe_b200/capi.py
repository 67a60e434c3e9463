"""ctypes binding of libmfb.so (include/mfb.h) + the host-side mirror of the reference's call sites.

The reference is a Fortran program without an FFI; its two seams on this path are the subroutines
`build_lse_mechanics_bem_harela(kf,kr)` (src/build_lse_mechanics_bem_harela.f90:22) and `solve_lse_c(...)`
(src/solve_lse_c.f90:25).  `Problem` exposes methods with those names and argument meaning so the parity tests read like
the reference's own call sequence (src/multifebe.f90:107-124):

    prob = Problem(ctx, model)                       # build_data + build_auxiliary_variables (once)
    A, b = prob.build_lse_mechanics_bem_harela(omega, material)   # A_c, b_c of this frequency
    x = prob.solve_lse_c()                           # zgetrf + zgetrs on the device-resident system

There is no CPU fallback: loading fails loudly if libmfb.so is missing and mfb_init fails without a CUDA device.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

STAT = dict(PAIRS_REGULAR=0, POINTS_REGULAR=1, PAIRS_ADAPTIVE=2, LEAVES=3, POINTS_ADAPTIVE=4, PAIRS_SINGULAR=5,
            POINTS_SINGULAR=6, NEAR_PAIRS=7, MS_ZERO=8, MS_REGULAR=9, MS_ADAPTIVE=10, MS_SINGULAR=11, MS_FREETERM=12,
            MS_LU=13, MS_SOLVE=14, MS_GEMM=15, MS_PANEL=16, LAUNCHES=17, MS_SETUP_HOST=18, MS_ASSEMBLE=19,
            FLOPS_REGULAR=20, MS_TRSM=21, MS_SWAP=22, LU_LAUNCHES=23, GEMM_LAUNCHES=24, GEMM_FLOPS=25, GEMM_EXEC_FLOPS=26,
            MS_REDIST=27, MS_DIST_LU=28, MS_DIST_SOLVE=29, MS_DIST_TOTAL=30)
STAT_COUNT = 32


class MfbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mfb error %d: %s" % (code, msg))
        self.code = code


# A process that keeps several problems in flight (ProblemLanes) uses ~3 streams per problem; the default of 8 hardware work queues makes streams share
# a queue and serialises them (measured on ME-TH-EL-001: 297 -> 466 solves/s with 8 lanes).  The variable is read when the CUDA context is created, so it
# is set at import, before anything touches the device; an explicit setting of the user wins.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def lib():
    """Load libmfb.so (built in-tree by multifebe_b200.build).  Raises if absent -- never falls back to another path."""
    global _LIB
    if _LIB is None:
        so = os.environ.get("MFB_LIB") or os.path.join(_HERE, "libmfb.so")   # MFB_LIB: development builds (kernel variants)
        if not os.path.exists(so):
            raise ImportError("multifebe_b200/libmfb.so is not built; run `python -m multifebe_b200.build` "
                              "(there is no CPU fallback for this path)")
        L = C.CDLL(so)
        L.mfb_last_error.restype = C.c_char_p
        _LIB = L
    return _LIB


def _check(code):
    if code != 0:
        raise MfbError(code, lib().mfb_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _z(v):
    v = complex(v)
    return np.array([v.real, v.imag], dtype=np.float64)


def dist_unique_id():
    """128-byte NCCL id created by rank 0 (the host broadcasts it to the other ranks, e.g. with torch.distributed)."""
    buf = C.create_string_buffer(128)
    _check(lib().mfb_dist_unique_id(buf))
    return buf.raw


def dist_layout(n, nb, nranks, rank):
    """Global column of every local column of `rank` in the block-cyclic layout of the distributed LU (host only)."""
    ncl = C.c_int()
    _check(lib().mfb_dist_layout(C.c_int(n), C.c_int(nb), C.c_int(nranks), C.c_int(rank), C.byref(ncl), None))
    out = np.zeros(ncl.value, dtype=np.int32)
    _check(lib().mfb_dist_layout(C.c_int(n), C.c_int(nb), C.c_int(nranks), C.c_int(rank), C.byref(ncl), _p(out)))
    return out


def dist_partition_tiles(tile_row0, tile_nbytes, n_dof, nranks):
    """(tile_rank, row_bounds): owner rank of every collocation tile and the internal-row range of every rank (host only)."""
    r0 = np.ascontiguousarray(tile_row0, dtype=np.int32); nby = np.ascontiguousarray(tile_nbytes, dtype=np.int32)
    tr = np.zeros(len(r0), dtype=np.int32); rb = np.zeros(nranks + 1, dtype=np.int32)
    _check(lib().mfb_dist_partition_tiles(C.c_int(len(r0)), _p(r0), _p(nby), C.c_int(n_dof), C.c_int(nranks), _p(tr), _p(rb)))
    return tr, rb


def freeterm(normals, tangents, nu, tol=1e-6):
    """(Mantic's matrix c (3 x 3, complex), scalar free term cp) of a boundary node from the unit normals / boundary tangents of its elements
    (mfb_freeterm_terms; host only)."""
    n = np.ascontiguousarray(normals, dtype=np.float64); t = np.ascontiguousarray(tangents, dtype=np.float64)
    cp = C.c_double(0.0); sb = np.zeros(9)
    _check(lib().mfb_freeterm_terms(C.c_int(len(n)), _p(n), _p(t), C.c_double(tol), C.byref(cp), _p(sb)))
    return cp.value * np.eye(3) - sb.reshape(3, 3) / (8.0 * np.pi * (1.0 - complex(nu))), cp.value


class CoupledProblem:
    """Several BE regions (elastic solids, inviscid fluids, poroelastic media) coupled through be-be interfaces, on the GPU through the
    single-region kernels: two auxiliary single-region assemblies per region (host/coupled.py), the column combination on the host, the
    coupled system factorised and solved on the device (mfb_zsolve with host arrays).  First functional version: the local matrices travel
    through host memory; the resident combination kernel is the next step."""

    def __init__(self, ctx, mrm):
        from .host.coupled import local_models
        self.ctx, self.m = ctx, mrm
        self.locals = [local_models(mrm, kr) for kr in range(len(mrm.regions))]      # kept between frequencies, like the plans of their problems
        self.problems = {}
        for mH, mG, _ in self.locals:
            self.problems[id(mH)] = Problem(ctx, mH); self.problems[id(mG)] = Problem(ctx, mG)
        self._solver = None
        self._terms = {}                                 # host/coupled.py::TermArrays per region, kept between frequencies

    def _assemble_local(self, model, region, omega, incident=None):
        pr = self.problems[id(model)]
        self._set_local_incident(pr, incident)
        if region.kind == "solid":
            A, b = pr.build_lse_mechanics_bem_harela(omega, region.material)
        elif region.kind == "fluid":
            A, b = pr.build_lse_mechanics_bem_harpot(omega, region.material)
        else:
            A, b = pr.build_lse_mechanics_bem_harpor(omega, region.material)
        return A if incident is None else (A, b)

    def _set_local_incident(self, pr, incident):
        if incident is None and not getattr(pr, "_has_incident", False):
            return                                       # nothing set, nothing to clear: models without incident fields make no extra call
        pr.set_incident(*(incident if incident is not None else (None, None)))
        pr._has_incident = incident is not None

    def set_incident(self, kr, u_inc=None, t_inc=None):
        """Incident wave field of region kr for the next assembly (MultiRegionModel.set_incident): it reaches the device through the region's H
        problem (mfb_har{ela,pot,por}3d_set_incident), whose right-hand side then is sum_e (hp u_inc - gp t_inc); the free-term part is added on the host."""
        self.m.set_incident(kr, u_inc, t_inc)

    def assemble(self, omega):
        """-> A, b of the coupled system (host arrays)."""
        from .host import coupled
        return coupled.assemble_coupled(self.m, omega, self._assemble_local, freeterm, locals_=self.locals)

    def solve_frequency_resident(self, omega):
        """The same without moving matrices through the host: the local systems stay on the device, mfb_combine_columns / mfb_add_entries build
        the coupled system in the resident matrix of a solver problem, mfb_zsolve factorises and solves it there (not yet run on hardware)."""
        from .host.coupled import TermArrays
        from .host import coupled as coupled_mod
        n = self.m.n_dof
        if self._solver is None or self._solver.m.n_dof != n:
            self._solver = _lu_only_problem(self.ctx, n)
        dst = self._solver
        _check(lib().mfb_system_zero(dst.h))
        for kr, (mH, mG, mp) in enumerate(self.locals):
            region = self.m.regions[kr]
            if kr not in self._terms:
                self._terms[kr] = TermArrays(self.m, kr, mp, freeterm)
            ta = self._terms[kr].at(omega)
            inc = self.m.incident.get(kr)
            for model, (sc, dc, cf) in ((mH, ta.H), (mG, ta.G)):
                pr = self.problems[id(model)]
                if model is mH:
                    self._set_local_incident(pr, inc)
                if region.kind == "solid":
                    pr.build_lse_mechanics_bem_harela(omega, region.material, want_host=False)
                elif region.kind == "fluid":
                    pr.build_lse_mechanics_bem_harpot(omega, region.material, want_host=False)
                else:
                    pr.build_lse_mechanics_bem_harpor(omega, region.material, want_host=False)
                _check(lib().mfb_combine_columns(dst.h, pr.h, C.c_int(len(ta.row_map)), _p(ta.row_map), C.c_int(len(sc)), _p(sc), _p(dc), _p(cf)))
                if model is mH and inc is not None:                      # b_loc of the H problem = sum_e (hp u_inc - gp t_inc): n_rows values through the host
                    bl = -pr.residual_vector(np.zeros(pr.m.n_dof, dtype=np.complex128))[:len(ta.row_map)]      # A 0 - b, host row order
                    fe = coupled_mod.incident_free_terms(self.m, kr, omega, freeterm)
                    rows = np.concatenate([ta.row_map, np.array([e[0] for e in fe], dtype=np.int32)]).astype(np.int32)
                    vals = np.concatenate([bl, np.array([e[1] for e in fe], dtype=np.complex128)]).astype(np.complex128)
                    cols = -np.ones(len(rows), dtype=np.int32)
                    _check(lib().mfb_add_entries(dst.h, C.c_int(len(rows)), _p(rows), _p(cols), _p(vals)))
            if len(ta.E[0]):
                _check(lib().mfb_add_entries(dst.h, C.c_int(len(ta.E[0])), _p(ta.E[0]), _p(ta.E[1]), _p(ta.E[2])))
        ipiv = np.zeros(n, dtype=np.int32)
        _check(lib().mfb_zsolve(dst.h, C.c_int(n), None, C.c_int(n), _p(ipiv), None, C.c_int(1), C.c_int(1)))
        return dst.get_solution()

    def solve_frequency(self, omega):
        """x of the coupled system: assembled as above, zgetrf + zgetrs on the device (LAPACK zgesv semantics through mfb_zsolve)."""
        A, b = self.assemble(omega)
        n = self.m.n_dof
        if self._solver is None or self._solver.m.n_dof != n:
            self._solver = _lu_only_problem(self.ctx, n)
        return self._solver.solve_lse_c(np.asfortranarray(A), b)

    def close(self):
        for pr in self.problems.values():
            pr.close()
        if self._solver is not None:
            self._solver.close()


def _lu_only_problem(ctx, n):
    """A minimal problem object whose only use is mfb_zsolve on host matrices of size n (mfb_zsolve is tied to a problem's n_dof): one tri3 element,
    three nodes, rows and columns spread over n."""
    from .host.coupled import LocalModel
    m = LocalModel()
    m.ndof, m.n_node, m.n_elem, m.n_colloc, m.n_dof = 1, 3, 1, 3, n
    m.node_x = np.array([[0.0, 0, 0], [1.0, 0, 0], [0.0, 1, 0]])
    m.etype = np.array([5], dtype=np.int32); m.elem_ptr = np.array([0, 3], dtype=np.int32); m.elem_node = np.array([0, 1, 2], dtype=np.int32)
    m.elem_reversed = np.zeros(1, dtype=np.uint8)
    m.colloc_x = np.array([[0.2, 0.2, 0.5], [0.6, 0.2, 0.5], [0.2, 0.6, 0.5]]); m.colloc_node = np.array([0, 1, 2], dtype=np.int32)
    m.colloc_elem = -np.ones(3, dtype=np.int32); m.colloc_kn = np.zeros(3, dtype=np.int32); m.colloc_xi = np.full((3, 2), -9.0)
    m.row = np.array([[0], [1], [2]], dtype=np.int32); m.col_u = np.array([[0], [1], [2]], dtype=np.int32); m.col_t = -np.ones((3, 1), dtype=np.int32)
    m.ctype = np.ones((3, 1), dtype=np.int32); m.cvalue = np.zeros((3, 1), dtype=np.complex128)
    m.qsi_relative_error, m.qsi_ns_max, m.precalset_gln, m.geometric_tolerance = 1e-6, 16, np.arange(2, 10, dtype=np.int32), 1e-6
    return Problem(ctx, m)


class Context:
    """One context per GPU (one process per GPU)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _check(lib().mfb_init(C.c_int(device), C.byref(self.h)))
        self.device = device

    def close(self):
        if self.h:
            lib().mfb_finalize(self.h)
            self.h = C.c_void_p()

    def mark(self, slot):
        _check(lib().mfb_stream_mark(self.h, C.c_int(slot)))

    def elapsed_ms(self, slot0, slot1):
        ms = C.c_double()
        _check(lib().mfb_stream_elapsed(self.h, C.c_int(slot0), C.c_int(slot1), C.byref(ms)))
        return ms.value

    def measure_peaks(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _check(lib().mfb_measure_peaks(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"dfma_tflops": a.value, "dmma_tflops": b.value, "copy_gbs": c.value}

    def zgemm_minus(self, Cm, A, B):
        """C -= A @ B on the device (FP64 tensor pipe); returns (C, ms)."""
        A = np.asfortranarray(A, dtype=np.complex128); B = np.asfortranarray(B, dtype=np.complex128)
        Cm = np.asfortranarray(Cm, dtype=np.complex128).copy(order="F")
        m, k = A.shape; n = B.shape[1]
        ms = C.c_double()
        _check(lib().mfb_zgemm_minus(self.h, C.c_int(m), C.c_int(n), C.c_int(k), _p(A), C.c_int(m), _p(B), C.c_int(k), _p(Cm),
                                     C.c_int(m), C.byref(ms)))
        return Cm, ms.value


class Problem:
    def __init__(self, ctx, model):
        self.ctx, self.m = ctx, model
        m = model
        k = self._keep = [np.ascontiguousarray(a) for a in (
            m.node_x, m.etype, m.elem_ptr, m.elem_node, m.elem_reversed, m.colloc_x, m.colloc_node, m.colloc_elem,
            m.colloc_kn, m.colloc_xi, m.row, m.col_u, m.col_t, m.ctype, m.precalset_gln)]
        self.h = C.c_void_p()
        cn = getattr(m, "colloc_n", None)
        self.ndof = int(getattr(m, "ndof", 3))
        has_sym = len(getattr(m, "symplane_eid", ())) > 0
        if has_sym and self.ndof in (1, 4):   # [symmetry planes] on a fluid / poroelastic region
            eid = np.ascontiguousarray(m.symplane_eid, dtype=np.int32); st = np.ascontiguousarray(m.symplane_t, dtype=np.float64)
            sc = np.ascontiguousarray(m.symplane_s, dtype=np.float64); k += [eid, st, sc]
            fn = lib().mfb_harpor3d_setup_sym if self.ndof == 4 else lib().mfb_harpot3d_setup_sym
            _check(fn(
                ctx.h, C.c_int(m.n_node), _p(k[0]), C.c_int(m.n_elem), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]),
                C.c_int(m.n_colloc), _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]), _p(k[10]), _p(k[11]), _p(k[12]), _p(k[13]),
                C.c_int(m.n_dof), C.c_double(m.qsi_relative_error), C.c_int(m.qsi_ns_max), C.c_int(len(m.precalset_gln)), _p(k[14]),
                C.c_double(m.geometric_tolerance), C.c_int(len(eid)), _p(eid), _p(sc), _p(st), C.byref(self.h)))
        elif self.ndof == 4:   # poroelastic region: four equations / unknowns per node (col_u = columns of tau, u_k; col_t = columns of Un, t_k)
            _check(lib().mfb_harpor3d_setup(
                ctx.h, C.c_int(m.n_node), _p(k[0]), C.c_int(m.n_elem), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]),
                C.c_int(m.n_colloc), _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]), _p(k[10]), _p(k[11]), _p(k[12]), _p(k[13]),
                C.c_int(m.n_dof), C.c_double(m.qsi_relative_error), C.c_int(m.qsi_ns_max), C.c_int(len(m.precalset_gln)), _p(k[14]),
                C.c_double(m.geometric_tolerance), C.byref(self.h)))
        elif self.ndof == 1:   # inviscid fluid region: one equation / unknown per node (col_u = column of p, col_t = column of Un)
            _check(lib().mfb_harpot3d_setup(
                ctx.h, C.c_int(m.n_node), _p(k[0]), C.c_int(m.n_elem), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]),
                C.c_int(m.n_colloc), _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]), _p(k[10]), _p(k[11]), _p(k[12]), _p(k[13]),
                C.c_int(m.n_dof), C.c_double(m.qsi_relative_error), C.c_int(m.qsi_ns_max), C.c_int(len(m.precalset_gln)), _p(k[14]),
                C.c_double(m.geometric_tolerance), C.byref(self.h)))
        elif len(getattr(m, "symplane_eid", ())) > 0:   # [symmetry planes]: every element is integrated with its mirror images
            if cn is not None:
                cn = np.ascontiguousarray(cn, dtype=np.float64); k.append(cn)
            eid = np.ascontiguousarray(m.symplane_eid, dtype=np.int32); st = np.ascontiguousarray(m.symplane_t, dtype=np.float64); k += [eid, st]
            _check(lib().mfb_harela3d_setup_sym(
                ctx.h, C.c_int(m.n_node), _p(k[0]), C.c_int(m.n_elem), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]),
                C.c_int(m.n_colloc), _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]), _p(cn) if cn is not None else None,
                _p(k[10]), _p(k[11]), _p(k[12]), _p(k[13]),
                C.c_int(m.n_dof), C.c_double(m.qsi_relative_error), C.c_int(m.qsi_ns_max), C.c_int(len(m.precalset_gln)), _p(k[14]),
                C.c_double(m.geometric_tolerance), C.c_int(len(eid)), _p(eid), _p(st), C.byref(self.h)))
        elif cn is None:
            _check(lib().mfb_harela3d_setup(
                ctx.h, C.c_int(m.n_node), _p(k[0]), C.c_int(m.n_elem), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]),
                C.c_int(m.n_colloc), _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]), _p(k[10]), _p(k[11]), _p(k[12]), _p(k[13]),
                C.c_int(m.n_dof), C.c_double(m.qsi_relative_error), C.c_int(m.qsi_ns_max), C.c_int(len(m.precalset_gln)), _p(k[14]),
                C.c_double(m.geometric_tolerance), C.byref(self.h)))
        else:   # hypersingular equation at points off the boundary (interior-point stresses)
            cn = np.ascontiguousarray(cn, dtype=np.float64); k.append(cn)
            _check(lib().mfb_harela3d_setup_hbie(
                ctx.h, C.c_int(m.n_node), _p(k[0]), C.c_int(m.n_elem), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]),
                C.c_int(m.n_colloc), _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]), _p(cn), _p(k[10]), _p(k[11]), _p(k[12]), _p(k[13]),
                C.c_int(m.n_dof), C.c_double(m.qsi_relative_error), C.c_int(m.qsi_ns_max), C.c_int(len(m.precalset_gln)), _p(k[14]),
                C.c_double(m.geometric_tolerance), C.byref(self.h)))
        self._cv = np.ascontiguousarray(m.cvalue, dtype=np.complex128)
        if self.ndof == 3 and np.isin(np.asarray(m.ctype), (2, 3)).any() and hasattr(m, "condition_rows"):   # local-axes conditions: the host's rows, added after every assembly
            rr, cc, vv = m.condition_rows(); k += [rr, cc, vv]
            _check(lib().mfb_set_condition_rows(self.h, C.c_int(len(rr)), _p(rr), _p(cc), _p(vv)))
        if self.ndof == 3 and (np.asarray(m.ctype) == 10).any():      # normal-pressure conditions need node()%n_fn
            nf = np.ascontiguousarray(m.n_fn, dtype=np.float64); k.append(nf)
            _check(lib().mfb_set_node_normals(self.h, _p(nf)))

    def close(self):
        if self.h:
            lib().mfb_problem_free(self.h)
            self.h = C.c_void_p()

    def set_incident(self, u_inc=None, t_inc=None):
        """Incident wave field at the nodes of every element ((sum nn, 3) complex each, element order; element()%incident_c of the reference);
        None clears.  Every later assembly adds hp u_inc - gp t_inc to b (assemble_bem_harela_equation.f90:651-666)."""
        # fluid region: p_inc, Un_inc, one value per element node; poroelastic region: (tau, u_k), (Un, t_k), four per element node
        fn = {1: lib().mfb_harpot3d_set_incident, 4: lib().mfb_harpor3d_set_incident}.get(self.ndof, lib().mfb_harela3d_set_incident)
        if u_inc is None:
            _check(fn(self.h, None, None))
            return
        nd = self.ndof
        u = np.ascontiguousarray(u_inc, dtype=np.complex128); t = np.ascontiguousarray(t_inc, dtype=np.complex128)
        if u.size != nd * int(self.m.elem_ptr[-1]) or t.size != u.size:
            raise ValueError("incident field: one row per element node")
        _check(fn(self.h, _p(u.reshape(-1, nd)), _p(t.reshape(-1, nd))))

    # ---- seam 1: build_lse_mechanics_bem_harela(kf,kr) after A_c=0; b_c=0 ----
    def build_lse_mechanics_bem_harela(self, omega, mat, want_host=True, out=None):
        """out = (A, b): the caller's own A_c, b_c (Fortran-ordered complex128), overwritten -- what the Fortran host passes."""
        n = self.m.n_dof
        A = b = None
        if want_host and out is not None:
            A, b = out
            assert A.flags.f_contiguous and A.dtype == np.complex128 and A.shape == (n, n) and b.dtype == np.complex128
        elif want_host:
            A = np.zeros((n, n), dtype=np.complex128, order="F")
            b = np.zeros(n, dtype=np.complex128)
        _check(lib().mfb_harela3d_assemble(self.h, C.c_double(omega), _p(_z(mat.lam)), _p(_z(mat.mu)), C.c_double(mat.rho),
                                           _p(_z(mat.nu)), _p(self._cv), _p(A) if want_host else None, _p(b) if want_host else None))
        return A, b

    # ---- seam 1 for an inviscid fluid region: build_lse_mechanics_bem_harpot(kf,kr) (src/build_lse_mechanics_bem_harpot.f90:22) ----
    def build_lse_mechanics_bem_harpot(self, omega, fluid, want_host=True):
        n = self.m.n_dof
        A = b = None
        if want_host:
            A = np.zeros((n, n), dtype=np.complex128, order="F")
            b = np.zeros(n, dtype=np.complex128)
        _check(lib().mfb_harpot3d_assemble(self.h, C.c_double(omega), C.c_double(fluid.rho), _p(_z(fluid.c)), _p(self._cv),
                                           _p(A) if want_host else None, _p(b) if want_host else None))
        return A, b

    def solve_frequency_fluid(self, omega, fluid, host=True):
        """One iteration of the frequency loop of a fluid region on the device: assemble, zgetrf, zgetrs."""
        x = np.zeros(self.m.n_dof, dtype=np.complex128) if host else None
        _check(lib().mfb_harpot3d_solve_frequency(self.h, C.c_double(omega), C.c_double(fluid.rho), _p(_z(fluid.c)),
                                                  _p(self._cv) if host else None, _p(x) if host else None))
        return x

    # ---- seam 1 for a poroelastic region: build_lse_mechanics_bem_harpor(kf,kr) (src/build_lse_mechanics_bem_harpor.f90:23) ----
    def _por_args(self, omega, po):
        return (C.c_double(omega), _p(_z(po.lam)), _p(_z(po.mu)), C.c_double(po.rho1), C.c_double(po.rho2), C.c_double(po.rhoa), _p(_z(po.R)), _p(_z(po.Q)),
                C.c_double(po.b))

    def build_lse_mechanics_bem_harpor(self, omega, poro, want_host=True):
        n = self.m.n_dof
        A = b = None
        if want_host:
            A = np.zeros((n, n), dtype=np.complex128, order="F")
            b = np.zeros(n, dtype=np.complex128)
        _check(lib().mfb_harpor3d_assemble(self.h, *self._por_args(omega, poro), _p(self._cv), _p(A) if want_host else None, _p(b) if want_host else None))
        return A, b

    def solve_frequency_poro(self, omega, poro, host=True):
        x = np.zeros(self.m.n_dof, dtype=np.complex128) if host else None
        _check(lib().mfb_harpor3d_solve_frequency(self.h, *self._por_args(omega, poro), _p(self._cv) if host else None, _p(x) if host else None))
        return x

    # ---- seam 2: solve_lse_c(n_dof,A,ipiv,...,n_rhs,b,factorize,scaling=F,condition=F,refine=F) ----
    def solve_lse_c(self, A=None, b=None, factorize=True, want_ipiv=False):
        """A=None, b=None: factorise/solve the device-resident system of the last assembly; returns x (and ipiv).
        With host A (n x n) and b (n or n x nrhs): LAPACK zgesv semantics, A is overwritten by the LU factors."""
        n = self.m.n_dof
        ipiv = np.zeros(n, dtype=np.int32)
        if A is not None and not (A.flags.f_contiguous and A.dtype == np.complex128):
            raise ValueError("A must be a Fortran-ordered complex128 array (it is overwritten by the LU factors)")
        if b is None:
            raise ValueError("pass the right-hand side b (use solve_frequency for the fully device-resident path)")
        bb = np.asfortranarray(b, dtype=np.complex128).reshape(n, -1, order="F").copy(order="F")
        _check(lib().mfb_zsolve(self.h, C.c_int(n), _p(A) if A is not None else None, C.c_int(n), _p(ipiv), _p(bb),
                                C.c_int(bb.shape[1]), C.c_int(int(factorize))))
        x = bb[:, 0] if np.ndim(b) == 1 else bb
        return (x, ipiv) if want_ipiv else x

    def solve_lse_c_ex(self, A=None, b=None, factorize=True, scaling=False, condition=False, refine=False, equed="N", r=None, c=None):
        """solve_lse_c with its optional stages (mfb_zsolve_ex): returns (x, info) with info = {equed, r, c, rcond, ferr, berr, ipiv}."""
        n = self.m.n_dof
        ipiv = np.zeros(n, dtype=np.int32)
        if A is not None and not (A.flags.f_contiguous and A.dtype == np.complex128):
            raise ValueError("A must be a Fortran-ordered complex128 array (it is overwritten by the LU factors)")
        bb = np.asfortranarray(b, dtype=np.complex128).reshape(n, -1, order="F").copy(order="F")
        nrhs = bb.shape[1]
        eq = C.create_string_buffer(equed.encode()[:1], 2)
        rr = np.ones(n) if r is None else np.ascontiguousarray(r, dtype=np.float64).copy()
        cc = np.ones(n) if c is None else np.ascontiguousarray(c, dtype=np.float64).copy()
        rcond = C.c_double(0.0); ferr = np.zeros(nrhs); berr = np.zeros(nrhs)
        _check(lib().mfb_zsolve_ex(self.h, C.c_int(n), _p(A) if A is not None else None, C.c_int(n), _p(ipiv), _p(bb), C.c_int(nrhs), C.c_int(int(factorize)),
                                   C.c_int(int(scaling)), C.c_int(int(condition)), C.c_int(int(refine)), eq, _p(rr), _p(cc), C.byref(rcond), _p(ferr), _p(berr)))
        x = bb[:, 0] if np.ndim(b) == 1 else bb
        return x, {"equed": eq.value.decode() or "N", "r": rr, "c": cc, "rcond": rcond.value, "ferr": ferr, "berr": berr, "ipiv": ipiv}

    # ---- one iteration of the frequency loop, device resident ----
    def solve_frequency(self, omega, mat, host=True):
        """host=True: prescribed values go up from host memory and the solution comes back to host memory (the call the
        Fortran loop body would make).  host=False: everything stays on the device (get_solution() fetches x)."""
        x = np.zeros(self.m.n_dof, dtype=np.complex128) if host else None
        _check(lib().mfb_harela3d_solve_frequency(self.h, C.c_double(omega), _p(_z(mat.lam)), _p(_z(mat.mu)), C.c_double(mat.rho),
                                                  _p(_z(mat.nu)), _p(self._cv) if host else None, _p(x) if host else None))
        return x

    def sweep(self, omegas, mat, rank=0, nranks=1, unique_id=None):
        """The whole frequency loop in one C-ABI call (mfb_harela3d_sweep): frequencies kf = rank, rank + nranks, ... on this GPU, NCCL all-reduce
        of the solutions inside the library; returns (X [n_freq, n_dof], info [n_freq]) on every rank."""
        om = np.ascontiguousarray(omegas, dtype=np.float64)
        X = np.zeros((len(om), self.m.n_dof), dtype=np.complex128)
        info = np.zeros(len(om), dtype=np.int32)
        uid = (C.c_char * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        _check(lib().mfb_harela3d_sweep(self.h, C.c_int(len(om)), _p(om), _p(_z(mat.lam)), _p(_z(mat.mu)), C.c_double(mat.rho), _p(_z(mat.nu)), _p(self._cv),
                                        C.c_int(rank), C.c_int(nranks), uid, _p(X), _p(info)))
        return X, info

    def build_lse_accumulate(self, omega, mat, A, b):
        """seam 1 with the reference's `+=` on the caller's arrays (mfb_harela3d_assemble_acc, accumulate = 1): A (Fortran order, ld = A.shape[0]) and b
        are ADDED to."""
        assert A.flags.f_contiguous and A.dtype == np.complex128 and b.dtype == np.complex128
        _check(lib().mfb_harela3d_assemble_acc(self.h, C.c_double(omega), _p(_z(mat.lam)), _p(_z(mat.mu)), C.c_double(mat.rho), _p(_z(mat.nu)), _p(self._cv),
                                               _p(A), C.c_int(A.shape[0]), _p(b), C.c_int(1)))

    def get_solution(self):
        x = np.zeros(self.m.n_dof, dtype=np.complex128)
        _check(lib().mfb_get_solution(self.h, _p(x)))
        return x

    def residual(self, x):
        """(berr, rel_resid) of x against the assembled, unfactorised device-resident system."""
        xx = np.ascontiguousarray(x, dtype=np.complex128)
        a, b = C.c_double(), C.c_double()
        _check(lib().mfb_residual(self.h, _p(xx), C.byref(a), C.byref(b)))
        return a.value, b.value

    def residual_vector(self, x):
        """r = A x - b of the assembled, unfactorised device-resident system (host order)."""
        xx = np.ascontiguousarray(x, dtype=np.complex128)
        r = np.zeros(self.m.n_dof, dtype=np.complex128)
        _check(lib().mfb_residual_vector(self.h, _p(xx), _p(r)))
        return r

    def get_entries(self, rows, cols):
        r = np.ascontiguousarray(rows, dtype=np.int32); c = np.ascontiguousarray(cols, dtype=np.int32)
        out = np.zeros(len(r), dtype=np.complex128)
        _check(lib().mfb_get_entries(self.h, C.c_int(len(r)), _p(r), _p(c), _p(out)))
        return out

    # ---- static elasticity: build_lse_mechanics_bem_staela(kr) / solve_lse_r(...) ----
    def build_lse_mechanics_bem_staela(self, mat, want_host=True):
        """A_r, b_r of the static analysis (real); mat.mu_r, mat.nu_r = region%property_r(2,3)."""
        n = self.m.n_dof
        A = np.zeros((n, n), dtype=np.float64, order="F") if want_host else None
        b = np.zeros(n, dtype=np.float64) if want_host else None
        cv = np.ascontiguousarray(self._cv.real, dtype=np.float64)
        _check(lib().mfb_staela3d_assemble(self.h, C.c_double(mat.mu_r), C.c_double(mat.nu_r), _p(cv),
                                           _p(A) if want_host else None, _p(b) if want_host else None))
        return A, b

    def solve_lse_r(self, A=None, b=None, factorize=True, want_ipiv=False):
        """LAPACK dgesv semantics with host A (overwritten by the LU factors) and b; A=None uses the resident static system."""
        n = self.m.n_dof
        ipiv = np.zeros(n, dtype=np.int32)
        if A is not None and not (A.flags.f_contiguous and A.dtype == np.float64):
            raise ValueError("A must be a Fortran-ordered float64 array (it is overwritten by the LU factors)")
        if b is None:
            raise ValueError("pass the right-hand side b (use solve_static for the fully device-resident path)")
        bb = np.asfortranarray(b, dtype=np.float64).reshape(n, -1, order="F").copy(order="F")
        _check(lib().mfb_dsolve(self.h, C.c_int(n), _p(A) if A is not None else None, C.c_int(n), _p(ipiv), _p(bb),
                                C.c_int(bb.shape[1]), C.c_int(int(factorize))))
        x = bb[:, 0] if np.ndim(b) == 1 else bb
        return (x, ipiv) if want_ipiv else x

    def solve_static(self, mat):
        x = np.zeros(self.m.n_dof, dtype=np.float64)
        cv = np.ascontiguousarray(self._cv.real, dtype=np.float64)
        _check(lib().mfb_staela3d_solve(self.h, C.c_double(mat.mu_r), C.c_double(mat.nu_r), _p(cv), _p(x)))
        return x

    # ---- one frequency over several GPUs (mfb_dist_*; collective over the ranks) ----
    def dist_init(self, rank, nranks, unique_id, nb=0):
        _check(lib().mfb_dist_init(self.h, C.c_int(rank), C.c_int(nranks), C.c_char_p(unique_id), C.c_int(nb)))

    def dist_init_loopback(self, nranks, nb=0):
        """Test mode: `nranks` virtual ranks on this GPU (the collectives become device copies)."""
        _check(lib().mfb_dist_init_loopback(self.h, C.c_int(nranks), C.c_int(nb)))

    def dist_info(self):
        r, n, ncl = C.c_int(), C.c_int(), C.c_int()
        _check(lib().mfb_dist_info(self.h, C.byref(r), C.byref(n), None, C.byref(ncl)))
        rb = np.zeros(n.value + 1, dtype=np.int32)
        _check(lib().mfb_dist_info(self.h, None, None, _p(rb), None))
        return {"rank": r.value, "nranks": n.value, "row_bounds": rb, "n_local_cols": ncl.value}

    def dist_solve_frequency(self, omega, mat, host=True):
        x = np.zeros(self.m.n_dof, dtype=np.complex128)
        _check(lib().mfb_dist_solve_frequency(self.h, C.c_double(omega), _p(_z(mat.lam)), _p(_z(mat.mu)), C.c_double(mat.rho),
                                              _p(_z(mat.nu)), _p(self._cv) if host else None, _p(x)))
        return x

    def dist_zsolve(self, A, b, want_ipiv=False):
        n = self.m.n_dof
        A = np.asfortranarray(A, dtype=np.complex128); b = np.ascontiguousarray(b, dtype=np.complex128)
        x = np.zeros(n, dtype=np.complex128); ipiv = np.zeros(n, dtype=np.int32)
        _check(lib().mfb_dist_zsolve(self.h, C.c_int(n), _p(A), C.c_int(n), _p(b), _p(x), _p(ipiv)))
        return (x, ipiv) if want_ipiv else x

    def stats(self):
        s = np.zeros(STAT_COUNT)
        _check(lib().mfb_get_stats(self.h, _p(s)))
        return {k: float(s[i]) for k, i in STAT.items()}

    def plan_modes(self, colloc, elem):
        c = np.ascontiguousarray(colloc, dtype=np.int32); e = np.ascontiguousarray(elem, dtype=np.int32)
        out = np.zeros(len(c), dtype=np.int32)
        _check(lib().mfb_plan_modes(self.h, C.c_int(len(c)), _p(c), _p(e), _p(out)))
        return out


class ProblemLanes:
    """Several frequencies of ONE mesh in flight on one GPU: `n_lanes` independent (mfb_ctx, mfb_problem) pairs, each with its own stream, system matrix,
    LU workspace and kernel-launch state, driven by one host thread each (the C-ABI calls release the GIL).  A small system cannot fill a B200 -- at the
    1386 DOF of the reference's ME-TH-EL-001 input a frequency is a chain of ~1400 dependent pivot steps of a few microseconds and kernels of a few
    CTAs -- but the frequencies of a sweep are independent (src/multifebe.f90:107-124), so the lanes overlap each other's latencies.  The kernel parameters
    travel as launch arguments and every context owns its launch state, which is what makes concurrent assemblies at different frequencies safe."""

    def __init__(self, model, device=0, n_lanes=8):
        self.m = model
        self.ctxs = [Context(device) for _ in range(n_lanes)]
        self.prs = [Problem(c, model) for c in self.ctxs]

    def run(self, omegas, mat, host=True):
        """solutions X[len(omegas), n_dof] (frequency kf on lane kf % n_lanes, in sweep order on each lane)"""
        import threading
        X = np.zeros((len(omegas), self.m.n_dof), dtype=np.complex128)
        errs = []

        def work(q):
            try:
                pr = self.prs[q]
                for kf in range(q, len(omegas), len(self.prs)):
                    X[kf] = pr.solve_frequency(float(omegas[kf]), mat, host=True) if host else (pr.solve_frequency(float(omegas[kf]), mat, host=False), pr.get_solution())[1]
            except Exception as e:      # re-raised on the calling thread
                errs.append(e)
        th = [threading.Thread(target=work, args=(q,)) for q in range(len(self.prs))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        return X

    def close(self):
        for pr in self.prs:
            pr.close()
        for c in self.ctxs:
            c.close()


class InternalPoints:
    """Displacements at points inside the region from the boundary solution (the displacement part of the reference's
    calculate_internal_points_mechanics_bem_harela, src/calculate_internal_points_mechanics_bem_harela.f90:160-178): a second
    problem on the same mesh whose collocation points are the interior points (host.InternalPointsModel)."""

    def __init__(self, ctx, model, points):
        from .host import InternalPointsModel
        self.ctx, self.base, self._points = ctx, model, points
        self.ipm = InternalPointsModel(model, points)
        self.pr = Problem(ctx, self.ipm)
        self.ipm_s = self.pr_s = None      # the hypersingular problem of the stresses, built on first use

    def close(self):
        self.pr.close()
        if self.pr_s is not None:
            self.pr_s.close()

    def stresses(self, omega, mat, x):
        """sigma (n_points, 3, 3) complex: sigma[p, l, k] = component l of the traction on the plane with normal e_k at point p
        (internalpoint%value_c(l,k), src/calculate_internal_points_mechanics_bem_harela.f90:404-470), harmonic analysis."""
        from .host import InternalPointsModel
        if self.pr_s is None:
            self.ipm_s = InternalPointsModel(self.base, self._points, stress=True)
            self.pr_s = Problem(self.ctx, self.ipm_s)
        self.pr_s.build_lse_mechanics_bem_harela(omega, mat, want_host=False)
        return self._sigma(x)

    def _sigma(self, x):
        xa = np.zeros(self.ipm_s.n_dof, dtype=np.complex128); xa[:self.base.n_dof] = x
        r = self.pr_s.residual_vector(xa)
        t = -r[self.base.n_dof:].reshape(-1, 3, 3)      # [point][plane k][component l]
        return np.transpose(t, (0, 2, 1))

    def stresses_static(self, mat, x):
        """The same for the static analysis (fbem_bem_staela3d_hbie_*): real sigma (n_points, 3, 3)."""
        from .host import InternalPointsModel
        if self.pr_s is None:
            self.ipm_s = InternalPointsModel(self.base, self._points, stress=True)
            self.pr_s = Problem(self.ctx, self.ipm_s)
        self.pr_s.build_lse_mechanics_bem_staela(mat, want_host=False)
        return self._sigma(np.asarray(x, dtype=np.complex128)).real

    def _u(self, x):
        xa = np.zeros(self.ipm.n_dof, dtype=np.complex128); xa[:self.base.n_dof] = x
        r = self.pr.residual_vector(xa)
        return -r[self.base.n_dof:].reshape(-1, 3)

    def displacements(self, omega, mat, x):
        """u (n_points, 3) complex at frequency omega for the boundary solution vector x of that frequency."""
        self.pr.build_lse_mechanics_bem_harela(omega, mat, want_host=False)
        return self._u(x)

    def pressures(self, omega, fluid, x):
        """p (n_points) complex inside an inviscid fluid region for the boundary solution x (internalpoint%value_c(1,0) of
        src/calculate_internal_points_mechanics_bem_harpot.f90): p(x_ip) = sum_e (g rho omega^2 Un - h p) = -(A x - b) at the interior rows."""
        self.pr.build_lse_mechanics_bem_harpot(omega, fluid, want_host=False)
        xa = np.zeros(self.ipm.n_dof, dtype=np.complex128); xa[:self.base.n_dof] = x
        return -self.pr.residual_vector(xa)[self.base.n_dof:]

    def displacements_static(self, mat, x):
        self.pr.build_lse_mechanics_bem_staela(mat, want_host=False)
        return self._u(np.asarray(x, dtype=np.complex128)).real
