/* mfb.h -- C ABI of the B200-native hot path of MultiFEBE's time-harmonic 3D elastodynamic BEM.
 *
 * The reference (MultiFEBE v2.0.1, Fortran 2003) has no FFI/plugin interface; the two seams this
 * library replaces are Fortran call sites operating on `module problem_variables` globals:
 *
 *   seam 1  subroutine build_lse_mechanics_bem_harela(kf,kr)      src/build_lse_mechanics_bem_harela.f90:22-968
 *           called from build_lse_mechanics_harmonic              src/build_lse_mechanics_harmonic.f90:88
 *           after `A_c=0; b_c=0` (:73-74).  Replaced by mfb_harela3d_setup (once per run: everything that does
 *           not depend on omega) + mfb_harela3d_assemble (once per frequency).
 *   seam 2  subroutine solve_lse_c(n_dof,A,ipiv,...,n_rhs,b,factorize,scaling,condition,refine)
 *                                                                 src/solve_lse_c.f90:25-219 (zgetrf :124, zgetrs :176)
 *           called from src/multifebe.f90:111-112.  Replaced by mfb_zsolve.
 *   loop    do kf=1,n_frequencies (src/multifebe.f90:107-124)     -> mfb_harela3d_solve_frequency (device-resident
 *           assemble + LU + solve of one frequency; the frequency shard of a multi-GPU sweep calls it per owned kf).
 *
 * Conventions: plain C, host pointers, no derived types.  All arrays are flat, 0-based indices, column-major
 * matrices, complex numbers interleaved (re,im) == complex(real64) == `mfb_z`.  Every function returns 0 on success,
 * <0 for an invalid argument / missing device, >0 for a numerical failure (e.g. singular pivot = LAPACK info);
 * mfb_last_error() returns the message (the Fortran shim maps nonzero to fbem_error_message + stop, as the reference
 * does at src/solve_lse_c.f90:129-133).  One caller thread per context; one context per GPU (one process per GPU).
 * There is no CPU fallback: every compute entry point fails with MFB_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef MFB_H
#define MFB_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } mfb_z;
typedef struct mfb_ctx mfb_ctx;
typedef struct mfb_problem mfb_problem;

#define MFB_OK 0
#define MFB_ERR_ARG (-1)
#define MFB_ERR_NO_DEVICE (-2)
#define MFB_ERR_CUDA (-3)
#define MFB_ERR_UNSUPPORTED (-4)

/* element type ids == lib/fbem/src/shape_functions.f90:198-206 */
#define MFB_TRI3 5
#define MFB_TRI6 6
#define MFB_QUAD4 7
#define MFB_QUAD8 8
#define MFB_QUAD9 9

const char* mfb_last_error(void);
int mfb_version(void);

/* Bind a context to CUDA device `device` (LOCAL_RANK under torchrun).  Fails (MFB_ERR_NO_DEVICE) without a GPU. */
int mfb_init(int device, mfb_ctx** ctx);
void mfb_finalize(mfb_ctx* ctx);

/* Once per run.  Replaces the omega-independent part of build_lse_mechanics_bem_harela + the per-element data the
 * reference recomputes every frequency (fbem_bem_element / init_precalculated_datasets, lib/fbem/src/bem_general.f90:163-759;
 * csize, n_phi, bounding ball, src/build_data_of_be_elements.f90:62-110) and builds the quadrature plan
 * (fbem_bem_harela3d_sbie_auto decisions, lib/fbem/src/bem_harela3d.f90:1474-1538, incl. the adaptive-subdivision leaf list
 * of _sbie_ext_adp :1050-1172 and the polar/line-integral data of _sbie_int :1174-1472) and the Mantic free terms' geometry.
 *   node_x[3*n_node]; etype[n_elem] in MFB_TRI3..MFB_QUAD9; elem_ptr[n_elem+1], elem_node[] = element(:)%node (0-based,
 *   isoparametric: geometric == functional nodes, continuous); elem_reversed[n_elem] = region%boundary_reversion;
 *   collocation points in the order of the reference's kb_col/ke_col/kn_col loop (build_lse_mechanics_bem_harela.f90:1118-1136):
 *   colloc_x[3*n_colloc] = x_i_sbie / x_i_sbie_mca, colloc_node = sn_col (its 3 rows receive the equation), colloc_elem /
 *   colloc_kn = owning element and local node (free-term pass :273-747; colloc_elem = -1: a point off the boundary, no
 *   free term), colloc_xi[2*n_colloc] = xi_i_sbie_mca, or (-9,-9) for a nodal SBIE point; row/col_u/col_t/ctype[3*n_node] = node%row(k,1), node%col(k,1), node%col(3+k,1), node%ctype(k,1)
 *   (0-based, -1 = none; ctype 0: u known, 1: t known) from build_auxiliary_variables_mechanics_harmonic.f90:151-198;
 *   settings = [settings] section defaults of src/read_settings.f90:74-216. */
int mfb_harela3d_setup(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                       const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                       const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                       const int* row, const int* col_u, const int* col_t, const int* ctype, int n_dof,
                       double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                       double geometric_tolerance, mfb_problem** problem);
/* The same set-up for the HYPERSINGULAR equation at points off the boundary (interior-point stresses):
 * fbem_bem_harela3d_hbie_auto (lib/fbem/src/bem_harela3d.f90:3632-3697) with its exterior branches _ext_pre :2573-2662 and
 * _ext_adp :3044-3167 (estimator order 7 instead of 5, kernels d*, s*).  colloc_n[3*n_colloc] = unit normal n_i of every
 * collocation point; colloc_elem must be -1 (a point on an element would need fbem_bem_harela3d_hbie_int: unsupported).
 * With three collocation points per interior point (n_i = e_1, e_2, e_3) the rows of -(A x - b) are the traction vectors on
 * the three coordinate planes, i.e. the stress tensor (src/calculate_internal_points_mechanics_bem_harela.f90:404-470). */
int mfb_harela3d_setup_hbie(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                            const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                            const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                            const double* colloc_n, const int* row, const int* col_u, const int* col_t, const int* ctype, int n_dof,
                            double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                            double geometric_tolerance, mfb_problem** problem);
/* The set-up of a region with SYMMETRY PLANES (the reference's [symmetry planes] section, src/read_symmetry_planes.f90:76-283; image loop
 * build_lse_mechanics_bem_harela.f90:1098-1107 with the multipliers of fbem_symmetry_multipliers, lib/fbem/src/symmetry.f90:60-171).
 * Only a half, quarter or octant of the boundary is meshed; every element is integrated 2, 4 or 8 times -- itself and its mirror images,
 * whose contributions take the sign symconf_t(k) on the columns of dof k and land on the columns of the root element.
 *   n_symplanes = 0..3; symplane_eid[i] = axis the plane is normal to (1 = x: plane_n1 / plane_yz, 2 = y: plane_n2 / plane_zx,
 *   3 = z: plane_n3 / plane_xy), ascending, all planes through the origin; symplane_t[3*i+k] = symplane_t(k,i): "symmetry" = -1 on the
 *   normal axis and +1 on the others, "antisymmetry" = the opposite signs.
 *   colloc_n = NULL for the displacement equation, or the unit normals of the hypersingular equation (interior points, as _setup_hbie).
 * As in the reference the far test of an image uses the bounding ball of the ROOT element (the calculation element's bball_centre is not
 * reflected), and a nodal collocation point lying in one or two planes gets its free term from the fan of elements completed with
 * their mirror images (:496-555).  mfb_plan_modes addresses image ks of element r as element ks*n_elem + r (ks in the step order of
 * fbem_symmetry_multipliers: 1 = plane 1, 2 = planes 1+2, 3 = plane 2, 4 = plane 3, 5 = 1+3, 6 = 1+2+3, 7 = 2+3). */
int mfb_harela3d_setup_sym(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                           const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                           const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                           const double* colloc_n, const int* row, const int* col_u, const int* col_t, const int* ctype, int n_dof,
                           double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                           double geometric_tolerance, int n_symplanes, const int* symplane_eid, const double* symplane_t,
                           mfb_problem** problem);
/* Symmetry planes for an inviscid fluid and for a poroelastic region (arguments of mfb_harpot3d_setup / mfb_harpor3d_setup, then the planes as in
 * mfb_harela3d_setup_sym plus symplane_s[i] = symplane_s(i), the multiplier of scalar variables: +1 symmetry, -1 antisymmetry): image loops of
 * src/build_lse_mechanics_bem_harpot.f90:790-800 (h, g times symconf_s) and _harpor.f90:855-865 (dof 0 times symconf_s, dofs 1..3 times symconf_t). */
int mfb_harpot3d_setup_sym(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                           const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                           const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                           const int* row, const int* col_p, const int* col_un, const int* ctype, int n_dof,
                           double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                           double geometric_tolerance, int n_symplanes, const int* symplane_eid, const double* symplane_s,
                           const double* symplane_t, mfb_problem** problem);
int mfb_harpor3d_setup_sym(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                           const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                           const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                           const int* row, const int* col_p, const int* col_s, const int* ctype, int n_dof,
                           double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                           double geometric_tolerance, int n_symplanes, const int* symplane_eid, const double* symplane_s,
                           const double* symplane_t, mfb_problem** problem);
void mfb_problem_free(mfb_problem* problem);

/* Once per frequency.  == `A_c=0; b_c=0` + build_lse_mechanics_bem_harela(kf,kr) for one elastic BE region with
 * ordinary `be` boundaries: integrates every (collocation point, element) pair, adds the free terms and scatters per
 * assemble_bem_harela_equation.f90:78-113 into the device-resident system.  lambda/mu/nu = region%property_c(6,2,3),
 * rho = property_r(1); cvalue[3*n_node] = node%cvalue_c(k,1,1).  If A (n_dof x n_dof col-major, ld = n_dof) / b are non-NULL
 * the assembled system is also copied to the host (the reference's A_c, b_c). */
int mfb_harela3d_assemble(mfb_problem* problem, double omega, const mfb_z* lambda, const mfb_z* mu, double rho,
                          const mfb_z* nu, const mfb_z* cvalue, mfb_z* A, mfb_z* b);

/* Boundary conditions ctype = 2 / 3 of an elastic region (local axes: u.l = U / t.l = T; src/read_conditions_bem_boundaries_mechanics_harmonic.f90:124-128):
 * accepted by the set-up calls in ctype[]; for such a dof BOTH u_k and t_k are unknowns (col_u AND col_t must be given) and the scatter is
 * A(row, col_u) += h, A(row, col_t) -= g (assemble_bem_harela_equation.f90:107-112).  The three condition rows node%row(k,0) of such a node belong to the host
 * (src/build_lse_mechanics_harmonic.f90:204-258): with the two-seam path it writes them into its A_c, b_c after mfb_harela3d_assemble as it does today; for the
 * resident path (mfb_harela3d_solve_frequency, mfb_staela3d_solve) it hands them over once with mfb_set_condition_rows -- entries (rows[i], cols[i], values[i]),
 * cols = -1 meaning the right-hand side -- and the library adds them after every assembly.  n = 0 clears them. */
int mfb_set_condition_rows(mfb_problem* problem, int n, const int* rows, const int* cols, const mfb_z* values);

/* Boundary condition ctype = 10 of an elastic region ("normal pressure known", assemble_bem_harela_equation.f90:97-106): accepted by the set-up calls in
 * ctype[]; cvalue then holds the pressure p and the library forms t_k = p n_fn(k) (negated on a reversed boundary) with the nodal unit normals
 * node(sn)%n_fn given here, n_fn[3 * n_node] (src/build_data_at_functional_nodes.f90:355-400).  Once, before the first assembly; a no-op for models without
 * such conditions.  The unknown of such a dof is u_k in column col_u, as for ctype 1. */
int mfb_set_node_normals(mfb_problem* problem, const double* n_fn);

/* Incident wave field of the region (the reference's region%n_incidentfields > 0 path): u_inc, t_inc at the nodes of every element as held in
 * element(se)%incident_c(1:3,kn,1) / (4:6,kn,1) (src/build_lse_mechanics_bem_harela.f90:291-304), index [(elem_ptr[e] + j) * 3 + k].  While set, every
 * pair and every free term adds hp u_inc - gp t_inc to b (src/assemble_bem_harela_equation.f90:651-666).  Both NULL clears it.  The field depends on
 * omega: call it before the assembly / solve_frequency of each frequency.  Displacement equation of elastic regions, harmonic analysis. */
int mfb_harela3d_set_incident(mfb_problem* problem, const mfb_z* u_inc, const mfb_z* t_inc);
/* The same for an inviscid fluid region (src/assemble_bem_harpot_equation.f90:471-481): p_inc, Un_inc at the nodes of every element, index [elem_ptr[e] + j]. */
int mfb_harpot3d_set_incident(mfb_problem* problem, const mfb_z* p_inc, const mfb_z* un_inc);
/* and for a poroelastic region (src/assemble_bem_harpor_equation.f90:1277-1289): (tau, u_k)_inc and (Un, t_k)_inc, index [(elem_ptr[e] + j) * 4 + k]. */
int mfb_harpor3d_set_incident(mfb_problem* problem, const mfb_z* u_inc, const mfb_z* t_inc);

/* == solve_lse_c(n_dof,A,ipiv,..,n_rhs,b,factorize,scaling=F,condition=F,refine=F): in-place LU with partial pivoting
 * (zgetrf) + triangular solves (zgetrs).  A == NULL: use the device-resident system of the last mfb_harela3d_assemble (and
 * keep the factors on the device); otherwise A (lda x n, host) is uploaded, overwritten by the LU factors on return.
 * ipiv (1-based, LAPACK convention) may be NULL.  b (ldb = n, nrhs columns; NULL with A == NULL: device-resident rhs) is
 * overwritten by the solution.  factorize = 0 re-uses the factors of the previous call (multifebe.f90:119-120).
 * Returns >0 (= LAPACK info) on an exactly singular pivot. */
int mfb_zsolve(mfb_problem* problem, int n, mfb_z* A, int lda, int* ipiv, mfb_z* b, int nrhs, int factorize);

/* == solve_lse_c with ALL its options (src/solve_lse_c.f90:25-46): scaling = zgeequ + zlaqge (:81-117), condition = zgecon (:140-165),
 * refine = zgerfs (:191-206), computed on the device around the LU of mfb_zsolve (the reference's unfactorised copy `Ao` lives in device memory).
 * A / lda / ipiv / factorize as in mfb_zsolve; b (host, ldb = n, nrhs columns) is overwritten by the solution.  equed (one character, 'N' 'R' 'C'
 * 'B'), r[n], c[n]: the reference's arguments of the same names -- written when factorize && scaling, read when !factorize && scaling.
 * rcond (written when condition && factorize), ferr[nrhs], berr[nrhs] (written when refine) may be NULL.  With all three flags 0 this is mfb_zsolve. */
int mfb_zsolve_ex(mfb_problem* problem, int n, mfb_z* A, int lda, int* ipiv, mfb_z* b, int nrhs, int factorize, int scaling, int condition,
                  int refine, char* equed, double* r, double* c, double* rcond, double* ferr, double* berr);

/* mfb_harela3d_assemble with the seam's `+=` semantics on the HOST side (src/build_lse_mechanics_harmonic.f90:73-95: A_c and b_c are zeroed once and every
 * region ADDS its rows): accumulate != 0 adds this region's assembled system to the caller's A (lda >= n_dof) and b instead of overwriting them, so that a
 * host with several regions (one mfb_problem each, all numbered in the same global n_dof) or with finite-element rows written before the call keeps them.
 * accumulate = 0 and lda = n_dof is mfb_harela3d_assemble. */
int mfb_harela3d_assemble_acc(mfb_problem* problem, double omega, const mfb_z* lambda, const mfb_z* mu, double rho, const mfb_z* nu,
                              const mfb_z* cvalue, mfb_z* A, int lda, mfb_z* b, int accumulate);

/* The whole frequency loop (src/multifebe.f90:107-124) as one call, sharded over the ranks of a job: rank q assembles, factorises and
 * solves the frequencies q, q + nranks, ... on its own GPU; one NCCL all-reduce (inside the library) gathers the solutions, and EVERY rank returns
 * X[n_dof x n_freq] (column kf = solution of omega[kf]).  nranks = 1 uses no NCCL; otherwise nccl_id128 is the id rank 0 made with mfb_dist_unique_id.
 * info[n_freq] (may be NULL): LAPACK info of a singular pivot per frequency (that column of X is zero). */
int mfb_harela3d_sweep(mfb_problem* problem, int n_freq, const double* omega, const mfb_z* lambda, const mfb_z* mu, double rho, const mfb_z* nu,
                       const mfb_z* cvalue, int rank, int nranks, const char* nccl_id128, mfb_z* X, int* info);

/* One iteration of the frequency loop (src/multifebe.f90:107-124) kept on the device: assemble, factorise, solve;
 * x[n_dof] receives the solution (what assign_solution_mechanics_harmonic.f90:192-205 reads from b_c); x == NULL keeps it on
 * the device (mfb_get_solution).  cvalue == NULL (here and in mfb_harela3d_assemble) re-uses the prescribed values of the
 * previous call (boundary conditions do not change along a frequency sweep). */
int mfb_harela3d_solve_frequency(mfb_problem* problem, double omega, const mfb_z* lambda, const mfb_z* mu, double rho,
                                 const mfb_z* nu, const mfb_z* cvalue, mfb_z* x);

/* Solution of the last mfb_harela3d_solve_frequency / device-resident mfb_zsolve (x[n_dof]). */
int mfb_get_solution(mfb_problem* problem, mfb_z* x);

/* Diagnostics on the assembled, not yet factorised, device-resident system (the analogue of the reference's optional
 * zgerfs report, src/solve_lse_c.f90:191-206): componentwise backward error berr = max_i |Ax-b|_i / (|A||x|+|b|)_i and
 * max|Ax-b| / max(|A||x|+|b|) for a candidate solution x; selected entries A(rows[i], cols[i]) (0-based). */
int mfb_residual(mfb_problem* problem, const mfb_z* x, double* berr, double* rel_resid);
/* r[n_dof] = A x - b of the same system.  Interior points (displacement part): a problem whose collocation
 * points lie inside the region (colloc_elem = -1 marks them: no free term; their rows belong to dummy nodes that no element
 * uses) assembles Somigliana's identity, so u(x_ip) = -(A x - b) at those rows with x = the boundary solution -- the values
 * the reference stores in internalpoint%value_c(k,0) (src/calculate_internal_points_mechanics_bem_harela.f90:160-178). */
int mfb_residual_vector(mfb_problem* problem, const mfb_z* x, mfb_z* r);
int mfb_get_entries(mfb_problem* problem, int n, const int* rows, const int* cols, mfb_z* out);

/* Plan statistics and per-phase device timings (CUDA events) of the last call; see mfb_stat_id. */
enum mfb_stat_id {
  MFB_STAT_PAIRS_REGULAR = 0, MFB_STAT_POINTS_REGULAR = 1, MFB_STAT_PAIRS_ADAPTIVE = 2, MFB_STAT_LEAVES = 3,
  MFB_STAT_POINTS_ADAPTIVE = 4, MFB_STAT_PAIRS_SINGULAR = 5, MFB_STAT_POINTS_SINGULAR = 6, MFB_STAT_NEAR_PAIRS = 7,
  MFB_STAT_MS_ZERO = 8, MFB_STAT_MS_REGULAR = 9, MFB_STAT_MS_ADAPTIVE = 10, MFB_STAT_MS_SINGULAR = 11,
  MFB_STAT_MS_FREETERM = 12, MFB_STAT_MS_LU = 13, MFB_STAT_MS_SOLVE = 14, MFB_STAT_MS_GEMM = 15, MFB_STAT_MS_PANEL = 16,
  MFB_STAT_LAUNCHES = 17, MFB_STAT_MS_SETUP_HOST = 18, MFB_STAT_MS_ASSEMBLE = 19, MFB_STAT_FLOPS_REGULAR = 20,
  MFB_STAT_MS_TRSM = 21, MFB_STAT_MS_SWAP = 22, MFB_STAT_LU_LAUNCHES = 23, MFB_STAT_GEMM_LAUNCHES = 24,
  MFB_STAT_GEMM_FLOPS = 25, MFB_STAT_GEMM_EXEC_FLOPS = 26,
  /* single-frequency multi-GPU mode (mfb_dist_*): redistribution row slabs -> column owners, distributed LU, back substitution, whole call */
  MFB_STAT_MS_REDIST = 27, MFB_STAT_MS_DIST_LU = 28, MFB_STAT_MS_DIST_SOLVE = 29, MFB_STAT_MS_DIST_TOTAL = 30, MFB_STAT_COUNT = 32
};
int mfb_get_stats(mfb_problem* problem, double* stats /* MFB_STAT_COUNT doubles */);

/* Per-pair integration mode chosen by the plan (for parity tests of the discrete decisions):
 * 2..30 = regular rule gln of the precalculated set, 100 = adaptive (Telles + subdivision), 200 = singular. */
int mfb_plan_modes(mfb_problem* problem, int n_pairs, const int* colloc, const int* elem, int* mode);

/* CUDA-event marks on the context's own stream (the stream every kernel of this library is launched on), so that a caller
 * can time a region on the device: mark(slot) records, elapsed(slot0, slot1) synchronises on slot1 and returns milliseconds. */
int mfb_stream_mark(mfb_ctx* ctx, int slot /* 0..7 */);
int mfb_stream_elapsed(mfb_ctx* ctx, int slot0, int slot1, double* ms);

/* Micro-benchmarks used to measure the roofline denominators on the box (FP64 FMA pipe, FP64 tensor (DMMA) pipe,
 * device copy bandwidth): returns TFLOP/s or GB/s. */
int mfb_measure_peaks(mfb_ctx* ctx, double* dfma_tflops, double* dmma_tflops, double* copy_gbs);

/* Standalone complex GEMM used by the LU trailing update (C -= A*B on planar re/im storage); exposed for tests/bench.
 * C (m x n), A (m x k), B (k x n), all host col-major interleaved complex. */
int mfb_zgemm_minus(mfb_ctx* ctx, int m, int n, int k, const mfb_z* A, int lda, const mfb_z* B, int ldb, mfb_z* C, int ldc,
                    double* ms);

/* ---- Static 3D elasticity (BASELINE config 2) ------------------------------------------------
 * The same mfb_problem serves the static analysis: mesh, collocation points, DOF maps and the quadrature plan do not depend on
 * the kernel (fbem_bem_staela3d_sbie_auto takes the decisions of its harmonic twin, lib/fbem/src/bem_staela3d.f90:524-589).
 *   mfb_staela3d_assemble  == `A_r=0; b_r=0` + build_lse_mechanics_bem_staela(kr) (src/build_lse_mechanics_static.f90:60-66,
 *                          src/build_lse_mechanics_bem_staela.f90) with the Kelvin kernels of fbem_bem_staela3d_sbie_ext_pre /
 *                          _ext_adp / _int (bem_staela3d.f90:592-1381) and the scatter of assemble_bem_staela_equation.f90.
 *                          mu, nu = region%property_r(2,3); cvalue[3*n_node] = node%cvalue_r(k,1,1); A (n_dof x n_dof col-major) /
 *                          b real, or NULL to keep the system on the device only.  Real arithmetic, real plane only.
 *   mfb_dsolve             == solve_lse_r (src/solve_lse_r.f90:25-232, dgetrf :137 + dgetrs :189), conventions of mfb_zsolve.
 *   mfb_staela3d_solve     assemble + factorise + solve on the device, x[n_dof] = solution (host column order). */
int mfb_staela3d_assemble(mfb_problem* problem, double mu, double nu, const double* cvalue, double* A, double* b);
int mfb_dsolve(mfb_problem* problem, int n, double* A, int lda, int* ipiv, double* b, int nrhs, int factorize);
int mfb_staela3d_solve(mfb_problem* problem, double mu, double nu, const double* cvalue, double* x);

/* ---- Inviscid fluid (acoustic) BE region (first brick) ----------------------------------------
 * One fluid region with ordinary `be` boundaries: scalar wave propagation, ONE equation and ONE unknown per node.
 *   mfb_harpot3d_setup     the arguments of mfb_harela3d_setup with one entry per node: row[n_node] = node%row(1,1), col_p[n_node] =
 *                          node%col(1,1) (column of p where Un is prescribed), col_un[n_node] = node%col(2,1) (column of Un where p is
 *                          prescribed), ctype[n_node] = node%ctype(1,1): 0 = p known, 1 = Un known (ctype 2, 3 -- impedance / radiation
 *                          conditions of assemble_bem_harpot_equation.f90:97-110 -- are not built: MFB_ERR_UNSUPPORTED).  The quadrature
 *                          plan uses the estimator order f = 3 of fbem_bem_harpot3d_sbie_auto (lib/fbem/src/bem_harpot3d.f90:961-1025).
 *   mfb_harpot3d_assemble  == `A_c=0; b_c=0` + build_lse_mechanics_bem_harpot(kf,kr) (src/build_lse_mechanics_bem_harpot.f90: element
 *                          loop :211-217, collocation loop :722-1133, free terms :243-660 with fbem_bem_pot3d_sbie_freeterm) with the
 *                          kernels of fbem_bem_harpot3d_sbie_ext_pre / _ext_adp / _int (:276-959) and the scatter of
 *                          assemble_bem_harpot_equation.f90:78-96.  rho = region%property_r(1), c = region%property_c(4);
 *                          cvalue[n_node] = node%cvalue_c(1,1,1) (prescribed p or Un).  The flux unknown is the normal displacement
 *                          Un = (dp/dn)/(rho omega^2) (:751, :1104).  A, b as in mfb_harela3d_assemble.
 *   mfb_harpot3d_solve_frequency  assemble + zgetrf + zgetrs on the device; mfb_zsolve / mfb_get_solution / mfb_residual* / mfb_get_entries
 *                          / mfb_plan_modes / mfb_get_stats work on such a problem as they do on an elastic one. */
int mfb_harpot3d_setup(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                       const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                       const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                       const int* row, const int* col_p, const int* col_un, const int* ctype, int n_dof,
                       double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                       double geometric_tolerance, mfb_problem** problem);
int mfb_harpot3d_assemble(mfb_problem* problem, double omega, double rho, const mfb_z* c, const mfb_z* cvalue, mfb_z* A, mfb_z* b);
int mfb_harpot3d_solve_frequency(mfb_problem* problem, double omega, double rho, const mfb_z* c, const mfb_z* cvalue, mfb_z* x);

/* ---- Biot poroelastic BE region ------------------------------------------------------------
 * One poroelastic region with ordinary `be` boundaries: FOUR equations and unknowns per node, component 0 = fluid phase (fluid equivalent
 * stress tau | fluid normal displacement Un), components 1..3 = solid skeleton (u_k | t_k).
 *   mfb_harpor3d_setup     the arguments of mfb_harela3d_setup with four entries per node: row[4*n_node] = node%row(0:3,1), col_p = node%col(0:3,1)
 *                          (tau, u_k), col_s = node%col(4:7,1) (Un, t_k), ctype[4*n_node] = node%ctype(0:3,1): the open-pore conditions 0 (tau / u_k
 *                          known) and 1 (Un / t_k known) of assemble_bem_harpor_equation.f90:78-110, :140-170 (the close-pore types 2..7 are refused).
 *   mfb_harpor3d_assemble  == `A_c=0; b_c=0` + build_lse_mechanics_bem_harpor(kf,kr) (src/build_lse_mechanics_bem_harpor.f90) with the kernels
 *                          of fbem_bem_harpor3d_sbie_ext_pre / _ext_adp / _int (lib/fbem/src/bem_harpor3d.f90:906-1890).  lambda, mu (drained, with
 *                          damping) = region%property_c(3:4), rho1, rho2 = property_r(13:14), rhoa = property_r(9), R, Q = property_c(10:11),
 *                          b = property_r(12); cvalue[4*n_node] = node%cvalue_c(0:3,1,1).
 *   mfb_harpor3d_solve_frequency  assemble + zgetrf + zgetrs on the device.
 * Parity suite: tests/test_gpu_poroelastic.py (A, b <= 1e-11 and x <= 1e-8 against the oracle on a B200). */
int mfb_harpor3d_setup(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                       const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                       const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                       const int* row, const int* col_p, const int* col_s, const int* ctype, int n_dof,
                       double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                       double geometric_tolerance, mfb_problem** problem);
int mfb_harpor3d_assemble(mfb_problem* problem, double omega, const mfb_z* lambda, const mfb_z* mu, double rho1, double rho2, double rhoa,
                          const mfb_z* R, const mfb_z* Q, double b, const mfb_z* cvalue, mfb_z* A, mfb_z* bvec);
int mfb_harpor3d_solve_frequency(mfb_problem* problem, double omega, const mfb_z* lambda, const mfb_z* mu, double rho1, double rho2, double rhoa,
                                 const mfb_z* R, const mfb_z* Q, double b, const mfb_z* cvalue, mfb_z* x);

/* ---- One frequency over several GPUs (shard 2) --------------------------------------------------
 * One process per GPU, every process holds the same mfb_problem (mesh + plan replicated).  Rank r assembles a contiguous
 * run of collocation-row blocks (the rows of build_lse_mechanics_bem_harela's kn_col loop it owns, all columns), the row
 * slabs move once to the owners of the matrix columns (block-cyclic, block = nb columns, NCCL send/recv), and the LU
 * (solve_lse_c -> zgetrf/zgetrs) runs distributed: per block column the owner factorises the panel, NCCL broadcasts it,
 * every rank interchanges / solves / updates its own columns on the FP64 tensor pipe; the right-hand side rides along as
 * a replicated extra column and the back substitution reduces one block of partial sums per step.  The solution is
 * returned on every rank.  All mfb_dist_* calls are collective over the ranks of the communicator.
 *   mfb_dist_unique_id   rank 0 creates the 128-byte NCCL id; the host distributes it (MPI_Bcast / torch.distributed)
 *   mfb_dist_init        joins the communicator (nb <= 0: default 256) and allocates the local column storage
 *   mfb_dist_init_loopback  test mode: `nranks` virtual ranks on ONE GPU, collectives are device copies (no NCCL)
 *   mfb_dist_solve_frequency  == mfb_harela3d_solve_frequency, distributed
 *   mfb_dist_zsolve      == zgesv on a host matrix every rank passes in full (tests of the distributed LU)
 * Host-only helpers (no GPU needed): mfb_dist_layout (local -> global column map of a rank) and
 * mfb_dist_partition_tiles (collocation tiles -> ranks, row_bounds[nranks+1]). */
int mfb_dist_unique_id(char* id128);
int mfb_dist_init(mfb_problem* problem, int rank, int nranks, const char* id128, int nb);
int mfb_dist_init_loopback(mfb_problem* problem, int nranks, int nb);
int mfb_dist_info(mfb_problem* problem, int* rank, int* nranks, int* row_bounds, int* n_local_cols);
int mfb_dist_solve_frequency(mfb_problem* problem, double omega, const mfb_z* lambda, const mfb_z* mu, double rho,
                             const mfb_z* nu, const mfb_z* cvalue, mfb_z* x);
int mfb_dist_zsolve(mfb_problem* problem, int n, const mfb_z* A, int lda, const mfb_z* b, mfb_z* x, int* ipiv);
int mfb_dist_layout(int n, int nb, int nranks, int rank, int* n_local_cols, int* local_to_global);
int mfb_dist_partition_tiles(int n_tiles, const int* tile_row0, const int* tile_nbytes, int n_dof, int nranks, int* tile_rank,
                             int* row_bounds);

/* ---- Resident combination of assembled systems (coupled regions from single-region assemblies; see api.cu; tests/test_gpu_coupled.py) ----
 * mfb_system_zero(dst) zeroes the resident system of dst and marks it assembled in host order; mfb_combine_columns adds
 * coef[i] * src(r, src_col[i]) to dst(row_map[r], dst_col[i]) for the first n_rows host rows of the assembled system of src (dst_col = -1: the
 * right-hand side); mfb_add_entries adds single entries (the free terms).  mfb_zsolve(dst, n, NULL, n, ipiv, NULL, 1, 1) then factorises and
 * solves the accumulated system and mfb_get_solution returns x. */
int mfb_system_zero(mfb_problem* problem);
int mfb_combine_columns(mfb_problem* dst, mfb_problem* src, int n_rows, const int* row_map, int n_terms, const int* src_col, const int* dst_col,
                        const mfb_z* coef);
int mfb_add_entries(mfb_problem* dst, int n, const int* rows, const int* cols, const mfb_z* values);

/* Host-only helper: geometry-only pieces of the free term at a boundary node (see api.cu); c_lk = cp delta_lk - sum_b[3*l+k] / (8 pi (1 - nu))
 * for a solid (Mantic), c = cp for a fluid, c_00 = J cp for the fluid phase of a poroelastic medium.  normals / tangents: 3 per element. */
int mfb_freeterm_terms(int n_elements, const double* normals, const double* tangents, double tol, double* cp, double* sum_b /* 9 */);

#ifdef __cplusplus
}
#endif
#endif
