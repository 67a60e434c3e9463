// api.cu -- C ABI (include/mfb.h) of the B200-native harmonic 3D BEM hot path: contexts, problem set-up
// (host planning + device residency), per-frequency assembly, LU solve.  No CPU fallback: without a CUDA device
// every entry point fails with MFB_ERR_NO_DEVICE.
#include "../../include/mfb.h"
#include "assembly.cuh"
#include "potential.cuh"
#include "poro.cuh"
#include "combine.cuh"
#include "lu.cuh"
#include "solve_ex.cuh"
#include "dist.cuh"
#include "plan_host.h"
#include "../../data/quad_tables.h"
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <omp.h>

using namespace mfbd;
typedef std::complex<double> cd;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                                        \
  do {                                                                                                                  \
    cudaError_t e__ = (call);                                                                                           \
    if (e__ != cudaSuccess) return fail(MFB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));             \
  } while (0)

struct mfb_ctx {
  int device; cudaStream_t stream; DevTables tables; double* tables_buf; cudaEvent_t marks[8];
  K1Launch k1;     // launch state of the regular kernel on this context's stream
};

struct GroupHost {
  int et, nn, n_elem, slot0;
  std::vector<int> elem_ids;
  DevGroup dev;
  std::vector<void*> owned;
  // work lists
  DevAdaptive adp; DevSingular sing;
};

// single-frequency multi-GPU mode (mfb_dist_*): row-block assembly + block-cyclic column LU (lu.cuh, dist.cuh)
struct DistState {
  bool on = false, loopback = false;
  int rank = 0, P = 1, nb = 256;
  DistComm* comm = nullptr;
  DistLU lu;
  std::vector<int> rb, tile_rank;          // internal-row boundaries [P+1]; owner rank of every collocation tile
  unsigned char* d_mask = nullptr;         // [n_tiles] tile_active of the rank being assembled
  double *sendbuf = nullptr, *recvbuf = nullptr, *bsum = nullptr;
  std::vector<size_t> soff, roff, ns, nr;
  cudaEvent_t ev[6];
};

struct mfb_problem {
  mfb_ctx* ctx;
  int n_node, n_elem, n_colloc, n_dof, ldp;
  long long lda;
  mfbh::Settings settings;
  std::vector<mfbh::Elem> elems;       // by original element id
  std::vector<int> slot_of_elem, elem_of_slot, cpos_of_colloc;
  std::vector<GroupHost> groups;
  std::vector<void*> owned;
  DevColloc colloc; DevSystem sys; DevClassify cls; DevFreeTerm ft;
  DevFreeTerm ft0;                     // poroelastic regions: the fluid-phase free terms, whose multiplier is J (ft: skeleton, multiplier F)
  unsigned char* plan;
  double* d_cvalue;
  int* d_ipiv; int* d_perm; std::vector<int> h_ipiv;
  LuWork lu; bool lu_ready; bool factored; bool have_cvalue; bool assembled; int asm_launches;
  double stats[MFB_STAT_COUNT];
  cudaEvent_t ev[8];
  std::vector<int> set_gln;
  std::vector<int> rowperm, colperm;   // host row / column -> internal row / column of the device-resident assembled system
  int *d_rowperm, *d_colperm; bool rows_permuted;
  std::vector<int> h_tile_row0, h_tile_nbytes;   // host copies of DevColloc::tile_row0 / tile_nbytes (row partition of the multi-GPU mode)
  DistState dist;
  alignas(64) unsigned char tmapS[128]; bool have_tmapS;   // one-plane box: K1 flush of the static (real) assembly
  int ndof;                                                // equations / unknowns per node: 3 (elastic solid), 1 (inviscid fluid, mfb_harpot3d_*)
  // incident field (mfb_harela3d_set_incident): flat [slot_off[n_elem]][4] device array in slot order, images carry the root's values times symconf_t(k)
  std::vector<unsigned char> node_rev;   // node belongs to a reversed boundary (from the root elements around it)
  // rows written by the host after every assembly (local-axes conditions of ctype 2 / 3 nodes): internal row / column (-1: rhs) and value
  int n_cond = 0; int *d_cond_row = nullptr, *d_cond_col = nullptr; double* d_cond_val = nullptr;
  bool skip_cond = false;   // multi-GPU row-block assembly: only the rank that owns the last rows (where the condition rows live) adds them
  double* d_nfn = nullptr; bool need_normals = false, have_normals = false;   // ctype 10: nodal normals n_fn (mfb_set_node_normals)
  double* d_einc = nullptr; bool have_inc = false; std::vector<int> slot_off_h, root_elem_ptr; std::vector<unsigned char> elem_symbits; int n_elem_root = 0;
  bool hbie;                                               // hypersingular equation at points off the boundary (interior-point stresses)
  bool real_resident;                                      // the resident system / factors are real (static path): Are only
  // optional stages of solve_lse_c (mfb_zsolve_ex): unfactorised (scaled) copy of A, scale factors in the order of the resident system
  // factorise + solve of a SMALL resident system as one CUDA graph (launch-bound regime: ~300 launches per frequency at 1386 DOF)
  int lu_plain_calls = 0;
  cudaGraphExec_t lu_graph = nullptr; int lu_graph_nodes = 0; int* d_graph_flags = nullptr; bool lu_graph_failed = false;
  double* d_vstage = nullptr;                              // 2 n_dof doubles: staging of solution / right-hand-side downloads
  double* Ao = nullptr; double *d_rs = nullptr, *d_cs = nullptr, *d_xtmp = nullptr; char equed = 'N'; std::vector<double> rs_int, cs_int;
  alignas(64) unsigned char tmapA[128]; bool have_tmap;   // CUtensorMap of the planar system matrix (K1 flush)   // rows_permuted: the resident matrix/factors are in the internal order
};

extern "C" const char* mfb_last_error(void) { return g_err.c_str(); }
extern "C" int mfb_version(void) { return 100; }

struct DevBuf {   // scoped device allocation: released on every exit path of the function that owns it
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
};

template <class T>
static int upload(std::vector<void*>& owned, const std::vector<T>& h, T** d, cudaStream_t st) {
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  CK(cudaMalloc((void**)d, bytes));
  owned.push_back(*d);
  if (!h.empty()) CK(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return 0;
}
#define UP(owned, h, d)                                        \
  do {                                                         \
    int r__ = upload(owned, h, d, st);                         \
    if (r__) return r__;                                       \
  } while (0)

extern "C" int mfb_init(int device, mfb_ctx** out) {
  if (!out) return fail(MFB_ERR_ARG, "mfb_init: null output");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return fail(MFB_ERR_NO_DEVICE, "mfb_init: no CUDA device (this library has no CPU path)");
  if (device < 0 || device >= n) return fail(MFB_ERR_ARG, "mfb_init: device index out of range");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(MFB_ERR_NO_DEVICE, "mfb_init: kernels are built for sm_100a only");
  mfb_ctx* c = new mfb_ctx();
  c->device = device;
  CK(cudaStreamCreate(&c->stream));   // a blocking stream on purpose: set-up code reads results back with plain cudaMemcpy (legacy stream) and relies on its implicit ordering
  for (int i = 0; i < 8; i++) CK(cudaEventCreate(&c->marks[i]));
  // Gauss-Legendre tables on the device (packed: rule n starts at n(n-1)/2)
  CK(cudaMalloc((void**)&c->tables_buf, 4 * 528 * sizeof(double)));
  CK(cudaMemcpy(c->tables_buf, QT_GL11_X, 528 * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->tables_buf + 528, QT_GL11_W, 528 * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->tables_buf + 2 * 528, QT_GL01_X, 528 * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->tables_buf + 3 * 528, QT_GL01_W, 528 * 8, cudaMemcpyHostToDevice));
  if (k1_launch_create(c->k1)) return fail(MFB_ERR_CUDA, "mfb_init: K1 launch state");
  c->tables.gl11_x = c->tables_buf; c->tables.gl11_w = c->tables_buf + 528; c->tables.gl01_x = c->tables_buf + 2 * 528; c->tables.gl01_w = c->tables_buf + 3 * 528;
  *out = c;
  return MFB_OK;
}
extern "C" void mfb_finalize(mfb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaFree(c->tables_buf);
  k1_launch_destroy(c->k1);
  for (int i = 0; i < 8; i++) cudaEventDestroy(c->marks[i]);
  cudaStreamDestroy(c->stream);
  delete c;
}

static void dist_release(mfb_problem* p) {
  DistState& d = p->dist;
  if (!d.on) return;
  for (auto& R : d.lu.r) dist_rank_free(R);
  d.lu.r.clear();
  cudaFree(d.d_mask); cudaFree(d.sendbuf); cudaFree(d.recvbuf); cudaFree(d.bsum);
  for (int i = 0; i < 6; i++) cudaEventDestroy(d.ev[i]);
  delete d.comm; d.comm = nullptr; d.on = false;
  p->colloc.tile_active = nullptr;
}

extern "C" void mfb_problem_free(mfb_problem* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  for (void* q : p->owned) cudaFree(q);
  for (auto& g : p->groups) for (void* q : g.owned) cudaFree(q);
  if (p->lu_ready) lu_work_free(p->lu);
  if (p->lu_graph) cudaGraphExecDestroy(p->lu_graph);
  cudaFree(p->d_cond_row); cudaFree(p->d_cond_col); cudaFree(p->d_cond_val);
  cudaFree(p->d_graph_flags); cudaFree(p->d_vstage); cudaFree(p->Ao); cudaFree(p->d_rs); cudaFree(p->d_cs); cudaFree(p->d_xtmp);
  for (int i = 0; i < 8; i++) cudaEventDestroy(p->ev[i]);
  dist_release(p);
  delete p;
}

// 30-bit Morton key of a point inside the box [lo, lo+ext]^3 (spatial clustering of tiles and element chunks)
static unsigned morton3(const double* x, const double* lo, double inv_ext) {
  unsigned key = 0;
  unsigned q[3];
  for (int c = 0; c < 3; c++) { double t = (x[c] - lo[c]) * inv_ext; t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t); q[c] = (unsigned)(t * 1023.0 + 0.5); }
  for (int b = 9; b >= 0; b--) for (int c = 0; c < 3; c++) key = (key << 1) | ((q[c] >> b) & 1u);
  return key;
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Symmetry planes through the origin (src/read_symmetry_planes.f90:76-283): plane i is normal to axis eid[i] (1..3, ascending) and multiplies
// the translation dof k of a reflected element by t[3*i+k] (symmetry: -1 on the normal axis, +1 elsewhere; antisymmetry: the opposite)
struct SymSpec { int n_planes; int eid[3]; double t[9]; double s[3]; };   // s: multiplier of the scalar variables (fluid pressure / fluid phase), symplane_s
static int setup_impl(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                      const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                      const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                      const int* row, const int* col_u, const int* col_t, const int* ctype, int n_dof,
                      double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                      double geometric_tolerance, const double* colloc_n /* NULL, or 3 per collocation point: hypersingular equation */,
                      int ndof /* 3: elastic solid, 1: inviscid fluid */, mfb_problem** out, const SymSpec* sym = nullptr);
extern "C" int mfb_harela3d_setup(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                                  const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                                  const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                                  const int* row, const int* col_u, const int* col_t, const int* ctype, int n_dof,
                                  double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                                  double geometric_tolerance, mfb_problem** out) {
  return setup_impl(ctx, n_node, node_x, n_elem, etype, elem_ptr, elem_node, elem_reversed, n_colloc, colloc_x, colloc_node, colloc_elem, colloc_kn, colloc_xi,
                    row, col_u, col_t, ctype, n_dof, qsi_relative_error, qsi_ns_max, n_precalsets, precalset_gln, geometric_tolerance, nullptr, 3, out);
}
// The same with symmetry planes: the reference's [symmetry planes] section (src/read_symmetry_planes.f90) -- up to three planes through the origin,
// normal to the axes symplane_eid[i] (1 = x, 2 = y, 3 = z, ascending), symplane_t[3*i+k] = multiplier of translation dof k across plane i.
// colloc_n may be NULL (displacement equation) or the unit normals of the hypersingular equation at interior points.
extern "C" int mfb_harela3d_setup_sym(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                                      const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                                      const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi, const double* colloc_n,
                                      const int* row, const int* col_u, const int* col_t, const int* ctype, int n_dof,
                                      double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                                      double geometric_tolerance, int n_symplanes, const int* symplane_eid, const double* symplane_t, mfb_problem** out) {
  if (n_symplanes < 0 || n_symplanes > 3 || (n_symplanes > 0 && (!symplane_eid || !symplane_t))) return fail(MFB_ERR_ARG, "mfb_harela3d_setup_sym: invalid symmetry planes");
  SymSpec sp; sp.n_planes = n_symplanes;
  for (int i = 0; i < n_symplanes; i++) { sp.eid[i] = symplane_eid[i]; sp.s[i] = 1.0; for (int k = 0; k < 3; k++) sp.t[3 * i + k] = symplane_t[3 * i + k]; }
  return setup_impl(ctx, n_node, node_x, n_elem, etype, elem_ptr, elem_node, elem_reversed, n_colloc, colloc_x, colloc_node, colloc_elem, colloc_kn, colloc_xi,
                    row, col_u, col_t, ctype, n_dof, qsi_relative_error, qsi_ns_max, n_precalsets, precalset_gln, geometric_tolerance, colloc_n, 3, out, &sp);
}
// Symmetry planes for the other two region types: the image loops of build_lse_mechanics_bem_harpot.f90:790-800 (h, g times symconf_s, :947-948) and
// build_lse_mechanics_bem_harpor.f90:855-865 (dof 0 times symconf_s, dofs 1..3 times symconf_t(k), :971-975).  symplane_s[i] = symplane_s(i): +1 symmetry, -1 antisymmetry.
static int fill_sym(SymSpec& sp, int n_symplanes, const int* symplane_eid, const double* symplane_s, const double* symplane_t) {
  if (n_symplanes < 0 || n_symplanes > 3 || (n_symplanes > 0 && (!symplane_eid || !symplane_s || !symplane_t))) return fail(MFB_ERR_ARG, "setup_sym: invalid symmetry planes");
  sp.n_planes = n_symplanes;
  for (int i = 0; i < n_symplanes; i++) { sp.eid[i] = symplane_eid[i]; sp.s[i] = symplane_s[i]; for (int k = 0; k < 3; k++) sp.t[3 * i + k] = symplane_t[3 * i + k]; }
  return MFB_OK;
}
extern "C" int mfb_harpot3d_setup_sym(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                                      const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                                      const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                                      const int* row, const int* col_p, const int* col_un, const int* ctype, int n_dof,
                                      double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                                      double geometric_tolerance, int n_symplanes, const int* symplane_eid, const double* symplane_s, const double* symplane_t, mfb_problem** out) {
  SymSpec sp; int r = fill_sym(sp, n_symplanes, symplane_eid, symplane_s, symplane_t); if (r) return r;
  return setup_impl(ctx, n_node, node_x, n_elem, etype, elem_ptr, elem_node, elem_reversed, n_colloc, colloc_x, colloc_node, colloc_elem, colloc_kn, colloc_xi,
                    row, col_p, col_un, ctype, n_dof, qsi_relative_error, qsi_ns_max, n_precalsets, precalset_gln, geometric_tolerance, nullptr, 1, out, &sp);
}
extern "C" int mfb_harpor3d_setup_sym(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                                      const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                                      const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                                      const int* row, const int* col_p, const int* col_s, const int* ctype, int n_dof,
                                      double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                                      double geometric_tolerance, int n_symplanes, const int* symplane_eid, const double* symplane_s, const double* symplane_t, mfb_problem** out) {
  SymSpec sp; int r = fill_sym(sp, n_symplanes, symplane_eid, symplane_s, symplane_t); if (r) return r;
  return setup_impl(ctx, n_node, node_x, n_elem, etype, elem_ptr, elem_node, elem_reversed, n_colloc, colloc_x, colloc_node, colloc_elem, colloc_kn, colloc_xi,
                    row, col_p, col_s, ctype, n_dof, qsi_relative_error, qsi_ns_max, n_precalsets, precalset_gln, geometric_tolerance, nullptr, 4, out, &sp);
}
// Hypersingular equation for points OFF the boundary (interior-point stresses): fbem_bem_harela3d_hbie_auto with its exterior
// branches (_ext_pre :2573-2662, _ext_adp :3044-3167); colloc_n[3*n_colloc] = unit normal n_i of each collocation point.
extern "C" int mfb_harela3d_setup_hbie(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                                       const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                                       const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi, const double* colloc_n,
                                       const int* row, const int* col_u, const int* col_t, const int* ctype, int n_dof,
                                       double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                                       double geometric_tolerance, mfb_problem** out) {
  if (!colloc_n) return fail(MFB_ERR_ARG, "mfb_harela3d_setup_hbie: null colloc_n");
  return setup_impl(ctx, n_node, node_x, n_elem, etype, elem_ptr, elem_node, elem_reversed, n_colloc, colloc_x, colloc_node, colloc_elem, colloc_kn, colloc_xi,
                    row, col_u, col_t, ctype, n_dof, qsi_relative_error, qsi_ns_max, n_precalsets, precalset_gln, geometric_tolerance, colloc_n, 3, out);
}
static int setup_impl(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                      const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                      const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                      const int* row, const int* col_u, const int* col_t, const int* ctype, int n_dof,
                      double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                      double geometric_tolerance, const double* colloc_n, int ndof, mfb_problem** out, const SymSpec* sym) {
  if (!ctx || !out || !node_x || !etype || !elem_ptr || !elem_node || !colloc_x || !colloc_node || !colloc_elem || !colloc_kn || !colloc_xi ||
      !row || !col_u || !col_t || !ctype || !precalset_gln)
    return fail(MFB_ERR_ARG, "mfb_harela3d_setup: null argument");
  if (n_node <= 0 || n_elem <= 0 || n_colloc <= 0 || n_dof <= 0 || n_precalsets <= 0 || n_precalsets > MAX_SETS)
    return fail(MFB_ERR_ARG, "mfb_harela3d_setup: invalid size");
  for (int e = 0; e < n_elem; e++) {
    if (etype[e] < MFB_TRI3 || etype[e] > MFB_QUAD9) return fail(MFB_ERR_UNSUPPORTED, "mfb_harela3d_setup: element type must be tri3/tri6/quad4/quad8/quad9");
    if (elem_ptr[e + 1] - elem_ptr[e] != mfbh::nodes_of(etype[e])) return fail(MFB_ERR_ARG, "mfb_harela3d_setup: elem_ptr inconsistent with etype");
  }
  if (ndof != 1 && ndof != 3 && ndof != 4) return fail(MFB_ERR_ARG, "setup: ndof must be 3 (elastic solid), 1 (inviscid fluid) or 4 (poroelastic medium)");
  for (int i = 0; i < ndof * n_node; i++)
    if (ctype[i] != 0 && ctype[i] != 1 && !((ctype[i] == 10 || ctype[i] == 2 || ctype[i] == 3) && ndof == 3 && !colloc_n))
      return fail(MFB_ERR_UNSUPPORTED, "mfb_harela3d_setup: ctype 0 (u / p known), 1 (t / Un known) and, for elastic regions, 2 / 3 (local axes) and 10 (normal pressure known) are supported");
  // ctype 10 (assemble_bem_harela_equation.f90:97-106): the pressure p on the node is known, t_k = p n_fn(k): a traction-known dof (kind 1, unknown u_k, column
  // col(k,1)) whose prescribed value is cvalue * n_fn(k), negated on a reversed boundary.  The nodal normals come with mfb_set_node_normals.
  // ctype 2 / 3 (local-axes conditions u.l = U / t.l = T, assemble_bem_harela_equation.f90:107-112): u_k AND t_k are unknowns, h goes to the column of u_k and -g
  // to the column of t_k (kind 2 of the scatter); the three condition rows of such a node are the host's (src/build_lse_mechanics_harmonic.f90:204-258,
  // mfb_set_condition_rows).  Groups with such dofs run the general K1 kernel (their columns are not three consecutive ones).
  std::vector<int> kind_of; std::vector<unsigned char> c10;
  bool any_kind2 = false;
  for (int i = 0; i < ndof * n_node; i++) if (ctype[i] == 10 || ctype[i] == 2 || ctype[i] == 3) {
    if (kind_of.empty()) kind_of.assign(ctype, ctype + (size_t)ndof * n_node);
    if (ctype[i] == 10) { if (c10.empty()) c10.assign((size_t)ndof * n_node, 0); c10[i] = 1; kind_of[i] = 1; }
    else { kind_of[i] = 2; any_kind2 = true; if (col_u[i] < 0 || col_u[i] >= n_dof || col_t[i] < 0 || col_t[i] >= n_dof) return fail(MFB_ERR_ARG, "mfb_harela3d_setup: a ctype 2 / 3 dof needs the columns of both u_k and t_k"); }
  }
  if (!kind_of.empty()) ctype = kind_of.data();
  double t_host0 = now_ms();
  // ---- symmetry images (lib/fbem/src/symmetry.f90:60-171; image loop build_lse_mechanics_bem_harela.f90:1098-1107) ----
  // Every root element gets n_sym - 1 image elements: element ks * n_root + r is image ks of root r, with the root's nodes (hence its columns and
  // prescribed values), reflected coordinates, the orientation flipped by an odd number of reflections, and the sign symconf_t(k) on its dof-k
  // columns (bits 5-7 of einfo).  From here on an image is an element like any other: the classifier, the planner and K1/K2/K3 never know.
  const int n_root = n_elem;
  int n_sym = 1;
  double conf_m[8][3], conf_t[8][3], conf_s[8]; bool conf_rev[8];
  for (int ks = 0; ks < 8; ks++) { conf_rev[ks] = false; conf_s[ks] = 1.0; for (int c = 0; c < 3; c++) { conf_m[ks][c] = 1.0; conf_t[ks][c] = 1.0; } }
  double plane_m[3][3] = {{1, 1, 1}, {1, 1, 1}, {1, 1, 1}};
  std::vector<int> x_etype, x_ptr, x_node; std::vector<unsigned char> x_rev;
  if (sym && sym->n_planes > 0) {
    if (sym->n_planes > 3) return fail(MFB_ERR_ARG, "setup: at most three symmetry planes");

    for (int i = 0; i < sym->n_planes; i++) {
      if (sym->eid[i] < 1 || sym->eid[i] > 3 || (i > 0 && sym->eid[i] <= sym->eid[i - 1])) return fail(MFB_ERR_ARG, "setup: symmetry plane axes must be 1..3 in ascending order");
      for (int c = 0; c < 3; c++) {
        if (fabs(sym->t[3 * i + c]) != 1.0 || fabs(sym->s[i]) != 1.0) return fail(MFB_ERR_ARG, "setup: symmetry multipliers must be +1 or -1");
        plane_m[i][c] = (c == sym->eid[i] - 1) ? -1.0 : 1.0;
      }
    }
    n_sym = 1 << sym->n_planes;
    if ((long long)n_root * n_sym > 0x7fffffffLL / 64) return fail(MFB_ERR_ARG, "setup: too many elements");
    static const int steps[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};   // fbem_symmetry_multipliers :81-170
    for (int ks = 0; ks < n_sym; ks++) {
      int cnt = 0;
      for (int i = 0; i < sym->n_planes; i++) if (steps[ks][i]) { cnt++; conf_s[ks] *= sym->s[i]; for (int c = 0; c < 3; c++) { conf_m[ks][c] *= plane_m[i][c]; conf_t[ks][c] *= sym->t[3 * i + c]; } }
      conf_rev[ks] = (cnt & 1) != 0;
    }
    const int nen = elem_ptr[n_root];
    x_etype.resize((size_t)n_root * n_sym); x_rev.resize((size_t)n_root * n_sym); x_ptr.resize((size_t)n_root * n_sym + 1); x_node.resize((size_t)nen * n_sym);
    for (int ks = 0; ks < n_sym; ks++) {
      for (int r = 0; r < n_root; r++) {
        const size_t e = (size_t)ks * n_root + r;
        x_etype[e] = etype[r]; x_ptr[e] = ks * nen + elem_ptr[r];
        x_rev[e] = (unsigned char)(((elem_reversed && elem_reversed[r]) != conf_rev[ks]) ? 1 : 0);
      }
      for (int q = 0; q < nen; q++) x_node[(size_t)ks * nen + q] = elem_node[q];
    }
    x_ptr[(size_t)n_root * n_sym] = n_sym * nen;
    etype = x_etype.data(); elem_ptr = x_ptr.data(); elem_node = x_node.data(); elem_reversed = x_rev.data(); n_elem = n_root * n_sym;
  }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  mfb_problem* p = new mfb_problem();
  *out = nullptr;
  p->ctx = ctx; p->n_node = n_node; p->n_elem = n_elem; p->n_colloc = n_colloc; p->n_dof = n_dof; p->ndof = ndof;
  p->lu_ready = false; p->factored = false; p->plan = nullptr; p->have_cvalue = false; p->assembled = false;
  memset(p->stats, 0, sizeof(p->stats));
  for (int i = 0; i < 8; i++) cudaEventCreate(&p->ev[i]);
  mfbh::Settings& S = p->settings;
  S.qsi_relative_error = qsi_relative_error; S.qsi_ns_max = qsi_ns_max; S.geometric_tolerance = geometric_tolerance;
  S.ps_gln.assign(precalset_gln, precalset_gln + n_precalsets);
  p->set_gln = S.ps_gln;
  // order of the estimator's model function: fbem_bem_harela3d_sbie_auto :1522 / _hbie_auto :3677 / fbem_bem_harpot3d_sbie_auto (bem_harpot3d.f90:1009)
  S.f = colloc_n ? 7 : (ndof == 1 ? 3 : 5);
  p->hbie = colloc_n != nullptr;
  mfbh::init_settings(S);

  // ---- per-element data (csize, n_phi, bounding ball) ----
  p->elems.resize(n_elem);
#pragma omp parallel for schedule(dynamic, 64)
  for (int e = 0; e < n_elem; e++) {
    mfbh::Elem& el = p->elems[e];
    el.et = etype[e]; el.nn = mfbh::nodes_of(el.et); el.reversed = elem_reversed && elem_reversed[e];
    const double* cm = conf_m[e / n_root];
    for (int k = 0; k < el.nn; k++) for (int c = 0; c < 3; c++) el.x[3 * k + c] = cm[c] * node_x[3 * (size_t)elem_node[elem_ptr[e] + k] + c];
    mfbh::element_data(el, S);
  }
  // the reference reflects the nodal coordinates of the calculation element only: csize, n_phi and the bounding ball -- its CENTRE included -- stay
  // those of the root element (build_lse_mechanics_bem_harela.f90:1052-1069 set them once, :1103-1105 touch x alone), and they feed the far test
  for (int e = n_root; e < n_elem; e++) {
    mfbh::Elem& el = p->elems[e]; const mfbh::Elem& ro = p->elems[e % n_root];
    el.cl = ro.cl; el.gln_far = ro.gln_far; el.br = ro.br; for (int c = 0; c < 3; c++) el.bc[c] = ro.bc[c];
  }
  // ---- bounding box of the mesh (Morton keys) ----
  double bb_lo[3] = {1e300, 1e300, 1e300}, bb_hi[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < n_node; i++) for (int c = 0; c < 3; c++) { bb_lo[c] = std::min(bb_lo[c], node_x[3 * (size_t)i + c]); bb_hi[c] = std::max(bb_hi[c], node_x[3 * (size_t)i + c]); }
  const double bb_ext = std::max(std::max(bb_hi[0] - bb_lo[0], bb_hi[1] - bb_lo[1]), std::max(bb_hi[2] - bb_lo[2], 1e-300));
  const double bb_inv = 1.0 / bb_ext;
  // ---- element slots: sorted by type, spatially clustered inside a type (a chunk of consecutive slots is a compact patch:
  //      its columns stay in L2 between the elements that share them, and a tile sees one quadrature rule for most chunks) ----
  p->slot_of_elem.assign(n_elem, -1); p->elem_of_slot.clear();
  const int types[5] = {MFB_TRI3, MFB_TRI6, MFB_QUAD4, MFB_QUAD8, MFB_QUAD9};
  for (int t = 0; t < 5; t++) {
    GroupHost g; g.et = types[t]; g.nn = mfbh::nodes_of(g.et); g.slot0 = (int)p->elem_of_slot.size();
    std::vector<std::pair<unsigned long long, int>> keyed;
    for (int e = 0; e < n_elem; e++) if (etype[e] == g.et) {
      double ctr[3] = {0, 0, 0};
      for (int k = 0; k < g.nn; k++) for (int c = 0; c < 3; c++) ctr[c] += p->elems[e].x[3 * k + c] / g.nn;
      // boundary-condition signature first (elements of one K1 class end up in the same ranges), then space
      unsigned sig = 0;
      for (int k = 0; k < ndof; k++) {
        const int ct0 = ctype[ndof * elem_node[elem_ptr[e]] + k];
        bool uni = true;
        for (int j = 1; j < g.nn; j++) if (ctype[ndof * elem_node[elem_ptr[e] + j] + k] != ct0) uni = false;
        sig = sig * 3 + (uni ? (unsigned)ct0 : 2u);
      }
      keyed.push_back(std::make_pair((unsigned long long)sig << 32 | morton3(ctr, bb_lo, bb_inv), e));
    }
    std::stable_sort(keyed.begin(), keyed.end());
    for (auto& ke : keyed) { p->slot_of_elem[ke.second] = (int)p->elem_of_slot.size(); p->elem_of_slot.push_back(ke.second); g.elem_ids.push_back(ke.second); }
    g.n_elem = (int)g.elem_ids.size();
    memset(&g.dev, 0, sizeof(g.dev)); memset(&g.adp, 0, sizeof(g.adp)); memset(&g.sing, 0, sizeof(g.sing));
    if (g.n_elem > 0) p->groups.push_back(g);
  }
  // ---- collocation tiles and the internal row order (DevColloc) ----
  // Row nodes = nodes that own collocation points.  They are ordered by (number of collocation points, Morton key); the
  // internal matrix rows follow that order (3 consecutive rows per node), so that every block of <= 32 consecutive row
  // nodes of one multiplicity class is a 16-byte aligned run of rows, spatially compact, with all its layers full.
  // A class with an odd number of nodes gives its last node to the loose tiles (keeps every later run aligned).
  std::vector<std::vector<int>> node_collocs(n_node);
  for (int c = 0; c < n_colloc; c++) {
    int nd = colloc_node[c];
    if (nd < 0 || nd >= n_node) { mfb_problem_free(p); return fail(MFB_ERR_ARG, "mfb_harela3d_setup: colloc_node out of range"); }
    for (int k = 0; k < ndof; k++) { int r = row[ndof * nd + k]; if (r < 0 || r >= n_dof) { mfb_problem_free(p); return fail(MFB_ERR_ARG, "mfb_harela3d_setup: collocation node without a valid row"); } }
    node_collocs[nd].push_back(c);
  }
  struct RowNode { int mult; unsigned key; int node; };
  std::vector<RowNode> rn;
  for (int nd = 0; nd < n_node; nd++) if (!node_collocs[nd].empty()) rn.push_back({(int)node_collocs[nd].size(), morton3(&node_x[3 * (size_t)nd], bb_lo, bb_inv), nd});
  std::stable_sort(rn.begin(), rn.end(), [](const RowNode& a, const RowNode& b) { return a.mult != b.mult ? a.mult < b.mult : a.key < b.key; });
  std::vector<int> bulk_nodes, loose_nodes;
  {
    std::vector<char> row_used(n_dof, 0);
    size_t i = 0;
    while (i < rn.size()) {
      size_t j = i; while (j < rn.size() && rn[j].mult == rn[i].mult) j++;
      size_t cnt = j - i;
      for (size_t q = i; q < j; q++) {
        const int nd = rn[q].node;
        bool dup = false;
        for (int k = 0; k < ndof; k++) { if (row_used[row[ndof * nd + k]]) dup = true; row_used[row[ndof * nd + k]] = 1; }
        if (dup) { mfb_problem_free(p); return fail(MFB_ERR_ARG, "mfb_harela3d_setup: two collocation nodes share a matrix row"); }
        if ((cnt & 1) && q == j - 1) loose_nodes.push_back(nd); else bulk_nodes.push_back(nd);
      }
      i = j;
    }
  }
  p->rowperm.assign(n_dof, -1);
  {
    int next = 0;
    for (int nd : bulk_nodes) for (int k = 0; k < ndof; k++) p->rowperm[row[ndof * nd + k]] = next++;
    for (int nd : loose_nodes) for (int k = 0; k < ndof; k++) p->rowperm[row[ndof * nd + k]] = next++;
    for (int r = 0; r < n_dof; r++) if (p->rowperm[r] < 0) p->rowperm[r] = next++;   // rows that no collocation point feeds
  }
  std::vector<int> t_row0, t_nbytes, lane_colloc;   // lane_colloc[32*tile + lane] = host collocation index or -1
  for (size_t i = 0; i < bulk_nodes.size();) {
    const int mult = (int)node_collocs[bulk_nodes[i]].size();
    size_t j = i; while (j < bulk_nodes.size() && j - i < 32 && (int)node_collocs[bulk_nodes[j]].size() == mult) j++;
    for (int layer = 0; layer < mult; layer++) {
      t_row0.push_back(ndof * (int)i); t_nbytes.push_back(8 * ndof * (int)(j - i));
      for (size_t q = i; q < i + 32; q++) lane_colloc.push_back(q < j ? node_collocs[bulk_nodes[q]][layer] : -1);
    }
    i = j;
  }
  {
    std::vector<int> loose;
    for (int nd : loose_nodes) for (int c : node_collocs[nd]) loose.push_back(c);
    for (size_t i = 0; i < loose.size(); i += 32) {
      t_row0.push_back(0); t_nbytes.push_back(0);
      for (size_t q = i; q < i + 32; q++) lane_colloc.push_back(q < loose.size() ? loose[q] : -1);
    }
  }
  const int n_tiles = (int)t_row0.size();
  const int ldp = 32 * n_tiles; p->ldp = ldp;
  p->cpos_of_colloc.assign(n_colloc, 0);
  std::vector<double> h_cx(3 * (size_t)ldp, 0.0); std::vector<int> h_crow((size_t)std::max(3, ndof) * ldp, -1);   // a poroelastic node has four rows
  for (int q = 0; q < ldp; q++) {
    const int c = lane_colloc[q];
    if (c < 0) continue;
    p->cpos_of_colloc[c] = q;
    for (int k = 0; k < 3; k++) h_cx[(size_t)k * ldp + q] = colloc_x[3 * (size_t)c + k];
    for (int k = 0; k < ndof; k++) h_crow[(size_t)k * ldp + q] = p->rowperm[row[ndof * colloc_node[c] + k]];   // a fluid node has one row: planes 1, 2 stay -1
  }
  double* d_cn = nullptr;
  if (colloc_n) {
    std::vector<double> h_cn(3 * (size_t)ldp, 0.0);
    for (int q = 0; q < ldp; q++) { const int c = lane_colloc[q]; if (c >= 0) for (int k = 0; k < 3; k++) h_cn[(size_t)k * ldp + q] = colloc_n[3 * (size_t)c + k]; }
    UP(p->owned, h_cn, &d_cn);
  }
  double* d_cx; int *d_crow, *d_trow0, *d_tnbytes;
  UP(p->owned, h_cx, &d_cx); UP(p->owned, h_crow, &d_crow); UP(p->owned, t_row0, &d_trow0); UP(p->owned, t_nbytes, &d_tnbytes);
  // Columns: when every collocation node pairs row k with the column of its unknown k (the reference's numbering,
  // build_auxiliary_variables_mechanics_harmonic.f90:151-198) the columns follow the same permutation, so that the strong
  // diagonal stays on the diagonal and partial pivoting keeps finding its pivot in place; otherwise columns keep their order.
  {
    bool paired = true;
    for (const RowNode& q : rn) for (int k = 0; k < ndof; k++) {
      const int nd = q.node, ct = ctype[ndof * nd + k];
      const int col = (ct == 0) ? col_t[ndof * nd + k] : col_u[ndof * nd + k];
      if (col != row[ndof * nd + k]) paired = false;
    }
    p->colperm.resize(n_dof);
    for (int i = 0; i < n_dof; i++) p->colperm[i] = paired ? p->rowperm[i] : i;
  }
  UP(p->owned, p->rowperm, &p->d_rowperm); UP(p->owned, p->colperm, &p->d_colperm);
  p->colloc.n_colloc = ldp; p->colloc.ldp = ldp; p->colloc.cx = d_cx; p->colloc.crow = d_crow;
  p->colloc.n_tiles = n_tiles; p->colloc.tile_row0 = d_trow0; p->colloc.tile_nbytes = d_tnbytes; p->colloc.tile_active = nullptr; p->colloc.cn = d_cn;
  p->h_tile_row0 = t_row0; p->h_tile_nbytes = t_nbytes;
  p->rows_permuted = false;

  // ---- flat scatter descriptors over all slots ----
  std::vector<int> slot_off(n_elem + 1, 0);
  for (int s = 0; s < n_elem; s++) slot_off[s + 1] = slot_off[s] + ndof * p->elems[p->elem_of_slot[s]].nn;
  p->slot_off_h = slot_off; p->n_elem_root = n_root; p->root_elem_ptr.assign(elem_ptr, elem_ptr + n_root + 1);
  p->elem_symbits.assign(n_elem, 0);   // bit k: multiplier -1 of a symmetry image on dof k of the node (solid: symconf_t; fluid: symconf_s; poroelastic: s, then t)
  for (int e = 0; e < n_elem; e++) {
    const int ks = e / n_root;
    if (ndof == 3) { for (int k = 0; k < 3; k++) if (conf_t[ks][k] < 0.0) p->elem_symbits[e] |= (unsigned char)(1u << k); }
    else { if (conf_s[ks] < 0.0) p->elem_symbits[e] |= 1u; if (ndof == 4) for (int k = 0; k < 3; k++) if (conf_t[ks][k] < 0.0) p->elem_symbits[e] |= (unsigned char)(2u << k); }
  }
  std::vector<int> h_ecol(slot_off[n_elem]); std::vector<unsigned char> h_ekind(slot_off[n_elem]);
  std::vector<int> h_ecol2; if (any_kind2) h_ecol2.assign(slot_off[n_elem], -1);
  for (int s = 0; s < n_elem; s++) {
    int e = p->elem_of_slot[s], nn = p->elems[e].nn;
    for (int j = 0; j < nn; j++) for (int k = 0; k < ndof; k++) {
      int node = elem_node[elem_ptr[e] + j];
      int ct = ctype[ndof * node + k];
      int col = (ct == 0) ? col_t[ndof * node + k] : col_u[ndof * node + k];
      if (col < 0 || col >= n_dof) { mfb_problem_free(p); return fail(MFB_ERR_ARG, "mfb_harela3d_setup: missing column for an unknown"); }
      h_ecol[slot_off[s] + j * ndof + k] = p->colperm[col]; h_ekind[slot_off[s] + j * ndof + k] = (unsigned char)ct;
      if (ct == 2) h_ecol2[slot_off[s] + j * ndof + k] = p->colperm[col_t[ndof * node + k]];
    }
  }
  int *d_ecol, *d_slot_off, *d_ecol2 = nullptr; unsigned char* d_ekind; double* d_ecv;
  UP(p->owned, h_ecol, &d_ecol); UP(p->owned, h_ekind, &d_ekind); UP(p->owned, slot_off, &d_slot_off);
  if (any_kind2) UP(p->owned, h_ecol2, &d_ecol2);
  CK(cudaMalloc((void**)&d_ecv, (size_t)slot_off[n_elem] * 2 * sizeof(double))); p->owned.push_back(d_ecv);

  // ---- groups: geometry, point sets ----
  for (auto& g : p->groups) {
    DevGroup& D = g.dev;
    D.et = g.et; D.nn = g.nn; D.n_elem = g.n_elem; D.slot0 = g.slot0; D.ndof = ndof;
    std::vector<double> h_xn((size_t)g.n_elem * 3 * g.nn), h_ball((size_t)g.n_elem * 5);
    std::vector<int> h_enode((size_t)g.n_elem * g.nn), h_glnfar(g.n_elem);
    std::vector<unsigned char> h_rev(g.n_elem), h_info(g.n_elem);
    for (int i = 0; i < g.n_elem; i++) {
      int e = g.elem_ids[i]; const mfbh::Elem& el = p->elems[e];
      {
        unsigned info = 8u | ((el.reversed && ndof != 4) ? 16u : 0u);
        // multipliers of a symmetry image: elastic bits 5-7 = symconf_t(k); fluid bit 5 = symconf_s; poroelastic bits 4-7 = symconf_s, symconf_t(1:3)
        if (ndof == 3) { for (int k = 0; k < 3; k++) if (conf_t[e / n_root][k] < 0.0) info |= 32u << k; }
        else if (ndof == 1) { if (conf_s[e / n_root] < 0.0) info |= 32u; }
        else { if (conf_s[e / n_root] < 0.0) info |= 16u; for (int k = 0; k < 3; k++) if (conf_t[e / n_root][k] < 0.0) info |= 32u << k; }
        for (int k = 0; k < ndof; k++) {
          const int ct0 = ctype[ndof * elem_node[elem_ptr[e]] + k];
          for (int j = 1; j < g.nn; j++) if (ctype[ndof * elem_node[elem_ptr[e] + j] + k] != ct0) info &= ~8u;
          if (ct0 == 1 && k < 3) info |= (1u << k);   // bits 0-2 only (bit 3 is the uniform-kinds flag; the class bits serve the elastic K1 kernel)
        }
        h_info[i] = (unsigned char)info;
      }
      for (int q = 0; q < 3 * g.nn; q++) h_xn[(size_t)i * 3 * g.nn + q] = el.x[q];
      for (int j = 0; j < g.nn; j++) h_enode[(size_t)i * g.nn + j] = elem_node[elem_ptr[e] + j];
      h_ball[5 * (size_t)i] = el.bc[0]; h_ball[5 * (size_t)i + 1] = el.bc[1]; h_ball[5 * (size_t)i + 2] = el.bc[2]; h_ball[5 * (size_t)i + 3] = el.br; h_ball[5 * (size_t)i + 4] = el.cl;
      h_glnfar[i] = el.gln_far; h_rev[i] = el.reversed ? 1 : 0;
    }
    double *d_xn, *d_ball; int *d_enode, *d_glnfar; unsigned char *d_rev, *d_info, *d_cvnz;
    UP(g.owned, h_xn, &d_xn); UP(g.owned, h_ball, &d_ball); UP(g.owned, h_enode, &d_enode); UP(g.owned, h_glnfar, &d_glnfar); UP(g.owned, h_rev, &d_rev);
    UP(g.owned, h_info, &d_info);
    CK(cudaMalloc((void**)&d_cvnz, (size_t)g.n_elem)); g.owned.push_back(d_cvnz);
    {
      // K1 element ranges: K1_ERANGE elements for the first 4/5 of the group, a quarter of that for the rest (the dynamic
      // schedule hands out the short tasks last, which trims the tail of the kernel)
      // Small meshes: shorter ranges, so that the (tile x range) task list still covers the device several times over (BASELINE config C1, 744
      // elements x 20-odd tiles, had 90 tasks for 1184 warp slots with 512-element ranges: K1 took 7 ms of latency-bound single warps).
      int erange = K1_ERANGE;
      {
        const long long want = (long long)g.n_elem * n_tiles / 8192;      // ~8k tasks per group
        while (erange > 16 && erange > want) erange >>= 1;
        if (const char* e_r = getenv("MFB_K1_ERANGE")) { const int v = atoi(e_r); if (v >= 8 && v <= K1_ERANGE) erange = v; }
      }
      std::vector<int> rs(1, 0), rof(g.n_elem);
      const int tail_from = g.n_elem - g.n_elem / 5;
      for (int e = 0; e < g.n_elem;) {
        int len = (e < tail_from) ? erange : std::max(erange / 4, 8);
        if (e < tail_from && e + len > tail_from) len = tail_from - e;
        const int e1 = std::min(g.n_elem, e + len);
        for (int q = e; q < e1; q++) rof[q] = (int)rs.size() - 1;
        rs.push_back(e1); e = e1;
      }
      int *d_rs, *d_rof, *d_rm;
      UP(g.owned, rs, &d_rs); UP(g.owned, rof, &d_rof);
      D.n_ranges = (int)rs.size() - 1; D.range_start = d_rs; D.range_of = d_rof;
      CK(cudaMalloc((void**)&d_rm, sizeof(int) * D.n_ranges)); g.owned.push_back(d_rm); D.range_modes = d_rm;
    }
    D.einc = nullptr; D.c10 = nullptr; D.nfn = nullptr; D.ecol2 = d_ecol2 ? d_ecol2 + slot_off[g.slot0] : nullptr;
    D.xn = d_xn; D.ball = d_ball; D.enode = d_enode; D.gln_far = d_glnfar; D.erev = d_rev; D.einfo = d_info; D.ecvnz = d_cvnz;
    D.ecol = d_ecol + slot_off[g.slot0]; D.ekind = d_ekind + slot_off[g.slot0]; D.ecv = d_ecv + 2 * (size_t)slot_off[g.slot0];
    D.has_mixed = 0;
    for (int i = 0; i < g.n_elem; i++) if (!(h_info[i] & 8u)) D.has_mixed = 1;
    D.cols3 = (ndof == 3 && !any_kind2) ? 1 : 0;
    for (int i = 0; i < g.n_elem && D.cols3; i++) {
      const int so = slot_off[g.slot0 + i];
      for (int j = 0; j < g.nn; j++) for (int k = 1; k < 3; k++) if (h_ecol[so + j * 3 + k] != h_ecol[so + j * 3] + k) D.cols3 = 0;
    }
    D.n_sets = n_precalsets;
    for (int s = 0; s < n_precalsets; s++) {
      int gln = precalset_gln[s];
      if (gln < 1 || gln > 30) { mfb_problem_free(p); return fail(MFB_ERR_ARG, "mfb_harela3d_setup: precalset gln out of range"); }
      int ngp = mfbh::pointset_size(g.et, gln), rec = 6 + g.nn;
      D.set_gln[s] = gln; D.ngp[s] = ngp;
      std::vector<double> h_pts((size_t)g.n_elem * ngp * rec);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < g.n_elem; i++) mfbh::build_pointset(p->elems[g.elem_ids[i]], gln, &h_pts[(size_t)i * ngp * rec]);
      double* d_pts; UP(g.owned, h_pts, &d_pts); D.pts[s] = d_pts;
      CK(cudaStreamSynchronize(st));  // h_pts goes out of scope
    }
    CK(cudaStreamSynchronize(st));
  }

  if (!c10.empty()) {
    unsigned char* d_c10; UP(p->owned, c10, &d_c10);
    CK(cudaMalloc((void**)&p->d_nfn, (size_t)3 * n_node * sizeof(double))); p->owned.push_back(p->d_nfn);
    for (auto& g : p->groups) { g.dev.c10 = d_c10; g.dev.nfn = p->d_nfn; }
    p->need_normals = true;
    p->node_rev.assign(n_node, 0);
    for (int e = 0; e < n_root; e++) if (p->elems[e].reversed) for (int k = elem_ptr[e]; k < elem_ptr[e + 1]; k++) p->node_rev[elem_node[k]] = 1;
  }
  // ---- K0: classify all pairs on the device, fetch the near list ----
  for (int n = 0; n < 32; n++) p->cls.far_thr[n] = S.far_thr[n];
  p->cls.far_dmax = S.far_dmax;
  p->cls.ps_gln_max = *std::max_element(S.ps_gln.begin(), S.ps_gln.end());
  CK(cudaMalloc((void**)&p->plan, (size_t)n_elem * ldp)); p->owned.push_back(p->plan);
  for (auto& g : p->groups) launch_classify(g.dev, p->colloc, p->cls, p->plan, st);
  unsigned long long* d_counter; CK(cudaMalloc((void**)&d_counter, 8)); p->owned.push_back(d_counter);
  CK(cudaMemsetAsync(d_counter, 0, 8, st));
  launch_count_near(p->plan, n_elem, p->colloc, d_counter, nullptr, 0, st);
  unsigned long long n_near = 0;
  CK(cudaMemcpyAsync(&n_near, d_counter, 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
  std::vector<int2> near(n_near);
  if (n_near > 0) {
    int2* d_list; CK(cudaMalloc((void**)&d_list, n_near * sizeof(int2)));
    CK(cudaMemsetAsync(d_counter, 0, 8, st));
    launch_count_near(p->plan, n_elem, p->colloc, d_counter, d_list, n_near, st);
    CK(cudaMemcpyAsync(near.data(), d_list, n_near * sizeof(int2), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    cudaFree(d_list);
  }
  std::sort(near.begin(), near.end(), [](const int2& a, const int2& b) { return a.y != b.y ? a.y < b.y : a.x < b.x; });

  // ---- host planning of the near pairs ----
  std::vector<mfbh::NearPlan> plans(n_near);
#pragma omp parallel for schedule(dynamic, 16)
  for (long long i = 0; i < (long long)n_near; i++) {
    int cpos = near[i].x, slot = near[i].y;
    double xi[3] = {h_cx[cpos], h_cx[(size_t)ldp + cpos], h_cx[2 * (size_t)ldp + cpos]};
    mfbh::plan_near_pair(p->elems[p->elem_of_slot[slot]], xi, S, plans[i]);
  }
  std::vector<int> pc_cpos, pc_slot; std::vector<unsigned char> pc_val;
  long long pts_regular_near = 0, n_adp = 0, n_leaves = 0, pts_adp = 0, n_sing = 0, pts_sing = 0;
  for (auto& g : p->groups) {
    std::vector<int> a_cpos, a_elem, a_leaf0(1, 0), a_gln; std::vector<double> a_leafd;
    std::vector<int> s_cpos, s_elem, s_ray0(1, 0); std::vector<double> s_d, s_rays;
    for (unsigned long long i = 0; i < n_near; i++) {
      int cpos = near[i].x, slot = near[i].y;
      if (slot < g.slot0 || slot >= g.slot0 + g.n_elem) continue;
      const mfbh::NearPlan& np = plans[i];
      pc_cpos.push_back(cpos); pc_slot.push_back(slot);
      if (np.mode == 0) { pc_val.push_back((unsigned char)np.set); pts_regular_near += np.points; }
      else if (np.mode == 1) {
        pc_val.push_back(PLAN_ADAPTIVE);
        a_cpos.push_back(cpos); a_elem.push_back(slot - g.slot0);
        for (const auto& lf : np.leaves) {
          for (int q = 0; q < 8; q++) a_leafd.push_back(lf.xi_s[q]);
          for (int q = 0; q < 4; q++) a_leafd.push_back(lf.tp1[q]);
          for (int q = 0; q < 4; q++) a_leafd.push_back(lf.tp2[q]);
          a_gln.push_back(lf.gln);
        }
        a_leaf0.push_back((int)a_gln.size());
        n_adp++; n_leaves += (long long)np.leaves.size(); pts_adp += np.points;
      } else {
        pc_val.push_back(PLAN_SINGULAR);
        s_cpos.push_back(cpos); s_elem.push_back(slot - g.slot0);
        s_d.push_back(np.xi_i[0]); s_d.push_back(np.xi_i[1]);
        for (int q = 0; q < 3; q++) s_d.push_back(np.x_i[q]);
        for (int q = 0; q < 9; q++) s_d.push_back(np.hli[q]);
        for (const auto& r : np.rays) { s_rays.push_back(r.ct); s_rays.push_back(r.st); s_rays.push_back(r.rhoij); s_rays.push_back(r.w); }
        s_ray0.push_back((int)(s_rays.size() / 4));
        n_sing++; pts_sing += np.points;
      }
    }
    int *d1, *d2, *d3, *d4; double* d5;
    UP(g.owned, a_cpos, &d1); UP(g.owned, a_elem, &d2); UP(g.owned, a_leaf0, &d3); UP(g.owned, a_gln, &d4); UP(g.owned, a_leafd, &d5);
    g.adp.n_pairs = (int)a_cpos.size(); g.adp.pair_cpos = d1; g.adp.pair_elem = d2; g.adp.pair_leaf0 = d3; g.adp.leaf_gln = d4; g.adp.leaf_d = d5;
    int *e1, *e2, *e3; double *e4, *e5;
    UP(g.owned, s_cpos, &e1); UP(g.owned, s_elem, &e2); UP(g.owned, s_ray0, &e3); UP(g.owned, s_d, &e4); UP(g.owned, s_rays, &e5);
    g.sing.n_pairs = (int)s_cpos.size(); g.sing.pair_cpos = e1; g.sing.pair_elem = e2; g.sing.pair_ray0 = e3; g.sing.pair_d = e4; g.sing.rays = e5;
    CK(cudaStreamSynchronize(st));
  }
  {
    int *d1, *d2; unsigned char* d3;
    std::vector<void*> tmp;
    UP(tmp, pc_cpos, &d1); UP(tmp, pc_slot, &d2); UP(tmp, pc_val, &d3);
    launch_patch_plan(p->plan, p->colloc, (int)pc_cpos.size(), d1, d2, d3, st);
    CK(cudaStreamSynchronize(st));
    for (void* q : tmp) cudaFree(q);
  }

  if (p->hbie && n_sing > 0) { mfb_problem_free(p); return fail(MFB_ERR_UNSUPPORTED, "mfb_harela3d_setup_hbie: a collocation point lies on an element (hypersingular interior integration is not built)"); }
  // ---- free-term entries (geometry only): src/build_lse_mechanics_bem_harela.f90:273-747 ----
  {
    // node -> (element, local node) incidences
    std::vector<int> cnt(n_node + 1, 0);
    for (int e = 0; e < n_root; e++) for (int k = elem_ptr[e]; k < elem_ptr[e + 1]; k++) cnt[elem_node[k] + 1]++;
    for (int i = 0; i < n_node; i++) cnt[i + 1] += cnt[i];
    std::vector<int> n2e(cnt[n_node]), n2k(cnt[n_node]), pos(cnt.begin(), cnt.end() - 1);
    for (int e = 0; e < n_root; e++) for (int k = elem_ptr[e]; k < elem_ptr[e + 1]; k++) { int nd = elem_node[k]; n2e[pos[nd]] = e; n2k[pos[nd]] = k - elem_ptr[e]; pos[nd]++; }
    std::vector<int> f_cpos, f_slot, f_jk, f_l; std::vector<double> f_val;
    std::vector<int> f0_cpos, f0_slot, f0_jk, f0_l; std::vector<double> f0_val;      // poroelastic fluid phase: value = beta * J
    for (int c = 0; c < n_colloc; c++) {
      int e = colloc_elem[c], kn = colloc_kn[c], sn = colloc_node[c];
      if (e == -1) continue;   // a point off the boundary (interior point of the region): Somigliana's identity has no free term there
      if (e < 0 || e >= n_root || kn < 0 || kn >= p->elems[e].nn) { mfb_problem_free(p); return fail(MFB_ERR_ARG, "mfb_harela3d_setup: invalid colloc_elem/colloc_kn"); }
      const mfbh::Elem& el = p->elems[e];
      int cpos = p->cpos_of_colloc[c], slot = p->slot_of_elem[e];
      bool mca = !(colloc_xi[2 * c] == -9.0 && colloc_xi[2 * c + 1] == -9.0);
      if (!mca) {
        double xi_i[2]; mfbh::node_xi(el.et, kn, xi_i);
        double cp = 0.5, sum_b[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (mfbh::xi_on_element_boundary(el.et, xi_i)) {
          int b0 = cnt[sn], ne = cnt[sn + 1] - b0;
          // planes that contain the node (fbem_node_symplanes_connectivity, lib/fbem/src/data_structures.f90:1116-1153): the fan of elements around it is
          // completed with the mirrored normals and tangents (build_lse_mechanics_bem_harela.f90:496-555; a single reflection takes the tangent of the
          // opposite orientation, a double one keeps it)
          int planes[3], npl = 0;
          if (sym) for (int i = 0; i < sym->n_planes; i++) if (fabs(node_x[3 * (size_t)sn + sym->eid[i] - 1]) <= geometric_tolerance) planes[npl++] = i;
          if (npl > 2) { mfb_problem_free(p); return fail(MFB_ERR_UNSUPPORTED, "mfb_harela3d_setup: a nodal collocation point lies in three symmetry planes"); }
          const int fan = ne << npl;
          std::vector<double> ns(3 * fan), ts(3 * fan), tr(3 * ne), nr(3);
          for (int k = 0; k < ne; k++) {
            const mfbh::Elem& ee = p->elems[n2e[b0 + k]];
            mfbh::node_normal_tangent(ee.et, ee.x, n2k[b0 + k], el.reversed, &ns[3 * k], &ts[3 * k]);
            if (npl) mfbh::node_normal_tangent(ee.et, ee.x, n2k[b0 + k], !el.reversed, nr.data(), &tr[3 * k]);
          }
          for (int k = 0; k < ne && npl; k++) for (int cc = 0; cc < 3; cc++) {
            const double m1 = plane_m[planes[0]][cc];
            ns[3 * (k + ne) + cc] = m1 * ns[3 * k + cc]; ts[3 * (k + ne) + cc] = m1 * tr[3 * k + cc];
            if (npl == 2) {
              const double m2 = plane_m[planes[1]][cc];
              ns[3 * (k + 2 * ne) + cc] = m1 * m2 * ns[3 * k + cc]; ts[3 * (k + 2 * ne) + cc] = m1 * m2 * ts[3 * k + cc];
              ns[3 * (k + 3 * ne) + cc] = m2 * ns[3 * k + cc]; ts[3 * (k + 3 * ne) + cc] = m2 * tr[3 * k + cc];
            }
          }
          if (mfbh::mantic_terms(fan, ns.data(), ts.data(), geometric_tolerance, &cp, sum_b)) { mfb_problem_free(p); return fail(2, "mfb_harela3d_setup: the normals/tangents configuration is not valid (free term)"); }
        }
        if (ndof == 4) {   // c(0,0) = J c_pot, c(1:3,1:3) = Mantic's matrix of the drained skeleton (build_lse_mechanics_bem_harpor.f90:583-615)
          f0_cpos.push_back(cpos); f0_slot.push_back(slot); f0_jk.push_back(kn * 4); f0_l.push_back(0); f0_val.push_back(0.0); f0_val.push_back(cp);
          for (int l = 0; l < 3; l++) for (int k = 0; k < 3; k++) {
            f_cpos.push_back(cpos); f_slot.push_back(slot); f_jk.push_back(kn * 4 + k + 1); f_l.push_back(l + 1);
            f_val.push_back(l == k ? cp : 0.0); f_val.push_back(sum_b[3 * l + k]);
          }
        } else if (ndof == 1) {   // scalar free term c = cp (fbem_bem_pot3d_sbie_freeterm == the isotropic part of Mantic's matrix; build_lse_mechanics_bem_harpot.f90:533,549)
          f_cpos.push_back(cpos); f_slot.push_back(slot); f_jk.push_back(kn); f_l.push_back(0);
          f_val.push_back(cp); f_val.push_back(0.0);
        } else
        for (int l = 0; l < 3; l++) for (int k = 0; k < 3; k++) {
          f_cpos.push_back(cpos); f_slot.push_back(slot); f_jk.push_back(kn * 3 + k); f_l.push_back(l);
          f_val.push_back(l == k ? cp : 0.0); f_val.push_back(sum_b[3 * l + k]);
        }
      } else {
        double phi[9]; mfbh::shape_values(el.et, &colloc_xi[2 * c], phi);
        if (ndof == 4) {   // hp(:,0,0) += J phi/2, hp(:,l,l) += phi/2 (build_lse_mechanics_bem_harpor.f90:366-370)
          for (int j = 0; j < el.nn; j++) {
            f0_cpos.push_back(cpos); f0_slot.push_back(slot); f0_jk.push_back(j * 4); f0_l.push_back(0); f0_val.push_back(0.0); f0_val.push_back(0.5 * phi[j]);
            for (int l = 1; l < 4; l++) { f_cpos.push_back(cpos); f_slot.push_back(slot); f_jk.push_back(j * 4 + l); f_l.push_back(l); f_val.push_back(0.5 * phi[j]); f_val.push_back(0.0); }
          }
        } else
        for (int l = 0; l < ndof; l++) for (int j = 0; j < el.nn; j++) {   // hp(:)=hp(:)+0.5d0*pphi_i (build_lse_mechanics_bem_harpot.f90:573)
          f_cpos.push_back(cpos); f_slot.push_back(slot); f_jk.push_back(j * ndof + l); f_l.push_back(l);
          f_val.push_back(0.5 * phi[j]); f_val.push_back(0.0);
        }
      }
    }
    int *d1, *d2, *d3, *d4; double* d5;
    UP(p->owned, f_cpos, &d1); UP(p->owned, f_slot, &d2); UP(p->owned, f_jk, &d3); UP(p->owned, f_l, &d4); UP(p->owned, f_val, &d5);
    p->ft.n = (int)f_cpos.size(); p->ft.cpos = d1; p->ft.slot = d2; p->ft.jk = d3; p->ft.l = d4; p->ft.val = d5;
    p->ft.slot_off = d_slot_off; p->ft.ecol = d_ecol; p->ft.ekind = d_ekind; p->ft.ecv = d_ecv; p->ft.einc = nullptr;
    {
      int *e1, *e2, *e3, *e4; double* e5;
      UP(p->owned, f0_cpos, &e1); UP(p->owned, f0_slot, &e2); UP(p->owned, f0_jk, &e3); UP(p->owned, f0_l, &e4); UP(p->owned, f0_val, &e5);
      p->ft0 = p->ft; p->ft0.n = (int)f0_cpos.size(); p->ft0.cpos = e1; p->ft0.slot = e2; p->ft0.jk = e3; p->ft0.l = e4; p->ft0.val = e5;
    }
    CK(cudaStreamSynchronize(st));
  }

  // ---- device-resident system ----
  p->lda = ((long long)n_dof + 31) / 32 * 32;
  double* dA; CK(cudaMalloc((void**)&dA, (size_t)2 * p->lda * n_dof * sizeof(double))); p->owned.push_back(dA);
  double* db; CK(cudaMalloc((void**)&db, (size_t)2 * p->lda * sizeof(double))); p->owned.push_back(db);
  CK(cudaMalloc((void**)&p->d_vstage, (size_t)2 * std::max(n_dof, 1) * sizeof(double)));
  p->sys.Are = dA; p->sys.Aim = dA + (size_t)p->lda * n_dof; p->sys.lda = p->lda; p->sys.n_dof = n_dof; p->sys.bre = db; p->sys.bim = db + p->lda;
  p->have_tmap = make_matrix_tensor_map(p->tmapA, p->sys.Are, p->lda, n_dof, 2) == 0;
  p->have_tmapS = make_matrix_tensor_map(p->tmapS, p->sys.Are, p->lda, n_dof, 1) == 0;
  p->real_resident = false;
  if (ndof == 3 && !p->have_tmap) { mfb_problem_free(p); return fail(MFB_ERR_CUDA, "mfb_harela3d_setup: cuTensorMapEncodeTiled failed (driver too old for sm_100a TMA?)"); }
  CK(cudaMalloc((void**)&p->d_cvalue, (size_t)2 * ndof * n_node * sizeof(double))); p->owned.push_back(p->d_cvalue);
  CK(cudaMalloc((void**)&p->d_ipiv, (size_t)n_dof * sizeof(int))); p->owned.push_back(p->d_ipiv);
  CK(cudaMalloc((void**)&p->d_perm, (size_t)n_dof * sizeof(int))); p->owned.push_back(p->d_perm);
  p->h_ipiv.assign(n_dof, 0);

  // statistics of the plan
  {
    std::vector<unsigned char> h_plan((size_t)n_elem * ldp);
    CK(cudaMemcpy(h_plan.data(), p->plan, h_plan.size(), cudaMemcpyDeviceToHost));
    long long pairs = 0, pts = 0; double flops = 0.0;
    for (auto& g : p->groups)
      for (int i = 0; i < g.n_elem; i++)
        for (int q = 0; q < ldp; q++) {
          unsigned char m = h_plan[(size_t)(g.slot0 + i) * ldp + q];
          // algorithmic flops per Gauss point / pair: elastic 585 + 72 n / 144 n; scalar 96 + 8 n / 12 n
          if (m < MAX_SETS) { pairs++; pts += g.dev.ngp[m]; flops += (ndof == 4) ? (double)g.dev.ngp[m] * (1500.0 + 128.0 * g.nn) + 256.0 * g.nn : (ndof == 1) ? (double)g.dev.ngp[m] * (96.0 + 8.0 * g.nn) + 12.0 * g.nn : (double)g.dev.ngp[m] * (585.0 + 72.0 * g.nn) + 144.0 * g.nn; }
        }
    p->stats[MFB_STAT_PAIRS_REGULAR] = (double)pairs; p->stats[MFB_STAT_POINTS_REGULAR] = (double)pts; p->stats[MFB_STAT_FLOPS_REGULAR] = flops;
    p->stats[MFB_STAT_PAIRS_ADAPTIVE] = (double)n_adp; p->stats[MFB_STAT_LEAVES] = (double)n_leaves; p->stats[MFB_STAT_POINTS_ADAPTIVE] = (double)pts_adp;
    p->stats[MFB_STAT_PAIRS_SINGULAR] = (double)n_sing; p->stats[MFB_STAT_POINTS_SINGULAR] = (double)pts_sing; p->stats[MFB_STAT_NEAR_PAIRS] = (double)n_near;
  }
  p->stats[MFB_STAT_MS_SETUP_HOST] = now_ms() - t_host0;
  *out = p;
  return MFB_OK;
}

// fbem_bem_harela3d_calculate_parameters (SBIE subset): lib/fbem/src/bem_harela3d.f90:143-287
static void host_kparams(cd lambda, cd mu, double rho, double omega, KParams& K) {
  const cd im(0.0, 1.0);
  cd c1 = std::sqrt((lambda + 2.0 * mu) / rho), c2 = std::sqrt(mu / rho);
  cd k1 = omega / c1, k2 = omega / c2;
  cd c1_2 = c1 * c1, c2_2 = c2 * c2, c1_3 = c1_2 * c1, c1_4 = c1_2 * c1_2, k2_2 = k2 * k2;
  cd ik1 = im * k1, ik2 = im * k2, ik1_2 = ik1 * ik1, ik2_2 = ik2 * ik2, r = c2_2 / c1_2;
  double om2 = omega * omega;
  cd psi[7], chi[7], T1[11], T2[10], T3[10];
  psi[1] = 0.5 * (1.0 + r); psi[2] = -1.0 / 3.0 * (2.0 / c2 + c2_2 / c1_3) * im * omega; psi[3] = im * k1 / k2_2;
  psi[4] = 1.0 / ik2; psi[5] = 1.0 / k2_2; psi[6] = -1.0 / k2_2;
  chi[1] = -0.5 * (1.0 - r); chi[2] = -r; chi[3] = -3.0 * r / ik1; chi[4] = 3.0 / ik2; chi[5] = -3.0 * r / ik1_2; chi[6] = 3.0 / ik2_2;
  T1[1] = 3.0 * (r - 1.0); T1[2] = -0.25 * (1.0 / c2_2 - c2_2 / c1_4) * om2; T1[3] = -2.0 * im * k1 * r; T1[4] = 2.0 * im * k2;
  T1[5] = -12.0 * r; T1[6] = 12.0; T1[7] = -r * 30.0 / ik1; T1[8] = 30.0 / ik2; T1[9] = -r * 30.0 / ik1_2; T1[10] = 30.0 / ik2_2;
  T2[1] = -r; T2[2] = -0.25 * (1.0 / c2_2 + c2_2 / c1_4) * om2; T2[3] = -im * k2; T2[4] = 2.0 * r; T2[5] = -3.0;
  T2[6] = r * 6.0 / ik1; T2[7] = -6.0 / ik2; T2[8] = r * 6.0 / ik1_2; T2[9] = -6.0 / ik2_2;
  T3[1] = r; T3[2] = 0.25 * (3.0 * c2_2 / c1_4 - 2.0 / c1_2 + 1.0 / c2_2) * om2; T3[3] = (2.0 * r - 1.0) * im * k1; T3[4] = 4.0 * r - 1.0;
  T3[5] = -2.0; T3[6] = r * 6.0 / ik1; T3[7] = -6.0 / ik2; T3[8] = r * 6.0 / ik1_2; T3[9] = -6.0 / ik2_2;
  auto cv = [](cd z) { return mk(z.real(), z.imag()); };
  memset(&K, 0, sizeof(K));
  K.k1 = cv(k1); K.k2 = cv(k2);
  for (int i = 1; i <= 6; i++) { K.psi[i] = cv(psi[i]); K.chi[i] = cv(chi[i]); }
  for (int i = 1; i <= 10; i++) K.T1[i] = cv(T1[i]);
  for (int i = 1; i <= 9; i++) { K.T2[i] = cv(T2[i]); K.T3[i] = cv(T3[i]); }
  const double c_1_4pi = 0.07957747154594767280411105048;
  K.cte_u = cv(c_1_4pi / mu); K.cte_t = c_1_4pi;
  // hypersingular kernels: S1..S5, cte_d, cte_s (bem_harela3d.f90:219-287)
  {
    const cd k1_2 = k1 * k1, c1_5 = c1_4 * c1, c2_3 = c2_2 * c2; const double om3 = om2 * omega;
    cd S1[12], S2[12], S3[13], S4[11], S5[12];
    S1[1] = 3.0 * (1.0 - 2.0 * c2_2 / c1_2); S1[2] = -0.5 * c2_2 / c1_4 * om2; S1[3] = k2_2; S1[4] = 4.0 * im * k1 * k1_2 / k2_2;
    S1[5] = -7.0 * im * k2; S1[6] = 24.0 * k1_2 / k2_2; S1[7] = -27.0; S1[8] = 60.0 * k1_2 / k2_2 / ik1; S1[9] = -60.0 / ik2;
    S1[10] = 60.0 * k1_2 / k2_2 / ik1_2; S1[11] = -60.0 / ik2_2;
    S2[1] = 6.0 * c2_2 / c1_2; S2[2] = (0.5 / c2_2 + 1.5 * c2_2 / c1_4 - 1.0 / c1_2) * om2; S2[3] = 2.0 * c2_2 / c1_4 * (c1_2 / c2_2 - 2.0) * om2;
    S2[4] = 2.0 * im * c2_2 / c1_3 * (8.0 - 3.0 * c1_2 / c2_2) * omega; S2[5] = -4.0 * im * k2; S2[6] = 6.0 * c2_2 / c1_2 * (6.0 - c1_2 / c2_2);
    S2[7] = -24.0; S2[8] = c2_2 / c1_2 * 60.0 / ik1; S2[9] = -60.0 / ik2; S2[10] = 60.0 / ik2_2; S2[11] = -60.0 / ik2_2;
    S3[1] = 30.0 * (1.0 - c2_2 / c1_2); S3[2] = 1.5 * (1.0 / c2_2 - c2_2 / c1_4) * om2; S3[3] = -4.0 * c2_2 / c1_2 * k1_2; S3[4] = 4.0 * k2_2;
    S3[5] = 40.0 * c2_2 / c1_2 * im * k1; S3[6] = -40.0 * im * k2; S3[7] = 180.0 * c2_2 / c1_2; S3[8] = -180.0;
    S3[9] = 420.0 * c2_2 / c1_2 / ik1; S3[10] = -420.0 / ik2; S3[11] = 420.0 / ik2_2; S3[12] = -420.0 / ik2_2;
    S4[1] = 2.0 * c2_2 / c1_2; S4[2] = 0.5 * (c2_2 / c1_4 + 1.0 / c2_2) * om2; S4[3] = -2.0 / 5.0 * (1.0 / c2_3 + 2.0 / 3.0 * c2_2 / c1_5) * im * om3;
    S4[4] = 2.0 * im * k2; S4[5] = -4.0 * c2_2 / c1_2; S4[6] = 6.0; S4[7] = -12.0 * c2_2 / c1_2 / ik1; S4[8] = 12.0 / ik2;
    S4[9] = -12.0 / ik2_2; S4[10] = 12.0 / ik2_2;
    S5[1] = 2.0 * (1.0 - 3.0 * c2_2 / c1_2); S5[2] = (-2.0 / c1_2 + 0.5 / c2_2 + 0.5 * c2_2 / c1_4) * om2;
    S5[3] = (8.0 / 3.0 / c1_3 + 4.0 / 15.0 / c2_3 - 1.0 / c1 / c2_2 - 24.0 / 15.0 * c2_2 / c1_5) * im * om3;
    S5[4] = (-4.0 / c1_2 + 1.0 / c2_2 + 4.0 * c2_2 / c1_4) * om2; S5[5] = 4.0 * im * (1.0 / c1 - 2.0 * c2_2 / c1_3) * omega;
    S5[6] = 4.0 * (1.0 - 3.0 * c2_2 / c1_2); S5[7] = 4.0; S5[8] = 12.0 * im * k1 / k2_2; S5[9] = -12.0 * im / k2; S5[10] = 12.0 / k2_2; S5[11] = -12.0 / k2_2;
    for (int i = 1; i <= 11; i++) { K.S1[i] = cv(S1[i]); K.S2[i] = cv(S2[i]); K.S5[i] = cv(S5[i]); }
    for (int i = 1; i <= 12; i++) K.S3[i] = cv(S3[i]);
    for (int i = 1; i <= 10; i++) K.S4[i] = cv(S4[i]);
    K.cte_d = c_1_4pi; K.cte_s = cv(c_1_4pi * mu);
  }
}
// pre-scaled copy for the regular kernel: psi, chi times -cte_u (slot 0 = coefficient of the bare E2(z2)/r term), T times cte_t
static void scale_kparams(const KParams& K, KParams& Q) {
  Q = K;
  const cplx mu_ = mk(-K.cte_u.re, -K.cte_u.im);
  Q.psi[0] = mu_; Q.chi[0] = mu_;
  for (int i = 1; i <= 6; i++) { Q.psi[i] = K.psi[i] * mu_; Q.chi[i] = K.chi[i] * mu_; }
  for (int i = 1; i <= 10; i++) Q.T1[i] = K.T1[i] * K.cte_t;
  for (int i = 1; i <= 9; i++) { Q.T2[i] = K.T2[i] * K.cte_t; Q.T3[i] = K.T3[i] * K.cte_t; }
}

// Static elasticity as the parameter set of the harmonic kernels: only the 1/r (psi, chi) and 1/r^2 (T1..T3) coefficients
// survive, k1 = k2 = 0 (every E_m term vanishes), all real.  With r_c = c2^2/c1^2 = (1-2nu)/(2(1-nu)) these are Kelvin's
// u*, t* exactly as fbem_bem_staela3d_sbie_u/_t write them (lib/fbem/src/bem_staela3d.f90:408-461):
// cte_u psi(1) = (3-4nu)/(16 pi mu (1-nu)), -cte_u chi(1) = 1/(16 pi mu (1-nu)), cte_t T1(1) = -3/(8 pi (1-nu)), cte_t T2(1) = -(1-2nu)/(8 pi (1-nu)) = -cte_t T3(1).
static void host_kparams_static(double mu, double nu, KParams& K) {
  memset(&K, 0, sizeof(K));
  const double rc = (1.0 - 2.0 * nu) / (2.0 * (1.0 - nu));
  K.psi[1] = mk(0.5 * (1.0 + rc), 0.0); K.chi[1] = mk(-0.5 * (1.0 - rc), 0.0);
  K.T1[1] = mk(3.0 * (rc - 1.0), 0.0); K.T2[1] = mk(-rc, 0.0); K.T3[1] = mk(rc, 0.0);
  const double c_1_4pi = 0.07957747154594767280411105048;
  K.cte_u = mk(c_1_4pi / mu, 0.0); K.cte_t = c_1_4pi;
  // hypersingular kernels: the 1/r^3 coefficients alone (bem_harela3d.f90:219-280 at omega = 0) are d*, s* of
  // fbem_bem_staela3d_hbie_ext_pre (bem_staela3d.f90:4428-4443): S1 = 3 nu/(1-nu), S2 = 3(1-2nu)/(1-nu), S3 = 15/(1-nu), ...
  K.S1[1] = mk(3.0 * (1.0 - 2.0 * rc), 0.0); K.S2[1] = mk(6.0 * rc, 0.0); K.S3[1] = mk(30.0 * (1.0 - rc), 0.0);
  K.S4[1] = mk(2.0 * rc, 0.0); K.S5[1] = mk(2.0 * (1.0 - 3.0 * rc), 0.0);
  K.cte_d = c_1_4pi; K.cte_s = mk(c_1_4pi * mu, 0.0);
}
static int assemble_device_k(mfb_problem* p, const KParams& K, const KParams& Q, cd nu, const mfb_z* cvalue, bool statics);
static bool finite_c(cd z) { return std::isfinite(z.real()) && std::isfinite(z.imag()); }
static int assemble_device(mfb_problem* p, double omega, cd lambda, cd mu, double rho, cd nu, const mfb_z* cvalue) {
  // omega = 0 gives 1/(i k)^2 = inf in the kernel parameters and NaN in every entry: refuse what the reference's own input checks refuse
  if (!(omega > 0.0) || !std::isfinite(omega)) return fail(MFB_ERR_ARG, "harmonic assembly: omega must be positive and finite (the static problem is mfb_staela3d_*)");
  if (!(rho > 0.0) || !std::isfinite(rho) || !finite_c(lambda) || !finite_c(mu) || !finite_c(nu) || mu == cd(0.0, 0.0) || lambda + 2.0 * mu == cd(0.0, 0.0) || nu == cd(1.0, 0.0))
    return fail(MFB_ERR_ARG, "harmonic assembly: rho must be positive, lambda, mu, nu finite, mu and lambda + 2 mu nonzero, nu != 1");
  KParams K, Q; host_kparams(lambda, mu, rho, omega, K); scale_kparams(K, Q);
  return assemble_device_k(p, K, Q, nu, cvalue, false);
}
static int assemble_device_k(mfb_problem* p, const KParams& K, const KParams& Q, cd nu, const mfb_z* cvalue, bool statics) {
  if (p->ndof != 3) return fail(MFB_ERR_ARG, "this problem was set up for an inviscid fluid region (mfb_harpot3d_setup): use mfb_harpot3d_assemble / _solve_frequency");
  if (statics && p->have_inc) return fail(MFB_ERR_ARG, "static assembly: an incident field is set (harmonic analysis only); clear it with mfb_harela3d_set_incident(problem, NULL, NULL)");
  cudaStream_t st = p->ctx->stream;
  if (p->need_normals && !p->have_normals) return fail(MFB_ERR_ARG, "assembly: the model has ctype 10 conditions (normal pressure): give the nodal normals with mfb_set_node_normals first");
  if (cvalue) {
    CK(cudaMemcpyAsync(p->d_cvalue, cvalue, (size_t)6 * p->n_node * sizeof(double), cudaMemcpyHostToDevice, st));
    for (auto& g : p->groups) launch_gather_cv(g.dev, p->d_cvalue, st);
    p->have_cvalue = true;
  } else if (!p->have_cvalue) return fail(MFB_ERR_ARG, "cvalue == NULL but no prescribed values are resident yet");
  CK(cudaEventRecord(p->ev[0], st));
  CK(cudaMemsetAsync(p->sys.Are, 0, (size_t)2 * p->lda * p->n_dof * sizeof(double), st));
  CK(cudaMemsetAsync(p->sys.bre, 0, (size_t)2 * p->lda * sizeof(double), st));
  CK(cudaEventRecord(p->ev[1], st));
  {
    const void* tm = statics ? (p->have_tmapS ? p->tmapS : nullptr) : (p->have_tmap ? p->tmapA : nullptr);
    for (auto& g : p->groups) launch_regular(g.dev, p->colloc, p->sys, p->plan, tm, statics, K, Q, p->ctx->k1, st);
  }
  CK(cudaEventRecord(p->ev[2], st));
  for (auto& g : p->groups) launch_adaptive(g.dev, p->colloc, p->sys, g.adp, p->ctx->tables, K, st);
  CK(cudaEventRecord(p->ev[3], st));
  for (auto& g : p->groups) launch_singular(g.dev, p->colloc, p->sys, g.sing, p->ctx->tables, K, st);
  CK(cudaEventRecord(p->ev[4], st));
  const double c_pi = 3.14159265358979323846264338328;
  cd F = -1.0 / (8.0 * c_pi * (1.0 - nu));
  launch_freeterm(p->colloc, p->sys, p->ft, mk(F.real(), F.imag()), st);
  if (p->n_cond > 0 && !p->skip_cond) launch_add_entries(p->sys, p->n_cond, p->d_cond_row, p->d_cond_col, p->d_cond_val, st);   // the host's condition rows (mfb_set_condition_rows)
  CK(cudaEventRecord(p->ev[5], st));
  CK(cudaGetLastError());
  p->factored = false; p->assembled = true; p->rows_permuted = true; p->real_resident = statics;
  // our kernels only (memsets/copies are not counted): free term + per group regular, adaptive, singular (+ gather_cv)
  p->asm_launches = 1;
  for (auto& g : p->groups)   // K1: one kernel per element class on 3/4-node elements (classes 0, 1 and, if present, 2)
    p->asm_launches += (((g.et == 5 || g.et == 7) && g.dev.cols3) ? 2 + g.dev.has_mixed : 1) + (cvalue ? 1 : 0) + (g.adp.n_pairs > 0) + (g.sing.n_pairs > 0);
  return MFB_OK;
}
// Rows that the HOST writes into the system after the BEM assembly: the local-axes conditions of ctype 2 / 3 nodes (src/build_lse_mechanics_harmonic.f90:204-258:
// A_c(row(k,0), col) = n_fn / t1_fn / t2_fn components, b_c(row(k,0)) = cvalue).  entries: rows[i], cols[i] (host indices; cols = -1: the right-hand side),
// values[i]; they are ADDED after every later assembly of this problem (elastic harmonic or static), so that the fused solve_frequency sees them.  n = 0 clears.
extern "C" int mfb_set_condition_rows(mfb_problem* p, int n, const int* rows, const int* cols, const mfb_z* values) {
  if (!p || n < 0 || (n > 0 && (!rows || !cols || !values))) return fail(MFB_ERR_ARG, "mfb_set_condition_rows: invalid argument");
  for (int i = 0; i < n; i++) if (rows[i] < 0 || rows[i] >= p->n_dof || cols[i] < -1 || cols[i] >= p->n_dof) return fail(MFB_ERR_ARG, "mfb_set_condition_rows: index out of range");
  CK(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  CK(cudaStreamSynchronize(st));
  if (p->d_cond_row) { cudaFree(p->d_cond_row); cudaFree(p->d_cond_col); cudaFree(p->d_cond_val); p->d_cond_row = p->d_cond_col = nullptr; p->d_cond_val = nullptr; }
  p->n_cond = 0;
  if (n == 0) return MFB_OK;
  std::vector<int> r(n), c(n); std::vector<double> v(2 * (size_t)n);
  for (int i = 0; i < n; i++) { r[i] = p->rowperm[rows[i]]; c[i] = cols[i] < 0 ? -1 : p->colperm[cols[i]]; v[2 * i] = values[i].re; v[2 * i + 1] = values[i].im; }
  CK(cudaMalloc((void**)&p->d_cond_row, n * sizeof(int))); CK(cudaMalloc((void**)&p->d_cond_col, n * sizeof(int))); CK(cudaMalloc((void**)&p->d_cond_val, 2 * (size_t)n * sizeof(double)));
  CK(cudaMemcpy(p->d_cond_row, r.data(), n * sizeof(int), cudaMemcpyHostToDevice)); CK(cudaMemcpy(p->d_cond_col, c.data(), n * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p->d_cond_val, v.data(), 2 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  p->n_cond = n; p->assembled = false;
  return MFB_OK;
}
// Nodal unit normals node(sn)%n_fn(1:3) (src/build_data_at_functional_nodes.f90:355-400: the normalised sum of the normals of the elements around the node, mirror
// images included) for the nodes whose ctype is 10; n_fn[3 * n_node].  Needed once, before the first assembly of a model with such conditions.
extern "C" int mfb_set_node_normals(mfb_problem* p, const double* n_fn) {
  if (!p || !n_fn) return fail(MFB_ERR_ARG, "mfb_set_node_normals: null argument");
  if (!p->need_normals) return MFB_OK;   // no ctype 10 in this model: nothing uses them
  for (size_t i = 0; i < (size_t)3 * p->n_node; i++) if (!std::isfinite(n_fn[i])) return fail(MFB_ERR_ARG, "mfb_set_node_normals: non-finite value");
  CK(cudaSetDevice(p->ctx->device));
  std::vector<double> h(n_fn, n_fn + (size_t)3 * p->n_node);
  for (int v = 0; v < p->n_node; v++) if (p->node_rev[v]) for (int k = 0; k < 3; k++) h[3 * (size_t)v + k] = -h[3 * (size_t)v + k];   // `if (sb_int_reversion) b -= ...` (:101-105)
  CK(cudaMemcpyAsync(p->d_nfn, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, p->ctx->stream));
  CK(cudaStreamSynchronize(p->ctx->stream));
  p->have_normals = true;
  if (p->have_cvalue) for (auto& g : p->groups) launch_gather_cv(g.dev, p->d_cvalue, p->ctx->stream);
  return MFB_OK;
}
// Incident wave field of the region (region%n_incidentfields > 0): u_inc, t_inc at the nodes of every element, as the reference holds them in
// element(se)%incident_c(1:3,kn,1) and (4:6,kn,1) (t_inc belongs to the element: it is formed with the element's normal).  Every pair then adds
// hp u_inc - gp t_inc to b (assemble_bem_harela_equation.f90:651-666), the free term included (it is part of hp, build_lse_mechanics_bem_harela.f90:715).
// Arrays: [(elem_ptr[e] + j) * 3 + k] complex, e over the elements given to the set-up; both NULL: no incident field.  Valid until the next call
// (set it before the assembly of every frequency: the field depends on omega).  An image of a symmetric model takes the root's values times symconf_t(k).
static int set_incident_impl(mfb_problem* p, const mfb_z* u_inc, const mfb_z* t_inc, int want_ndof);
extern "C" int mfb_harela3d_set_incident(mfb_problem* p, const mfb_z* u_inc, const mfb_z* t_inc) { return set_incident_impl(p, u_inc, t_inc, 3); }
// The same for an inviscid fluid region: p_inc, Un_inc at the nodes of every element, index [elem_ptr[e] + j] (element()%incident_c(1,kn,1) / (2,kn,1));
// every pair and free term adds hp p_inc - gp Un_inc to b (src/assemble_bem_harpot_equation.f90:471-481; gp already carries rho omega^2).
extern "C" int mfb_harpot3d_set_incident(mfb_problem* p, const mfb_z* p_inc, const mfb_z* un_inc) { return set_incident_impl(p, p_inc, un_inc, 1); }
// and for a poroelastic region (src/assemble_bem_harpor_equation.f90:1277-1289): (tau, u_k)_inc and (Un, t_k)_inc, index [(elem_ptr[e] + j) * 4 + k]
extern "C" int mfb_harpor3d_set_incident(mfb_problem* p, const mfb_z* u_inc, const mfb_z* t_inc) { return set_incident_impl(p, u_inc, t_inc, 4); }
static int set_incident_impl(mfb_problem* p, const mfb_z* u_inc, const mfb_z* t_inc, int want_ndof) {
  if (!p) return fail(MFB_ERR_ARG, "mfb_harela3d_set_incident: null problem");
  if (p->ndof != want_ndof || p->hbie) return fail(MFB_ERR_UNSUPPORTED, "set_incident: built for the displacement equation of elastic regions (mfb_harela3d_set_incident) and for fluid regions (mfb_harpot3d_set_incident)");
  if ((u_inc == nullptr) != (t_inc == nullptr)) return fail(MFB_ERR_ARG, "mfb_harela3d_set_incident: give both u_inc and t_inc, or neither");
  CK(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  const int n_elem = p->n_elem, n_root = p->n_elem_root;
  const size_t total = (size_t)p->slot_off_h[n_elem];
  if (u_inc) {
    std::vector<double> h(4 * total);
    for (int s = 0; s < n_elem; s++) {
      const int e = p->elem_of_slot[s], r = e % n_root, nn = p->elems[e].nn, o = p->slot_off_h[s];
      const unsigned bits = p->elem_symbits[e];
      const int nd = p->ndof;
      for (int j = 0; j < nn; j++) for (int k = 0; k < nd; k++) {
        const size_t q = ((size_t)p->root_elem_ptr[r] + j) * nd + k;
        const double sg = ((bits >> k) & 1u) ? -1.0 : 1.0;
        if (!std::isfinite(u_inc[q].re) || !std::isfinite(u_inc[q].im) || !std::isfinite(t_inc[q].re) || !std::isfinite(t_inc[q].im))
          return fail(MFB_ERR_ARG, "mfb_harela3d_set_incident: non-finite value");
        double* d = &h[4 * ((size_t)o + j * nd + k)];
        d[0] = sg * u_inc[q].re; d[1] = sg * u_inc[q].im; d[2] = sg * t_inc[q].re; d[3] = sg * t_inc[q].im;
      }
    }
    if (!p->d_einc) { CK(cudaMalloc((void**)&p->d_einc, std::max<size_t>(total, 1) * 4 * sizeof(double))); p->owned.push_back(p->d_einc); }
    CK(cudaMemcpyAsync(p->d_einc, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // h goes out of scope
  }
  p->have_inc = u_inc != nullptr;
  for (auto& g : p->groups) g.dev.einc = p->have_inc ? p->d_einc + 4 * (size_t)p->slot_off_h[g.slot0] : nullptr;
  p->ft.einc = p->have_inc ? p->d_einc : nullptr; p->ft0.einc = p->ft.einc;
  // the element classes of K1 depend on it (every element runs as the general class while a field is set)
  if (p->have_cvalue) for (auto& g : p->groups) launch_gather_cv(g.dev, p->d_cvalue, st);
  p->assembled = false;
  return MFB_OK;
}
static int collect_assembly_times(mfb_problem* p) {
  CK(cudaEventSynchronize(p->ev[5]));
  float t;
  cudaEventElapsedTime(&t, p->ev[0], p->ev[1]); p->stats[MFB_STAT_MS_ZERO] = t;
  cudaEventElapsedTime(&t, p->ev[1], p->ev[2]); p->stats[MFB_STAT_MS_REGULAR] = t;
  cudaEventElapsedTime(&t, p->ev[2], p->ev[3]); p->stats[MFB_STAT_MS_ADAPTIVE] = t;
  cudaEventElapsedTime(&t, p->ev[3], p->ev[4]); p->stats[MFB_STAT_MS_SINGULAR] = t;
  cudaEventElapsedTime(&t, p->ev[4], p->ev[5]); p->stats[MFB_STAT_MS_FREETERM] = t;
  cudaEventElapsedTime(&t, p->ev[0], p->ev[5]); p->stats[MFB_STAT_MS_ASSEMBLE] = t;
  p->stats[MFB_STAT_LAUNCHES] = p->asm_launches;
  return MFB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Inviscid fluid (acoustic) BE region (first brick): fbem_bem_harpot3d_calculate_parameters
// (lib/fbem/src/bem_harpot3d.f90:123-165) + build_lse_mechanics_bem_harpot (src/build_lse_mechanics_bem_harpot.f90) with the
// scatter of assemble_bem_harpot_equation.f90:78-96, on a problem set up with ndof = 1.
// ---------------------------------------------------------------------------------------------------------------------
static int assemble_pot_device(mfb_problem* p, double omega, double rho, cd c, const mfb_z* cvalue) {
  if (p->ndof != 1) return fail(MFB_ERR_ARG, "this problem was not set up for an inviscid fluid region (mfb_harpot3d_setup)");
  if (!(omega > 0.0) || !(rho > 0.0) || c == cd(0.0, 0.0)) return fail(MFB_ERR_ARG, "mfb_harpot3d: omega, rho must be positive and c nonzero");
  cudaStream_t st = p->ctx->stream;
  const double c_pi = 3.14159265358979323846264338328;
  const cd im(0.0, 1.0), k = omega / c;
  PotParams pp;
  pp.k = mk(k.real(), k.imag());
  { cd v = -im * k; pp.P1 = mk(v.real(), v.imag()); }
  { cd v = 0.5 * (k * k); pp.Q1 = mk(v.real(), v.imag()); }
  { cd v = im * k; pp.Q2 = mk(v.real(), v.imag()); }
  pp.c4pi = 1.0 / (4.0 * c_pi); pp.d1J = rho * (omega * omega);
  set_pot_params(pp, st);
  if (cvalue) {
    CK(cudaMemcpyAsync(p->d_cvalue, cvalue, (size_t)2 * p->n_node * sizeof(double), cudaMemcpyHostToDevice, st));
    for (auto& g : p->groups) launch_gather_cv(g.dev, p->d_cvalue, st);
    p->have_cvalue = true;
  } else if (!p->have_cvalue) return fail(MFB_ERR_ARG, "cvalue == NULL but no prescribed values are resident yet");
  CK(cudaEventRecord(p->ev[0], st));
  CK(cudaMemsetAsync(p->sys.Are, 0, (size_t)2 * p->lda * p->n_dof * sizeof(double), st));
  CK(cudaMemsetAsync(p->sys.bre, 0, (size_t)2 * p->lda * sizeof(double), st));
  CK(cudaEventRecord(p->ev[1], st));
  for (auto& g : p->groups) launch_pot_regular(g.dev, p->colloc, p->sys, p->plan, st);
  CK(cudaEventRecord(p->ev[2], st));
  for (auto& g : p->groups) launch_pot_adaptive(g.dev, p->colloc, p->sys, g.adp, p->ctx->tables, st);
  CK(cudaEventRecord(p->ev[3], st));
  for (auto& g : p->groups) launch_pot_singular(g.dev, p->colloc, p->sys, g.sing, p->ctx->tables, st);
  CK(cudaEventRecord(p->ev[4], st));
  launch_freeterm(p->colloc, p->sys, p->ft, mk(0.0, 0.0), st);
  CK(cudaEventRecord(p->ev[5], st));
  CK(cudaGetLastError());
  p->factored = false; p->assembled = true; p->rows_permuted = true; p->real_resident = false;
  p->asm_launches = 1;
  for (auto& g : p->groups) p->asm_launches += 1 + (cvalue ? 1 : 0) + (g.adp.n_pairs > 0) + (g.sing.n_pairs > 0);
  return MFB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Biot poroelastic BE region: fbem_bem_harpor3d_calculate_parameters (por_params_host) + build_lse_mechanics_bem_harpor
// with the open-pore scatter of assemble_bem_harpor_equation.f90:78-110, :140-170, on a problem set up with ndof = 4.
// ---------------------------------------------------------------------------------------------------------------------
static int assemble_por_device(mfb_problem* p, double omega, cd lambda, cd mu, double rho1, double rho2, double rhoa, cd R, cd Q, double b, const mfb_z* cvalue) {
  if (p->ndof != 4) return fail(MFB_ERR_ARG, "this problem was not set up for a poroelastic region (mfb_harpor3d_setup)");
  if (!(omega > 0.0) || !(rho1 > 0.0) || !(rho2 > 0.0) || R == cd(0.0, 0.0) || mu == cd(0.0, 0.0)) return fail(MFB_ERR_ARG, "mfb_harpor3d: omega, rho1, rho2 must be positive, mu and R nonzero");
  cudaStream_t st = p->ctx->stream;
  PorParams pp; por_params_host(lambda, mu, rho1, rho2, rhoa, R, Q, b, omega, pp);
  set_por_params(pp, st);
  if (cvalue) {
    CK(cudaMemcpyAsync(p->d_cvalue, cvalue, (size_t)8 * p->n_node * sizeof(double), cudaMemcpyHostToDevice, st));
    for (auto& g : p->groups) launch_gather_cv(g.dev, p->d_cvalue, st);
    p->have_cvalue = true;
  } else if (!p->have_cvalue) return fail(MFB_ERR_ARG, "cvalue == NULL but no prescribed values are resident yet");
  CK(cudaEventRecord(p->ev[0], st));
  CK(cudaMemsetAsync(p->sys.Are, 0, (size_t)2 * p->lda * p->n_dof * sizeof(double), st));
  CK(cudaMemsetAsync(p->sys.bre, 0, (size_t)2 * p->lda * sizeof(double), st));
  CK(cudaEventRecord(p->ev[1], st));
  for (auto& g : p->groups) launch_por_regular(g.dev, p->colloc, p->sys, p->plan, st);
  CK(cudaEventRecord(p->ev[2], st));
  for (auto& g : p->groups) launch_por_adaptive(g.dev, p->colloc, p->sys, g.adp, p->ctx->tables, st);
  CK(cudaEventRecord(p->ev[3], st));
  for (auto& g : p->groups) launch_por_singular(g.dev, p->colloc, p->sys, g.sing, p->ctx->tables, st);
  CK(cudaEventRecord(p->ev[4], st));
  const double c_pi = 3.14159265358979323846264338328;
  const cd nu = 0.5 * lambda / (lambda + mu);
  const cd F = -1.0 / (8.0 * c_pi * (1.0 - nu));
  launch_freeterm(p->colloc, p->sys, p->ft, mk(F.real(), F.imag()), st);       // skeleton: Mantic's matrix
  launch_freeterm(p->colloc, p->sys, p->ft0, pp.J, st);                        // fluid phase: J c_pot
  CK(cudaEventRecord(p->ev[5], st));
  CK(cudaGetLastError());
  p->factored = false; p->assembled = true; p->rows_permuted = true; p->real_resident = false;
  p->asm_launches = 2;
  for (auto& g : p->groups) p->asm_launches += 1 + (cvalue ? 1 : 0) + (g.adp.n_pairs > 0) + (g.sing.n_pairs > 0);
  return MFB_OK;
}

// copy a planar device matrix to an interleaved host matrix in column chunks (bounded staging buffer)
static int download_matrix(mfb_problem* p, const double* re, const double* im, long long ld, int rows, int cols, mfb_z* host, long long ldh, const int* rowperm = nullptr,
                           const int* colperm = nullptr, bool accumulate = false) {
  cudaStream_t st = p->ctx->stream;
  int chunk = (int)std::max<long long>(1, std::min<long long>(cols, (256ll << 20) / (16ll * rows)));
  DevBuf sb; double* stage;
  if (cols == 1 && rows <= p->n_dof && p->d_vstage) stage = p->d_vstage;   // a vector: persistent staging (cudaMalloc / cudaFree synchronise the whole device and would
  else { CK(cudaMalloc(&sb.p, (size_t)chunk * rows * 16)); stage = (double*)sb.p; }   // serialise problems that work side by side on different streams)
  std::vector<mfb_z> hstage;                       // accumulate: the chunk lands here and is ADDED to the caller's matrix (the seam's `+=`)
  if (accumulate) hstage.resize((size_t)chunk * rows);
  for (int c0 = 0; c0 < cols; c0 += chunk) {
    int nc = std::min(chunk, cols - c0);
    launch_interleave(re, im, ld, rows, nc, stage, rows, rowperm, colperm, c0, st);
    if (!accumulate) CK(cudaMemcpy2DAsync(host + (long long)c0 * ldh, (size_t)ldh * 16, stage, (size_t)rows * 16, (size_t)rows * 16, nc, cudaMemcpyDeviceToHost, st));
    else CK(cudaMemcpyAsync(hstage.data(), stage, (size_t)nc * rows * 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (accumulate) {
#pragma omp parallel for schedule(static)
      for (int c = 0; c < nc; c++) {
        mfb_z* dst = host + (long long)(c0 + c) * ldh; const mfb_z* src = hstage.data() + (size_t)c * rows;
        for (int i = 0; i < rows; i++) { dst[i].re += src[i].re; dst[i].im += src[i].im; }
      }
    }
  }
  return MFB_OK;
}
static int upload_matrix(mfb_problem* p, const mfb_z* host, long long ldh, int rows, int cols, double* re, double* im, long long ld, const int* rowperm = nullptr) {
  cudaStream_t st = p->ctx->stream;
  int chunk = (int)std::max<long long>(1, std::min<long long>(cols, (256ll << 20) / (16ll * rows)));
  DevBuf sb; CK(cudaMalloc(&sb.p, (size_t)chunk * rows * 16)); double* stage = (double*)sb.p;
  for (int c0 = 0; c0 < cols; c0 += chunk) {
    int nc = std::min(chunk, cols - c0);
    CK(cudaMemcpy2DAsync(stage, (size_t)rows * 16, host + (long long)c0 * ldh, (size_t)ldh * 16, (size_t)rows * 16, nc, cudaMemcpyHostToDevice, st));
    launch_deinterleave(stage, rows, rows, nc, re + (long long)c0 * ld, im + (long long)c0 * ld, ld, rowperm, st);
    CK(cudaStreamSynchronize(st));
  }
  return MFB_OK;
}

extern "C" int mfb_harela3d_assemble_acc(mfb_problem* p, double omega, const mfb_z* lambda, const mfb_z* mu, double rho, const mfb_z* nu,
                                         const mfb_z* cvalue, mfb_z* A, int lda, mfb_z* b, int accumulate) {
  if (!p || !lambda || !mu || !nu) return fail(MFB_ERR_ARG, "mfb_harela3d_assemble: null argument");
  if (A && lda < p->n_dof) return fail(MFB_ERR_ARG, "mfb_harela3d_assemble: lda < n_dof");
  CK(cudaSetDevice(p->ctx->device));
  int r = assemble_device(p, omega, cd(lambda->re, lambda->im), cd(mu->re, mu->im), rho, cd(nu->re, nu->im), cvalue);
  if (r) return r;
  r = collect_assembly_times(p);
  if (r) return r;
  // the device-resident system is in internal row order; the host sees the reference's row order
  if (A) { r = download_matrix(p, p->sys.Are, p->sys.Aim, p->lda, p->n_dof, p->n_dof, A, lda, p->d_rowperm, p->d_colperm, accumulate != 0); if (r) return r; }
  if (b) { r = download_matrix(p, p->sys.bre, p->sys.bim, p->lda, p->n_dof, 1, b, p->n_dof, p->d_rowperm, nullptr, accumulate != 0); if (r) return r; }
  return MFB_OK;
}
extern "C" int mfb_harela3d_assemble(mfb_problem* p, double omega, const mfb_z* lambda, const mfb_z* mu, double rho, const mfb_z* nu,
                                     const mfb_z* cvalue, mfb_z* A, mfb_z* b) {
  return mfb_harela3d_assemble_acc(p, omega, lambda, mu, rho, nu, cvalue, A, p ? p->n_dof : 0, b, 0);
}

extern "C" int mfb_harpot3d_setup(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                                  const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                                  const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                                  const int* row, const int* col_p, const int* col_un, const int* ctype, int n_dof,
                                  double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                                  double geometric_tolerance, mfb_problem** out) {
  return setup_impl(ctx, n_node, node_x, n_elem, etype, elem_ptr, elem_node, elem_reversed, n_colloc, colloc_x, colloc_node, colloc_elem, colloc_kn, colloc_xi,
                    row, col_p, col_un, ctype, n_dof, qsi_relative_error, qsi_ns_max, n_precalsets, precalset_gln, geometric_tolerance, nullptr, 1, out);
}
extern "C" int mfb_harpot3d_assemble(mfb_problem* p, double omega, double rho, const mfb_z* c, const mfb_z* cvalue, mfb_z* A, mfb_z* b) {
  if (!p || !c) return fail(MFB_ERR_ARG, "mfb_harpot3d_assemble: null argument");
  CK(cudaSetDevice(p->ctx->device));
  int r = assemble_pot_device(p, omega, rho, cd(c->re, c->im), cvalue);
  if (r) return r;
  r = collect_assembly_times(p);
  if (r) return r;
  if (A) { r = download_matrix(p, p->sys.Are, p->sys.Aim, p->lda, p->n_dof, p->n_dof, A, p->n_dof, p->d_rowperm, p->d_colperm); if (r) return r; }
  if (b) { r = download_matrix(p, p->sys.bre, p->sys.bim, p->lda, p->n_dof, 1, b, p->n_dof, p->d_rowperm); if (r) return r; }
  return MFB_OK;
}

extern "C" int mfb_harpor3d_setup(mfb_ctx* ctx, int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr,
                                  const int* elem_node, const unsigned char* elem_reversed, int n_colloc, const double* colloc_x,
                                  const int* colloc_node, const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                                  const int* row, const int* col_p, const int* col_s, const int* ctype, int n_dof,
                                  double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln,
                                  double geometric_tolerance, mfb_problem** out) {
  return setup_impl(ctx, n_node, node_x, n_elem, etype, elem_ptr, elem_node, elem_reversed, n_colloc, colloc_x, colloc_node, colloc_elem, colloc_kn, colloc_xi,
                    row, col_p, col_s, ctype, n_dof, qsi_relative_error, qsi_ns_max, n_precalsets, precalset_gln, geometric_tolerance, nullptr, 4, out);
}
extern "C" int mfb_harpor3d_assemble(mfb_problem* p, double omega, const mfb_z* lambda, const mfb_z* mu, double rho1, double rho2, double rhoa,
                                     const mfb_z* R, const mfb_z* Q, double b, const mfb_z* cvalue, mfb_z* A, mfb_z* bb) {
  if (!p || !lambda || !mu || !R || !Q) return fail(MFB_ERR_ARG, "mfb_harpor3d_assemble: null argument");
  CK(cudaSetDevice(p->ctx->device));
  int r = assemble_por_device(p, omega, cd(lambda->re, lambda->im), cd(mu->re, mu->im), rho1, rho2, rhoa, cd(R->re, R->im), cd(Q->re, Q->im), b, cvalue);
  if (r) return r;
  r = collect_assembly_times(p);
  if (r) return r;
  if (A) { r = download_matrix(p, p->sys.Are, p->sys.Aim, p->lda, p->n_dof, p->n_dof, A, p->n_dof, p->d_rowperm, p->d_colperm); if (r) return r; }
  if (bb) { r = download_matrix(p, p->sys.bre, p->sys.bim, p->lda, p->n_dof, 1, bb, p->n_dof, p->d_rowperm); if (r) return r; }
  return MFB_OK;
}

// per-phase LU timings use pre-created events recorded on the stream without synchronising; MFB_LU_TIMING=0 disables them
static bool lu_timing() { const char* e = getenv("MFB_LU_TIMING"); return !(e && e[0] == '0'); }
static int ensure_lu(mfb_problem* p) {
  if (!p->lu_ready) {
    int nb = 256;
    if (const char* e = getenv("MFB_LU_NB")) { int v = atoi(e); if (v >= 32 && v <= 1024 && v % 32 == 0) nb = v; }
    if (lu_work_alloc(p->lu, p->n_dof, nb) != 0) return fail(MFB_ERR_CUDA, "LU workspace allocation failed");
    p->lu_ready = true;
  }
  return MFB_OK;
}
// factorise the device-resident system and bring the pivots back (permutation vector for the solves)
static int factor_device(mfb_problem* p, int n, bool timing) {
  cudaStream_t st = p->ctx->stream;
  int r = ensure_lu(p); if (r) return r;
  CK(cudaEventRecord(p->ev[6], st));
  int e = zgetrf_planar(p->sys.Are, p->real_resident ? nullptr : p->sys.Aim, p->lda, n, p->d_ipiv, p->lu, st, timing);
  if (e) return fail(MFB_ERR_CUDA, std::string("zgetrf_planar: ") + cudaGetErrorString((cudaError_t)e));
  CK(cudaEventRecord(p->ev[7], st));
  int info = 0;
  CK(cudaMemcpyAsync(p->h_ipiv.data(), p->d_ipiv, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&info, p->lu.info, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  float t; cudaEventElapsedTime(&t, p->ev[6], p->ev[7]); p->stats[MFB_STAT_MS_LU] = t;
  if (timing) lu_collect_times(p->lu);
  p->stats[MFB_STAT_LU_LAUNCHES] = (double)p->lu.launches;
  p->stats[MFB_STAT_GEMM_LAUNCHES] = (double)p->lu.gemm_launches; p->stats[MFB_STAT_GEMM_FLOPS] = p->lu.gemm_flops; p->stats[MFB_STAT_GEMM_EXEC_FLOPS] = p->lu.gemm_exec_flops;
  p->stats[MFB_STAT_MS_PANEL] = p->lu.ms_panel; p->stats[MFB_STAT_MS_SWAP] = p->lu.ms_swap; p->stats[MFB_STAT_MS_TRSM] = p->lu.ms_trsm; p->stats[MFB_STAT_MS_GEMM] = p->lu.ms_gemm;
  for (int i = 0; i < n; i++)   // never index with a pivot the device did not produce (a fault upstream must not become a host out-of-bounds swap)
    if (p->h_ipiv[i] < i + 1 || p->h_ipiv[i] > n) { p->factored = false; return fail(MFB_ERR_CUDA, "zgetrf_planar: pivot index out of range (device fault during the factorisation?)"); }
  std::vector<int> perm(n);
  for (int i = 0; i < n; i++) perm[i] = i;
  for (int i = 0; i < n; i++) { int q = p->h_ipiv[i] - 1; if (q != i) std::swap(perm[i], perm[q]); }
  CK(cudaMemcpyAsync(p->d_perm, perm.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  p->factored = true;
  if (info > 0) { char buf[128]; snprintf(buf, sizeof(buf), "zgetrf: U(%d,%d) is exactly zero, the matrix is singular", info, info); return fail(info, buf); }
  return MFB_OK;
}

// Factorise the resident complex system and solve its resident right-hand side: the tail shared by the *_solve_frequency entry points.
// Small systems (n <= MFB_LU_GRAPH_MAX_N, default 4096) run the whole sequence -- ~230 panel / update launches on two streams, the diagonal-block
// inverses, the permutation (formed on the device), ~50 substitution launches -- as ONE CUDA graph captured at the first call: at 1386 DOF the
// host spent ~12 us per launch when several problems were driven side by side (capi.ProblemLanes), i.e. 3.6 ms of launches per frequency.
static int lu_graph_max_n() { static int v = -1; if (v < 0) { const char* e = getenv("MFB_LU_GRAPH_MAX_N"); v = e ? atoi(e) : 4096; } return v; }
static int factor_and_solve_resident(mfb_problem* p) {
  cudaStream_t st = p->ctx->stream;
  const int n = p->n_dof;
  int r = ensure_lu(p); if (r) return r;
  // the first factorisation of a problem always takes the plain path: it performs the one-time set-up that must not happen inside a stream
  // capture (function attributes, tensor maps, cluster-launch probing) and leaves per-phase timings of this size in the statistics
  const bool want_graph = !p->real_resident && n <= lu_graph_max_n() && p->lu.inv && !p->lu_graph_failed && p->lu_plain_calls >= 1;
  if (!want_graph) p->lu_plain_calls++;
  if (!want_graph) {
    r = factor_device(p, n, lu_timing());
    if (r) return r;
    CK(cudaEventRecord(p->ev[6], st));
    int e = zgetrs_planar(p->sys.Are, p->real_resident ? nullptr : p->sys.Aim, p->lda, n, p->d_perm, p->sys.bre, p->real_resident ? nullptr : p->sys.bim, p->lda, 1, st, p->lu.inv,
                          p->lu.solve_ws);
    if (e) return fail(MFB_ERR_CUDA, std::string("zgetrs_planar: ") + cudaGetErrorString((cudaError_t)e));
    CK(cudaEventRecord(p->ev[7], st)); CK(cudaEventSynchronize(p->ev[7]));
    float t; cudaEventElapsedTime(&t, p->ev[6], p->ev[7]); p->stats[MFB_STAT_MS_SOLVE] = t;
    return MFB_OK;
  }
  if (!p->d_graph_flags) { CK(cudaMalloc((void**)&p->d_graph_flags, 2 * sizeof(int))); }
  if (!p->lu_graph) {
    // capture (thread-local mode: other host threads keep issuing CUDA calls for their own problems meanwhile)
    cudaGraph_t g = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    int e = 0;
    if (ce == cudaSuccess) {
      cudaMemsetAsync(p->d_graph_flags, 0, 2 * sizeof(int), st);
      e = zgetrf_planar(p->sys.Are, p->sys.Aim, p->lda, n, p->d_ipiv, p->lu, st, false);
      if (!e) e = launch_perm_from_ipiv(p->d_ipiv, p->d_perm, n, p->d_graph_flags, st);
      if (!e) e = zgetrs_planar(p->sys.Are, p->sys.Aim, p->lda, n, p->d_perm, p->sys.bre, p->sys.bim, p->lda, 1, st, p->lu.inv, p->lu.solve_ws);
      ce = cudaStreamEndCapture(st, &g);
    }
    if (ce != cudaSuccess || e || !g) {      // not capturable on this configuration: plain launches from now on
      cudaGetLastError(); if (g) cudaGraphDestroy(g);
      p->lu_graph_failed = true;
      return factor_and_solve_resident(p);
    }
    size_t nn = 0; cudaGraphGetNodes(g, nullptr, &nn); p->lu_graph_nodes = (int)nn;
    ce = cudaGraphInstantiate(&p->lu_graph, g, 0);
    cudaGraphDestroy(g);
    if (ce != cudaSuccess) { cudaGetLastError(); p->lu_graph = nullptr; p->lu_graph_failed = true; return factor_and_solve_resident(p); }
  }
  CK(cudaEventRecord(p->ev[6], st));
  CK(cudaGraphLaunch(p->lu_graph, st));
  CK(cudaEventRecord(p->ev[7], st));
  int flags[2] = {0, 0}, info = 0;
  CK(cudaMemcpyAsync(flags, p->d_graph_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&info, p->lu.info, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  float t; cudaEventElapsedTime(&t, p->ev[6], p->ev[7]);
  p->stats[MFB_STAT_MS_LU] = t; p->stats[MFB_STAT_MS_SOLVE] = 0.0;           // one graph: the split is not observable
  p->stats[MFB_STAT_LU_LAUNCHES] = (double)p->lu_graph_nodes;
  p->stats[MFB_STAT_GEMM_LAUNCHES] = (double)p->lu.gemm_launches; p->stats[MFB_STAT_GEMM_FLOPS] = p->lu.gemm_flops; p->stats[MFB_STAT_GEMM_EXEC_FLOPS] = p->lu.gemm_exec_flops;
  p->stats[MFB_STAT_MS_PANEL] = p->stats[MFB_STAT_MS_SWAP] = p->stats[MFB_STAT_MS_TRSM] = p->stats[MFB_STAT_MS_GEMM] = 0.0;
  p->factored = true;
  if (flags[0]) return fail(MFB_ERR_CUDA, "zgetrf (graph): pivot index out of range");
  if (info > 0) { char buf[128]; snprintf(buf, sizeof(buf), "zgetrf: U(%d,%d) is exactly zero, the matrix is singular", info, info); return fail(info, buf); }
  return MFB_OK;
}

extern "C" int mfb_zsolve(mfb_problem* p, int n, mfb_z* A, int lda, int* ipiv, mfb_z* b, int nrhs, int factorize) {
  if (!p) return fail(MFB_ERR_ARG, "mfb_zsolve: null problem");
  if (n != p->n_dof) return fail(MFB_ERR_ARG, "mfb_zsolve: n must equal the problem's n_dof");
  if (nrhs < 0 || (A && lda < n)) return fail(MFB_ERR_ARG, "mfb_zsolve: invalid nrhs/lda");
  CK(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  int r;
  if (factorize) {
    if (A) { r = upload_matrix(p, A, lda, n, n, p->sys.Are, p->sys.Aim, p->lda); if (r) return r; p->rows_permuted = false; p->real_resident = false; }
    else if (!p->assembled) return fail(MFB_ERR_ARG, "mfb_zsolve: factorize=1 with A == NULL but no assembled system is resident (nothing assembled yet, or the resident "
                                                     "matrix already holds the factors of an earlier solve: assemble again or pass factorize=0)");
    else if (p->real_resident) return fail(MFB_ERR_ARG, "mfb_zsolve: the resident system is real (static assembly); use mfb_dsolve");
    p->assembled = false;
    r = factor_device(p, n, lu_timing());
    if (ipiv) memcpy(ipiv, p->h_ipiv.data(), (size_t)n * sizeof(int));
    if (A) { int r2 = download_matrix(p, p->sys.Are, p->sys.Aim, p->lda, n, n, A, lda); if (r2) return r2; }
    if (r) return r;
  } else if (!p->factored) return fail(MFB_ERR_ARG, "mfb_zsolve: factorize=0 but no factors are resident");
  else if (p->real_resident) return fail(MFB_ERR_ARG, "mfb_zsolve: factorize=0 but the resident factors are real (mfb_dsolve / static path); use mfb_dsolve");
  if (nrhs == 0) return MFB_OK;
  double *bre = p->sys.bre, *bim = p->sys.bim; long long ldb = p->lda;
  DevBuf tmp;   // freed on every exit path
  if (b) {
    if (nrhs > 1) { CK(cudaMalloc((void**)&tmp.p, (size_t)2 * p->lda * nrhs * sizeof(double))); bre = (double*)tmp.p; bim = bre + (size_t)p->lda * nrhs; }
    r = upload_matrix(p, b, n, n, nrhs, bre, bim, ldb, p->rows_permuted ? p->d_rowperm : nullptr); if (r) return r;
  } else if (nrhs != 1) return fail(MFB_ERR_ARG, "mfb_zsolve: device-resident rhs has a single column");
  CK(cudaEventRecord(p->ev[6], st));
  int e = zgetrs_planar(p->sys.Are, p->sys.Aim, p->lda, n, p->d_perm, bre, bim, ldb, nrhs, st, p->lu.inv);
  if (e) return fail(MFB_ERR_CUDA, std::string("zgetrs_planar: ") + cudaGetErrorString((cudaError_t)e));
  CK(cudaEventRecord(p->ev[7], st)); CK(cudaEventSynchronize(p->ev[7]));
  float t; cudaEventElapsedTime(&t, p->ev[6], p->ev[7]); p->stats[MFB_STAT_MS_SOLVE] = t;
  if (b) { r = download_matrix(p, bre, bim, ldb, n, nrhs, b, n, p->rows_permuted ? p->d_colperm : nullptr); if (r) return r; }
  return MFB_OK;
}

// solve_lse_c with ALL its options (src/solve_lse_c.f90:25-219): scaling (zgeequ + zlaqge :81-117), condition (zgecon :140-165), refine (zgerfs :191-206)
// around the LU of mfb_zsolve, everything on the device (solve_ex.cu).  Conventions of mfb_zsolve for A / ipiv / b / factorize.  equed (1 char, in/out),
// r, c (n doubles each, in/out) are the reference's arguments of the same names: written when factorize && scaling, read when !factorize && scaling.
// rcond (out, when condition && factorize), ferr / berr (out, nrhs each, when refine) may be NULL.  The unfactorised copy `Ao` of the reference lives
// on the device.
extern "C" int mfb_zsolve_ex(mfb_problem* p, int n, mfb_z* A, int lda, int* ipiv, mfb_z* b, int nrhs, int factorize, int scaling, int condition, int refine,
                             char* equed, double* r, double* c, double* rcond, double* ferr, double* berr) {
  if (!p) return fail(MFB_ERR_ARG, "mfb_zsolve_ex: null problem");
  if (!scaling && !condition && !refine) return mfb_zsolve(p, n, A, lda, ipiv, b, nrhs, factorize);
  if (n != p->n_dof) return fail(MFB_ERR_ARG, "mfb_zsolve_ex: n must equal the problem's n_dof");
  if (nrhs < 0 || (A && lda < n)) return fail(MFB_ERR_ARG, "mfb_zsolve_ex: invalid nrhs/lda");
  if (scaling && (!equed || !r || !c)) return fail(MFB_ERR_ARG, "mfb_zsolve_ex: scaling needs equed, r and c");
  if (nrhs > 0 && !b) return fail(MFB_ERR_ARG, "mfb_zsolve_ex: pass the right-hand side b (host)");
  CK(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  const size_t plane = (size_t)p->lda * n;
  int rr;
  if (!p->d_rs) { CK(cudaMalloc((void**)&p->d_rs, (size_t)n * 8)); CK(cudaMalloc((void**)&p->d_cs, (size_t)n * 8)); CK(cudaMalloc((void**)&p->d_xtmp, ((size_t)12 * p->lda + 256) * 8)); }
  if ((condition || refine) && !p->Ao) CK(cudaMalloc((void**)&p->Ao, 2 * plane * 8));
  if (factorize) {
    if (A) { rr = upload_matrix(p, A, lda, n, n, p->sys.Are, p->sys.Aim, p->lda); if (rr) return rr; p->rows_permuted = false; p->real_resident = false; }
    else if (!p->assembled) return fail(MFB_ERR_ARG, "mfb_zsolve_ex: factorize=1 with A == NULL but no assembled system is resident");
    else if (p->real_resident) return fail(MFB_ERR_ARG, "mfb_zsolve_ex: the resident system is real; use mfb_dsolve");
    p->assembled = false; p->equed = 'N';
    if (scaling) {
      double rowcnd, colcnd, amax; int info = 0;
      int e = zequilibrate(p->sys.Are, p->sys.Aim, p->lda, n, p->d_rs, p->d_cs, p->rs_int, p->cs_int, &rowcnd, &colcnd, &amax, &p->equed, &info, st);
      if (e) return fail(MFB_ERR_CUDA, "equilibration kernels failed");
      if (info) { char buf[96]; snprintf(buf, sizeof(buf), "zgeequ: %s %d of A is exactly zero", info <= n ? "row" : "column", info <= n ? info : info - n); return fail(info, buf); }
      *equed = p->equed;
      for (int i = 0; i < n; i++) { r[i] = p->rs_int[p->rows_permuted ? p->rowperm[i] : i]; c[i] = p->cs_int[p->rows_permuted ? p->colperm[i] : i]; }
    }
    if (condition || refine) CK(cudaMemcpyAsync(p->Ao, p->sys.Are, 2 * plane * 8, cudaMemcpyDeviceToDevice, st));   // Are and Aim are one allocation (planes back to back)
  } else {
    if (!p->factored || p->real_resident) return fail(MFB_ERR_ARG, "mfb_zsolve_ex: factorize=0 but no complex factors are resident");
    if (scaling) {
      if (*equed != 'N' && *equed != 'R' && *equed != 'C' && *equed != 'B') return fail(MFB_ERR_ARG, "mfb_zsolve_ex: invalid value of equed");
      p->equed = *equed; p->rs_int.assign(n, 1.0); p->cs_int.assign(n, 1.0);
      for (int i = 0; i < n; i++) { p->rs_int[p->rows_permuted ? p->rowperm[i] : i] = r[i]; p->cs_int[p->rows_permuted ? p->colperm[i] : i] = c[i]; }
      CK(cudaMemcpyAsync(p->d_rs, p->rs_int.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(p->d_cs, p->cs_int.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    }
  }
  if (factorize) {
    rr = factor_device(p, n, lu_timing());
    if (ipiv) memcpy(ipiv, p->h_ipiv.data(), (size_t)n * sizeof(int));
    if (A) { int r2 = download_matrix(p, p->sys.Are, p->sys.Aim, p->lda, n, n, A, lda); if (r2) return r2; }
    if (rr) return rr;
  }
  if (!p->lu.inv) return fail(MFB_ERR_UNSUPPORTED, "mfb_zsolve_ex needs the diagonal-block inverses of the factorisation (MFB_LU_SOLVE_INV=0 is set)");
  typedef std::complex<double> zc;
  // device scratch: v (2 ld) | transposed-solve workspace (4 ld + 128) | residual r (2 ld) + s (ld) | spare
  double *vre = p->d_xtmp, *vim = vre + p->lda, *tws = vim + p->lda, *res = tws + 4 * p->lda + 128;
  auto dev_solve = [&](std::vector<zc>& v, bool conjt) -> int {          // v := inv(A) v or inv(A)^H v (internal order of the resident factors)
    std::vector<double> h(2 * (size_t)n);
    for (int i = 0; i < n; i++) { h[i] = v[i].real(); h[n + i] = v[i].imag(); }
    cudaMemcpyAsync(vre, h.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st); cudaMemcpyAsync(vim, h.data() + n, (size_t)n * 8, cudaMemcpyHostToDevice, st);
    int e = conjt ? zgetrs_conjtrans_planar(p->sys.Are, p->sys.Aim, p->lda, n, p->d_perm, p->lu.inv, vre, vim, tws, st)
                  : zgetrs_planar(p->sys.Are, p->sys.Aim, p->lda, n, p->d_perm, vre, vim, p->lda, 1, st, p->lu.inv, p->lu.solve_ws);
    if (e) return e;
    cudaMemcpyAsync(h.data(), vre, (size_t)n * 8, cudaMemcpyDeviceToHost, st); cudaMemcpyAsync(h.data() + n, vim, (size_t)n * 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    for (int i = 0; i < n; i++) v[i] = zc(h[i], h[n + i]);
    return (int)cudaGetLastError();
  };
  int dev_err = 0;
  if (condition && factorize) {
    const double anorm = matrix_norm1(p->Ao, p->Ao + plane, p->lda, n, res, st);
    const double ainvnm = norm1_estimate(n, [&](std::vector<zc>& v, bool conjt) { if (!dev_err) dev_err = dev_solve(v, conjt); });
    if (dev_err) return fail(MFB_ERR_CUDA, "condition estimate: a device solve failed");
    if (rcond) *rcond = (anorm > 0.0 && ainvnm > 0.0) ? (1.0 / ainvnm) / anorm : 0.0;
  }
  const double eps = DBL_EPSILON * 0.5, safmin = DBL_MIN;
  for (int col = 0; col < nrhs; col++) {
    // internal-order right-hand side (row scaling applied), solution, and the refinement of zgerfs
    std::vector<zc> bb(n), x(n);
    for (int i = 0; i < n; i++) { const int q = p->rows_permuted ? p->rowperm[i] : i; bb[q] = zc(b[(size_t)col * n + i].re, b[(size_t)col * n + i].im); }
    if (scaling && (p->equed == 'R' || p->equed == 'B')) for (int i = 0; i < n; i++) bb[i] *= p->rs_int[i];
    x = bb;
    dev_err = dev_solve(x, false); if (dev_err) return fail(MFB_ERR_CUDA, "zgetrs_planar failed");
    if (refine) {
      DevSystem so = p->sys; so.Are = p->Ao; so.Aim = p->Ao + plane; so.bre = res + 3 * p->lda; so.bim = res + 4 * p->lda;   // b of the residual kernel
      std::vector<double> hb(2 * (size_t)n), hx(2 * (size_t)n), hr(3 * (size_t)n);
      for (int i = 0; i < n; i++) { hb[i] = bb[i].real(); hb[n + i] = bb[i].imag(); }
      cudaMemcpyAsync(so.bre, hb.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st); cudaMemcpyAsync(so.bim, hb.data() + n, (size_t)n * 8, cudaMemcpyHostToDevice, st);
      const int nz = n + 1; const double safe1 = nz * safmin, safe2 = safe1 / eps;
      double lstres = 3.0, be = 0.0; int count = 1;
      std::vector<zc> rv(n); std::vector<double> rw(n);
      for (;;) {
        for (int i = 0; i < n; i++) { hx[i] = x[i].real(); hx[n + i] = x[i].imag(); }
        cudaMemcpyAsync(vre, hx.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st); cudaMemcpyAsync(vim, hx.data() + n, (size_t)n * 8, cudaMemcpyHostToDevice, st);
        cudaMemsetAsync(res, 0, (size_t)3 * p->lda * 8, st);
        launch_residual(so, vre, vim, res, res + p->lda, res + 2 * p->lda, st);       // A x - b and |A||x| + |b| (1-norm moduli, as zgerfs)
        cudaMemcpyAsync(hr.data(), res, (size_t)n * 8, cudaMemcpyDeviceToHost, st); cudaMemcpyAsync(hr.data() + n, res + p->lda, (size_t)n * 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(hr.data() + 2 * (size_t)n, res + 2 * p->lda, (size_t)n * 8, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        be = 0.0;
        for (int i = 0; i < n; i++) {
          rv[i] = zc(-hr[i], -hr[n + i]); rw[i] = hr[2 * (size_t)n + i];
          const double num = std::fabs(rv[i].real()) + std::fabs(rv[i].imag());
          be = std::max(be, rw[i] > safe2 ? num / rw[i] : (num + safe1) / (rw[i] + safe1));
        }
        if (be > eps && 2.0 * be <= lstres && count <= 5) {
          std::vector<zc> dx = rv; dev_err = dev_solve(dx, false); if (dev_err) return fail(MFB_ERR_CUDA, "refinement: a device solve failed");
          for (int i = 0; i < n; i++) x[i] += dx[i];
          lstres = be; count++;
        } else break;
      }
      if (berr) berr[col] = be;
      if (ferr) {
        for (int i = 0; i < n; i++) { const double num = std::fabs(rv[i].real()) + std::fabs(rv[i].imag()); rw[i] = rw[i] > safe2 ? num + nz * eps * rw[i] : num + nz * eps * rw[i] + safe1; }
        const double est = norm1_estimate(n, [&](std::vector<zc>& v, bool conjt) {      // operator diag(W) inv(A)^H and its adjoint inv(A) diag(W)
          if (dev_err) return;
          if (!conjt) { dev_err = dev_solve(v, true); for (int i = 0; i < n; i++) v[i] *= rw[i]; }
          else { for (int i = 0; i < n; i++) v[i] *= rw[i]; dev_err = dev_solve(v, false); }
        });
        if (dev_err) return fail(MFB_ERR_CUDA, "forward error estimate: a device solve failed");
        double xmax = 0.0; for (int i = 0; i < n; i++) xmax = std::max(xmax, std::fabs(x[i].real()) + std::fabs(x[i].imag()));
        ferr[col] = xmax != 0.0 ? est / xmax : est;
      }
    }
    if (scaling && (p->equed == 'C' || p->equed == 'B')) for (int i = 0; i < n; i++) x[i] *= p->cs_int[i];
    for (int i = 0; i < n; i++) { const zc v = x[p->rows_permuted ? p->colperm[i] : i]; b[(size_t)col * n + i].re = v.real(); b[(size_t)col * n + i].im = v.imag(); }
  }
  return MFB_OK;
}

extern "C" int mfb_harela3d_solve_frequency(mfb_problem* p, double omega, const mfb_z* lambda, const mfb_z* mu, double rho, const mfb_z* nu,
                                            const mfb_z* cvalue, mfb_z* x) {
  if (!p || !lambda || !mu || !nu) return fail(MFB_ERR_ARG, "mfb_harela3d_solve_frequency: null argument");
  CK(cudaSetDevice(p->ctx->device));
  int r = assemble_device(p, omega, cd(lambda->re, lambda->im), cd(mu->re, mu->im), rho, cd(nu->re, nu->im), cvalue);
  if (r) return r;
  r = factor_and_solve_resident(p);
  int r2 = collect_assembly_times(p); if (r2) return r2;
  if (r) return r;
  p->assembled = false;
  if (!x) return MFB_OK;   // solution stays on the device (mfb_get_solution)
  return download_matrix(p, p->sys.bre, p->sys.bim, p->lda, p->n_dof, 1, x, p->n_dof, p->d_colperm);
}

// ---------------------------------------------------------------------------------------------------------------------
// The frequency loop of src/multifebe.f90:107-124 as ONE C-ABI call, sharded over the ranks of a job: rank q owns the
// frequencies kf = q, q + nranks, ...; each one is assembled, factorised and solved without leaving the device; the solutions are gathered
// with one NCCL all-reduce (every rank contributes its own columns, zeros elsewhere) so that every rank returns the whole n_dof x n_freq
// block X (column kf = solution of omega[kf], host column order).  No collective on the data path.  nranks = 1: no NCCL at all.
// nccl_id128: the 128-byte unique id made by rank 0 with mfb_dist_unique_id and handed to the other ranks by the host's own means.
// info[kf] (may be NULL): 0, or the LAPACK info of a singular pivot at that frequency (the sweep goes on).
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int mfb_harela3d_sweep(mfb_problem* p, int n_freq, const double* omega, const mfb_z* lambda, const mfb_z* mu, double rho, const mfb_z* nu,
                                  const mfb_z* cvalue, int rank, int nranks, const char* nccl_id128, mfb_z* X, int* info) {
  if (!p || !omega || !lambda || !mu || !nu || !X || n_freq < 1) return fail(MFB_ERR_ARG, "mfb_harela3d_sweep: invalid argument");
  if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !nccl_id128)) return fail(MFB_ERR_ARG, "mfb_harela3d_sweep: invalid rank / nranks / unique id");
  if (p->have_inc && n_freq > 1) return fail(MFB_ERR_ARG, "mfb_harela3d_sweep: an incident field is set; it depends on the frequency, so such a model is swept with one mfb_harela3d_set_incident + mfb_harela3d_solve_frequency per frequency");
  CK(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  const int n = p->n_dof;
  const size_t plane = (size_t)p->lda * n_freq;
  DevBuf xd; CK(cudaMalloc(&xd.p, 2 * plane * sizeof(double)));
  double* Xre = (double*)xd.p; double* Xim = Xre + plane;
  CK(cudaMemsetAsync(Xre, 0, 2 * plane * sizeof(double), st));
  std::vector<int> linfo(n_freq, 0);
  for (int kf = rank; kf < n_freq; kf += nranks) {
    int r = mfb_harela3d_solve_frequency(p, omega[kf], lambda, mu, rho, nu, kf == rank ? cvalue : nullptr, nullptr);
    if (r > 0) { linfo[kf] = r; continue; }          // singular at this frequency: reported, column left at zero
    if (r) return r;
    CK(cudaMemcpyAsync(Xre + (size_t)kf * p->lda, p->sys.bre, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(Xim + (size_t)kf * p->lda, p->sys.bim, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
  }
  if (nranks > 1) {
    std::string err;
    DistComm* comm = make_nccl_comm(rank, nranks, nccl_id128, err);
    if (!comm) return fail(MFB_ERR_CUDA, "mfb_harela3d_sweep: " + err);
    double* bufs[1] = {Xre}; int ranks[1] = {rank}; cudaStream_t sts[1] = {st};
    int e = comm->allreduce_sum(ranks, bufs, 2 * plane, sts, 1);
    std::string ce = e ? std::string(comm->last_error()) : std::string();
    DevBuf di; int e2 = 0;
    if (!e && info) {       // the per-frequency info flags travel the same way
      std::vector<double> fi(n_freq); for (int k = 0; k < n_freq; k++) fi[k] = linfo[k];
      if (cudaMalloc(&di.p, (size_t)n_freq * 8) == cudaSuccess) {
        cudaMemcpyAsync(di.p, fi.data(), (size_t)n_freq * 8, cudaMemcpyHostToDevice, st);
        double* b2[1] = {(double*)di.p}; e2 = comm->allreduce_sum(ranks, b2, (size_t)n_freq, sts, 1);
        cudaMemcpyAsync(fi.data(), di.p, (size_t)n_freq * 8, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        for (int k = 0; k < n_freq; k++) linfo[k] = (int)fi[k];
      }
    }
    CK(cudaStreamSynchronize(st));
    delete comm;
    if (e || e2) return fail(MFB_ERR_CUDA, "mfb_harela3d_sweep: NCCL all-reduce failed: " + ce);
  }
  if (info) for (int k = 0; k < n_freq; k++) info[k] = linfo[k];
  return download_matrix(p, Xre, Xim, p->lda, n, n_freq, X, n, p->d_colperm);
}

extern "C" int mfb_harpot3d_solve_frequency(mfb_problem* p, double omega, double rho, const mfb_z* c, const mfb_z* cvalue, mfb_z* x) {
  if (!p || !c) return fail(MFB_ERR_ARG, "mfb_harpot3d_solve_frequency: null argument");
  CK(cudaSetDevice(p->ctx->device));
  int r = assemble_pot_device(p, omega, rho, cd(c->re, c->im), cvalue);
  if (r) return r;
  r = factor_and_solve_resident(p);
  int r2 = collect_assembly_times(p); if (r2) return r2;
  if (r) return r;
  p->assembled = false;
  if (!x) return MFB_OK;
  return download_matrix(p, p->sys.bre, p->sys.bim, p->lda, p->n_dof, 1, x, p->n_dof, p->d_colperm);
}

extern "C" int mfb_harpor3d_solve_frequency(mfb_problem* p, double omega, const mfb_z* lambda, const mfb_z* mu, double rho1, double rho2, double rhoa,
                                            const mfb_z* R, const mfb_z* Q, double b, const mfb_z* cvalue, mfb_z* x) {
  if (!p || !lambda || !mu || !R || !Q) return fail(MFB_ERR_ARG, "mfb_harpor3d_solve_frequency: null argument");
  CK(cudaSetDevice(p->ctx->device));
  int r = assemble_por_device(p, omega, cd(lambda->re, lambda->im), cd(mu->re, mu->im), rho1, rho2, rhoa, cd(R->re, R->im), cd(Q->re, Q->im), b, cvalue);
  if (r) return r;
  r = factor_and_solve_resident(p);
  int r2 = collect_assembly_times(p); if (r2) return r2;
  if (r) return r;
  p->assembled = false;
  if (!x) return MFB_OK;
  return download_matrix(p, p->sys.bre, p->sys.bim, p->lda, p->n_dof, 1, x, p->n_dof, p->d_colperm);
}

extern "C" int mfb_get_solution(mfb_problem* p, mfb_z* x) {
  if (!p || !x) return fail(MFB_ERR_ARG, "mfb_get_solution: null argument");
  CK(cudaSetDevice(p->ctx->device));
  return download_matrix(p, p->sys.bre, p->sys.bim, p->lda, p->n_dof, 1, x, p->n_dof, p->rows_permuted ? p->d_colperm : nullptr);
}

extern "C" int mfb_get_entries(mfb_problem* p, int n, const int* rows, const int* cols, mfb_z* out) {
  if (!p || !rows || !cols || !out || n < 0) return fail(MFB_ERR_ARG, "mfb_get_entries: invalid argument");
  if (!p->assembled) return fail(MFB_ERR_ARG, "mfb_get_entries: no assembled (unfactorised) system is resident");
  for (int i = 0; i < n; i++) if (rows[i] < 0 || rows[i] >= p->n_dof || cols[i] < 0 || cols[i] >= p->n_dof) return fail(MFB_ERR_ARG, "mfb_get_entries: index out of range");
  CK(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  int *dr, *dc; double* dout;
  CK(cudaMalloc((void**)&dr, (size_t)std::max(n, 1) * 4)); CK(cudaMalloc((void**)&dc, (size_t)std::max(n, 1) * 4)); CK(cudaMalloc((void**)&dout, (size_t)std::max(n, 1) * 16));
  std::vector<int> irows(rows, rows + n), icols(cols, cols + n);
  if (p->rows_permuted) for (int i = 0; i < n; i++) { irows[i] = p->rowperm[rows[i]]; icols[i] = p->colperm[cols[i]]; }
  CK(cudaMemcpyAsync(dr, irows.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(dc, icols.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
  launch_get_entries(p->sys, n, dr, dc, dout, st);
  CK(cudaMemcpyAsync(out, dout, (size_t)n * 16, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
  cudaFree(dr); cudaFree(dc); cudaFree(dout);
  return MFB_OK;
}

extern "C" int mfb_residual(mfb_problem* p, const mfb_z* x, double* berr, double* rel_resid) {
  if (!p || !x) return fail(MFB_ERR_ARG, "mfb_residual: null argument");
  if (!p->assembled) return fail(MFB_ERR_ARG, "mfb_residual: no assembled (unfactorised) system is resident");
  CK(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  const int n = p->n_dof;
  double* d; CK(cudaMalloc((void**)&d, (size_t)5 * n * sizeof(double)));
  std::vector<double> hx(2 * (size_t)n);
  for (int i = 0; i < n; i++) { const int q = p->rows_permuted ? p->colperm[i] : i; hx[q] = x[i].re; hx[n + q] = x[i].im; }
  CK(cudaMemcpyAsync(d, hx.data(), (size_t)2 * n * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(d + 2 * (size_t)n, 0, (size_t)3 * n * 8, st));
  launch_residual(p->sys, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, d + 4 * (size_t)n, st);
  std::vector<double> h(3 * (size_t)n);
  CK(cudaMemcpyAsync(h.data(), d + 2 * (size_t)n, (size_t)3 * n * 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
  cudaFree(d);
  double be = 0.0, rmax = 0.0, smax = 0.0;
  for (int i = 0; i < n; i++) {
    double r = fabs(h[i]) + fabs(h[n + i]), sc = h[2 * (size_t)n + i];
    if (sc > 0.0 && r / sc > be) be = r / sc;
    rmax = std::max(rmax, r); smax = std::max(smax, sc);
  }
  if (berr) *berr = be;
  if (rel_resid) *rel_resid = smax > 0.0 ? rmax / smax : 0.0;
  return MFB_OK;
}

extern "C" int mfb_get_stats(mfb_problem* p, double* stats) {
  if (!p || !stats) return fail(MFB_ERR_ARG, "mfb_get_stats: null argument");
  memcpy(stats, p->stats, sizeof(p->stats));
  return MFB_OK;
}

extern "C" int mfb_plan_modes(mfb_problem* p, int n_pairs, const int* colloc, const int* elem, int* mode) {
  if (!p || !colloc || !elem || !mode) return fail(MFB_ERR_ARG, "mfb_plan_modes: null argument");
  CK(cudaSetDevice(p->ctx->device));
  std::vector<unsigned char> h((size_t)p->n_elem * p->ldp);
  CK(cudaMemcpy(h.data(), p->plan, h.size(), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n_pairs; i++) {
    if (colloc[i] < 0 || colloc[i] >= p->n_colloc || elem[i] < 0 || elem[i] >= p->n_elem) return fail(MFB_ERR_ARG, "mfb_plan_modes: index out of range");
    unsigned char m = h[(size_t)p->slot_of_elem[elem[i]] * p->ldp + p->cpos_of_colloc[colloc[i]]];
    mode[i] = (m < MAX_SETS) ? p->set_gln[m] : (m == PLAN_ADAPTIVE ? 100 : (m == PLAN_SINGULAR ? 200 : -1));
  }
  return MFB_OK;
}

extern "C" int mfb_stream_mark(mfb_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot >= 8) return fail(MFB_ERR_ARG, "mfb_stream_mark: invalid argument");
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->marks[slot], ctx->stream));
  return MFB_OK;
}
extern "C" int mfb_stream_elapsed(mfb_ctx* ctx, int slot0, int slot1, double* ms) {
  if (!ctx || !ms || slot0 < 0 || slot0 >= 8 || slot1 < 0 || slot1 >= 8) return fail(MFB_ERR_ARG, "mfb_stream_elapsed: invalid argument");
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventSynchronize(ctx->marks[slot1]));
  float t; CK(cudaEventElapsedTime(&t, ctx->marks[slot0], ctx->marks[slot1])); *ms = t;
  return MFB_OK;
}

extern "C" int mfb_measure_peaks(mfb_ctx* ctx, double* dfma_tflops, double* dmma_tflops, double* copy_gbs) {
  if (!ctx) return fail(MFB_ERR_ARG, "mfb_measure_peaks: null context");
  CK(cudaSetDevice(ctx->device));
  if (dfma_tflops) *dfma_tflops = bench_dfma(ctx->stream);
  if (dmma_tflops) *dmma_tflops = bench_dmma(ctx->stream);
  if (copy_gbs) *copy_gbs = bench_copy(ctx->stream);
  CK(cudaGetLastError());
  return MFB_OK;
}

extern "C" int mfb_zgemm_minus(mfb_ctx* ctx, int m, int n, int k, const mfb_z* A, int lda, const mfb_z* B, int ldb, mfb_z* C, int ldc, double* ms) {
  if (!ctx || !A || !B || !C) return fail(MFB_ERR_ARG, "mfb_zgemm_minus: null argument");
  if (m <= 0 || n <= 0 || k <= 0 || (k & 1)) return fail(MFB_ERR_ARG, "mfb_zgemm_minus: k must be even and sizes positive");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  long long la = (m + 31) / 32 * 32, lb = (k + 31) / 32 * 32, lc = la;
  double *dA, *dB, *dC, *stage;
  size_t smax = std::max(std::max((size_t)m * k, (size_t)k * n), (size_t)m * n);
  CK(cudaMalloc((void**)&dA, 2 * la * k * 8)); CK(cudaMalloc((void**)&dB, 2 * lb * n * 8)); CK(cudaMalloc((void**)&dC, 2 * lc * n * 8));
  CK(cudaMalloc((void**)&stage, smax * 16));
  CK(cudaMemsetAsync(dA, 0, 2 * la * k * 8, st)); CK(cudaMemsetAsync(dB, 0, 2 * lb * n * 8, st));
  CK(cudaMemcpy2DAsync(stage, (size_t)m * 16, A, (size_t)lda * 16, (size_t)m * 16, k, cudaMemcpyHostToDevice, st));
  launch_deinterleave(stage, m, m, k, dA, dA + la * k, la, nullptr, st);
  CK(cudaMemcpy2DAsync(stage, (size_t)k * 16, B, (size_t)ldb * 16, (size_t)k * 16, n, cudaMemcpyHostToDevice, st));
  launch_deinterleave(stage, k, k, n, dB, dB + lb * n, lb, nullptr, st);
  CK(cudaMemcpy2DAsync(stage, (size_t)m * 16, C, (size_t)ldc * 16, (size_t)m * 16, n, cudaMemcpyHostToDevice, st));
  launch_deinterleave(stage, m, m, n, dC, dC + lc * n, lc, nullptr, st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  CK(cudaEventRecord(e0, st));
  GemmTmaMaps tm; tm.ok = 0;
  gemm_tma_make_maps(tm, dA, dA + la * k, la, m, k, dB, dB + lb * n, lb, k, n);
  CK(cudaEventRecord(e0, st));   // the encoding of the maps (host work) is outside the timed launch
  if (gemm_tma_usable(tm, m, n, k)) zgemm_minus_planar_tma(tm, m, n, k, 0, 0, 0, 0, dC, dC + lc * n, lc, st);
  else zgemm_minus_planar(m, n, k, dA, dA + la * k, la, dB, dB + lb * n, lb, dC, dC + lc * n, lc, st);
  CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
  float t; cudaEventElapsedTime(&t, e0, e1); if (ms) *ms = t;
  launch_interleave(dC, dC + lc * n, lc, m, n, stage, m, nullptr, nullptr, 0, st);
  CK(cudaMemcpy2DAsync(C, (size_t)ldc * 16, stage, (size_t)m * 16, (size_t)m * 16, n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(stage); cudaEventDestroy(e0); cudaEventDestroy(e1);
  CK(cudaGetLastError());
  return MFB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Single-frequency multi-GPU mode: collocation-row blocks for the assembly, block-cyclic columns for the
// LU, NCCL to move row slabs to their column owners once and to broadcast each factorised panel.
// ---------------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------------
// Resident combination of assembled systems (coupled regions from single-region assemblies, multifebe_b200/host/coupled.py).
//   mfb_system_zero      dst: zero the resident matrix and right-hand side and mark it "assembled, host order" (no row permutation), so that
//                        mfb_zsolve(dst, n, NULL, ..., NULL, 1, 1) factorises and solves what the calls below accumulate
//   mfb_combine_columns  dst(row_map[r], dst_col[i]) += coef[i] * src(r, src_col[i]) for the first n_rows host rows of the ASSEMBLED system of src
//                        (host row / column indices on both sides; dst_col = -1: the right-hand side of dst); both problems on the same context
//   mfb_add_entries      dst(rows[i], cols[i]) += values[i] (cols = -1: right-hand side) -- the free terms
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int mfb_system_zero(mfb_problem* p) {
  if (!p) return fail(MFB_ERR_ARG, "mfb_system_zero: null problem");
  CK(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  CK(cudaMemsetAsync(p->sys.Are, 0, (size_t)2 * p->lda * p->n_dof * sizeof(double), st));
  CK(cudaMemsetAsync(p->sys.bre, 0, (size_t)2 * p->lda * sizeof(double), st));
  p->factored = false; p->assembled = true; p->rows_permuted = false; p->real_resident = false;
  return MFB_OK;
}
extern "C" int mfb_combine_columns(mfb_problem* dst, mfb_problem* src, int n_rows, const int* row_map, int n_terms, const int* src_col, const int* dst_col,
                                   const mfb_z* coef) {
  if (!dst || !src || !row_map || !src_col || !dst_col || !coef || n_rows < 0 || n_terms < 0) return fail(MFB_ERR_ARG, "mfb_combine_columns: invalid argument");
  if (dst->ctx != src->ctx) return fail(MFB_ERR_ARG, "mfb_combine_columns: both problems must live on the same context");
  if (!src->assembled || src->real_resident) return fail(MFB_ERR_ARG, "mfb_combine_columns: src holds no assembled complex system");
  if (!dst->assembled || dst->rows_permuted) return fail(MFB_ERR_ARG, "mfb_combine_columns: call mfb_system_zero(dst) first");
  if (n_rows > src->n_dof) return fail(MFB_ERR_ARG, "mfb_combine_columns: n_rows exceeds the size of src");
  std::vector<int> sr(n_rows), dr(n_rows), sc(n_terms), dc(n_terms);
  for (int r = 0; r < n_rows; r++) {
    if (row_map[r] < 0 || row_map[r] >= dst->n_dof) return fail(MFB_ERR_ARG, "mfb_combine_columns: row_map out of range");
    sr[r] = src->rows_permuted ? src->rowperm[r] : r; dr[r] = row_map[r];
  }
  for (int i = 0; i < n_terms; i++) {
    if (src_col[i] < 0 || src_col[i] >= src->n_dof || dst_col[i] < -1 || dst_col[i] >= dst->n_dof) return fail(MFB_ERR_ARG, "mfb_combine_columns: column out of range");
    sc[i] = src->rows_permuted ? src->colperm[src_col[i]] : src_col[i]; dc[i] = dst_col[i];
  }
  CK(cudaSetDevice(dst->ctx->device));
  cudaStream_t st = dst->ctx->stream;
  std::vector<void*> tmp;
  int *d_sr, *d_dr, *d_sc, *d_dc; double* d_cf;
  std::vector<double> cf(2 * (size_t)std::max(n_terms, 1));
  for (int i = 0; i < n_terms; i++) { cf[2 * i] = coef[i].re; cf[2 * i + 1] = coef[i].im; }
  UP(tmp, sr, &d_sr); UP(tmp, dr, &d_dr); UP(tmp, sc, &d_sc); UP(tmp, dc, &d_dc); UP(tmp, cf, &d_cf);
  launch_combine(src->sys, dst->sys, n_rows, d_sr, d_dr, n_terms, d_sc, d_dc, d_cf, st);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  for (void* q : tmp) cudaFree(q);
  return MFB_OK;
}
extern "C" int mfb_add_entries(mfb_problem* dst, int n, const int* rows, const int* cols, const mfb_z* values) {
  if (!dst || !rows || !cols || !values || n < 0) return fail(MFB_ERR_ARG, "mfb_add_entries: invalid argument");
  if (!dst->assembled || dst->rows_permuted) return fail(MFB_ERR_ARG, "mfb_add_entries: call mfb_system_zero(dst) first");
  for (int i = 0; i < n; i++) if (rows[i] < 0 || rows[i] >= dst->n_dof || cols[i] < -1 || cols[i] >= dst->n_dof) return fail(MFB_ERR_ARG, "mfb_add_entries: index out of range");
  CK(cudaSetDevice(dst->ctx->device));
  cudaStream_t st = dst->ctx->stream;
  std::vector<void*> tmp;
  std::vector<int> r(rows, rows + n), c(cols, cols + n); std::vector<double> v(2 * (size_t)std::max(n, 1));
  for (int i = 0; i < n; i++) { v[2 * i] = values[i].re; v[2 * i + 1] = values[i].im; }
  int *d_r, *d_c; double* d_v;
  UP(tmp, r, &d_r); UP(tmp, c, &d_c); UP(tmp, v, &d_v);
  launch_add_entries(dst->sys, n, d_r, d_c, d_v, st);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  for (void* q : tmp) cudaFree(q);
  return MFB_OK;
}

// Host-only helper (no GPU needed): the geometry-only pieces of the free term of a boundary node from the unit normals and the unit boundary
// tangents of the elements that meet there -- cp = (2 pi + sum of the signed dihedral angles) / 4 pi (fbem_bem_pot3d_sbie_freeterm,
// lib/fbem/src/bem_stapot3d.f90:155-296) and the tensor sum_b[9] of Mantic's formula, so that c_lk = cp delta_lk - sum_b[l][k] / (8 pi (1 - nu))
// (fbem_bem_harela3d_sbie_freeterm, lib/fbem/src/bem_harela3d.f90:365-542).  Used by the coupled-region assembly of multifebe_b200/host/coupled.py,
// whose auxiliary single-region problems are set up without free terms.  Returns 2 when the normals / tangents configuration is not valid.
extern "C" int mfb_freeterm_terms(int n_elements, const double* normals, const double* tangents, double tol, double* cp, double* sum_b) {
  if (n_elements <= 0 || !normals || !tangents || !cp || !sum_b) return fail(MFB_ERR_ARG, "mfb_freeterm_terms: invalid argument");
  if (mfbh::mantic_terms(n_elements, normals, tangents, tol, cp, sum_b)) return fail(2, "mfb_freeterm_terms: the normals/tangents configuration is not valid");
  return MFB_OK;
}

extern "C" int mfb_dist_layout(int n, int nb, int nranks, int rank, int* n_local_cols, int* local_to_global) {
  if (n <= 0 || nb <= 0 || nranks <= 0 || rank < 0 || rank >= nranks) return fail(MFB_ERR_ARG, "mfb_dist_layout: invalid argument");
  const int ncl = dist_ncols_local(n, nb, nranks, rank);
  if (n_local_cols) *n_local_cols = ncl;
  if (local_to_global) for (int lc = 0; lc < ncl; lc++) local_to_global[lc] = (lc / nb * nranks + rank) * nb + lc % nb;
  return MFB_OK;
}

// Tiles -> ranks.  Bulk tiles come in row order and the layers of one row block share row0, so a rank gets a contiguous
// run of row blocks (about n_tiles / nranks tiles, layers counted); loose tiles (arbitrary rows of the last internal
// rows) go to the last rank, whose row range extends to n_dof.  row_bounds[nranks + 1].
extern "C" int mfb_dist_partition_tiles(int n_tiles, const int* tile_row0, const int* tile_nbytes, int n_dof, int nranks, int* tile_rank, int* row_bounds) {
  if (n_tiles <= 0 || !tile_row0 || !tile_nbytes || nranks <= 0 || !tile_rank || !row_bounds) return fail(MFB_ERR_ARG, "mfb_dist_partition_tiles: invalid argument");
  int r = 0, acc = 0, bulk_end = 0;
  row_bounds[0] = 0;
  for (int t = 0; t < n_tiles;) {
    if (tile_nbytes[t] == 0) { tile_rank[t] = nranks - 1; t++; continue; }
    int t1 = t; while (t1 < n_tiles && tile_nbytes[t1] > 0 && tile_row0[t1] == tile_row0[t]) t1++;
    if (r < nranks - 1 && acc > 0 && (long long)acc * nranks >= (long long)n_tiles * (r + 1)) { r++; row_bounds[r] = tile_row0[t]; }
    for (int q = t; q < t1; q++) tile_rank[q] = r;
    acc += t1 - t;
    bulk_end = std::max(bulk_end, tile_row0[t] + tile_nbytes[t] / 8);
    t = t1;
  }
  for (int q = r + 1; q < nranks; q++) row_bounds[q] = bulk_end;
  row_bounds[nranks] = n_dof;
  return MFB_OK;
}

extern "C" int mfb_dist_unique_id(char* id128) {
  if (!id128) return fail(MFB_ERR_ARG, "mfb_dist_unique_id: null argument");
  std::string err;
  if (nccl_unique_id(id128, err)) return fail(MFB_ERR_CUDA, "mfb_dist_unique_id: " + err);
  return MFB_OK;
}

static int dist_setup(mfb_problem* p, int rank, int nranks, bool loopback, const char* id128, int nb) {
  if (!p) return fail(MFB_ERR_ARG, "mfb_dist_init: null problem");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(MFB_ERR_ARG, "mfb_dist_init: invalid rank / nranks");
  if (nb <= 0) nb = 256;
  if (nb % 32 || nb > 1024) return fail(MFB_ERR_ARG, "mfb_dist_init: block size must be a multiple of 32, at most 1024");
  CK(cudaSetDevice(p->ctx->device));
  dist_release(p);
  DistState& d = p->dist;
  d.loopback = loopback; d.rank = rank; d.P = nranks; d.nb = nb;
  if (loopback) d.comm = make_loopback_comm(nranks);
  else { std::string err; d.comm = make_nccl_comm(rank, nranks, id128, err); if (!d.comm) return fail(MFB_ERR_CUDA, "mfb_dist_init: " + err); }
  const int n = p->n_dof, n_tiles = p->colloc.n_tiles;
  d.rb.assign(nranks + 1, 0); d.tile_rank.assign(n_tiles, 0);
  int r = mfb_dist_partition_tiles(n_tiles, p->h_tile_row0.data(), p->h_tile_nbytes.data(), n, nranks, d.tile_rank.data(), d.rb.data());
  if (r) return r;
  d.on = true;
  for (int i = 0; i < 6; i++) cudaEventCreate(&d.ev[i]);
  CK(cudaMalloc((void**)&d.d_mask, (size_t)n_tiles));
  CK(cudaMalloc((void**)&d.bsum, (size_t)2 * p->lda * sizeof(double)));
  d.lu.n = n; d.lu.nb = nb; d.lu.nblk = (n + nb - 1) / nb; d.lu.P = nranks; d.lu.lda = p->lda; d.lu.comm = d.comm; d.lu.ms_lu = d.lu.ms_solve = 0.f;
  { const char* e = getenv("MFB_DIST_OWNER_WAITS"); d.lu.owner_waits_for_panel = e ? atoi(e) : 1; }
  const int n_local = loopback ? nranks : 1;
  d.lu.r.resize(n_local);
  for (int i = 0; i < n_local; i++) {
    int e = dist_rank_alloc(d.lu.r[i], loopback ? i : rank, n, p->lda, nb, nranks, p->ctx->stream, !loopback);
    if (e) return fail(MFB_ERR_CUDA, "mfb_dist_init: allocation of the distributed LU workspace failed");
  }
  if (!loopback) {
    std::vector<unsigned char> mask(n_tiles);
    for (int t = 0; t < n_tiles; t++) mask[t] = d.tile_rank[t] == rank;
    CK(cudaMemcpy(d.d_mask, mask.data(), n_tiles, cudaMemcpyHostToDevice));
    const size_t my_rows = (size_t)(d.rb[rank + 1] - d.rb[rank]), my_cols = (size_t)d.lu.r[0].ncl;
    d.soff.assign(nranks, 0); d.roff.assign(nranks, 0); d.ns.assign(nranks, 0); d.nr.assign(nranks, 0);
    size_t so = 0, ro = 0;
    for (int q = 0; q < nranks; q++) {
      if (q == rank) continue;
      d.ns[q] = 2 * (size_t)dist_ncols_local(n, nb, nranks, q) * my_rows; d.soff[q] = so; so += d.ns[q];
      d.nr[q] = 2 * my_cols * (size_t)(d.rb[q + 1] - d.rb[q]); d.roff[q] = ro; ro += d.nr[q];
    }
    CK(cudaMalloc((void**)&d.sendbuf, std::max<size_t>(so, 1) * sizeof(double)));
    CK(cudaMalloc((void**)&d.recvbuf, std::max<size_t>(ro, 1) * sizeof(double)));
  }
  return MFB_OK;
}
extern "C" int mfb_dist_init(mfb_problem* p, int rank, int nranks, const char* id128, int nb) {
  if (!id128) return fail(MFB_ERR_ARG, "mfb_dist_init: null unique id");
  return dist_setup(p, rank, nranks, false, id128, nb);
}
extern "C" int mfb_dist_init_loopback(mfb_problem* p, int nranks, int nb) { return dist_setup(p, 0, nranks, true, nullptr, nb); }

extern "C" int mfb_dist_info(mfb_problem* p, int* rank, int* nranks, int* row_bounds /* nranks + 1 */, int* n_local_cols) {
  if (!p || !p->dist.on) return fail(MFB_ERR_ARG, "mfb_dist_info: the problem is not in multi-GPU mode");
  const DistState& d = p->dist;
  if (rank) *rank = d.rank;
  if (nranks) *nranks = d.P;
  if (row_bounds) for (int q = 0; q <= d.P; q++) row_bounds[q] = d.rb[q];
  if (n_local_cols) *n_local_cols = d.lu.r[0].ncl;
  return MFB_OK;
}

// factorise + solve the distributed system whose local columns (and right-hand side column) are in place; x (host) or NULL
static int dist_factor_solve(mfb_problem* p, mfb_z* x) {
  DistState& d = p->dist;
  cudaStream_t st = p->ctx->stream;
  CK(cudaEventRecord(d.ev[2], st));
  int e = zgetrf_dist(d.lu);
  if (e) return fail(MFB_ERR_CUDA, std::string("zgetrf_dist: ") + (e > 0 ? cudaGetErrorString((cudaError_t)e) : d.comm->last_error()));
  CK(cudaEventRecord(d.ev[3], st));
  e = zgetrs_dist(d.lu);
  if (e) return fail(MFB_ERR_CUDA, std::string("zgetrs_dist: ") + (e > 0 ? cudaGetErrorString((cudaError_t)e) : d.comm->last_error()));
  CK(cudaEventRecord(d.ev[4], st));
  CK(cudaStreamSynchronize(st));
  float t;
  cudaEventElapsedTime(&t, d.ev[2], d.ev[3]); p->stats[MFB_STAT_MS_DIST_LU] = t;
  cudaEventElapsedTime(&t, d.ev[3], d.ev[4]); p->stats[MFB_STAT_MS_DIST_SOLVE] = t;
  double fl = 0.0; long long launches = 0; int info = 0;
  for (auto& R : d.lu.r) {
    fl += R.gemm_flops; launches += R.w.launches;
    int i1 = 0; CK(cudaMemcpy(&i1, R.w.info, sizeof(int), cudaMemcpyDeviceToHost));
    if (i1 > 0 && (info == 0 || i1 < info)) info = i1;
  }
  p->stats[MFB_STAT_GEMM_FLOPS] = fl; p->stats[MFB_STAT_LU_LAUNCHES] = (double)launches;
  p->factored = false; p->assembled = false;
  if (info > 0) { char buf[128]; snprintf(buf, sizeof(buf), "zgetrf (distributed): U(%d,%d) is exactly zero, the matrix is singular", info, info); return fail(info, buf); }
  if (!x) return MFB_OK;
  DistRank& R0 = d.lu.r[0];
  return download_matrix(p, R0.xfin, R0.xfin + p->lda, p->lda, p->n_dof, 1, x, p->n_dof, p->rows_permuted ? p->d_colperm : nullptr);
}

extern "C" int mfb_dist_solve_frequency(mfb_problem* p, double omega, const mfb_z* lambda, const mfb_z* mu, double rho, const mfb_z* nu,
                                        const mfb_z* cvalue, mfb_z* x) {
  if (!p || !lambda || !mu || !nu) return fail(MFB_ERR_ARG, "mfb_dist_solve_frequency: null argument");
  DistState& d = p->dist;
  if (!d.on) return fail(MFB_ERR_ARG, "mfb_dist_solve_frequency: call mfb_dist_init first");
  CK(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  const int n = p->n_dof, P = d.P, nb = d.nb; const long long lda = p->lda;
  const cd la(lambda->re, lambda->im), m_(mu->re, mu->im), nu_(nu->re, nu->im);
  CK(cudaEventRecord(d.ev[0], st));
  if (!d.loopback) {
    DistRank& R = d.lu.r[0];
    p->colloc.tile_active = d.d_mask; p->skip_cond = d.rank != P - 1;   // rows that no collocation point feeds belong to the last rank's row range
    int r = assemble_device(p, omega, la, m_, rho, nu_, cvalue);
    p->colloc.tile_active = nullptr; p->skip_cond = false;
    if (r) return r;
    CK(cudaEventRecord(d.ev[1], st));
    // the right-hand side: every rank holds the rows it assembled; sum -> replicated
    int rk = d.rank; double* bb = p->sys.bre;
    if (P > 1 && d.comm->allreduce_sum(&rk, &bb, (size_t)2 * lda, &st, 1)) return fail(MFB_ERR_CUDA, std::string("allreduce of the right-hand side: ") + d.comm->last_error());
    // row slab -> column owners
    std::vector<double*> sp(P, nullptr), rp(P, nullptr);
    for (int q = 0; q < P; q++) {
      if (q == d.rank) continue;
      sp[q] = d.sendbuf + d.soff[q]; rp[q] = d.recvbuf + d.roff[q];
      launch_pack_slab(p->sys.Are, p->sys.Aim, lda, n, nb, P, q, d.rb[d.rank], d.rb[d.rank + 1], sp[q], st);
    }
    if (P > 1 && d.comm->exchange(sp.data(), d.ns.data(), rp.data(), d.nr.data(), st)) return fail(MFB_ERR_CUDA, std::string("redistribution: ") + d.comm->last_error());
    launch_copy_own(p->sys.Are, p->sys.Aim, lda, n, nb, P, d.rank, d.rb[d.rank], d.rb[d.rank + 1], R.Lre, R.Lim, st);
    for (int q = 0; q < P; q++) if (q != d.rank) launch_unpack_slab(rp[q], R.ncl, d.rb[q], d.rb[q + 1], R.Lre, R.Lim, lda, st);
    CK(cudaMemcpyAsync(R.Lre + (size_t)R.ncl * lda, p->sys.bre, (size_t)lda * 8, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(R.Lim + (size_t)R.ncl * lda, p->sys.bim, (size_t)lda * 8, cudaMemcpyDeviceToDevice, st));
  } else {
    // virtual ranks, one after the other on this GPU: assemble the rank's row blocks, hand the slab to every column owner
    CK(cudaMemsetAsync(d.bsum, 0, (size_t)2 * lda * 8, st));
    std::vector<unsigned char> mask(p->colloc.n_tiles);
    for (int r = 0; r < P; r++) {
      for (int t = 0; t < p->colloc.n_tiles; t++) mask[t] = d.tile_rank[t] == r;
      CK(cudaMemcpyAsync(d.d_mask, mask.data(), mask.size(), cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st));
      p->colloc.tile_active = d.d_mask; p->skip_cond = r != P - 1;
      int rr = assemble_device(p, omega, la, m_, rho, nu_, r == 0 ? cvalue : nullptr);
      p->colloc.tile_active = nullptr; p->skip_cond = false;
      if (rr) return rr;
      for (int q = 0; q < P; q++) launch_copy_own(p->sys.Are, p->sys.Aim, lda, n, nb, P, q, d.rb[r], d.rb[r + 1], d.lu.r[q].Lre, d.lu.r[q].Lim, st);
      launch_add_into(d.bsum, p->sys.bre, (size_t)2 * lda, st);
    }
    CK(cudaEventRecord(d.ev[1], st));
    for (int q = 0; q < P; q++) {
      DistRank& R = d.lu.r[q];
      CK(cudaMemcpyAsync(R.Lre + (size_t)R.ncl * lda, d.bsum, (size_t)lda * 8, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(R.Lim + (size_t)R.ncl * lda, d.bsum + lda, (size_t)lda * 8, cudaMemcpyDeviceToDevice, st));
    }
  }
  p->rows_permuted = true;
  int r = dist_factor_solve(p, x);
  float t;
  cudaEventElapsedTime(&t, d.ev[0], d.ev[1]); p->stats[MFB_STAT_MS_ASSEMBLE] = t;
  cudaEventElapsedTime(&t, d.ev[1], d.ev[2]); p->stats[MFB_STAT_MS_REDIST] = t;
  cudaEventElapsedTime(&t, d.ev[0], d.ev[4]); p->stats[MFB_STAT_MS_DIST_TOTAL] = t;
  return r;
}

// Test / seam-2 entry of the multi-GPU mode: factorise and solve a host matrix that EVERY rank passes in full (each rank
// keeps only its block-cyclic columns).  A (lda x n, host, not overwritten), b (n) -> x (n), host row / column order.
extern "C" int mfb_dist_zsolve(mfb_problem* p, int n, const mfb_z* A, int lda_h, const mfb_z* b, mfb_z* x, int* ipiv) {
  if (!p || !A || !b || !x) return fail(MFB_ERR_ARG, "mfb_dist_zsolve: null argument");
  DistState& d = p->dist;
  if (!d.on) return fail(MFB_ERR_ARG, "mfb_dist_zsolve: call mfb_dist_init first");
  if (n != p->n_dof || lda_h < n) return fail(MFB_ERR_ARG, "mfb_dist_zsolve: n must equal the problem's n_dof");
  CK(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  const long long lda = p->lda;
  int r = upload_matrix(p, A, lda_h, n, n, p->sys.Are, p->sys.Aim, lda); if (r) return r;
  r = upload_matrix(p, b, n, n, 1, p->sys.bre, p->sys.bim, lda); if (r) return r;
  p->rows_permuted = false; p->assembled = false; p->factored = false;
  CK(cudaEventRecord(d.ev[0], st)); CK(cudaEventRecord(d.ev[1], st));
  for (auto& R : d.lu.r) {
    launch_copy_own(p->sys.Are, p->sys.Aim, lda, n, d.nb, d.P, R.rank, 0, n, R.Lre, R.Lim, st);
    CK(cudaMemcpyAsync(R.Lre + (size_t)R.ncl * lda, p->sys.bre, (size_t)lda * 8, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(R.Lim + (size_t)R.ncl * lda, p->sys.bim, (size_t)lda * 8, cudaMemcpyDeviceToDevice, st));
  }
  r = dist_factor_solve(p, x);
  if (ipiv) CK(cudaMemcpy(ipiv, d.lu.r[0].ipiv, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// Static 3D elasticity: build_lse_mechanics_bem_staela + assemble_bem_staela_equation +
// solve_lse_r (dgetrf / dgetrs) on the same problem object (mesh, collocation tiles and quadrature plan do not depend on
// the kernel: fbem_bem_staela3d_sbie_auto takes the decisions of fbem_bem_harela3d_sbie_auto, bem_staela3d.f90:524-589).
// ---------------------------------------------------------------------------------------------------------------------
static int download_real(mfb_problem* p, const double* re, long long ld, int rows, int cols, double* host, long long ldh, const int* rowperm, const int* colperm) {
  cudaStream_t st = p->ctx->stream;
  int chunk = (int)std::max<long long>(1, std::min<long long>(cols, (256ll << 20) / (8ll * rows)));
  DevBuf sb; CK(cudaMalloc(&sb.p, (size_t)chunk * rows * 8)); double* stage = (double*)sb.p;
  for (int c0 = 0; c0 < cols; c0 += chunk) {
    int nc = std::min(chunk, cols - c0);
    launch_gather_real(re, ld, rows, nc, stage, rows, rowperm, colperm, c0, st);
    CK(cudaMemcpy2DAsync(host + (long long)c0 * ldh, (size_t)ldh * 8, stage, (size_t)rows * 8, (size_t)rows * 8, nc, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  return MFB_OK;
}
static int upload_real(mfb_problem* p, const double* host, long long ldh, int rows, int cols, double* re, long long ld, const int* rowperm) {
  cudaStream_t st = p->ctx->stream;
  int chunk = (int)std::max<long long>(1, std::min<long long>(cols, (256ll << 20) / (8ll * rows)));
  DevBuf sb; CK(cudaMalloc(&sb.p, (size_t)chunk * rows * 8)); double* stage = (double*)sb.p;
  for (int c0 = 0; c0 < cols; c0 += chunk) {
    int nc = std::min(chunk, cols - c0);
    CK(cudaMemcpy2DAsync(stage, (size_t)rows * 8, host + (long long)c0 * ldh, (size_t)ldh * 8, (size_t)rows * 8, nc, cudaMemcpyHostToDevice, st));
    launch_scatter_real(stage, rows, rows, nc, re + (long long)c0 * ld, ld, rowperm, st);
    CK(cudaStreamSynchronize(st));
  }
  return MFB_OK;
}
static int assemble_static_device(mfb_problem* p, double mu, double nu, const double* cvalue) {

  if (!(mu > 0.0) || !(nu > -1.0 && nu < 0.5)) return fail(MFB_ERR_ARG, "static assembly: mu must be positive and nu in (-1, 0.5)");
  KParams K, Q; host_kparams_static(mu, nu, K); scale_kparams(K, Q);
  std::vector<mfb_z> cv;
  if (cvalue) { cv.resize((size_t)3 * p->n_node); for (size_t i = 0; i < cv.size(); i++) { cv[i].re = cvalue[i]; cv[i].im = 0.0; } }
  int r = assemble_device_k(p, K, Q, cd(nu, 0.0), cvalue ? cv.data() : nullptr, true);
  if (r) return r;
  if (cvalue) CK(cudaStreamSynchronize(p->ctx->stream));   // cv is a local staging buffer
  return MFB_OK;
}
extern "C" int mfb_staela3d_assemble(mfb_problem* p, double mu, double nu, const double* cvalue, double* A, double* b) {
  if (!p) return fail(MFB_ERR_ARG, "mfb_staela3d_assemble: null problem");
  CK(cudaSetDevice(p->ctx->device));
  int r = assemble_static_device(p, mu, nu, cvalue); if (r) return r;
  r = collect_assembly_times(p); if (r) return r;
  if (A) { r = download_real(p, p->sys.Are, p->lda, p->n_dof, p->n_dof, A, p->n_dof, p->d_rowperm, p->d_colperm); if (r) return r; }
  if (b) { r = download_real(p, p->sys.bre, p->lda, p->n_dof, 1, b, p->n_dof, p->d_rowperm, nullptr); if (r) return r; }
  return MFB_OK;
}
// solve_lse_r(n_dof,A,ipiv,...,n_rhs,b,factorize,...) (src/solve_lse_r.f90:25-232: dgetrf :137, dgetrs :189): same conventions as
// mfb_zsolve with real arrays.  A == NULL: the device-resident system of the last mfb_staela3d_assemble.
extern "C" int mfb_dsolve(mfb_problem* p, int n, double* A, int lda, int* ipiv, double* b, int nrhs, int factorize) {
  if (!p) return fail(MFB_ERR_ARG, "mfb_dsolve: null problem");
  if (n != p->n_dof) return fail(MFB_ERR_ARG, "mfb_dsolve: n must equal the problem's n_dof");
  if (nrhs < 0 || (A && lda < n)) return fail(MFB_ERR_ARG, "mfb_dsolve: invalid nrhs/lda");
  CK(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  int r;
  if (factorize) {
    if (A) { r = upload_real(p, A, lda, n, n, p->sys.Are, p->lda, nullptr); if (r) return r; p->rows_permuted = false; p->real_resident = true; }
    else if (!p->assembled) return fail(MFB_ERR_ARG, "mfb_dsolve: factorize=1 with A == NULL but no assembled system is resident");
    else if (!p->real_resident) return fail(MFB_ERR_ARG, "mfb_dsolve: the resident system is complex (harmonic assembly); use mfb_zsolve");
    p->assembled = false;
    r = factor_device(p, n, lu_timing());
    if (ipiv) memcpy(ipiv, p->h_ipiv.data(), (size_t)n * sizeof(int));
    if (A) { int r2 = download_real(p, p->sys.Are, p->lda, n, n, A, lda, nullptr, nullptr); if (r2) return r2; }
    if (r) return r;
  } else if (!p->factored || !p->real_resident) return fail(MFB_ERR_ARG, "mfb_dsolve: factorize=0 but no real factors are resident");
  if (nrhs == 0) return MFB_OK;
  double* bre = p->sys.bre; long long ldb = p->lda;
  DevBuf tmp;   // freed on every exit path
  if (b) {
    if (nrhs > 1) { CK(cudaMalloc((void**)&tmp.p, (size_t)p->lda * nrhs * sizeof(double))); bre = (double*)tmp.p; }
    r = upload_real(p, b, n, n, nrhs, bre, ldb, p->rows_permuted ? p->d_rowperm : nullptr); if (r) return r;
  } else if (nrhs != 1) return fail(MFB_ERR_ARG, "mfb_dsolve: device-resident rhs has a single column");
  CK(cudaEventRecord(p->ev[6], st));
  int e = zgetrs_planar(p->sys.Are, nullptr, p->lda, n, p->d_perm, bre, nullptr, ldb, nrhs, st, p->lu.inv);
  if (e) return fail(MFB_ERR_CUDA, std::string("dgetrs (planar): ") + cudaGetErrorString((cudaError_t)e));
  CK(cudaEventRecord(p->ev[7], st)); CK(cudaEventSynchronize(p->ev[7]));
  float t; cudaEventElapsedTime(&t, p->ev[6], p->ev[7]); p->stats[MFB_STAT_MS_SOLVE] = t;
  if (b) { r = download_real(p, bre, ldb, n, nrhs, b, n, p->rows_permuted ? p->d_colperm : nullptr, nullptr); if (r) return r; }
  return MFB_OK;
}
// assemble + dgetrf + dgetrs without leaving the device (the body of the static driver, src/multifebe.f90: build_lse_mechanics_static,
// solve_lse_r, assign_solution): x[n_dof] = solution in the host's column order
extern "C" int mfb_staela3d_solve(mfb_problem* p, double mu, double nu, const double* cvalue, double* x) {
  if (!p) return fail(MFB_ERR_ARG, "mfb_staela3d_solve: null problem");
  CK(cudaSetDevice(p->ctx->device));
  int r = assemble_static_device(p, mu, nu, cvalue); if (r) return r;
  r = factor_device(p, p->n_dof, lu_timing());
  int r2 = collect_assembly_times(p); if (r2) return r2;
  if (r) return r;
  cudaStream_t st = p->ctx->stream;
  CK(cudaEventRecord(p->ev[6], st));
  int e = zgetrs_planar(p->sys.Are, nullptr, p->lda, p->n_dof, p->d_perm, p->sys.bre, nullptr, p->lda, 1, st, p->lu.inv, p->lu.solve_ws);
  if (e) return fail(MFB_ERR_CUDA, std::string("dgetrs (planar): ") + cudaGetErrorString((cudaError_t)e));
  CK(cudaEventRecord(p->ev[7], st)); CK(cudaEventSynchronize(p->ev[7]));
  float t; cudaEventElapsedTime(&t, p->ev[6], p->ev[7]); p->stats[MFB_STAT_MS_SOLVE] = t;
  p->assembled = false;
  if (!x) return MFB_OK;
  return download_real(p, p->sys.bre, p->lda, p->n_dof, 1, x, p->n_dof, p->d_colperm, nullptr);
}

// r = A x - b of the assembled, not yet factorised, device-resident system (host row and column order).  With collocation points
// placed INSIDE the region (colloc_elem = -1 at set-up: no free term) the rows of those points are Somigliana's identity, i.e.
// u(x_ip) = sum_e (g t - h u) = -(A x - b) at those rows: the displacement part of the reference's interior-point pass
// (src/calculate_internal_points_mechanics_bem_harela.f90:160-178 with fbem_bem_harela3d_sbie_auto, :372-420).
extern "C" int mfb_residual_vector(mfb_problem* p, const mfb_z* x, mfb_z* r) {
  if (!p || !x || !r) return fail(MFB_ERR_ARG, "mfb_residual_vector: null argument");
  if (!p->assembled) return fail(MFB_ERR_ARG, "mfb_residual_vector: no assembled (unfactorised) system is resident");
  CK(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  const int n = p->n_dof;
  double* d; CK(cudaMalloc((void**)&d, (size_t)5 * n * sizeof(double)));
  std::vector<double> hx(2 * (size_t)n);
  for (int i = 0; i < n; i++) { const int q = p->rows_permuted ? p->colperm[i] : i; hx[q] = x[i].re; hx[n + q] = x[i].im; }
  CK(cudaMemcpyAsync(d, hx.data(), (size_t)2 * n * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(d + 2 * (size_t)n, 0, (size_t)3 * n * 8, st));
  launch_residual(p->sys, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, d + 4 * (size_t)n, st);
  std::vector<double> h(2 * (size_t)n);
  CK(cudaMemcpyAsync(h.data(), d + 2 * (size_t)n, (size_t)2 * n * 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
  cudaFree(d);
  for (int i = 0; i < n; i++) { const int q = p->rows_permuted ? p->rowperm[i] : i; r[i].re = h[q]; r[i].im = h[n + q]; }
  return MFB_OK;
}
