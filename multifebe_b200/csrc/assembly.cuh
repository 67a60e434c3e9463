// assembly.cuh -- device-side data layout and launchers of the B200 assembly kernels.
#pragma once
#include <cuda_runtime.h>
#include "bem_math.cuh"

namespace mfbd {

const int MAX_SETS = 16;
const int K1_ERANGE = 512;                // elements per K1 task at most (<= 2048: queue entries carry an 11-bit element offset)
const unsigned char PLAN_NEAR = 250;      // classifier could not settle the pair with the ball test -> host planner
const unsigned char PLAN_ADAPTIVE = 254;  // Telles + subdivision leaf list (kernel K2)
const unsigned char PLAN_SINGULAR = 253;  // polar transformation (kernel K3)
const unsigned char PLAN_NONE = 255;

// One element-type group (elements sorted by type; "slot" = position in that order).
struct DevGroup {
  int et, nn, n_elem, slot0;     // slot0: first element slot of the group
  int ndof;                      // dofs per node: 3 (elastic; the strides below are written for it), 1 (inviscid fluid, potential.cuh: stride nn)
  const double* xn;              // [n_elem][3*nn] node coordinates
  const int* ecol;               // [n_elem][3*nn] column of A for (node j, dof k); index j*3+k
  const int* ecol2;              // NULL, or [n_elem][3*nn]: column of t_k for a dof whose u_k AND t_k are unknown (kind 2: local-axes conditions), -1 elsewhere
  const unsigned char* ekind;    // [n_elem][3*nn] 0: u known (A -= g, b -= h*u), 1: t known (A += h, b += g*t), 2: both unknown (A(:,ecol) += h, A(:,ecol2) -= g)
  const int* enode;              // [n_elem][nn]   global node ids
  const unsigned char* erev;     // [n_elem] reversed orientation (h -> -h)
  double* ecv;                   // [n_elem][3*nn][2] prescribed value of (j,k), refreshed per frequency
  const unsigned char* einfo;    // [n_elem] bits 0-2: kind of dof k (when the same for every node j), bit 3: kinds uniform over j, bit 4: reversed,
                                 // bits 5-7: multiplier -1 of a symmetry image on dof k (elastic: symconf_t(k); fluid: bit 5 = symconf_s); poroelastic: bits 4-7 = dofs 0-3, no reversed bit
  unsigned char* ecvnz;          // [n_elem] 1 when some prescribed value of the element is nonzero (refreshed per frequency)
  const unsigned char* c10;      // NULL, or [3*n_node]: 1 where ctype = 10 (normal pressure known): the prescribed value is cvalue * nfn
  const double* nfn;             // [3*n_node] nodal unit normals, negated for the nodes of a reversed boundary (with c10)
  const double* einc;            // NULL, or [n_elem][3*nn][4]: incident field (u_k re, im, t_k re, im) at node j of the element (element()%incident_c): every
                                 // pair then adds h u_inc - g t_inc to b (assemble_bem_harela_equation.f90:651-666); elements with uniform kinds run as K1 mode 1, the others as mode 2
  const double* ball;            // [n_elem][5]: centre(3), radius, characteristic length
  const int* gln_far;            // [n_elem]
  int n_ranges;                  // K1 tasks = (collocation tile) x (element range); long ranges first, short ones last (load balance)
  const int* range_start;        // [n_ranges+1]
  const int* range_of;           // [n_elem] range of an element
  int* range_modes;              // [n_ranges] bit m set: the range holds elements of K1 mode m (refreshed per frequency)
  int has_mixed;                 // 1: some element has boundary-condition kinds that differ between its nodes (K1 mode 2)
  int cols3;                     // 1: the three dof columns of every element node are consecutive (col(j,k) = col(j,0) + k)
  int n_sets; int set_gln[MAX_SETS]; int ngp[MAX_SETS];
  const double* pts[MAX_SETS];   // [n_elem][ngp][6+nn]: x(3), n(3), phi_j*J*w
};

struct DevSystem {
  double *Are, *Aim; long long lda; int n_dof;   // planar (split re/im) column-major system matrix
  double *bre, *bim;
};

// Collocation points live in TILES of 32 lanes (position cpos = 32*tile + lane, padding lanes have crow = -1).  A tile
// is one layer of a block of <= 32 row nodes whose 3 x cnt matrix rows are consecutive in the library's INTERNAL row order
// (row0 even, cnt even => the block's rows of one matrix column are one 16-byte aligned run of nbytes = 24*cnt bytes that
// a single bulk reduce can update); layer l holds the l-th collocation point of every node of the block (MCA nodes have
// several).  Loose tiles (nbytes = 0) hold arbitrary points and are flushed with per-lane RED.
struct DevColloc {
  int n_colloc, ldp;             // ldp = 32 * n_tiles = number of lane positions (n_colloc == ldp on the device)
  const double* cx;              // [3][ldp] collocation points (SoA)
  const int* crow;               // [3][ldp] internal A rows of the three equations of each collocation point, -1 = padding lane
  int n_tiles;
  const int* tile_row0;          // [n_tiles] first internal row of the tile's run
  const int* tile_nbytes;        // [n_tiles] 24*cnt, or 0 for a loose tile
  const double* cn;              // [3][ldp] unit normal at the collocation point, or NULL: hypersingular equation (interior-point stresses)
  const unsigned char* tile_active;   // [n_tiles] or NULL = all: tiles (row blocks) this rank assembles (single-frequency multi-GPU mode)
};

struct DevClassify {
  double far_thr[32]; double far_dmax; int ps_gln_max;
};

// adaptive (quasi-singular) work lists of one group
struct DevAdaptive {
  int n_pairs; const int* pair_cpos; const int* pair_elem; const int* pair_leaf0;  // [n_pairs+1]
  const double* leaf_d;          // [n_leaves][16]: xi_s(8), tp1(4), tp2(4)
  const int* leaf_gln;           // [n_leaves]
};
// singular work lists of one group
struct DevSingular {
  int n_pairs; const int* pair_cpos; const int* pair_elem; const int* pair_ray0;   // [n_pairs+1]
  const double* pair_d;          // [n_pairs][14]: xi_i(2), x_i(3), hli(9)
  const double* rays;            // [n_rays][4]: cos, sin, rhoij, w
};
// free-term entries: value added to h(j,l,k) of the collocation point before the BC-aware scatter
struct DevFreeTerm {
  int n; const int* cpos; const int* slot; const int* jk; const int* l; const double* val;  // val[n][2] = (alpha, beta): value = alpha + beta*F
  const int* slot_off;           // [n_slots+1] offset of each element slot in the flat (j,k) arrays below
  const int* ecol; const unsigned char* ekind; const double* ecv;   // flat over all groups
  const double* einc;            // flat [.][4] incident field per (element, j, k), or NULL
};
struct DevTables { const double *gl11_x, *gl11_w, *gl01_x, *gl01_w; };  // packed, rule n at offset n(n-1)/2

// per-context launch state of K1 (see assembly.cu)
struct K1Launch { int* counters = nullptr; cudaStream_t aux[2]; cudaEvent_t ev_fork, ev_join[2]; int n_sm = 148; };
int k1_launch_create(K1Launch& k);
void k1_launch_destroy(K1Launch& k);
void launch_classify(const DevGroup& g, const DevColloc& c, const DevClassify& k, unsigned char* plan, cudaStream_t st);
void launch_count_near(const unsigned char* plan, long long n_slots, const DevColloc& c, unsigned long long* counter, int2* list,
                       unsigned long long capacity, cudaStream_t st);
void launch_patch_plan(unsigned char* plan, const DevColloc& c, int n, const int* cpos, const int* slot, const unsigned char* val, cudaStream_t st);
void launch_gather_cv(const DevGroup& g, const double* cvalue, cudaStream_t st);
// kp: the region / frequency parameters as the reference defines them (K2, K3, the general K1); kq: the pre-scaled copy of the bulk K1 (kernel_scalars_scaled)
void launch_regular(const DevGroup& g, const DevColloc& c, const DevSystem& s, const unsigned char* plan, const void* tmap, bool statics, const KParams& kp, const KParams& kq,
                    K1Launch& k1, cudaStream_t st);
int make_matrix_tensor_map(void* out_128B, double* Are, long long lda, int n_dof, int box_planes);
// real plane only, host row/column order: out[:, 0:cols) = host columns [col0, col0 + cols)
void launch_gather_real(const double* re, long long ld, int rows, int cols, double* out, long long ldo, const int* rowperm, const int* colperm, int col0, cudaStream_t st);
void launch_scatter_real(const double* in, long long ldi, int rows, int cols, double* re, long long ld, const int* rowperm, cudaStream_t st);
void launch_adaptive(const DevGroup& g, const DevColloc& c, const DevSystem& s, const DevAdaptive& a, const DevTables& t, const KParams& kp, cudaStream_t st);
void launch_singular(const DevGroup& g, const DevColloc& c, const DevSystem& s, const DevSingular& a, const DevTables& t, const KParams& kp, cudaStream_t st);
void launch_freeterm(const DevColloc& c, const DevSystem& s, const DevFreeTerm& f, cplx F, cudaStream_t st);
void launch_residual(const DevSystem& s, const double* xre, const double* xim, double* rr, double* ri, double* ss, cudaStream_t st);
void launch_get_entries(const DevSystem& s, int n, const int* rows, const int* cols, double* out, cudaStream_t st);
void launch_interleave(const double* re, const double* im, long long ld, int rows, int cols, double* out, long long ldo, const int* rowperm,
                       const int* colperm, int col0, cudaStream_t st);
void launch_deinterleave(const double* in, long long ldi, int rows, int cols, double* re, double* im, long long ld, const int* rowperm, cudaStream_t st);

}  // namespace mfbd
