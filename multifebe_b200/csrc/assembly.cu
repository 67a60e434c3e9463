// assembly.cu -- sm_100a kernels of the harmonic 3D SBIE influence-matrix assembly.
//
//   K0 k_classify   ball test + rule thresholds per (collocation point, element) pair -> 1 plan byte
//                   (far branch of fbem_bem_harela3d_sbie_auto, lib/fbem/src/bem_harela3d.f90:1501-1506,1522-1531)
//   K1 k_regular    regular quadrature with precalculated point sets (fbem_bem_harela3d_sbie_ext_pre, :628-700)
//                   fused with the BC-aware scatter (src/assemble_bem_harela_equation.f90:78-113)
//   K2 k_adaptive   Telles + subdivision leaves (fbem_bem_harela3d_sbie_ext_st, :702-1048), one warp per pair
//   K3 k_singular   polar-transformation interior integration (fbem_bem_harela3d_sbie_int, :1174-1472), one warp per pair
//   K5 k_freeterm   free-term entries (src/build_lse_mechanics_bem_harela.f90:273-747)
//
// Mapping of K1: a warp owns 32 consecutive collocation points (sorted by matrix row, so that for a fixed matrix
// column the 32 lanes x 3 load directions address 96 consecutive rows of the column-major planar matrix) and walks
// a chunk of elements; the element's point set is read through warp-uniform loads, the 9*n complex accumulators of
// the pair live in registers, and the result goes to the matrix with coalesced RED.ADD.F64 (several elements share a
// node/column and MCA points share rows, so plain stores are not possible).  The right-hand side contribution of each
// lane is reduced in registers over the whole element chunk and flushed once.
#include "assembly.cuh"
#include <cuda.h>
#include <cstdio>

namespace mfbd {

// The region / frequency parameters travel as __grid_constant__ kernel arguments (constant bank, like __constant__ memory, but private to the launch):
// two problems assembling different frequencies on different streams of one process do not share them (round 1 kept them in __constant__ symbols,
// which serialised a process to one frequency in flight).

// ------------------------------------------------------------------------------------------------------------------
// K0: classification.  d must be bit-identical to the host/reference value (it feeds a discrete decision), hence the
// explicit round-to-nearest intrinsics (no FMA contraction): r = c - x_i; rmin = sqrt(r.r) - R; far iff rmin > 4R;
// d = rmin/cl; gln_near = first n with d >= far_thr[n].
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_classify(DevGroup g, DevColloc c, DevClassify k, unsigned char* __restrict__ plan) {
  int cpos = blockIdx.x * blockDim.x + threadIdx.x;
  int e = blockIdx.y;
  if (cpos >= c.ldp) return;
  unsigned char out = PLAN_NONE;
  if (c.crow[cpos] >= 0) {
    const double* b = g.ball + 5 * (size_t)e;
    double r0 = __dsub_rn(b[0], c.cx[cpos]), r1 = __dsub_rn(b[1], c.cx[c.ldp + cpos]), r2 = __dsub_rn(b[2], c.cx[2 * c.ldp + cpos]);
    double rr = __dadd_rn(__dadd_rn(__dmul_rn(r0, r0), __dmul_rn(r1, r1)), __dmul_rn(r2, r2));
    double rmin = __dsub_rn(__dsqrt_rn(rr), b[3]);
    if (rmin > __dmul_rn(4.0, b[3])) {
      double d = __ddiv_rn(rmin, b[4]);
      int gn = 31;
      if (d >= k.far_dmax) gn = 2;
      else {
#pragma unroll 1
        for (int n = 2; n <= 30; n++) if (d >= k.far_thr[n]) { gn = n; break; }
      }
      if (!(d > 2.0)) gn = 31;  // cannot happen for a ball that contains the element (R >= cl/2); be safe -> host
      int gln = max(g.gln_far[e], gn);
      out = PLAN_NEAR;
      if (gn <= 30 && gln <= k.ps_gln_max) {
        for (int s = 0; s < g.n_sets; s++) if (g.set_gln[s] >= gln) { out = (unsigned char)s; break; }
      }
    } else out = PLAN_NEAR;
  }
  plan[(size_t)(g.slot0 + e) * c.ldp + cpos] = out;
}
void launch_classify(const DevGroup& g, const DevColloc& c, const DevClassify& k, unsigned char* plan, cudaStream_t st) {
  if (g.n_elem == 0) return;
  dim3 grid((c.ldp + 255) / 256, g.n_elem);
  k_classify<<<grid, 256, 0, st>>>(g, c, k, plan);
}

__global__ void k_collect_near(const unsigned char* __restrict__ plan, long long n_slots, DevColloc c, unsigned long long* counter,
                               int2* list, unsigned long long capacity) {
  long long total = n_slots * c.ldp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (plan[i] == PLAN_NEAR) {
      unsigned long long p = atomicAdd(counter, 1ull);
      if (list && p < capacity) list[p] = make_int2((int)(i % c.ldp), (int)(i / c.ldp));
    }
  }
}
void launch_count_near(const unsigned char* plan, long long n_slots, const DevColloc& c, unsigned long long* counter, int2* list,
                       unsigned long long capacity, cudaStream_t st) {
  k_collect_near<<<1184, 256, 0, st>>>(plan, n_slots, c, counter, list, capacity);
}
__global__ void k_patch_plan(unsigned char* plan, DevColloc c, int n, const int* cpos, const int* slot, const unsigned char* val) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) plan[(size_t)slot[i] * c.ldp + cpos[i]] = val[i];
}
void launch_patch_plan(unsigned char* plan, const DevColloc& c, int n, const int* cpos, const int* slot, const unsigned char* val, cudaStream_t st) {
  if (n > 0) k_patch_plan<<<(n + 255) / 256, 256, 0, st>>>(plan, c, n, cpos, slot, val);
}

// prescribed values per (element, j, k) and the "some value is nonzero" flag, refreshed once per frequency
__global__ void k_gather_cv(DevGroup g, const double* __restrict__ cvalue) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.n_elem) return;
  bool nz = false;
  const int nd = g.ndof;
  for (int j = 0; j < g.nn; j++) {
    const int node = g.enode[e * g.nn + j];
    for (int k = 0; k < nd; k++) {
      double vr = cvalue[2 * (nd * (size_t)node + k)], vi = cvalue[2 * (nd * (size_t)node + k) + 1];
      if (g.c10 && g.c10[nd * (size_t)node + k]) {   // p known: t_k = p n_fn(k), with the sign of the orientation (assemble_bem_harela_equation.f90:97-106)
        const double f = g.nfn[nd * (size_t)node + k];   // already negated for the nodes of a reversed boundary (the ROOT element's orientation counts, not an image's)
        vr *= f; vi *= f;
      }
      if ((g.einfo[e] >> ((nd == 4 ? 4 : 5) + k)) & 1u) { vr = -vr; vi = -vi; }   // symmetry image: multiplier -1 on dof k (the b term carries it through the value); bits 5-7 (4-7 for a poroelastic node)
      const size_t i = (size_t)e * nd * g.nn + j * nd + k;
      g.ecv[2 * i] = vr; g.ecv[2 * i + 1] = vi;
      nz = nz || vr != 0.0 || vi != 0.0;
    }
  }
  g.ecvnz[e] = nz ? 1 : 0;
  const int mode = !(g.einfo[e] & 8u) ? 2 : (g.einc ? 3 : (nz ? 1 : 0));   // an incident field needs both kernel combinations: class 3 (uniform kinds) or 2
  atomicOr(g.range_modes + g.range_of[e], 1 << mode);
}
void launch_gather_cv(const DevGroup& g, const double* cvalue, cudaStream_t st) {
  if (g.n_elem > 0) {
    cudaMemsetAsync(g.range_modes, 0, sizeof(int) * g.n_ranges, st);
    k_gather_cv<<<(g.n_elem + 127) / 128, 128, 0, st>>>(g, cvalue);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Scatter of one pair's accumulators (assemble_bem_harela_equation.f90:78-113).  `mine(j*3+k)` selects which entries
// this lane writes (all for K1; a lane-strided subset after the warp reduction of K2/K3).
// ------------------------------------------------------------------------------------------------------------------
template <int NN, int NL, class Pred>
__device__ __forceinline__ void scatter_pair(const Acc<NN, NL>& a, const int* __restrict__ ecol, const unsigned char* __restrict__ ekind,
                                             const double* __restrict__ ecv, bool rev, const DevSystem& s, int il,
                                             int r0, int r1, int r2, double* bre, double* bim, Pred mine, const KParams& c_kp, bool hbie = false,
                                             unsigned symbits = 0u /* bit k: symconf_t(k) = -1 of a symmetry image (ecv already carries it) */,
                                             const double* __restrict__ einc = nullptr /* incident field of the element's (j,k), or NULL */,
                                             const int* __restrict__ ecol2 = nullptr /* columns of t_k for the dofs of kind 2, or NULL */) {
  // h (or m of the hypersingular equation) is scaled by cte_t (cte_s) and changes sign on a reversed element; g (l) by cte_u (cte_d)
  const cplx ch0 = hbie ? c_kp.cte_s : mk(c_kp.cte_t, 0.0);
  const cplx ch = rev ? mk(-ch0.re, -ch0.im) : ch0;
  const cplx cu = hbie ? mk(c_kp.cte_d, 0.0) : c_kp.cte_u;
#pragma unroll
  for (int j = 0; j < NN; j++) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int jk = j * 3 + k;
      if (!mine(jk)) continue;
      const int col = ecol[jk];
      const int kind = ekind[jk];
      const double cvr = ecv[2 * jk], cvi = ecv[2 * jk + 1];
      double* Ar = s.Are + (size_t)col * s.lda;
      double* Ai = s.Aim + (size_t)col * s.lda;
#pragma unroll
      for (int ll = 0; ll < NL; ll++) {
        const int l = (NL == 3) ? ll : il;
        const int row = (l == 0) ? r0 : (l == 1 ? r1 : r2);
        const int q = (ll * 3 + k) * NN + j;
        double hr = ch.re * a.hr[q] - ch.im * a.hi[q], hi = ch.re * a.hi[q] + ch.im * a.hr[q];
        double gr = cu.re * a.gr[q] - cu.im * a.gi[q], gi = cu.re * a.gi[q] + cu.im * a.gr[q];
        double ar, ai, br, bi;
        if (kind == 0) { ar = -gr; ai = -gi; br = -(hr * cvr - hi * cvi); bi = -(hr * cvi + hi * cvr); }
        else if (kind == 2) {   // u_k and t_k both unknown (local-axes conditions, assemble_bem_harela_equation.f90:107-112): h to the column of u_k, -g to that of t_k
          ar = hr; ai = hi; br = 0.0; bi = 0.0;
          const double sg = ((symbits >> k) & 1u) ? 1.0 : -1.0;
          double* A2r = s.Are + (size_t)ecol2[jk] * s.lda; double* A2i = s.Aim + (size_t)ecol2[jk] * s.lda;
          atomicAdd(A2r + row, sg * gr); atomicAdd(A2i + row, sg * gi);
        }
        else { ar = hr; ai = hi; br = gr * cvr - gi * cvi; bi = gr * cvi + gi * cvr; }
        if ((symbits >> k) & 1u) { ar = -ar; ai = -ai; }
        if (einc) {   // b += h u_inc - g t_inc (assemble_bem_harela_equation.f90:651-666); a symmetry image's sign is in the values
          const double ur = einc[4 * jk], ui = einc[4 * jk + 1], tr = einc[4 * jk + 2], ti = einc[4 * jk + 3];
          br += (hr * ur - hi * ui) - (gr * tr - gi * ti); bi += (hr * ui + hi * ur) - (gr * ti + gi * tr);
        }
        atomicAdd(Ar + row, ar);
        atomicAdd(Ai + row, ai);
        bre[l] += br; bim[l] += bi;
      }
    }
  }
}
struct AllEntries { __device__ __forceinline__ bool operator()(int) const { return true; } };
struct LaneEntries { int lane; __device__ __forceinline__ bool operator()(int jk) const { return (jk & 31) == lane; } };

// ------------------------------------------------------------------------------------------------------------------
// K1: regular pairs.  NL = 3: all load directions at once (9*NN complex accumulators, elements with <= 4 nodes);
// NL = 1: one load direction per pass (3*NN accumulators) for 6/8/9-node elements.
// ------------------------------------------------------------------------------------------------------------------
const int K1_WARPS = 4;
const int K1_ECHUNK = 32;

// HB: hypersingular equation (interior-point stresses): the collocation point carries a unit normal (DevColloc::cn)
template <int ET, int NL, bool HB>
__global__ void __launch_bounds__(K1_WARPS * 32) k_regular(DevGroup g, DevColloc c, DevSystem s, const unsigned char* __restrict__ plan, const __grid_constant__ KParams c_kp) {
  const int NN = ElemTraits<ET>::NN;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cpos = (blockIdx.x * K1_WARPS + warp) * 32 + lane;
  if (cpos - lane >= c.ldp) return;
  const int r0 = c.crow[cpos], r1 = c.crow[c.ldp + cpos], r2 = c.crow[2 * c.ldp + cpos];
  const bool valid = r0 >= 0 && (!c.tile_active || c.tile_active[cpos >> 5]);
  const double xc[3] = {c.cx[cpos], c.cx[c.ldp + cpos], c.cx[2 * c.ldp + cpos]};
  double ni[3] = {0.0, 0.0, 0.0};
  if (HB) { ni[0] = c.cn[cpos]; ni[1] = c.cn[c.ldp + cpos]; ni[2] = c.cn[2 * c.ldp + cpos]; }
  double bre[3] = {0.0, 0.0, 0.0}, bim[3] = {0.0, 0.0, 0.0};
  const int e0 = blockIdx.y * K1_ECHUNK, e1 = min(e0 + K1_ECHUNK, g.n_elem);
  for (int e = e0; e < e1; e++) {
    unsigned char m = valid ? plan[(size_t)(g.slot0 + e) * c.ldp + cpos] : PLAN_NONE;
    unsigned todo = __ballot_sync(0xffffffffu, m < MAX_SETS);
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const int sset = __shfl_sync(0xffffffffu, (int)m, leader);
      const unsigned grp = __ballot_sync(0xffffffffu, (int)m == sset);
      todo &= ~grp;
      if ((int)m == sset) {
        const int ngp = g.ngp[sset];
        const double* P = g.pts[sset] + (size_t)e * ngp * (6 + NN);
        const int* ecol = g.ecol + (size_t)e * 3 * NN;
        const unsigned char* ekind = g.ekind + (size_t)e * 3 * NN;
        const double* ecv = g.ecv + (size_t)e * 6 * NN;
        const bool rev = g.erev[e] != 0;
#pragma unroll 1
        for (int il = 0; il < (NL == 3 ? 1 : 3); il++) {
          Acc<NN, NL> acc; acc.zero();
#pragma unroll 1
          for (int kp = 0; kp < ngp; kp++) {
            const double* q = P + (size_t)kp * (6 + NN);
            double x[3] = {__ldg(q), __ldg(q + 1), __ldg(q + 2)}, n[3] = {__ldg(q + 3), __ldg(q + 4), __ldg(q + 5)}, w[NN];
#pragma unroll
            for (int j = 0; j < NN; j++) w[j] = __ldg(q + 6 + j);
            if (HB) accumulate_exterior_hbie<NN, NL>(acc, c_kp, x, n, xc, ni, w, il);
            else accumulate_exterior<NN, NL>(acc, c_kp, x, n, xc, w, il);
          }
          scatter_pair<NN, NL>(acc, ecol, ekind, ecv, rev, s, il, r0, r1, r2, bre, bim, AllEntries(), c_kp, HB, (unsigned)g.einfo[e] >> 5,
                               (!HB && g.einc) ? g.einc + (size_t)e * 12 * NN : nullptr, g.ecol2 ? g.ecol2 + (size_t)e * 3 * NN : nullptr);
        }
      }
    }
  }
  if (valid) {
    if (bre[0] != 0.0 || bim[0] != 0.0) { atomicAdd(s.bre + r0, bre[0]); atomicAdd(s.bim + r0, bim[0]); }
    if (bre[1] != 0.0 || bim[1] != 0.0) { atomicAdd(s.bre + r1, bre[1]); atomicAdd(s.bim + r1, bim[1]); }
    if (bre[2] != 0.0 || bim[2] != 0.0) { atomicAdd(s.bre + r2, bre[2]); atomicAdd(s.bim + r2, bim[2]); }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K1 (bulk): regular pairs of 3- and 4-node elements, the kernel that carries the assembly.
//   * a warp owns one collocation TILE (32 lanes = 32 points whose 96 matrix rows are consecutive, see DevColloc) and
//     walks a chunk of elements; the element's point set is read with warp-uniform loads;
//   * per Gauss point only what the boundary conditions of the element need is formed: the traction-kernel combination
//     for a t-known dof (its h goes to A), the displacement-kernel combination for a u-known dof (its g goes to A);
//     the other one only if the element carries a nonzero prescribed value (it then goes to b).  Kernel parameters are
//     pre-scaled on the host so that the accumulators are matrix entries;
//   * the pair's 9*NN complex entries are staged in shared memory as [column][plane][96 rows] and added to the planar
//     matrix by the TMA: one cp.reduce.async.bulk (.add.f64, SASS UBLKRED) per column and plane, 768 bytes each, issued
//     by 2*3*NN lanes.  The LSU never sees the matrix update (RED.F64 from registers costs 2.3 SM-cycles per lane and
//     entry, 5x the arithmetic of a far pair; profiles/r01_microbench_red.md).
// ------------------------------------------------------------------------------------------------------------------
const int KB_WARPS = 4;
#ifndef MFB_KB_WARPS_FAST
#define MFB_KB_WARPS_FAST 4
#endif
const int KB_WARPS_FAST = MFB_KB_WARPS_FAST;   // warps per CTA of the MODE 0 kernel (2 CTAs per SM)

#ifndef MFB_KB_CACHE_GP
#define MFB_KB_CACHE_GP 4
#endif
const int KB_CACHE_GP = MFB_KB_CACHE_GP;   // in-place batches of up to this many points keep their kernel scalars in shared memory between node chunks
// 3/4-node elements are accumulated whole; 6/8/9-node elements in NCH chunks of NW = 3 nodes (27 complex accumulators)
template <int NN_>
struct K1Shape {
  static const int NN = NN_, NW = (NN_ <= 4) ? NN_ : 3, NCH = (NN_ + NW - 1) / NW;
  static const int STAGE = 2 * 3 * NW * 96;                            // doubles per warp: staging of NW node boxes
  static const int CACHE = (NCH > 1) ? KB_CACHE_GP * 10 * 32 : 0;      // doubles per warp
  static const int SMEM_PER_WARP = (STAGE + CACHE) * 8 + MAX_SETS * 64 * 2 + MAX_SETS * 4;
};

template <int NN>
struct AccA { double re[9 * NN], im[9 * NN]; };   // [(l*3+k)*NN + j]: entry of load direction l (row) and dof k of node j (column)

// One Gauss point of one pair, accumulated for NW nodes of the element (all of them for 3/4-node elements, a chunk of
// three for 6/8/9-node elements).  rec = (x[3], n[3]) of the point, w[NW] = phi_j*J*weight of the chunk's nodes.
// MODE 0: element whose boundary-condition kinds are the same for all its nodes and whose prescribed values are all zero
//         (the common case): only the combination that goes to the matrix is formed.
// MODE 1: uniform kinds, some prescribed value nonzero: the other combination times S_k = sum_j w_j cv_jk (sk, over ALL nodes
//         of the element, formed by the caller) goes to b when do_b.
// MODE 3: MODE 1 with an incident field: besides S_k the sums of the incident values over the nodes (see the caller) multiply both combinations.
// MODE 2: any element: both combinations, per (node, dof) one goes to A and the other, times the prescribed value, to b;
//         ekind / ecv point at the chunk's first node.
// have_ks: the kernel scalars of this point are already in ks (cached by the pass over the first chunk of nodes).
// ST: static elasticity (Kelvin kernels, lib/fbem/src/bem_staela3d.f90:629-645): the kernel scalars are the 1/r and 1/r^2 terms
//     alone and everything is real -- no imaginary accumulators, no exponentials.
template <int NW, int MODE, bool ST>
__device__ __forceinline__ void k1_point(AccA<NW>& a, double* bacc, const double* rec, const double* w, const double* sk, bool do_b, const double* xc,
                                         double sgn, unsigned info, const unsigned char* __restrict__ ekind, const double* __restrict__ ecv, bool have_ks,
                                         KScal& k, const KParams& c_kq, const double* __restrict__ einc = nullptr /* MODE 2: incident field of the chunk's (j,k) */) {
  const double n[3] = {sgn * rec[3], sgn * rec[4], sgn * rec[5]};
  const double rv0 = rec[0] - xc[0], rv1 = rec[1] - xc[1], rv2 = rec[2] - xc[2];
  const double r2 = fma(rv0, rv0, fma(rv1, rv1, rv2 * rv2));
  const double d1r1 = rsqrt(r2), r = r2 * d1r1;
  if (ST) {
    const double d1r2 = d1r1 * d1r1;
    k.psi = mk(c_kq.psi[1].re * d1r1, 0.0); k.chi = mk(c_kq.chi[1].re * d1r1, 0.0);
    k.T1 = mk(c_kq.T1[1].re * d1r2, 0.0); k.T2 = mk(c_kq.T2[1].re * d1r2, 0.0); k.T3 = mk(c_kq.T3[1].re * d1r2, 0.0);
  } else if (!have_ks) kernel_scalars_scaled(c_kq, r, d1r1, k, MODE != 0 || (info & 7u) != 7u, MODE != 0 || (info & 7u) != 0u);
  const double dx[3] = {rv0 * d1r1, rv1 * d1r1, rv2 * d1r1};
  const double drdn = fma(dx[0], n[0], fma(dx[1], n[1], dx[2] * n[2]));
  const cplx t1d = k.T1 * drdn;
  if (MODE != 2) {
#pragma unroll
    for (int kk = 0; kk < 3; kk++) {
      const bool tk = (info >> kk) & 1u;
      const double skr = (MODE == 1 || MODE == 3) ? sk[kk] : 0.0, ski = (MODE == 1 || MODE == 3) ? sk[3 + kk] : 0.0;
      if (tk) {
#pragma unroll
        for (int l = 0; l < 3; l++) {
          const double dd = dx[l] * dx[kk], c2 = (l == kk) ? fma(dx[kk], n[l], drdn) : dx[kk] * n[l], c3 = dx[l] * n[kk];
          const double fr = fma(t1d.re, dd, fma(k.T2.re, c2, k.T3.re * c3)), fi = ST ? 0.0 : fma(t1d.im, dd, fma(k.T2.im, c2, k.T3.im * c3));
#pragma unroll
          for (int j = 0; j < NW; j++) { a.re[(l * 3 + kk) * NW + j] = fma(fr, w[j], a.re[(l * 3 + kk) * NW + j]); if (!ST) a.im[(l * 3 + kk) * NW + j] = fma(fi, w[j], a.im[(l * 3 + kk) * NW + j]); }
          if ((MODE == 1 || MODE == 3) && do_b) {
            const double orr = (l == kk) ? fma(-k.chi.re, dd, k.psi.re) : -k.chi.re * dd;
            if (ST) bacc[l] -= orr * skr;
            else {
              const double oi = (l == kk) ? fma(-k.chi.im, dd, k.psi.im) : -k.chi.im * dd;
              bacc[l] -= orr * skr - oi * ski; bacc[3 + l] -= orr * ski + oi * skr;
              if (MODE == 3) { bacc[l] += fr * sk[6 + kk] - fi * sk[9 + kk]; bacc[3 + l] += fr * sk[9 + kk] + fi * sk[6 + kk]; }   // incident field: the matrix combination times M_k
            }
          }
        }
      } else {
#pragma unroll
        for (int l = 0; l < 3; l++) {
          const double dd = dx[l] * dx[kk];
          const double fr = (l == kk) ? fma(-k.chi.re, dd, k.psi.re) : -k.chi.re * dd, fi = ST ? 0.0 : ((l == kk) ? fma(-k.chi.im, dd, k.psi.im) : -k.chi.im * dd);
#pragma unroll
          for (int j = 0; j < NW; j++) { a.re[(l * 3 + kk) * NW + j] = fma(fr, w[j], a.re[(l * 3 + kk) * NW + j]); if (!ST) a.im[(l * 3 + kk) * NW + j] = fma(fi, w[j], a.im[(l * 3 + kk) * NW + j]); }
          if ((MODE == 1 || MODE == 3) && do_b) {
            const double c2 = (l == kk) ? fma(dx[kk], n[l], drdn) : dx[kk] * n[l], c3 = dx[l] * n[kk];
            const double orr = fma(t1d.re, dd, fma(k.T2.re, c2, k.T3.re * c3));
            if (ST) bacc[l] -= orr * skr;
            else {
              const double oi = fma(t1d.im, dd, fma(k.T2.im, c2, k.T3.im * c3));
              bacc[l] -= orr * skr - oi * ski; bacc[3 + l] -= orr * ski + oi * skr;
              if (MODE == 3) { bacc[l] += fr * sk[6 + kk] - fi * sk[9 + kk]; bacc[3 + l] += fr * sk[9 + kk] + fi * sk[6 + kk]; }
            }
          }
        }
      }
    }
  } else {
#pragma unroll
    for (int kk = 0; kk < 3; kk++) {
      double ftr[3], fti[3], fur[3], fui[3];
#pragma unroll
      for (int l = 0; l < 3; l++) {
        const double dd = dx[l] * dx[kk], c2 = (l == kk) ? fma(dx[kk], n[l], drdn) : dx[kk] * n[l], c3 = dx[l] * n[kk];
        ftr[l] = fma(t1d.re, dd, fma(k.T2.re, c2, k.T3.re * c3)); fti[l] = ST ? 0.0 : fma(t1d.im, dd, fma(k.T2.im, c2, k.T3.im * c3));
        fur[l] = (l == kk) ? fma(-k.chi.re, dd, k.psi.re) : -k.chi.re * dd; fui[l] = ST ? 0.0 : ((l == kk) ? fma(-k.chi.im, dd, k.psi.im) : -k.chi.im * dd);
      }
#pragma unroll
      for (int j = 0; j < NW; j++) {
        if (w[j] == 0.0) continue;   // padding node of a short last chunk (its ekind/ecv are out of range)
        const bool tk = ekind[j * 3 + kk] != 0;
        const double cvr = w[j] * __ldg(ecv + 2 * (j * 3 + kk)), cvi = w[j] * __ldg(ecv + 2 * (j * 3 + kk) + 1);
#pragma unroll
        for (int l = 0; l < 3; l++) {
          const double ar = tk ? ftr[l] : fur[l], orr = tk ? fur[l] : ftr[l];
          a.re[(l * 3 + kk) * NW + j] = fma(ar, w[j], a.re[(l * 3 + kk) * NW + j]);
          if (ST) bacc[l] -= orr * cvr;
          else {
            const double ai = tk ? fti[l] : fui[l], oi = tk ? fui[l] : fti[l];
            a.im[(l * 3 + kk) * NW + j] = fma(ai, w[j], a.im[(l * 3 + kk) * NW + j]);
            bacc[l] -= orr * cvr - oi * cvi; bacc[3 + l] -= orr * cvi + oi * cvr;
          }
        }
        if (!ST && einc) {   // b += h u_inc - g t_inc: ft is h, fu is -g in this kernel's scaling (assemble_bem_harela_equation.f90:651-666)
          const double ur = w[j] * __ldg(einc + 4 * (j * 3 + kk)), ui = w[j] * __ldg(einc + 4 * (j * 3 + kk) + 1);
          const double tr = w[j] * __ldg(einc + 4 * (j * 3 + kk) + 2), ti = w[j] * __ldg(einc + 4 * (j * 3 + kk) + 3);
#pragma unroll
          for (int l = 0; l < 3; l++) {
            bacc[l] += (ftr[l] * ur - fti[l] * ui) + (fur[l] * tr - fui[l] * ti);
            bacc[3 + l] += (ftr[l] * ui + fti[l] * ur) + (fur[l] * ti + fui[l] * tr);
          }
        }
      }
    }
  }
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// volatile so that the request for point kp+1 stays ahead of the arithmetic of point kp (the compiler otherwise sinks it)
__device__ __forceinline__ double ldg_ahead(const double* p) { double v; asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }

const int KB_QCAP = 64;        // entries per deferred queue (at most 31 waiting + 32 new)
const int KB_INPLACE_MIN = 10; // fewer lanes than this with the in-place set: defer them too
const int KB_SMEM_QUEUE = MAX_SETS * KB_QCAP * 2;   // bytes per warp

// Work of one warp: a stream of tasks (one collocation tile x one range of elements) fetched from a global counter.
// Pairs whose plan asks for the lowest rule (set 0, ~90 % of all pairs) are integrated IN PLACE, lane = collocation point
// of the tile, all lanes on the same element, and flushed by bulk reduce.  Pairs with any other rule are DEFERRED: pushed
// to a per-set queue in shared memory and integrated 32 at a time, one (point, element) pair per lane, flushed with
// per-lane RED, as soon as a queue holds a full warp: the lane occupancy of the high-order rules goes from ~25 % to
// ~100 % (a tile sees several rules for the elements around the switch distances of the rule estimator).  Both kinds of
// batch run through the same code (one copy of the point arithmetic: the instruction cache matters here).
// One instantiation per element class (MODE 0/1/2 of k1_point); an element is visited by the kernel of its class only.
// ST = static elasticity: real arithmetic, only the real plane of the matrix is updated (tensor map with a one-plane box).
template <int ET, int MODE, int WARPS, bool ST>
__global__ void __launch_bounds__(WARPS * 32, 2) k_regular_bulk(const __grid_constant__ CUtensorMap tmap, DevGroup g, DevColloc c, DevSystem s,
                                                                  const unsigned char* __restrict__ plan,
                                                                  int* __restrict__ task_counter, const __grid_constant__ KParams c_kq) {
  typedef K1Shape<ElemTraits<ET>::NN> SH;
  constexpr int NN = SH::NN, NC = 3 * NN, NW = SH::NW, NCH = SH::NCH, RECN = 6 + NN;
  extern __shared__ __align__(128) double k1_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr bool GEN = MODE != 0;
  double* buf = k1_smem + (size_t)warp * SH::STAGE;
  double* kcache = k1_smem + (size_t)WARPS * SH::STAGE + (size_t)warp * SH::CACHE;   // [point][10 scalars][32 lanes], chunked elements only
  unsigned short* queue = reinterpret_cast<unsigned short*>(k1_smem + (size_t)WARPS * (SH::STAGE + SH::CACHE)) + (size_t)warp * (MAX_SETS * KB_QCAP);
  int* qcnt = reinterpret_cast<int*>(reinterpret_cast<unsigned short*>(k1_smem + (size_t)WARPS * (SH::STAGE + SH::CACHE)) + (size_t)WARPS * (MAX_SETS * KB_QCAP)) + warp * MAX_SETS;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int n_tasks = c.n_tiles * g.n_ranges;
  bool pending = false;
  for (;;) {
    int task = 0;
    if (lane == 0) task = atomicAdd(task_counter, 1);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= n_tasks) break;
    const int range = task / c.n_tiles, tile = task - range * c.n_tiles;   // consecutive tasks: same elements, consecutive tiles
    if (!((g.range_modes[range] >> MODE) & 1)) continue;                   // no element of this kernel's class in the range
    if (c.tile_active && !c.tile_active[tile]) continue;                   // row block of another rank
    if (lane < MAX_SETS) qcnt[lane] = 0;
    __syncwarp();
    const int cpos = tile * 32 + lane;
    const int r0 = c.crow[cpos];
    const bool valid = r0 >= 0;
    const double xc[3] = {c.cx[cpos], c.cx[c.ldp + cpos], c.cx[2 * c.ldp + cpos]};
    const int row0 = c.tile_row0[tile], nbytes = c.tile_nbytes[tile];
    double bacc_t[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const int e0 = g.range_start[range], e1 = g.range_start[range + 1];
    const unsigned char* pl = plan + (size_t)g.slot0 * c.ldp + cpos;
    int e = e0, ecur = 0, drain = 0;
    unsigned char m = PLAN_NONE, m_next = (valid && e0 < e1) ? pl[(size_t)e0 * c.ldp] : PLAN_NONE;
    unsigned info_next = g.einfo[e0], cvnz_next = g.ecvnz[e0];   // element class of the next element, requested one element ahead like its plan byte
    unsigned fullsets = 0u;      // sets whose queue holds >= 32 entries
    unsigned inplace_sets = 0u;   // rules of the current element that enough lanes of the tile share to be integrated in place (bit = set index)
    for (;;) {
      // ---- select the next batch (all control flow here is warp-uniform) ----
      int sset, el, src; bool act, inplace;
      if (fullsets) {
        sset = __ffs(fullsets) - 1;
        int cnt = qcnt[sset];
        const unsigned short ent = queue[sset * KB_QCAP + cnt - 32 + lane];
        __syncwarp();
        cnt -= 32;
        if (lane == 0) qcnt[sset] = cnt;
        __syncwarp();
        if (cnt < 32) fullsets &= ~(1u << sset);
        src = ent & 31; el = e0 + (ent >> 5); act = true; inplace = false;
      } else if (inplace_sets) {
        sset = __ffs(inplace_sets) - 1; inplace_sets &= inplace_sets - 1u;
        el = ecur; src = lane; act = ((int)m == sset); inplace = true;
      } else if (e < e1) {
        m = m_next; ecur = e; e++;
        m_next = (valid && e < e1) ? pl[(size_t)e * c.ldp] : PLAN_NONE;
        const unsigned info_e = info_next, cvnz_e = cvnz_next;
        if (e < e1) { info_next = g.einfo[e]; cvnz_next = g.ecvnz[e]; }
        const int mode_e = !(info_e & 8u) ? 2 : (g.einc ? 3 : (cvnz_e != 0 ? 1 : 0));
        if (mode_e != MODE) continue;
        const unsigned reg = __ballot_sync(0xffffffffu, m < MAX_SETS);
        if (reg == 0u) continue;
        // Every rule that at least KB_INPLACE_MIN lanes of the tile ask for is integrated IN PLACE (lane = collocation point, bulk-reduce flush): a
        // tile is spatially compact, so near an element most of its 32 points want the same rule.  Round 1 did this for the lowest rule only and
        // sent every other rule through the per-lane RED flush of the deferred queues: on a quad9 m = 20 mesh, where more than half of all pairs sit
        // inside the gln >= 3 distance, that flush was the kernel (5.4 TFLOP/s algorithmic, 37.1 ms in round 1, profiles/r02_one_step_quad9_after_inplace.log after).
        unsigned todo = reg;
        inplace_sets = 0u;
        if (lane < 3) {   // pull the next element's first point set towards L1 while this one is integrated
          const double* Pn = g.pts[0] + (size_t)(ecur + 1) * g.ngp[0] * RECN;
          if (ecur + 1 < e1) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(Pn) + 128 * lane));
        }
        while (todo) {
          const int leader = __ffs(todo) - 1;
          const int qs = __shfl_sync(0xffffffffu, (int)m, leader);
          const unsigned grp = __ballot_sync(0xffffffffu, (int)m == qs) & todo;
          todo &= ~grp;
          // 3/4-node elements: only the lowest rule (their 27 / 36 entries per pair flush cheaply from a packed deferred batch, and a packed batch keeps
          // all 32 lanes busy: with every rule in place K1 of the tri3 m = 40 mesh went from 47 to 53 ms)
          if (__popc(grp) >= KB_INPLACE_MIN && nbytes > 0 && (NCH > 1 || qs == 0)) { inplace_sets |= 1u << qs; continue; }
          const int base = qcnt[qs];
          if ((grp >> lane) & 1u) queue[qs * KB_QCAP + base + __popc(grp & lt_mask)] = (unsigned short)(((ecur - e0) << 5) | lane);
          __syncwarp();
          const int cnt = base + __popc(grp);
          if (lane == 0) qcnt[qs] = cnt;
          if (cnt >= 32) fullsets |= 1u << qs;
          __syncwarp();
        }
        continue;
      } else if (drain < g.n_sets) {
        sset = drain++;
        const int cnt = qcnt[sset];
        if (cnt == 0) continue;
        act = lane < cnt;
        const unsigned short ent = act ? queue[sset * KB_QCAP + lane] : (unsigned short)0;
        src = ent & 31; el = e0 + (ent >> 5); inplace = false;
      } else break;

      // ---- integrate the batch: one pair per lane; 6/8/9-node elements in chunks of three nodes ----
      const double xs[3] = {__shfl_sync(0xffffffffu, xc[0], src), __shfl_sync(0xffffffffu, xc[1], src), __shfl_sync(0xffffffffu, xc[2], src)};
      const int rs = __shfl_sync(0xffffffffu, r0, src);
      const int* ecol = g.ecol + (size_t)el * NC;
      const int ngp = g.ngp[sset];
      // the kernel scalars of the first chunk's pass are kept in shared memory for the other chunks (in-place batches of few points)
      const bool use_cache = !ST && (NCH > 1) && inplace && ngp <= KB_CACHE_GP;
#pragma unroll 1
      for (int ch = 0; ch < NCH; ch++) {
        const int j0 = ch * NW;
        int mycol = 0;
        if (inplace && lane < NW && j0 + lane < NN) mycol = __ldg(ecol + 3 * (j0 + lane));   // first column of the chunk's nodes, needed by the flush only
        AccA<NW> acc;
#pragma unroll
        for (int i = 0; i < 9 * NW; i++) { acc.re[i] = 0.0; acc.im[i] = 0.0; }
        double bacc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (act) {
          const unsigned info = g.einfo[el];
          const double sgn = (info & 16u) ? -1.0 : 1.0;
          const double* P = g.pts[sset] + (size_t)el * ngp * RECN;
          const unsigned char* ekind = g.ekind + (size_t)el * NC + 3 * j0;
          const double* ecv = g.ecv + (size_t)el * 2 * NC + 6 * j0;
          // the record of point kp+1 is requested before point kp is integrated
          double rec[6 + NW];
#pragma unroll
          for (int i = 0; i < 6; i++) rec[i] = ldg_ahead(P + i);
#pragma unroll
          for (int i = 0; i < NW; i++) rec[6 + i] = (j0 + i < NN) ? ldg_ahead(P + 6 + j0 + i) : 0.0;
#pragma unroll 1
          for (int kp = 0; kp < ngp; kp++) {
            double cur[6 + NW];
#pragma unroll
            for (int i = 0; i < 6 + NW; i++) cur[i] = rec[i];
            if (kp + 1 < ngp) {
              const double* Pn = P + (size_t)(kp + 1) * RECN;
#pragma unroll
              for (int i = 0; i < 6; i++) rec[i] = ldg_ahead(Pn + i);
#pragma unroll
              for (int i = 0; i < NW; i++) rec[6 + i] = (j0 + i < NN) ? ldg_ahead(Pn + 6 + j0 + i) : 0.0;
            }
            // sk[0..5]: O_k (re | im), the multiplier of the combination that does NOT go to the matrix (b -= other O_k); sk[6..11]: M_k, the multiplier of the
            // one that does (b += main M_k).  Without an incident field O_k = S_k = sum_j w_j cv_jk, M_k = 0.  With one (assemble_bem_harela_equation.f90:651-666:
            // b += h u_inc - g t_inc; in this kernel's scaling ft = h, fu = -g): a t-known dof has main = ft, other = fu: O_k = S_k - ST_k, M_k = SU_k; a u-known dof
            // has main = fu, other = ft: O_k = S_k - SU_k, M_k = ST_k, with SU_k = sum_j w_j u_inc_jk, ST_k = sum_j w_j t_inc_jk.
            double sk[12] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            if ((MODE == 1 || MODE == 3) && ch == 0) {   // sums over all the nodes of the element
              const double* wall = P + (size_t)kp * RECN + 6;
              const double* ecv0 = g.ecv + (size_t)el * 2 * NC;
              const double* inc0 = (MODE == 3 && !ST && g.einc) ? g.einc + (size_t)el * 4 * NC : nullptr;   // class 3 = class 1 + incident field (its own instantiation: class 1 keeps its registers)
#pragma unroll
              for (int kk = 0; kk < 3; kk++) {
                double sur = 0.0, sui = 0.0, str_ = 0.0, sti = 0.0;
#pragma unroll
                for (int j = 0; j < NN; j++) {
                  const double wj = (NCH == 1) ? cur[6 + j < 6 + NW ? 6 + j : 6] : __ldg(wall + j);
                  sk[kk] = fma(wj, __ldg(ecv0 + 2 * (j * 3 + kk)), sk[kk]);
                  if (!ST) sk[3 + kk] = fma(wj, __ldg(ecv0 + 2 * (j * 3 + kk) + 1), sk[3 + kk]);
                  if (inc0) {
                    const double* q4 = inc0 + 4 * (j * 3 + kk);
                    sur = fma(wj, __ldg(q4), sur); sui = fma(wj, __ldg(q4 + 1), sui); str_ = fma(wj, __ldg(q4 + 2), str_); sti = fma(wj, __ldg(q4 + 3), sti);
                  }
                }
                if (inc0) {
                  if ((info >> kk) & 1u) { sk[kk] -= str_; sk[3 + kk] -= sti; sk[6 + kk] = sur; sk[9 + kk] = sui; }
                  else { sk[kk] -= sur; sk[3 + kk] -= sui; sk[6 + kk] = str_; sk[9 + kk] = sti; }
                }
              }
            }
            KScal ks;
            const bool have_ks = use_cache && ch > 0;
            if (have_ks) {
              const double* kc = kcache + (size_t)kp * 320 + lane;
              ks.psi = mk(kc[0], kc[32]); ks.chi = mk(kc[64], kc[96]); ks.T1 = mk(kc[128], kc[160]); ks.T2 = mk(kc[192], kc[224]); ks.T3 = mk(kc[256], kc[288]);
            }
            k1_point<NW, MODE, ST>(acc, bacc, cur, cur + 6, sk, ch == 0, xs, sgn, info, ekind, ecv, have_ks, ks, c_kq,
                                   (MODE == 2 && g.einc) ? g.einc + (size_t)el * 4 * NC + 12 * j0 : nullptr);
            if (use_cache && ch == 0) {
              double* kc = kcache + (size_t)kp * 320 + lane;
              kc[0] = ks.psi.re; kc[32] = ks.psi.im; kc[64] = ks.chi.re; kc[96] = ks.chi.im; kc[128] = ks.T1.re; kc[160] = ks.T1.im;
              kc[192] = ks.T2.re; kc[224] = ks.T2.im; kc[256] = ks.T3.re; kc[288] = ks.T3.im;
            }
          }
          if (info & 0xE0u) {   // symmetry image (build_lse_mechanics_bem_harela.f90:1203-1206): h(:,:,k), g(:,:,k) times symconf_t(k); b took it through ecv
#pragma unroll
            for (int k = 0; k < 3; k++)
              if ((info >> (5 + k)) & 1u) {
#pragma unroll
                for (int l = 0; l < 3; l++)
#pragma unroll
                  for (int j = 0; j < NW; j++) { acc.re[(l * 3 + k) * NW + j] = -acc.re[(l * 3 + k) * NW + j]; if (!ST) acc.im[(l * 3 + k) * NW + j] = -acc.im[(l * 3 + k) * NW + j]; }
              }
          }
        }
        // ---- flush of the chunk ----
        if (inplace && nbytes > 0) {
          if (pending && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous flush has left the buffer
          __syncwarp();
          // staging layout = the TMA box of one node: [node j][plane][dof k][96 rows] (static: one plane)
#pragma unroll
          for (int j = 0; j < NW; j++)
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
              for (int l = 0; l < 3; l++) {
                if (ST) buf[(j * 3 + k) * 96 + 3 * lane + l] = acc.re[(l * 3 + k) * NW + j];
                else {
                  buf[((j * 2 + 0) * 3 + k) * 96 + 3 * lane + l] = acc.re[(l * 3 + k) * NW + j];
                  buf[((j * 2 + 1) * 3 + k) * 96 + 3 * lane + l] = acc.im[(l * 3 + k) * NW + j];
                }
              }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int j = 0; j < NW; j++) {
            const int col = __shfl_sync(0xffffffffu, mycol, j);
            if (lane == 0 && j0 + j < NN)
              asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&tmap), "r"(row0), "r"(col), "r"(0),
                           "r"(smem_u32(buf + j * (ST ? 288 : 576)))
                           : "memory");
          }
          if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          pending = true;
          if (GEN) {
#pragma unroll
            for (int i = 0; i < 6; i++) bacc_t[i] += bacc[i];
          }
        } else if (act) {
#pragma unroll
          for (int j = 0; j < NW; j++) {
            if (j0 + j >= NN) continue;
#pragma unroll
            for (int k = 0; k < 3; k++) {
              const int col = __ldg(ecol + (j0 + j) * 3 + k);
              double* Ar = s.Are + (size_t)col * s.lda + rs; double* Ai = s.Aim + (size_t)col * s.lda + rs;
#pragma unroll
              for (int l = 0; l < 3; l++) { atomicAdd(Ar + l, acc.re[(l * 3 + k) * NW + j]); if (!ST) atomicAdd(Ai + l, acc.im[(l * 3 + k) * NW + j]); }
            }
          }
          if (GEN) {
#pragma unroll
            for (int l = 0; l < 3; l++)
              if (bacc[l] != 0.0 || bacc[3 + l] != 0.0) { atomicAdd(s.bre + rs + l, bacc[l]); if (!ST) atomicAdd(s.bim + rs + l, bacc[3 + l]); }
          }
        }
      }
    }
    if (GEN && valid) {
#pragma unroll
      for (int l = 0; l < 3; l++)
        if (bacc_t[l] != 0.0 || bacc_t[3 + l] != 0.0) { atomicAdd(s.bre + r0 + l, bacc_t[l]); if (!ST) atomicAdd(s.bim + r0 + l, bacc_t[3 + l]); }
    }
  }
  if (pending && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Per-context launch state of K1 (task counters of the persistent kernels, the two side streams on which the element classes run beside each
// other, fork / join events): owned by the mfb_ctx, so that contexts on different streams (several frequencies in flight in one process) never
// share it.  Round 1 kept it in function-local statics.
int k1_launch_create(K1Launch& k) {
  if (cudaMalloc((void**)&k.counters, 4 * sizeof(int)) != cudaSuccess) return 1;
  for (int i = 0; i < 2; i++) { cudaStreamCreateWithFlags(&k.aux[i], cudaStreamNonBlocking); cudaEventCreateWithFlags(&k.ev_join[i], cudaEventDisableTiming); }
  cudaEventCreateWithFlags(&k.ev_fork, cudaEventDisableTiming);
  int dev = 0; k.n_sm = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&k.n_sm, cudaDevAttrMultiProcessorCount, dev);
  return (int)cudaGetLastError();
}
void k1_launch_destroy(K1Launch& k) {
  if (!k.counters) return;
  cudaFree(k.counters); k.counters = nullptr;
  for (int i = 0; i < 2; i++) { cudaStreamDestroy(k.aux[i]); cudaEventDestroy(k.ev_join[i]); }
  cudaEventDestroy(k.ev_fork);
}
template <int ET, int MODE, int WARPS, bool ST>
static void launch_bulk_mode(const CUtensorMap& tmap, const DevGroup& g, const DevColloc& c, const DevSystem& s, const unsigned char* plan, const KParams& kq, K1Launch& k1,
                             cudaStream_t st) {
  const int smem = WARPS * K1Shape<ElemTraits<ET>::NN>::SMEM_PER_WARP;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_regular_bulk<ET, MODE, WARPS, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  const int n_tasks = c.n_tiles * g.n_ranges;
  int ctas = 2 * k1.n_sm; if (ctas * WARPS > n_tasks) ctas = (n_tasks + WARPS - 1) / WARPS;
  k_regular_bulk<ET, MODE, WARPS, ST><<<ctas, WARPS * 32, smem, st>>>(tmap, g, c, s, plan, k1.counters + MODE, kq);
}
template <int ET, bool ST>
static void launch_regular_bulk(const CUtensorMap& tmap, const DevGroup& g, const DevColloc& c, const DevSystem& s, const unsigned char* plan, const KParams& kq, K1Launch& k1,
                                cudaStream_t st) {
  cudaMemsetAsync(k1.counters, 0, 4 * sizeof(int), st);
  if constexpr (!ST) {
    if (g.einc) {   // incident field set: elements with uniform kinds are class 3, the others class 2
      launch_bulk_mode<ET, 3, KB_WARPS, ST>(tmap, g, c, s, plan, kq, k1, st);
      if (g.has_mixed) launch_bulk_mode<ET, 2, KB_WARPS, ST>(tmap, g, c, s, plan, kq, k1, st);
      return;
    }
  }
  if (c.n_tiles * g.n_ranges < 2048) {   // a small mesh: the classes one after the other on the caller's stream (no side streams: they only pay on long kernels,
    launch_bulk_mode<ET, 1, KB_WARPS, ST>(tmap, g, c, s, plan, kq, k1, st);        // and a process that keeps many small problems in flight runs out of hardware queues)
    if (g.has_mixed) launch_bulk_mode<ET, 2, KB_WARPS, ST>(tmap, g, c, s, plan, kq, k1, st);
    launch_bulk_mode<ET, 0, KB_WARPS_FAST, ST>(tmap, g, c, s, plan, kq, k1, st);
    return;
  }
  // the three element classes run concurrently (their tails overlap); everything joins the caller's stream again
  cudaEventRecord(k1.ev_fork, st);
  cudaStreamWaitEvent(k1.aux[0], k1.ev_fork, 0);
  launch_bulk_mode<ET, 1, KB_WARPS, ST>(tmap, g, c, s, plan, kq, k1, k1.aux[0]);
  cudaEventRecord(k1.ev_join[0], k1.aux[0]);
  if (g.has_mixed) {
    cudaStreamWaitEvent(k1.aux[1], k1.ev_fork, 0);
    launch_bulk_mode<ET, 2, KB_WARPS, ST>(tmap, g, c, s, plan, kq, k1, k1.aux[1]);
    cudaEventRecord(k1.ev_join[1], k1.aux[1]);
  }
  launch_bulk_mode<ET, 0, KB_WARPS_FAST, ST>(tmap, g, c, s, plan, kq, k1, st);
  cudaStreamWaitEvent(st, k1.ev_join[0], 0);
  if (g.has_mixed) cudaStreamWaitEvent(st, k1.ev_join[1], 0);
}

// 3-D tensor map of the planar system matrix for the K1 flush: (row, column, plane), box = 96 rows x 3 columns x 2 planes
// (the three dofs of one node, both planes, for the 32 collocation points of a tile), FLOAT64, no swizzle.
int make_matrix_tensor_map(void* out_128B, double* Are, long long lda, int n_dof, int box_planes) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return 1;
    encode = (EncodeFn)fn;
  }
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  cuuint64_t dims[3] = {(cuuint64_t)n_dof, (cuuint64_t)n_dof, 2};
  cuuint64_t strides[2] = {(cuuint64_t)lda * 8, (cuuint64_t)lda * (cuuint64_t)n_dof * 8};
  cuuint32_t box[3] = {96, 3, (cuuint32_t)box_planes}, estr[3] = {1, 1, 1};
  CUresult r = encode(reinterpret_cast<CUtensorMap*>(out_128B), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, Are, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

template <int ET>
static void launch_regular_et(const DevGroup& g, const DevColloc& c, const DevSystem& s, const unsigned char* plan, const void* tmap, bool statics, const KParams& kq, K1Launch& k1,
                              cudaStream_t st) {
  if (statics) launch_regular_bulk<ET, true>(*reinterpret_cast<const CUtensorMap*>(tmap), g, c, s, plan, kq, k1, st);
  else launch_regular_bulk<ET, false>(*reinterpret_cast<const CUtensorMap*>(tmap), g, c, s, plan, kq, k1, st);
}
// tmap: tensor map of the planar matrix whose box has two planes (harmonic) or one (statics = true); without it, or for
// elements whose dof columns are not consecutive, the general kernel runs (in a static run it works with the harmonic
// parameter set whose frequency-dependent coefficients are zero, see set_kparams).
void launch_regular(const DevGroup& g, const DevColloc& c, const DevSystem& s, const unsigned char* plan, const void* tmap, bool statics, const KParams& kp, const KParams& kq,
                    K1Launch& k1, cudaStream_t st) {
  if (g.n_elem == 0) return;
  dim3 grid((c.ldp + 32 * K1_WARPS - 1) / (32 * K1_WARPS), (g.n_elem + K1_ECHUNK - 1) / K1_ECHUNK);
  dim3 block(K1_WARPS * 32);
  if (c.cn) {   // hypersingular equation: the general kernel (a handful of interior points, not a hot path)
    switch (g.et) {
      case 5: k_regular<5, 3, true><<<grid, block, 0, st>>>(g, c, s, plan, kp); break;
      case 7: k_regular<7, 3, true><<<grid, block, 0, st>>>(g, c, s, plan, kp); break;
      case 6: k_regular<6, 1, true><<<grid, block, 0, st>>>(g, c, s, plan, kp); break;
      case 8: k_regular<8, 1, true><<<grid, block, 0, st>>>(g, c, s, plan, kp); break;
      case 9: k_regular<9, 1, true><<<grid, block, 0, st>>>(g, c, s, plan, kp); break;
    }
    return;
  }
  switch (g.et) {
    case 5: if (g.cols3 && tmap) { launch_regular_et<5>(g, c, s, plan, tmap, statics, kq, k1, st); break; }
            k_regular<5, 3, false><<<grid, block, 0, st>>>(g, c, s, plan, kp); break;
    case 7: if (g.cols3 && tmap) { launch_regular_et<7>(g, c, s, plan, tmap, statics, kq, k1, st); break; }
            k_regular<7, 3, false><<<grid, block, 0, st>>>(g, c, s, plan, kp); break;
    case 6: if (g.cols3 && tmap) { launch_regular_et<6>(g, c, s, plan, tmap, statics, kq, k1, st); break; }
            k_regular<6, 1, false><<<grid, block, 0, st>>>(g, c, s, plan, kp); break;
    case 8: if (g.cols3 && tmap) { launch_regular_et<8>(g, c, s, plan, tmap, statics, kq, k1, st); break; }
            k_regular<8, 1, false><<<grid, block, 0, st>>>(g, c, s, plan, kp); break;
    case 9: if (g.cols3 && tmap) { launch_regular_et<9>(g, c, s, plan, tmap, statics, kq, k1, st); break; }
            k_regular<9, 1, false><<<grid, block, 0, st>>>(g, c, s, plan, kp); break;
  }
}

// warp all-reduce of every accumulator
template <int NN, int NL>
__device__ __forceinline__ void warp_reduce(Acc<NN, NL>& a) {
#pragma unroll
  for (int i = 0; i < NL * 3 * NN; i++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a.hr[i] += __shfl_xor_sync(0xffffffffu, a.hr[i], o); a.hi[i] += __shfl_xor_sync(0xffffffffu, a.hi[i], o);
      a.gr[i] += __shfl_xor_sync(0xffffffffu, a.gr[i], o); a.gi[i] += __shfl_xor_sync(0xffffffffu, a.gi[i], o);
    }
  }
}
__device__ __forceinline__ void flush_b(const DevSystem& s, int r0, int r1, int r2, const double* bre, const double* bim) {
  if (bre[0] != 0.0 || bim[0] != 0.0) { atomicAdd(s.bre + r0, bre[0]); atomicAdd(s.bim + r0, bim[0]); }
  if (bre[1] != 0.0 || bim[1] != 0.0) { atomicAdd(s.bre + r1, bre[1]); atomicAdd(s.bim + r1, bim[1]); }
  if (bre[2] != 0.0 || bim[2] != 0.0) { atomicAdd(s.bre + r2, bre[2]); atomicAdd(s.bim + r2, bim[2]); }
}

// ------------------------------------------------------------------------------------------------------------------
// K2: adaptive pairs -- one warp per pair, lanes stride over the gln x gln points of every leaf.
// ------------------------------------------------------------------------------------------------------------------
template <int ET, int NL, bool HB>
__global__ void __launch_bounds__(128) k_adaptive(DevGroup g, DevColloc c, DevSystem s, DevAdaptive a, DevTables t, const __grid_constant__ KParams c_kp) {
  const int NN = ElemTraits<ET>::NN;
  const bool tri = (ElemTraits<ET>::NV == 3);
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= a.n_pairs) return;
  const int cpos = a.pair_cpos[p], e = a.pair_elem[p];
  if (c.tile_active && !c.tile_active[cpos >> 5]) return;
  const double xc[3] = {c.cx[cpos], c.cx[c.ldp + cpos], c.cx[2 * c.ldp + cpos]};
  const int r0 = c.crow[cpos], r1 = c.crow[c.ldp + cpos], r2 = c.crow[2 * c.ldp + cpos];
  double ni[3] = {0.0, 0.0, 0.0};
  if (HB) { ni[0] = c.cn[cpos]; ni[1] = c.cn[c.ldp + cpos]; ni[2] = c.cn[2 * c.ldp + cpos]; }
  double xn[3 * NN];
#pragma unroll
  for (int i = 0; i < 3 * NN; i++) xn[i] = g.xn[(size_t)e * 3 * NN + i];
  const double* gx = tri ? t.gl01_x : t.gl11_x;
  const double* gw = tri ? t.gl01_w : t.gl11_w;
  double bre[3] = {0.0, 0.0, 0.0}, bim[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
  for (int il = 0; il < (NL == 3 ? 1 : 3); il++) {
    Acc<NN, NL> acc; acc.zero();
#pragma unroll 1
    for (int lf = a.pair_leaf0[p]; lf < a.pair_leaf0[p + 1]; lf++) {
      const double* L = a.leaf_d + 16 * (size_t)lf;
      double xi_s[8], tp1[4], tp2[4];
#pragma unroll
      for (int i = 0; i < 8; i++) xi_s[i] = __ldg(L + i);
#pragma unroll
      for (int i = 0; i < 4; i++) { tp1[i] = __ldg(L + 8 + i); tp2[i] = __ldg(L + 12 + i); }
      const int gln = a.leaf_gln[lf], off = gln * (gln - 1) / 2;
#pragma unroll 1
      for (int idx = lane; idx < gln * gln; idx += 32) {
        const int k1 = idx / gln, k2 = idx - k1 * gln;
        double x[3], n[3], w[NN];
        leaf_point<ET>(xn, xi_s, tp1, tp2, __ldg(gx + off + k1), __ldg(gw + off + k1), __ldg(gx + off + k2), __ldg(gw + off + k2), x, n, w);
        if (HB) accumulate_exterior_hbie<NN, NL>(acc, c_kp, x, n, xc, ni, w, il);
        else accumulate_exterior<NN, NL>(acc, c_kp, x, n, xc, w, il);
      }
    }
    warp_reduce<NN, NL>(acc);
    LaneEntries le; le.lane = lane;
    scatter_pair<NN, NL>(acc, g.ecol + (size_t)e * 3 * NN, g.ekind + (size_t)e * 3 * NN, g.ecv + (size_t)e * 6 * NN, g.erev[e] != 0, s, il,
                         r0, r1, r2, bre, bim, le, c_kp, HB, (unsigned)g.einfo[e] >> 5, (!HB && g.einc) ? g.einc + (size_t)e * 12 * NN : nullptr, g.ecol2 ? g.ecol2 + (size_t)e * 3 * NN : nullptr);
  }
  flush_b(s, r0, r1, r2, bre, bim);
}
void launch_adaptive(const DevGroup& g, const DevColloc& c, const DevSystem& s, const DevAdaptive& a, const DevTables& t, const KParams& kp, cudaStream_t st) {
  if (a.n_pairs == 0) return;
  dim3 grid((a.n_pairs + 3) / 4), block(128);
  if (c.cn) {
    switch (g.et) {
      case 5: k_adaptive<5, 3, true><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
      case 7: k_adaptive<7, 3, true><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
      case 6: k_adaptive<6, 1, true><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
      case 8: k_adaptive<8, 1, true><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
      case 9: k_adaptive<9, 1, true><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
    }
    return;
  }
  switch (g.et) {
    case 5: k_adaptive<5, 3, false><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
    case 7: k_adaptive<7, 3, false><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
    case 6: k_adaptive<6, 1, false><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
    case 8: k_adaptive<8, 1, false><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
    case 9: k_adaptive<9, 1, false><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K3: singular pairs -- one warp per pair, lanes stride over (ray, radial point); 15 radial Gauss-Legendre points on
// [0,1] per ray (ngp_rho = 15, bem_harela3d.f90:1298); the line-integral term phi_i*T2(1)*hli (:1462-1466) is added
// after the reduction.
// ------------------------------------------------------------------------------------------------------------------
template <int ET, int NL>
__global__ void __launch_bounds__(128) k_singular(DevGroup g, DevColloc c, DevSystem s, DevSingular a, DevTables t, const __grid_constant__ KParams c_kp) {
  const int NN = ElemTraits<ET>::NN;
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= a.n_pairs) return;
  const int cpos = a.pair_cpos[p], e = a.pair_elem[p];
  if (c.tile_active && !c.tile_active[cpos >> 5]) return;
  const int r0 = c.crow[cpos], r1 = c.crow[c.ldp + cpos], r2 = c.crow[2 * c.ldp + cpos];
  const double* D = a.pair_d + 14 * (size_t)p;
  const double xi_i0 = D[0], xi_i1 = D[1];
  const double xc[3] = {D[2], D[3], D[4]};
  double xn[3 * NN];
#pragma unroll
  for (int i = 0; i < 3 * NN; i++) xn[i] = g.xn[(size_t)e * 3 * NN + i];
  double phi_i[NN];
  { double d1[NN], d2[NN]; shape<ET>(xi_i0, xi_i1, phi_i, d1, d2); }
  const int ray0 = a.pair_ray0[p], nray = a.pair_ray0[p + 1] - ray0;
  const double* gx = t.gl01_x + 15 * 14 / 2;
  const double* gw = t.gl01_w + 15 * 14 / 2;
  double bre[3] = {0.0, 0.0, 0.0}, bim[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
  for (int il = 0; il < (NL == 3 ? 1 : 3); il++) {
    Acc<NN, NL> acc; acc.zero();
#pragma unroll 1
    for (int idx = lane; idx < nray * 15; idx += 32) {
      const int kr_ = idx / 15, kk = idx - kr_ * 15;
      const double* R = a.rays + 4 * (size_t)(ray0 + kr_);
      const double ct = __ldg(R), sn = __ldg(R + 1), rhoij = __ldg(R + 2), wray = __ldg(R + 3);
      const double rho = rhoij * __ldg(gx + kk), wrad = __ldg(gw + kk);
      double phi[NN], x[3], n[3], jg;
      geometry_at<ET>(xn, xi_i0 + rho * ct, xi_i1 + rho * sn, phi, x, n, jg);
      const double jw = jg * rho * wray * wrad;
      accumulate_interior<NN, NL>(acc, c_kp, x, n, xc, phi, phi_i, jw, il);
    }
    warp_reduce<NN, NL>(acc);
    // + phi_i(j) * T2(1) * hli(l,k)
    const cplx t21 = c_kp.T2[1];
#pragma unroll
    for (int ll = 0; ll < NL; ll++) {
      const int l = (NL == 3) ? ll : il;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double hl = D[5 + 3 * l + k];
#pragma unroll
        for (int j = 0; j < NN; j++) { acc.hr[(ll * 3 + k) * NN + j] += phi_i[j] * t21.re * hl; acc.hi[(ll * 3 + k) * NN + j] += phi_i[j] * t21.im * hl; }
      }
    }
    LaneEntries le; le.lane = lane;
    scatter_pair<NN, NL>(acc, g.ecol + (size_t)e * 3 * NN, g.ekind + (size_t)e * 3 * NN, g.ecv + (size_t)e * 6 * NN, g.erev[e] != 0, s, il,
                         r0, r1, r2, bre, bim, le, c_kp, false, (unsigned)g.einfo[e] >> 5, g.einc ? g.einc + (size_t)e * 12 * NN : nullptr, g.ecol2 ? g.ecol2 + (size_t)e * 3 * NN : nullptr);
  }
  flush_b(s, r0, r1, r2, bre, bim);
}
void launch_singular(const DevGroup& g, const DevColloc& c, const DevSystem& s, const DevSingular& a, const DevTables& t, const KParams& kp, cudaStream_t st) {
  if (a.n_pairs == 0) return;
  dim3 grid((a.n_pairs + 3) / 4), block(128);
  switch (g.et) {
    case 5: k_singular<5, 3><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
    case 7: k_singular<7, 3><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
    case 6: k_singular<6, 1><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
    case 8: k_singular<8, 1><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
    case 9: k_singular<9, 1><<<grid, block, 0, st>>>(g, c, s, a, t, kp); break;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K5: free-term entries (h-type): ctype 1 -> A(row,col_u) += c ; ctype 0 -> b(row) -= c*u_prescribed
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_freeterm(DevColloc c, DevSystem s, DevFreeTerm f, cplx F) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= f.n) return;
  const int cpos = f.cpos[i], l = f.l[i];
  if (c.tile_active && !c.tile_active[cpos >> 5]) return;
  const int row = c.crow[l * c.ldp + cpos];
  const int o = f.slot_off[f.slot[i]] + f.jk[i];
  // value = alpha + beta*F, F = -1/(8 pi (1-nu)) (Mantic; bem_harela3d.f90:538) -- alpha,beta are geometry-only
  const double vr = f.val[2 * i] + f.val[2 * i + 1] * F.re, vi = f.val[2 * i + 1] * F.im;
  if (f.einc) {   // the free term is part of h: b += c u_inc at the collocation node
    const double ur = f.einc[4 * (size_t)o], ui = f.einc[4 * (size_t)o + 1];
    atomicAdd(s.bre + row, vr * ur - vi * ui);
    atomicAdd(s.bim + row, vr * ui + vi * ur);
  }
  if (f.ekind[o] != 0) {   // t_k known, or u_k and t_k both unknown: the free term multiplies the unknown u_k
    atomicAdd(s.Are + (size_t)f.ecol[o] * s.lda + row, vr);
    atomicAdd(s.Aim + (size_t)f.ecol[o] * s.lda + row, vi);
  } else {
    const double cvr = f.ecv[2 * (size_t)o], cvi = f.ecv[2 * (size_t)o + 1];
    atomicAdd(s.bre + row, -(vr * cvr - vi * cvi));
    atomicAdd(s.bim + row, -(vr * cvi + vi * cvr));
  }
}
void launch_freeterm(const DevColloc& c, const DevSystem& s, const DevFreeTerm& f, cplx F, cudaStream_t st) {
  if (f.n > 0) k_freeterm<<<(f.n + 255) / 256, 256, 0, st>>>(c, s, f, F);
}

// planar <-> interleaved complex (host interface format), column-major; rowperm (or NULL) maps a host row to the
// library's internal row
__global__ void k_interleave(const double* __restrict__ re, const double* __restrict__ im, long long ld, int rows, int cols, double* __restrict__ out, long long ldo,
                             const int* __restrict__ rowperm, const int* __restrict__ colperm, int col0) {
  long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long cidx = i / rows, r = i - cidx * rows;
    long long ri = rowperm ? rowperm[r] : r;
    long long ci = colperm ? colperm[col0 + cidx] : col0 + cidx;
    out[2 * (cidx * ldo + r)] = re[ci * ld + ri];
    out[2 * (cidx * ldo + r) + 1] = im[ci * ld + ri];
  }
}
__global__ void k_deinterleave(const double* __restrict__ in, long long ldi, int rows, int cols, double* __restrict__ re, double* __restrict__ im, long long ld,
                               const int* __restrict__ rowperm) {
  long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long cidx = i / rows, r = i - cidx * rows;
    long long ri = rowperm ? rowperm[r] : r;
    re[cidx * ld + ri] = in[2 * (cidx * ldi + r)];
    im[cidx * ld + ri] = in[2 * (cidx * ldi + r) + 1];
  }
}
// r = A x - b and the componentwise scale s_i = sum_j |A_ij||x_j| + |b_i| (zgerfs-style backward error), planar storage.
// grid = (row blocks of 256, column chunks); partial sums are combined with atomics (rr, ri, ss are zeroed by the caller).
__global__ void k_residual(DevSystem s, const double* __restrict__ xre, const double* __restrict__ xim, int cchunk, double* rr, double* ri, double* ss) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = blockIdx.y * cchunk, c1 = min(c0 + cchunk, s.n_dof);
  if (i >= s.n_dof) return;
  double ar = 0.0, ai = 0.0, as = 0.0;
  for (int j = c0; j < c1; j++) {
    const double a = s.Are[(size_t)j * s.lda + i], b = s.Aim[(size_t)j * s.lda + i], xr = xre[j], xi = xim[j];
    ar += a * xr - b * xi; ai += a * xi + b * xr;
    as += (fabs(a) + fabs(b)) * (fabs(xr) + fabs(xi));
  }
  if (blockIdx.y == 0) { ar -= s.bre[i]; ai -= s.bim[i]; as += fabs(s.bre[i]) + fabs(s.bim[i]); }
  atomicAdd(rr + i, ar); atomicAdd(ri + i, ai); atomicAdd(ss + i, as);
}
void launch_residual(const DevSystem& s, const double* xre, const double* xim, double* rr, double* ri, double* ss, cudaStream_t st) {
  const int cchunk = 512;
  dim3 grid((s.n_dof + 255) / 256, (s.n_dof + cchunk - 1) / cchunk);
  k_residual<<<grid, 256, 0, st>>>(s, xre, xim, cchunk, rr, ri, ss);
}
// selected entries A(rows[i], cols[i]) -> out[i] (interleaved complex)
__global__ void k_get_entries(DevSystem s, int n, const int* __restrict__ rows, const int* __restrict__ cols, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { out[2 * i] = s.Are[(size_t)cols[i] * s.lda + rows[i]]; out[2 * i + 1] = s.Aim[(size_t)cols[i] * s.lda + rows[i]]; }
}
void launch_get_entries(const DevSystem& s, int n, const int* rows, const int* cols, double* out, cudaStream_t st) {
  if (n > 0) k_get_entries<<<(n + 255) / 256, 256, 0, st>>>(s, n, rows, cols, out);
}

// out[:, 0:cols) = host columns [col0, col0+cols) of the planar matrix (re, im): internal column colperm[c] (or c), internal row rowperm[r] (or r)
void launch_interleave(const double* re, const double* im, long long ld, int rows, int cols, double* out, long long ldo, const int* rowperm,
                       const int* colperm, int col0, cudaStream_t st) {
  k_interleave<<<2368, 256, 0, st>>>(re, im, ld, rows, cols, out, ldo, rowperm, colperm, col0);
}
void launch_deinterleave(const double* in, long long ldi, int rows, int cols, double* re, double* im, long long ld, const int* rowperm, cudaStream_t st) {
  k_deinterleave<<<2368, 256, 0, st>>>(in, ldi, rows, cols, re, im, ld, rowperm);
}

// real matrices of the static path (host interface: plain column-major doubles)
__global__ void k_gather_real(const double* __restrict__ re, long long ld, int rows, int cols, double* __restrict__ out, long long ldo, const int* __restrict__ rowperm,
                              const int* __restrict__ colperm, int col0) {
  long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long cidx = i / rows, r = i - cidx * rows;
    long long ri = rowperm ? rowperm[r] : r, ci = colperm ? colperm[col0 + cidx] : col0 + cidx;
    out[cidx * ldo + r] = re[ci * ld + ri];
  }
}
__global__ void k_scatter_real(const double* __restrict__ in, long long ldi, int rows, int cols, double* __restrict__ re, long long ld, const int* __restrict__ rowperm) {
  long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long cidx = i / rows, r = i - cidx * rows;
    long long ri = rowperm ? rowperm[r] : r;
    re[cidx * ld + ri] = in[cidx * ldi + r];
  }
}
void launch_gather_real(const double* re, long long ld, int rows, int cols, double* out, long long ldo, const int* rowperm, const int* colperm, int col0, cudaStream_t st) {
  k_gather_real<<<2368, 256, 0, st>>>(re, ld, rows, cols, out, ldo, rowperm, colperm, col0);
}
void launch_scatter_real(const double* in, long long ldi, int rows, int cols, double* re, long long ld, const int* rowperm, cudaStream_t st) {
  k_scatter_real<<<2368, 256, 0, st>>>(in, ldi, rows, cols, re, ld, rowperm);
}

}  // namespace mfbd
