// poro.cu -- sm_100a kernels of the Biot poroelastic SBIE assembly: four equations and four unknowns per node (0 = fluid phase tau | Un,
// 1..3 = skeleton u_k | t_k), ordinary `be` boundaries with the open-pore conditions 0 / 1 per component.
//
//   R1 k_por_regular   regular quadrature with the precalculated point sets (fbem_bem_harpor3d_sbie_ext_pre, lib/fbem/src/bem_harpor3d.f90:906-1011)
//                      fused with the BC-aware scatter (assemble_bem_harpor_equation.f90:78-110, :140-170)
//   R2 k_por_adaptive  Telles + subdivision leaves (fbem_bem_harpor3d_sbie_ext_st :1013-1416, leaf list of _ext_adp :1418-1540)
//   R3 k_por_singular  polar-transformation interior integration (fbem_bem_harpor3d_sbie_int :1542-1890): skeleton block with the CPV term and
//                      the line integrals of the static solution, fluid and coupling blocks weakly singular
// Free terms (c_00 = J c_pot, c_lk = Mantic's matrix of the drained skeleton, or phi_j/2 at MCA points) go through k_freeterm with two
// lists (multipliers F and J).  The point arithmetic is por_math.cuh (checked on the host against the oracle).
//
// STATUS: first hardware run in round 2 (compute-sanitizer clean; tests/test_gpu_poroelastic.py green on a B200, profiles/r02_first_contact.log); R1 rewritten
// later in that round (BC-aware accumulators, scalar cache, packed batches: DESIGN.md section 9.10).
//
// Mapping: a warp owns a collocation tile and walks a chunk of elements; R1 packs its pairs per rule (one pair per lane, see k_por_regular), R2 / R3 give
// a warp to a pair.  ONE equation (row l of the node block) per pass, so that a pair's accumulators fit the register file; R1 keeps the radial scalars
// of the first points in shared memory between the passes, R2 / R3 recompute the 4 x 4 point blocks in every pass.
#include "poro.cuh"
#include "por_pair.cuh"
#include <cstdio>

namespace mfbd {

__constant__ PorParams c_por;

void set_por_params(const PorParams& pp, cudaStream_t st) { cudaMemcpyToSymbolAsync(c_por, &pp, sizeof(PorParams), 0, cudaMemcpyHostToDevice, st); }

// BC-aware scatter of equation l of one pair (assemble_bem_harpor_equation.f90:78-110, :140-170; open-pore conditions): entries selected by
// `mine(j * 4 + k)`; h is scaled by cte_t(l, k) and changes sign on a reversed element, g by cte_u(l, k).
template <int NN, class Pred>
__device__ __forceinline__ void por_scatter_row(const RAcc<NN>& a, int l, const int* __restrict__ ecol, const unsigned char* __restrict__ ekind,
                                                const double* __restrict__ ecv, bool rev, const DevSystem& s, int row, double& bre, double& bim, Pred mine,
                                                unsigned symbits = 0u /* bit k: multiplier -1 of a symmetry image on dof k (symconf_s for k = 0, symconf_t(k) else) */,
                                                const double* __restrict__ einc = nullptr /* incident field of the element's (j, k): (u re, im, t re, im), or NULL */) {
#pragma unroll
  for (int k = 0; k < 4; k++) {
#pragma unroll
    for (int j = 0; j < NN; j++) {
      const int jk = j * 4 + k;
      if (!mine(jk)) continue;
      double hr, hi, gr, gi;
      por_finished_entry<NN>(a, c_por, l, k, j, rev, hr, hi, gr, gi);
      const int col = ecol[jk];
      const double cvr = ecv[2 * jk], cvi = ecv[2 * jk + 1];
      double ar, ai;
      if (ekind[jk] == 0) { ar = -gr; ai = -gi; bre -= hr * cvr - hi * cvi; bim -= hr * cvi + hi * cvr; }
      else { ar = hr; ai = hi; bre += gr * cvr - gi * cvi; bim += gr * cvi + gi * cvr; }
      if ((symbits >> k) & 1u) { ar = -ar; ai = -ai; }   // build_lse_mechanics_bem_harpor.f90:971-975; the b terms carry the sign in ecv
      if (einc) {   // b += hp u_inc - gp t_inc (assemble_bem_harpor_equation.f90:1277-1289); a symmetry image's sign is in the values
        const double ur = einc[4 * jk], ui = einc[4 * jk + 1], tr = einc[4 * jk + 2], ti = einc[4 * jk + 3];
        bre += (hr * ur - hi * ui) - (gr * tr - gi * ti); bim += (hr * ui + hi * ur) - (gr * ti + gi * tr);
      }
      atomicAdd(s.Are + (size_t)col * s.lda + row, ar);
      atomicAdd(s.Aim + (size_t)col * s.lda + row, ai);
    }
  }
}
struct PorAll { __device__ __forceinline__ bool operator()(int) const { return true; } };
struct PorLane { int lane; __device__ __forceinline__ bool operator()(int jk) const { return (jk & 31) == lane; } };

// ------------------------------------------------------------------------------------------------------------------
// R1: regular pairs
// ------------------------------------------------------------------------------------------------------------------
const int R1_WARPS = 4;
const int R1_ECHUNK = 32;
const int R1_CACHE_GP = 4;   // points of a pair whose twelve radial scalars stay in shared memory between the four equation passes (24 doubles x 32 lanes each)

// Row l of the two 4 x 4 blocks from the radial scalars (the formulas of por_exterior_blocks, one row of them)
__device__ __forceinline__ void por_row_from_scalars(const PorScal& s, const double* dx, const double* n, double drdn, int l, cplx ur[4], cplx tr[4]) {
  if (l == 0) {
    ur[0] = s.eta; tr[0] = s.W0 * drdn;
#pragma unroll
    for (int c = 0; c < 3; c++) { ur[c + 1] = s.vartheta * dx[c]; tr[c + 1] = cfmar(s.T01, dx[c] * drdn, s.T02 * n[c]); }
  } else {
    const double dxl = (l == 1) ? dx[0] : (l == 2 ? dx[1] : dx[2]), nl = (l == 1) ? n[0] : (l == 2 ? n[1] : n[2]);
    ur[0] = s.vartheta * dxl; tr[0] = cfmar(s.W1, dxl * drdn, s.W2 * nl);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double dl = (l - 1 == k) ? 1.0 : 0.0, dd = dxl * dx[k];
      ur[k + 1] = mk(s.psi.re * dl - s.chi.re * dd, s.psi.im * dl - s.chi.im * dd);
      const double c1 = dd * drdn, c2 = drdn * dl + dx[k] * nl, c3 = dxl * n[k];
      tr[k + 1] = mk(s.T1.re * c1 + s.T2.re * c2 + s.T3.re * c3, s.T1.im * c1 + s.T2.im * c2 + s.T3.im * c3);
    }
  }
}
#define POR_SCAL_FIELDS(X) X(eta, 0) X(vartheta, 1) X(psi, 2) X(chi, 3) X(W0, 4) X(T01, 5) X(T02, 6) X(W1, 7) X(W2, 8) X(T1, 9) X(T2, 10) X(T3, 11)
__device__ __forceinline__ void por_scal_store(const PorScal& s, double* c) {
#define X(f, i) c[(2 * i) * 32] = s.f.re; c[(2 * i + 1) * 32] = s.f.im;
  POR_SCAL_FIELDS(X)
#undef X
}
__device__ __forceinline__ void por_scal_load(PorScal& s, const double* c) {
#define X(f, i) s.f = mk(c[(2 * i) * 32], c[(2 * i + 1) * 32]);
  POR_SCAL_FIELDS(X)
#undef X
}

const int R1_QCAP = 64;      // entries per deferred queue (at most 31 waiting + 32 new)

// One regular pair (collocation point xs with matrix rows rws, element el of the group, rule sset), all four equations; sc = this lane's column of the
// shared-memory scalar cache.  The pair's collocation point need not be the lane's own (packed batches), so the b terms go out with atomics.
template <int ET>
__device__ __forceinline__ void por_regular_pair(const DevGroup& g, const DevSystem& s, int sset, int el, const double* xs, const int* rws, double* sc) {
  constexpr int NN = ElemTraits<ET>::NN, RECN = 6 + NN;
  const int ngp = g.ngp[sset];
  const double* P = g.pts[sset] + (size_t)el * ngp * RECN;
  const int* ecol = g.ecol + (size_t)el * 4 * NN;
  const unsigned char* ekind = g.ekind + (size_t)el * 4 * NN;
  const double* ecv = g.ecv + (size_t)el * 8 * NN;
  const bool rev = g.erev[el] != 0;
  const unsigned info = g.einfo[el];
  // The common element: the same kind of condition on all its nodes, all prescribed values zero.  Only the combination that goes to the matrix is
  // accumulated (4 NN complex numbers per equation: they stay in registers; the general path below keeps h AND g, 16 NN doubles, and for 8/9-node
  // elements lives in local memory -- 4.6 KB of spills per thread, the kernel's cost in round 1), and the radial scalars of the first R1_CACHE_GP
  // points are computed in the pass of equation 0 and read back from shared memory by the other three.
  if ((info & 8u) && !g.ecvnz[el] && !g.einc) {
    unsigned kinds = 0u;
#pragma unroll
    for (int k = 0; k < 4; k++) kinds |= (ekind[k] != 0 ? 1u : 0u) << k;
#pragma unroll 1
    for (int l = 0; l < 4; l++) {
      double ar[4 * NN], ai[4 * NN];
#pragma unroll
      for (int i = 0; i < 4 * NN; i++) { ar[i] = 0.0; ai[i] = 0.0; }
#pragma unroll 1
      for (int kp = 0; kp < ngp; kp++) {
        const double* q = P + (size_t)kp * RECN;
        const double n[3] = {__ldg(q + 3), __ldg(q + 4), __ldg(q + 5)};
        const double rv0 = __ldg(q) - xs[0], rv1 = __ldg(q + 1) - xs[1], rv2 = __ldg(q + 2) - xs[2];
        const double r = sqrt(rv0 * rv0 + rv1 * rv1 + rv2 * rv2), d1r1 = 1.0 / r;
        const double dx[3] = {rv0 * d1r1, rv1 * d1r1, rv2 * d1r1};
        const double drdn = dx[0] * n[0] + dx[1] * n[1] + dx[2] * n[2];
        PorScal ps;
        if (l > 0 && kp < R1_CACHE_GP) por_scal_load(ps, sc + (size_t)kp * 24 * 32);
        else {
          por_scalars<false>(c_por, r, d1r1, ps);
          if (l == 0 && kp < R1_CACHE_GP) por_scal_store(ps, sc + (size_t)kp * 24 * 32);
        }
        cplx ur[4], tr[4];
        por_row_from_scalars(ps, dx, n, drdn, l, ur, tr);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const cplx f = ((kinds >> k) & 1u) ? tr[k] : ur[k];
#pragma unroll
          for (int j = 0; j < NN; j++) { const double wj = __ldg(q + 6 + j); ar[k * NN + j] = fma(f.re, wj, ar[k * NN + j]); ai[k * NN + j] = fma(f.im, wj, ai[k * NN + j]); }
        }
      }
      const int row = (l == 0) ? rws[0] : (l == 1 ? rws[1] : (l == 2 ? rws[2] : rws[3]));
#pragma unroll
      for (int k = 0; k < 4; k++) {
        // A += cte_t h (sign of the orientation) for a dof whose secondary variable is known, A -= cte_u g otherwise (assemble_bem_harpor_equation.f90:78-110)
        const bool tk = (kinds >> k) & 1u;
        const cplx c0 = tk ? ((l == 0) ? c_por.cte_t[0][k] : c_por.cte_t[1][k]) : ((l == 0) ? c_por.cte_u[0][k] : c_por.cte_u[1][k]);
        const double sg = (tk ? (rev ? -1.0 : 1.0) : -1.0) * (((info >> (4 + k)) & 1u) ? -1.0 : 1.0);   // sign of a symmetry image on dof k
#pragma unroll
        for (int j = 0; j < NN; j++) {
          const int col = ecol[j * 4 + k];
          atomicAdd(s.Are + (size_t)col * s.lda + row, sg * (c0.re * ar[k * NN + j] - c0.im * ai[k * NN + j]));
          atomicAdd(s.Aim + (size_t)col * s.lda + row, sg * (c0.re * ai[k * NN + j] + c0.im * ar[k * NN + j]));
        }
      }
    }
    return;
  }
#pragma unroll 1
  for (int l = 0; l < 4; l++) {
    RAcc<NN> acc; acc.zero();
#pragma unroll 1
    for (int kp = 0; kp < ngp; kp++) {
      const double* q = P + (size_t)kp * RECN;
      const double x[3] = {__ldg(q), __ldg(q + 1), __ldg(q + 2)}, n[3] = {__ldg(q + 3), __ldg(q + 4), __ldg(q + 5)};
      double w[NN];
#pragma unroll
      for (int j = 0; j < NN; j++) w[j] = __ldg(q + 6 + j);
      por_regular_point<NN>(acc, c_por, x, n, w, xs, l);
    }
    const int row = (l == 0) ? rws[0] : (l == 1 ? rws[1] : (l == 2 ? rws[2] : rws[3]));
    double br = 0.0, bi = 0.0;
    por_scatter_row<NN>(acc, l, ecol, ekind, ecv, rev, s, row, br, bi, PorAll(), info >> 4, g.einc ? g.einc + (size_t)el * 16 * NN : nullptr);
    if (br != 0.0 || bi != 0.0) { atomicAdd(s.bre + row, br); atomicAdd(s.bim + row, bi); }
  }
}

// A warp owns a collocation tile and walks a chunk of elements.  The 32 points of a tile ask for different rules near an element, so integrating "all lanes
// on element e" left 36 % of the lanes busy (ncu, profiles/r02_ncu_por_regular.txt).  Pairs are therefore pushed to one queue per rule in shared memory and
// integrated 32 at a time, one (point, element) pair per lane, as soon as a queue holds a full warp (the scheme of k_regular_bulk in assembly.cu).
template <int ET>
__global__ void __launch_bounds__(R1_WARPS * 32) k_por_regular(DevGroup g, DevColloc c, DevSystem s, const unsigned char* __restrict__ plan) {
  extern __shared__ __align__(16) double por_smem[];   // [warp][R1_CACHE_GP][24 scalars][32 lanes], then the queues
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = blockIdx.x * R1_WARPS + warp;
  if (tile >= c.n_tiles) return;
  if (c.tile_active && !c.tile_active[tile]) return;
  const int cpos = tile * 32 + lane;
  const int rows[4] = {c.crow[cpos], c.crow[c.ldp + cpos], c.crow[2 * c.ldp + cpos], c.crow[3 * c.ldp + cpos]};
  const bool valid = rows[0] >= 0;
  const double xc[3] = {c.cx[cpos], c.cx[c.ldp + cpos], c.cx[2 * c.ldp + cpos]};
  double* sc = por_smem + ((size_t)warp * R1_CACHE_GP * 24) * 32 + lane;
  unsigned short* queue = reinterpret_cast<unsigned short*>(por_smem + (size_t)R1_WARPS * R1_CACHE_GP * 24 * 32) + (size_t)warp * (MAX_SETS * R1_QCAP);
  int* qcnt = reinterpret_cast<int*>(reinterpret_cast<unsigned short*>(por_smem + (size_t)R1_WARPS * R1_CACHE_GP * 24 * 32) + (size_t)R1_WARPS * (MAX_SETS * R1_QCAP)) + warp * MAX_SETS;
  const unsigned lt_mask = (1u << lane) - 1u;
  if (lane < MAX_SETS) qcnt[lane] = 0;
  __syncwarp();
  const int e0 = blockIdx.y * R1_ECHUNK, e1 = min(e0 + R1_ECHUNK, g.n_elem);
  auto batch = [&](int sset, unsigned short ent, bool act) {
    const int src = ent & 31, el = e0 + (ent >> 5);
    const double xs[3] = {__shfl_sync(0xffffffffu, xc[0], src), __shfl_sync(0xffffffffu, xc[1], src), __shfl_sync(0xffffffffu, xc[2], src)};
    const int rws[4] = {__shfl_sync(0xffffffffu, rows[0], src), __shfl_sync(0xffffffffu, rows[1], src), __shfl_sync(0xffffffffu, rows[2], src), __shfl_sync(0xffffffffu, rows[3], src)};
    if (act) por_regular_pair<ET>(g, s, sset, el, xs, rws, sc);
    __syncwarp();
  };
  // one place where batches are integrated (a second call site would double the code of the pair body: the instruction cache is what this kernel waits for,
  // ncu stall reason no_instruction, profiles/r02_ncu_por_regular_after.txt)
  int e = e0, drain = 0;
  unsigned fullsets = 0u;      // rules whose queue holds >= 32 entries
  for (;;) {
    int sset; unsigned short ent; bool act;
    if (fullsets) {
      sset = __ffs(fullsets) - 1;
      int cnt = qcnt[sset];
      ent = queue[sset * R1_QCAP + cnt - 32 + lane];
      __syncwarp();
      cnt -= 32;
      if (lane == 0) qcnt[sset] = cnt;
      __syncwarp();
      if (cnt < 32) fullsets &= ~(1u << sset);
      act = true;
    } else if (e < e1) {
      const unsigned char m = valid ? plan[(size_t)(g.slot0 + e) * c.ldp + cpos] : PLAN_NONE;
      unsigned todo = __ballot_sync(0xffffffffu, m < MAX_SETS);
      while (todo) {
        const int leader = __ffs(todo) - 1;
        const int qs = __shfl_sync(0xffffffffu, (int)m, leader);
        const unsigned grp = __ballot_sync(0xffffffffu, (int)m == qs) & todo;
        todo &= ~grp;
        const int base = qcnt[qs];
        if ((grp >> lane) & 1u) queue[qs * R1_QCAP + base + __popc(grp & lt_mask)] = (unsigned short)(((e - e0) << 5) | lane);
        __syncwarp();
        const int cnt = base + __popc(grp);
        if (lane == 0) qcnt[qs] = cnt;
        if (cnt >= 32) fullsets |= 1u << qs;
        __syncwarp();
      }
      e++;
      continue;
    } else if (drain < g.n_sets) {   // what is left in the queues
      sset = drain++;
      const int cnt = qcnt[sset];
      if (cnt == 0) continue;
      act = lane < cnt;
      ent = act ? queue[sset * R1_QCAP + lane] : (unsigned short)0;
    } else break;
    batch(sset, ent, act);
  }
}
void launch_por_regular(const DevGroup& g, const DevColloc& c, const DevSystem& s, const unsigned char* plan, cudaStream_t st) {
  if (g.n_elem == 0) return;
  dim3 grid((c.n_tiles + R1_WARPS - 1) / R1_WARPS, (g.n_elem + R1_ECHUNK - 1) / R1_ECHUNK), block(R1_WARPS * 32);
  const int smem = R1_WARPS * R1_CACHE_GP * 24 * 32 * (int)sizeof(double) + R1_WARPS * MAX_SETS * (R1_QCAP * 2 + 4);   // 96 KB of scalar cache + 8.5 KB of queues: two CTAs per SM
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_por_regular<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); cudaFuncSetAttribute(k_por_regular<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_por_regular<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); cudaFuncSetAttribute(k_por_regular<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_por_regular<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  switch (g.et) {
    case 5: k_por_regular<5><<<grid, block, smem, st>>>(g, c, s, plan); break;
    case 6: k_por_regular<6><<<grid, block, smem, st>>>(g, c, s, plan); break;
    case 7: k_por_regular<7><<<grid, block, smem, st>>>(g, c, s, plan); break;
    case 8: k_por_regular<8><<<grid, block, smem, st>>>(g, c, s, plan); break;
    case 9: k_por_regular<9><<<grid, block, smem, st>>>(g, c, s, plan); break;
  }
}

template <int NN>
__device__ __forceinline__ void por_warp_reduce(RAcc<NN>& a) {
#pragma unroll
  for (int i = 0; i < 4 * NN; i++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a.hr[i] += __shfl_xor_sync(0xffffffffu, a.hr[i], o); a.hi[i] += __shfl_xor_sync(0xffffffffu, a.hi[i], o);
      a.gr[i] += __shfl_xor_sync(0xffffffffu, a.gr[i], o); a.gi[i] += __shfl_xor_sync(0xffffffffu, a.gi[i], o);
    }
  }
}

// Row l of the two blocks at a point of the element that holds the collocation point (por_interior_blocks, one row of it): ps = por_scalars<true>, the static
// 1/r^2 parts of W0, T1 and of the dr/dn delta term of T2 are added back here; fc[k] = T2(1)/r^2 (n_l r,k - n_k r,l) is the CPV kernel of the skeleton block.
__device__ __forceinline__ void por_interior_row_from_scalars(const PorParams& p, const PorScal& s, const double* dx, const double* n, double drdn, double d1r2, int l,
                                                              cplx ur[4], cplx tr[4], cplx fc[3]) {
  if (l == 0) {
    ur[0] = s.eta; tr[0] = cfmar(p.W0[1], d1r2, s.W0) * drdn;
#pragma unroll
    for (int c = 0; c < 3; c++) { ur[c + 1] = s.vartheta * dx[c]; tr[c + 1] = cfmar(s.T01, dx[c] * drdn, s.T02 * n[c]); fc[c] = mk(0.0, 0.0); }
  } else {
    const double dxl = (l == 1) ? dx[0] : (l == 2 ? dx[1] : dx[2]), nl = (l == 1) ? n[0] : (l == 2 ? n[1] : n[2]);
    ur[0] = s.vartheta * dxl; tr[0] = cfmar(s.W1, dxl * drdn, s.W2 * nl);
    const cplx T1s = cfmar(p.T1[1], d1r2, s.T1), T2s = cfmar(p.T2[1], d1r2, s.T2), T21 = p.T2[1] * d1r2;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double dl = (l - 1 == k) ? 1.0 : 0.0, dd = dxl * dx[k];
      ur[k + 1] = mk(s.psi.re * dl - s.chi.re * dd, s.psi.im * dl - s.chi.im * dd);
      const double c1 = dd * drdn, c2 = drdn * dl, c3 = dx[k] * nl, c4 = dxl * n[k];
      tr[k + 1] = mk(T1s.re * c1 + T2s.re * c2 + s.T2.re * c3 + s.T3.re * c4, T1s.im * c1 + T2s.im * c2 + s.T2.im * c3 + s.T3.im * c4);
      fc[k] = T21 * (nl * dx[k] - n[k] * dxl);
    }
  }
}
// warp sum of the BC-aware accumulators and their lane-strided scatter: entry (j, k) of equation l goes to the matrix with cte_t (and the sign of the
// orientation) when the secondary variable of dof k is known, with -cte_u otherwise (assemble_bem_harpor_equation.f90:78-110); info bits 4-7: symmetry image
template <int NN>
__device__ __forceinline__ void por_reduce_scatter_fast(double* ar, double* ai, unsigned kinds, unsigned info, bool rev, int l, const int* __restrict__ ecol,
                                                        const DevSystem& s, int row, int lane) {
#pragma unroll
  for (int i = 0; i < 4 * NN; i++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ar[i] += __shfl_xor_sync(0xffffffffu, ar[i], o); ai[i] += __shfl_xor_sync(0xffffffffu, ai[i], o); }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const bool tk = (kinds >> k) & 1u;
    const cplx c0 = tk ? ((l == 0) ? c_por.cte_t[0][k] : c_por.cte_t[1][k]) : ((l == 0) ? c_por.cte_u[0][k] : c_por.cte_u[1][k]);
    const double sg = (tk ? (rev ? -1.0 : 1.0) : -1.0) * (((info >> (4 + k)) & 1u) ? -1.0 : 1.0);
#pragma unroll
    for (int j = 0; j < NN; j++) {
      if (((j * 4 + k) & 31) != lane) continue;
      const int col = ecol[j * 4 + k];
      atomicAdd(s.Are + (size_t)col * s.lda + row, sg * (c0.re * ar[k * NN + j] - c0.im * ai[k * NN + j]));
      atomicAdd(s.Aim + (size_t)col * s.lda + row, sg * (c0.re * ai[k * NN + j] + c0.im * ar[k * NN + j]));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// R2: adaptive pairs -- one warp per pair, lanes stride over the gln x gln points of every leaf, one equation per pass
// ------------------------------------------------------------------------------------------------------------------
template <int ET>
__global__ void __launch_bounds__(128) k_por_adaptive(DevGroup g, DevColloc c, DevSystem s, DevAdaptive a, DevTables t) {
  constexpr int NN = ElemTraits<ET>::NN;
  constexpr bool tri = (ElemTraits<ET>::NV == 3);
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= a.n_pairs) return;
  const int cpos = a.pair_cpos[p], e = a.pair_elem[p];
  if (c.tile_active && !c.tile_active[cpos >> 5]) return;
  const double xc[3] = {c.cx[cpos], c.cx[c.ldp + cpos], c.cx[2 * c.ldp + cpos]};
  double xn[3 * NN];
#pragma unroll
  for (int i = 0; i < 3 * NN; i++) xn[i] = g.xn[(size_t)e * 3 * NN + i];
  const double* gx = tri ? t.gl01_x : t.gl11_x;
  const double* gw = tri ? t.gl01_w : t.gl11_w;
  const bool rev = g.erev[e] != 0;
  if ((g.einfo[e] & 8u) && !g.ecvnz[e] && !g.einc) {   // the common element (uniform kinds, no prescribed value): only what goes to the matrix, in registers (see k_por_regular)
    const unsigned info = g.einfo[e];
    const unsigned char* ekind = g.ekind + (size_t)e * 4 * NN;
    unsigned kinds = 0u;
#pragma unroll
    for (int k = 0; k < 4; k++) kinds |= (ekind[k] != 0 ? 1u : 0u) << k;
#pragma unroll 1
    for (int l = 0; l < 4; l++) {
      double ar[4 * NN], ai[4 * NN];
#pragma unroll
      for (int i = 0; i < 4 * NN; i++) { ar[i] = 0.0; ai[i] = 0.0; }
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
          const double rv0 = x[0] - xc[0], rv1 = x[1] - xc[1], rv2 = x[2] - xc[2];
          const double r = sqrt(rv0 * rv0 + rv1 * rv1 + rv2 * rv2), d1r1 = 1.0 / r;
          const double dx[3] = {rv0 * d1r1, rv1 * d1r1, rv2 * d1r1};
          const double drdn = dx[0] * n[0] + dx[1] * n[1] + dx[2] * n[2];
          PorScal ps; por_scalars<false>(c_por, r, d1r1, ps);
          cplx ur[4], tr[4];
          por_row_from_scalars(ps, dx, n, drdn, l, ur, tr);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const cplx f = ((kinds >> k) & 1u) ? tr[k] : ur[k];
#pragma unroll
            for (int j = 0; j < NN; j++) { ar[k * NN + j] = fma(f.re, w[j], ar[k * NN + j]); ai[k * NN + j] = fma(f.im, w[j], ai[k * NN + j]); }
          }
        }
      }
      por_reduce_scatter_fast<NN>(ar, ai, kinds, info, rev, l, g.ecol + (size_t)e * 4 * NN, s, c.crow[l * c.ldp + cpos], lane);
    }
    return;
  }
#pragma unroll 1
  for (int l = 0; l < 4; l++) {
    RAcc<NN> acc; acc.zero();
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
        por_leaf_point<ET>(acc, c_por, xn, xi_s, tp1, tp2, __ldg(gx + off + k1), __ldg(gw + off + k1), __ldg(gx + off + k2), __ldg(gw + off + k2), xc, l);
      }
    }
    por_warp_reduce<NN>(acc);
    const int row = c.crow[l * c.ldp + cpos];
    double br = 0.0, bi = 0.0;
    PorLane pl; pl.lane = lane;
    por_scatter_row<NN>(acc, l, g.ecol + (size_t)e * 4 * NN, g.ekind + (size_t)e * 4 * NN, g.ecv + (size_t)e * 8 * NN, rev, s, row, br, bi, pl, (unsigned)g.einfo[e] >> 4,
                         g.einc ? g.einc + (size_t)e * 16 * NN : nullptr);
    if (br != 0.0 || bi != 0.0) { atomicAdd(s.bre + row, br); atomicAdd(s.bim + row, bi); }
  }
}
void launch_por_adaptive(const DevGroup& g, const DevColloc& c, const DevSystem& s, const DevAdaptive& a, const DevTables& t, cudaStream_t st) {
  if (a.n_pairs == 0) return;
  dim3 grid((a.n_pairs + 3) / 4), block(128);
  switch (g.et) {
    case 5: k_por_adaptive<5><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 6: k_por_adaptive<6><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 7: k_por_adaptive<7><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 8: k_por_adaptive<8><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 9: k_por_adaptive<9><<<grid, block, 0, st>>>(g, c, s, a, t); break;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// R3: singular pairs -- one warp per pair, lanes stride over (ray, radial point), one equation per pass; the CPV kernel of the skeleton
// block is integrated against (phi_j - phi_j(xi_i)) and the line-integral term phi_j(xi_i) T2(1) hli(l,k) (bem_harpor3d.f90:1876-1880) is
// added after the reduction.
// ------------------------------------------------------------------------------------------------------------------
template <int ET>
__global__ void __launch_bounds__(128) k_por_singular(DevGroup g, DevColloc c, DevSystem s, DevSingular a, DevTables t) {
  constexpr int NN = ElemTraits<ET>::NN;
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= a.n_pairs) return;
  const int cpos = a.pair_cpos[p], e = a.pair_elem[p];
  if (c.tile_active && !c.tile_active[cpos >> 5]) return;
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
  const bool rev = g.erev[e] != 0;
  if ((g.einfo[e] & 8u) && !g.ecvnz[e] && !g.einc) {   // the common element: only what goes to the matrix, in registers
    const unsigned info = g.einfo[e];
    const unsigned char* ekind = g.ekind + (size_t)e * 4 * NN;
    unsigned kinds = 0u;
#pragma unroll
    for (int k = 0; k < 4; k++) kinds |= (ekind[k] != 0 ? 1u : 0u) << k;
#pragma unroll 1
    for (int l = 0; l < 4; l++) {
      double ar[4 * NN], ai[4 * NN];
#pragma unroll
      for (int i = 0; i < 4 * NN; i++) { ar[i] = 0.0; ai[i] = 0.0; }
#pragma unroll 1
      for (int idx = lane; idx < nray * 15; idx += 32) {
        const int kr_ = idx / 15, kk = idx - kr_ * 15;
        const double* R = a.rays + 4 * (size_t)(ray0 + kr_);
        const double ct = __ldg(R), sn = __ldg(R + 1), rhoij = __ldg(R + 2), wray = __ldg(R + 3);
        const double rho = rhoij * __ldg(gx + kk), wrad = __ldg(gw + kk);
        double phi[NN], x[3], n[3], jg;
        geometry_at<ET>(xn, xi_i0 + rho * ct, xi_i1 + rho * sn, phi, x, n, jg);
        const double jw = jg * rho * wray * wrad;
        const double rv0 = x[0] - xc[0], rv1 = x[1] - xc[1], rv2 = x[2] - xc[2];
        const double r = sqrt(rv0 * rv0 + rv1 * rv1 + rv2 * rv2), d1r1 = 1.0 / r;
        const double dx[3] = {rv0 * d1r1, rv1 * d1r1, rv2 * d1r1};
        const double drdn = dx[0] * n[0] + dx[1] * n[1] + dx[2] * n[2];
        PorScal ps; por_scalars<true>(c_por, r, d1r1, ps);
        cplx ur[4], tr[4], fc[3];
        por_interior_row_from_scalars(c_por, ps, dx, n, drdn, d1r1 * d1r1, l, ur, tr, fc);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const bool tk = (kinds >> k) & 1u;
          const cplx f = tk ? tr[k] : ur[k];
#pragma unroll
          for (int j = 0; j < NN; j++) {
            const double wj = phi[j] * jw;
            ar[k * NN + j] = fma(f.re, wj, ar[k * NN + j]); ai[k * NN + j] = fma(f.im, wj, ai[k * NN + j]);
          }
          if (tk && l > 0 && k > 0) {   // CPV kernel of the skeleton block against phi_j - phi_j(xi_i)
#pragma unroll
            for (int j = 0; j < NN; j++) {
              const double wc = (phi[j] - phi_i[j]) * jw;
              ar[k * NN + j] = fma(fc[k - 1].re, wc, ar[k * NN + j]); ai[k * NN + j] = fma(fc[k - 1].im, wc, ai[k * NN + j]);
            }
          }
        }
      }
      // + phi_j(xi_i) T2(1) hli(l, k) on the skeleton block of h (bem_harpor3d.f90:1876-1880), added once per warp: lane 0 before the warp sum
      if (l > 0 && lane == 0) {
        const cplx t21 = c_por.T2[1];
#pragma unroll
        for (int k = 0; k < 3; k++)
          if ((kinds >> (k + 1)) & 1u) {
            const double hl = D[5 + 3 * (l - 1) + k];
#pragma unroll
            for (int j = 0; j < NN; j++) { ar[(k + 1) * NN + j] += phi_i[j] * t21.re * hl; ai[(k + 1) * NN + j] += phi_i[j] * t21.im * hl; }
          }
      }
      por_reduce_scatter_fast<NN>(ar, ai, kinds, info, rev, l, g.ecol + (size_t)e * 4 * NN, s, c.crow[l * c.ldp + cpos], lane);
    }
    return;
  }
#pragma unroll 1
  for (int l = 0; l < 4; l++) {
    RAcc<NN> acc; acc.zero();
#pragma unroll 1
    for (int idx = lane; idx < nray * 15; idx += 32) {
      const int kr_ = idx / 15, kk = idx - kr_ * 15;
      const double* R = a.rays + 4 * (size_t)(ray0 + kr_);
      const double ct = __ldg(R), sn = __ldg(R + 1), rhoij = __ldg(R + 2), wray = __ldg(R + 3);
      const double rho = rhoij * __ldg(gx + kk), wrad = __ldg(gw + kk);
      por_singular_point<ET>(acc, c_por, xn, xi_i0, xi_i1, phi_i, ct, sn, rho, wray, wrad, xc, l);
    }
    por_warp_reduce<NN>(acc);
    por_singular_line_terms<NN>(acc, c_por, phi_i, D + 5, l);
    const int row = c.crow[l * c.ldp + cpos];
    double br = 0.0, bi = 0.0;
    PorLane pl; pl.lane = lane;
    por_scatter_row<NN>(acc, l, g.ecol + (size_t)e * 4 * NN, g.ekind + (size_t)e * 4 * NN, g.ecv + (size_t)e * 8 * NN, rev, s, row, br, bi, pl, (unsigned)g.einfo[e] >> 4,
                         g.einc ? g.einc + (size_t)e * 16 * NN : nullptr);
    if (br != 0.0 || bi != 0.0) { atomicAdd(s.bre + row, br); atomicAdd(s.bim + row, bi); }
  }
}
void launch_por_singular(const DevGroup& g, const DevColloc& c, const DevSystem& s, const DevSingular& a, const DevTables& t, cudaStream_t st) {
  if (a.n_pairs == 0) return;
  dim3 grid((a.n_pairs + 3) / 4), block(128);
  switch (g.et) {
    case 5: k_por_singular<5><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 6: k_por_singular<6><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 7: k_por_singular<7><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 8: k_por_singular<8><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 9: k_por_singular<9><<<grid, block, 0, st>>>(g, c, s, a, t); break;
  }
}

}  // namespace mfbd
