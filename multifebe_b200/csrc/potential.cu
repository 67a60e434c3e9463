// potential.cu -- sm_100a kernels of the scalar-wave (inviscid fluid / acoustic) SBIE assembly: one equation and one unknown
// per node.
//
//   P1 k_pot_regular   regular quadrature with the precalculated point sets (fbem_bem_harpot3d_sbie_ext_pre,
//                      lib/fbem/src/bem_harpot3d.f90:276-328) fused with the BC-aware scatter (assemble_bem_harpot_equation.f90:78-96)
//   P2 k_pot_adaptive  Telles + subdivision leaves (fbem_bem_harpot3d_sbie_ext_st :330-640, leaf list of _ext_adp :642-764)
//   P3 k_pot_singular  polar-transformation interior integration (fbem_bem_harpot3d_sbie_int :767-959; weakly singular: no
//                      CPV part, no line integrals)
// The free term (c = solid angle / 4 pi, fbem_bem_pot3d_sbie_freeterm; or phi_j(xi_i)/2 for an MCA point) goes through
// k_freeterm of assembly.cu with the scalar descriptors.
//
// Mapping of P1: a warp owns one collocation tile (32 lanes = 32 points whose matrix rows are consecutive in the library's
// internal row order) and walks a chunk of elements; the point set of the element is read with warp-uniform loads, the 2*NN
// complex accumulators of the pair live in registers and go to the planar matrix as coalesced RED.ADD.F64 (32 consecutive rows
// of one column per instruction).  The right-hand side of each lane is reduced in registers over the chunk and flushed once.
#include "potential.cuh"
#include <cstdio>

namespace mfbd {

__constant__ PotParams c_pp;

void set_pot_params(const PotParams& pp, cudaStream_t st) { cudaMemcpyToSymbolAsync(c_pp, &pp, sizeof(PotParams), 0, cudaMemcpyHostToDevice, st); }

// BC-aware scatter of node j of one pair (assemble_bem_harpot_equation.f90:78-96): raw sums (hr, hi), (gr, gi) -> h = -/+ c4pi * sum
// (sign flips on a reversed element), g = c4pi * d1J * sum; ctype 0 (p known): A(row, col_Un) -= g, b -= h p; ctype 1 (Un known):
// A(row, col_p) += h, b += g Un.
// neg: the element is a symmetry image whose scalar multiplier symconf_s is -1 (build_lse_mechanics_bem_harpot.f90:947-948): the matrix entry changes sign
// here, the b term through the prescribed value (k_gather_cv folds the sign into ecv)
__device__ __forceinline__ void pot_scatter_node(double hr, double hi, double gr, double gi, bool rev, int col, int kind, double cvr, double cvi,
                                                 const DevSystem& s, int row, double& bre, double& bim, bool neg = false,
                                                 const double* __restrict__ inc = nullptr /* (p_inc re, im, Un_inc re, im) of this element node, or NULL */) {
  const double sh = rev ? c_pp.c4pi : -c_pp.c4pi, sg = c_pp.c4pi * c_pp.d1J;
  hr *= sh; hi *= sh; gr *= sg; gi *= sg;
  double ar, ai;
  if (kind == 0) { ar = -gr; ai = -gi; bre -= hr * cvr - hi * cvi; bim -= hr * cvi + hi * cvr; }
  else { ar = hr; ai = hi; bre += gr * cvr - gi * cvi; bim += gr * cvi + gi * cvr; }
  if (neg) { ar = -ar; ai = -ai; }
  if (inc) {   // incident field: b += hp p_inc - gp Un_inc (assemble_bem_harpot_equation.f90:471-481); a symmetry image's sign is in the values
    bre += (hr * inc[0] - hi * inc[1]) - (gr * inc[2] - gi * inc[3]); bim += (hr * inc[1] + hi * inc[0]) - (gr * inc[3] + gi * inc[2]);
  }
  atomicAdd(s.Are + (size_t)col * s.lda + row, ar);
  atomicAdd(s.Aim + (size_t)col * s.lda + row, ai);
}

// ------------------------------------------------------------------------------------------------------------------
// P1: regular pairs
// ------------------------------------------------------------------------------------------------------------------
const int P1_WARPS = 4;
const int P1_ECHUNK = 64;

const int P1_QCAP = 64;   // entries per deferred queue (at most 31 waiting + 32 new)

// A warp owns a collocation tile and walks a chunk of elements.  Near an element the 32 points of a tile ask for different rules, and integrating "all lanes on
// element e" kept 22 of 32 lanes busy (ncu, round 1).  Pairs are pushed to one queue per rule in shared memory and integrated 32 at a time, one (point, element)
// pair per lane, as soon as a queue holds a full warp (the scheme of k_regular_bulk in assembly.cu and k_por_regular).
template <int ET>
__global__ void __launch_bounds__(P1_WARPS * 32) k_pot_regular(DevGroup g, DevColloc c, DevSystem s, const unsigned char* __restrict__ plan) {
  constexpr int NN = ElemTraits<ET>::NN, RECN = 6 + NN;
  __shared__ unsigned short queue_all[P1_WARPS][MAX_SETS * P1_QCAP];
  __shared__ int qcnt_all[P1_WARPS][MAX_SETS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = blockIdx.x * P1_WARPS + warp;
  if (tile >= c.n_tiles) return;
  if (c.tile_active && !c.tile_active[tile]) return;
  unsigned short* queue = queue_all[warp]; int* qcnt = qcnt_all[warp];
  const unsigned lt_mask = (1u << lane) - 1u;
  if (lane < MAX_SETS) qcnt[lane] = 0;
  __syncwarp();
  const int cpos = tile * 32 + lane;
  const int row = c.crow[cpos];
  const bool valid = row >= 0;
  const double xc[3] = {c.cx[cpos], c.cx[c.ldp + cpos], c.cx[2 * c.ldp + cpos]};
  const int e0 = blockIdx.y * P1_ECHUNK, e1 = min(e0 + P1_ECHUNK, g.n_elem);
  const unsigned char* pl = plan + (size_t)g.slot0 * c.ldp + cpos;
  unsigned char m_next = (valid && e0 < e1) ? pl[(size_t)e0 * c.ldp] : PLAN_NONE;
  // one batch: lane integrates the pair (collocation point of lane src, element e0 + offset); the point need not be the lane's own: b goes out with atomics
  auto batch = [&](int sset, unsigned short ent, bool act) {
    const int src = ent & 31, el = e0 + (ent >> 5);
    const double xs[3] = {__shfl_sync(0xffffffffu, xc[0], src), __shfl_sync(0xffffffffu, xc[1], src), __shfl_sync(0xffffffffu, xc[2], src)};
    const int rs = __shfl_sync(0xffffffffu, row, src);
    if (act) {
      const int ngp = g.ngp[sset];
      const double* P = g.pts[sset] + (size_t)el * ngp * RECN;
      PAcc<NN> acc; acc.zero();
#pragma unroll 1
      for (int kp = 0; kp < ngp; kp++) {
        const double* q = P + (size_t)kp * RECN;
        const double x[3] = {__ldg(q), __ldg(q + 1), __ldg(q + 2)}, n[3] = {__ldg(q + 3), __ldg(q + 4), __ldg(q + 5)};
        double w[NN];
#pragma unroll
        for (int j = 0; j < NN; j++) w[j] = __ldg(q + 6 + j);
        pot_accumulate<NN>(acc, c_pp, x, n, xs, w);
      }
      const bool rev = g.erev[el] != 0, neg = (g.einfo[el] & 32u) != 0;
      const int* ecol = g.ecol + (size_t)el * NN;
      const unsigned char* ekind = g.ekind + (size_t)el * NN;
      const double* ecv = g.ecv + (size_t)el * 2 * NN;
      double bre = 0.0, bim = 0.0;
#pragma unroll
      for (int j = 0; j < NN; j++)
        pot_scatter_node(acc.hr[j], acc.hi[j], acc.gr[j], acc.gi[j], rev, __ldg(ecol + j), ekind[j], ecv[2 * j], ecv[2 * j + 1], s, rs, bre, bim, neg,
                         g.einc ? g.einc + 4 * ((size_t)el * NN + j) : nullptr);
      if (bre != 0.0 || bim != 0.0) { atomicAdd(s.bre + rs, bre); atomicAdd(s.bim + rs, bim); }
    }
    __syncwarp();
  };
  int e = e0, drain = 0;
  unsigned fullsets = 0u;      // rules whose queue holds >= 32 entries
  for (;;) {                   // one call site of the batch body (instruction cache)
    int sset; unsigned short ent; bool act;
    if (fullsets) {
      sset = __ffs(fullsets) - 1;
      int cnt = qcnt[sset];
      ent = queue[sset * P1_QCAP + cnt - 32 + lane];
      __syncwarp();
      cnt -= 32;
      if (lane == 0) qcnt[sset] = cnt;
      __syncwarp();
      if (cnt < 32) fullsets &= ~(1u << sset);
      act = true;
    } else if (e < e1) {
      const unsigned char m = m_next;
      m_next = (valid && e + 1 < e1) ? pl[(size_t)(e + 1) * c.ldp] : PLAN_NONE;   // requested one element ahead
      unsigned todo = __ballot_sync(0xffffffffu, m < MAX_SETS);
      while (todo) {
        const int leader = __ffs(todo) - 1;
        const int qs = __shfl_sync(0xffffffffu, (int)m, leader);
        const unsigned grp = __ballot_sync(0xffffffffu, (int)m == qs) & todo;
        todo &= ~grp;
        const int base = qcnt[qs];
        if ((grp >> lane) & 1u) queue[qs * P1_QCAP + base + __popc(grp & lt_mask)] = (unsigned short)(((e - e0) << 5) | lane);
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
      ent = act ? queue[sset * P1_QCAP + lane] : (unsigned short)0;
    } else break;
    batch(sset, ent, act);
  }
}

void launch_pot_regular(const DevGroup& g, const DevColloc& c, const DevSystem& s, const unsigned char* plan, cudaStream_t st) {
  if (g.n_elem == 0) return;
  dim3 grid((c.n_tiles + P1_WARPS - 1) / P1_WARPS, (g.n_elem + P1_ECHUNK - 1) / P1_ECHUNK), block(P1_WARPS * 32);
  switch (g.et) {
    case 5: k_pot_regular<5><<<grid, block, 0, st>>>(g, c, s, plan); break;
    case 6: k_pot_regular<6><<<grid, block, 0, st>>>(g, c, s, plan); break;
    case 7: k_pot_regular<7><<<grid, block, 0, st>>>(g, c, s, plan); break;
    case 8: k_pot_regular<8><<<grid, block, 0, st>>>(g, c, s, plan); break;
    case 9: k_pot_regular<9><<<grid, block, 0, st>>>(g, c, s, plan); break;
  }
}

template <int NN>
__device__ __forceinline__ void pot_warp_reduce(PAcc<NN>& a) {
#pragma unroll
  for (int i = 0; i < NN; i++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a.hr[i] += __shfl_xor_sync(0xffffffffu, a.hr[i], o); a.hi[i] += __shfl_xor_sync(0xffffffffu, a.hi[i], o);
      a.gr[i] += __shfl_xor_sync(0xffffffffu, a.gr[i], o); a.gi[i] += __shfl_xor_sync(0xffffffffu, a.gi[i], o);
    }
  }
}
// after the warp reduction lane j (< NN) scatters node j
template <int NN>
__device__ __forceinline__ void pot_scatter_pair(const PAcc<NN>& a, const DevGroup& g, int e, const DevSystem& s, int row, int lane) {
  double bre = 0.0, bim = 0.0;
  const int* ecol = g.ecol + (size_t)e * NN;
  const unsigned char* ekind = g.ekind + (size_t)e * NN;
  const double* ecv = g.ecv + (size_t)e * 2 * NN;
  const bool rev = g.erev[e] != 0;
#pragma unroll
  for (int j = 0; j < NN; j++)
    if (lane == j) pot_scatter_node(a.hr[j], a.hi[j], a.gr[j], a.gi[j], rev, ecol[j], ekind[j], ecv[2 * j], ecv[2 * j + 1], s, row, bre, bim, (g.einfo[e] & 32u) != 0,
                                     g.einc ? g.einc + 4 * ((size_t)e * NN + j) : nullptr);
  if (bre != 0.0 || bim != 0.0) { atomicAdd(s.bre + row, bre); atomicAdd(s.bim + row, bim); }
}

// ------------------------------------------------------------------------------------------------------------------
// P2: adaptive pairs -- one warp per pair, lanes stride over the gln x gln points of every leaf
// ------------------------------------------------------------------------------------------------------------------
template <int ET>
__global__ void __launch_bounds__(128) k_pot_adaptive(DevGroup g, DevColloc c, DevSystem s, DevAdaptive a, DevTables t) {
  constexpr int NN = ElemTraits<ET>::NN;
  constexpr bool tri = (ElemTraits<ET>::NV == 3);
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= a.n_pairs) return;
  const int cpos = a.pair_cpos[p], e = a.pair_elem[p];
  if (c.tile_active && !c.tile_active[cpos >> 5]) return;
  const double xc[3] = {c.cx[cpos], c.cx[c.ldp + cpos], c.cx[2 * c.ldp + cpos]};
  const int row = c.crow[cpos];
  double xn[3 * NN];
#pragma unroll
  for (int i = 0; i < 3 * NN; i++) xn[i] = g.xn[(size_t)e * 3 * NN + i];
  const double* gx = tri ? t.gl01_x : t.gl11_x;
  const double* gw = tri ? t.gl01_w : t.gl11_w;
  PAcc<NN> acc; acc.zero();
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
      pot_accumulate<NN>(acc, c_pp, x, n, xc, w);
    }
  }
  pot_warp_reduce<NN>(acc);
  pot_scatter_pair<NN>(acc, g, e, s, row, lane);
}
void launch_pot_adaptive(const DevGroup& g, const DevColloc& c, const DevSystem& s, const DevAdaptive& a, const DevTables& t, cudaStream_t st) {
  if (a.n_pairs == 0) return;
  dim3 grid((a.n_pairs + 3) / 4), block(128);
  switch (g.et) {
    case 5: k_pot_adaptive<5><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 6: k_pot_adaptive<6><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 7: k_pot_adaptive<7><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 8: k_pot_adaptive<8><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 9: k_pot_adaptive<9><<<grid, block, 0, st>>>(g, c, s, a, t); break;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// P3: singular pairs -- one warp per pair, lanes stride over (ray, radial point); 15 radial Gauss-Legendre points on [0,1]
// per ray (ngp_rho = 15, bem_harpot3d.f90:866), jw = J * rho * (jthetap * w_angular) * w_radial (:918)
// ------------------------------------------------------------------------------------------------------------------
template <int ET>
__global__ void __launch_bounds__(128) k_pot_singular(DevGroup g, DevColloc c, DevSystem s, DevSingular a, DevTables t) {
  constexpr int NN = ElemTraits<ET>::NN;
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= a.n_pairs) return;
  const int cpos = a.pair_cpos[p], e = a.pair_elem[p];
  if (c.tile_active && !c.tile_active[cpos >> 5]) return;
  const int row = c.crow[cpos];
  const double* D = a.pair_d + 14 * (size_t)p;
  const double xi_i0 = D[0], xi_i1 = D[1];
  const double xc[3] = {D[2], D[3], D[4]};
  double xn[3 * NN];
#pragma unroll
  for (int i = 0; i < 3 * NN; i++) xn[i] = g.xn[(size_t)e * 3 * NN + i];
  const int ray0 = a.pair_ray0[p], nray = a.pair_ray0[p + 1] - ray0;
  const double* gx = t.gl01_x + 15 * 14 / 2;
  const double* gw = t.gl01_w + 15 * 14 / 2;
  PAcc<NN> acc; acc.zero();
#pragma unroll 1
  for (int idx = lane; idx < nray * 15; idx += 32) {
    const int kr_ = idx / 15, kk = idx - kr_ * 15;
    const double* R = a.rays + 4 * (size_t)(ray0 + kr_);
    const double ct = __ldg(R), sn = __ldg(R + 1), rhoij = __ldg(R + 2), wray = __ldg(R + 3);
    const double rho = rhoij * __ldg(gx + kk), wrad = __ldg(gw + kk);
    double phi[NN], x[3], n[3], jg;
    geometry_at<ET>(xn, xi_i0 + rho * ct, xi_i1 + rho * sn, phi, x, n, jg);
    const double jw = jg * rho * wray * wrad;
    double w[NN];
#pragma unroll
    for (int j = 0; j < NN; j++) w[j] = phi[j] * jw;
    pot_accumulate<NN>(acc, c_pp, x, n, xc, w);
  }
  pot_warp_reduce<NN>(acc);
  pot_scatter_pair<NN>(acc, g, e, s, row, lane);
}
void launch_pot_singular(const DevGroup& g, const DevColloc& c, const DevSystem& s, const DevSingular& a, const DevTables& t, cudaStream_t st) {
  if (a.n_pairs == 0) return;
  dim3 grid((a.n_pairs + 3) / 4), block(128);
  switch (g.et) {
    case 5: k_pot_singular<5><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 6: k_pot_singular<6><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 7: k_pot_singular<7><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 8: k_pot_singular<8><<<grid, block, 0, st>>>(g, c, s, a, t); break;
    case 9: k_pot_singular<9><<<grid, block, 0, st>>>(g, c, s, a, t); break;
  }
}

}  // namespace mfbd
