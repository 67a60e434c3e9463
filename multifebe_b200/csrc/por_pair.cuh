// por_pair.cuh -- the per-point and per-pair bodies of the poroelastic kernels (poro.cu), host+device inline so that the SAME code runs
// lane-serially on the host against the CPU oracle (tests/native/por_pair_host.cpp, tests/test_por_pair_host.py): what a lane does at one
// integration point of the regular (R1), adaptive (R2) and singular (R3) kernels, the line-integral terms of R3 and the constants and
// orientation of a finished pair (`h = cte_t h, g = cte_u g; if (reverse) h = -h`, bem_harpor3d.f90:1003-1010).  The kernels add the
// lane mapping, the warp reduction and the scatter.
#pragma once
#include "por_math.cuh"

namespace mfbd {

// accumulators of one equation l of one pair: h(j, l, k), g(j, l, k), k = 0..3
template <int NN>
struct RAcc {
  double hr[4 * NN], hi[4 * NN], gr[4 * NN], gi[4 * NN];     // [k * NN + j]
  MFB_HD void zero() {
#pragma unroll
    for (int i = 0; i < 4 * NN; i++) { hr[i] = 0.0; hi[i] = 0.0; gr[i] = 0.0; gi[i] = 0.0; }
  }
};

// row l of the 4 x 4 blocks, selected with compile-time indices only
MFB_HD void por_row(const cplx f[4][4], int l, cplx out[4]) {
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = (l == 0) ? f[0][k] : (l == 1 ? f[1][k] : (l == 2 ? f[2][k] : f[3][k]));
}

template <int NN>
MFB_HD void por_accumulate_row(RAcc<NN>& a, const cplx fu[4][4], const cplx ft[4][4], int l, const double* w) {
  cplx ur[4], tr[4];
  por_row(fu, l, ur); por_row(ft, l, tr);
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int j = 0; j < NN; j++) {
      a.hr[k * NN + j] = fma(tr[k].re, w[j], a.hr[k * NN + j]); a.hi[k * NN + j] = fma(tr[k].im, w[j], a.hi[k * NN + j]);
      a.gr[k * NN + j] = fma(ur[k].re, w[j], a.gr[k * NN + j]); a.gi[k * NN + j] = fma(ur[k].im, w[j], a.gi[k * NN + j]);
    }
}

// R1: one record (x[3], n[3], phi_j J w [NN]) of a precalculated point set (fbem_bem_harpor3d_sbie_ext_pre)
template <int NN>
MFB_HD void por_regular_point(RAcc<NN>& acc, const PorParams& P, const double* x, const double* n, const double* w, const double* xc, int l) {
  cplx fu[4][4], ft[4][4];
  por_exterior_blocks(P, x, n, xc, fu, ft);
  por_accumulate_row<NN>(acc, fu, ft, l, w);
}

// R2: one Gauss point (k1, k2) of one leaf (fbem_bem_harpor3d_sbie_ext_st)
template <int ET>
MFB_HD void por_leaf_point(RAcc<ElemTraits<ET>::NN>& acc, const PorParams& P, const double* xn, const double* xi_s, const double* tp1, const double* tp2,
                           double g1, double w1, double g2, double w2, const double* xc, int l) {
  constexpr int NN = ElemTraits<ET>::NN;
  double x[3], n[3], w[NN];
  leaf_point<ET>(xn, xi_s, tp1, tp2, g1, w1, g2, w2, x, n, w);
  por_regular_point<NN>(acc, P, x, n, w, xc, l);
}

// R3: one radial point of one ray (fbem_bem_harpor3d_sbie_int): weakly singular parts against phi_j, the CPV kernel of the skeleton block
// against phi_j - phi_j(xi_i)
template <int ET>
MFB_HD void por_singular_point(RAcc<ElemTraits<ET>::NN>& acc, const PorParams& P, const double* xn, double xi_i0, double xi_i1, const double* phi_i,
                               double ct, double sn, double rho, double wray, double wrad, const double* xc, int l) {
  constexpr int NN = ElemTraits<ET>::NN;
  double phi[NN], x[3], n[3], jg;
  geometry_at<ET>(xn, xi_i0 + rho * ct, xi_i1 + rho * sn, phi, x, n, jg);
  const double jw = jg * rho * wray * wrad;
  double w[NN];
#pragma unroll
  for (int j = 0; j < NN; j++) w[j] = phi[j] * jw;
  cplx fu[4][4], ft[4][4], fc[3][3];
  por_interior_blocks(P, x, n, xc, fu, ft, fc);
  por_accumulate_row<NN>(acc, fu, ft, l, w);
  if (l > 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const cplx f = (l == 1) ? fc[0][k] : (l == 2 ? fc[1][k] : fc[2][k]);
#pragma unroll
      for (int j = 0; j < NN; j++) {
        const double wc = (phi[j] - phi_i[j]) * jw;
        acc.hr[(k + 1) * NN + j] = fma(f.re, wc, acc.hr[(k + 1) * NN + j]); acc.hi[(k + 1) * NN + j] = fma(f.im, wc, acc.hi[(k + 1) * NN + j]);
      }
    }
  }
}

// R3 after the reduction: + phi_j(xi_i) T2(1) hli(l, k) on the skeleton block (bem_harpor3d.f90:1876-1880); hli = D[5..13], [l][k]
template <int NN>
MFB_HD void por_singular_line_terms(RAcc<NN>& acc, const PorParams& P, const double* phi_i, const double* hli, int l) {
  if (l == 0) return;
  const cplx t21 = P.T2[1];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double hl = hli[3 * (l - 1) + k];
#pragma unroll
    for (int j = 0; j < NN; j++) { acc.hr[(k + 1) * NN + j] += phi_i[j] * t21.re * hl; acc.hi[(k + 1) * NN + j] += phi_i[j] * t21.im * hl; }
  }
}

// the finished entry (j, l, k): h scaled by cte_t(l, k) with the sign of the orientation, g by cte_u(l, k)
template <int NN>
MFB_HD void por_finished_entry(const RAcc<NN>& a, const PorParams& P, int l, int k, int j, bool rev, double& hr, double& hi, double& gr, double& gi) {
  const cplx ct0 = (l == 0) ? P.cte_t[0][k] : P.cte_t[1][k];     // cte_t(l, k) depends on l only through l == 0
  const cplx cu = (l == 0) ? P.cte_u[0][k] : P.cte_u[1][k];
  const cplx ct = rev ? mk(-ct0.re, -ct0.im) : ct0;
  const int q = k * NN + j;
  hr = ct.re * a.hr[q] - ct.im * a.hi[q]; hi = ct.re * a.hi[q] + ct.im * a.hr[q];
  gr = cu.re * a.gr[q] - cu.im * a.gi[q]; gi = cu.re * a.gi[q] + cu.im * a.gr[q];
}

}  // namespace mfbd
