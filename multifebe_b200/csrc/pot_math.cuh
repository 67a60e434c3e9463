// pot_math.cuh -- per-Gauss-point arithmetic of the scalar-wave (inviscid fluid / acoustic) SBIE kernels (host+device inline, like
// bem_math.cuh; the host build is used by tests/test_pot_math_host.py to check this arithmetic against the oracle without a GPU).
//
// Regularised form of the reference (lib/fbem/src/bem_harpot3d.f90:313-318): with z = -i k r and E_m(z) = e^z - sum_{j<m} z^j/j!
//   fs_P = 1/r + P(1) + E_2/r                      (p* = fs_P / 4 pi)
//   fs_Q = 1/r^2 + Q(1) + Q(2) E_2/r + E_3/r^2     (q* = -fs_Q dr/dn / 4 pi)
#pragma once
#include "bem_math.cuh"

namespace mfbd {

// fbem_bem_harpot3d_parameters (lib/fbem/src/bem_harpot3d.f90:102-115, SBIE subset) with the constants of the assembly folded
// in: c4pi = 1/(4 pi) (`h=-h*c_1_4pi; g=g*c_1_4pi`, :322-323) and d1J = rho omega^2, the factor that turns the flux dp/dn into
// the reference's unknown Un (src/build_lse_mechanics_bem_harpot.f90:751, :1104).
struct PotParams {
  cplx k;          // wavenumber omega / c
  cplx P1;         // -i k
  cplx Q1, Q2;     // k^2 / 2, i k
  double c4pi, d1J;
};

MFB_HD double mfb_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}

// E_2(z), E_3(z) of z = -i k r.  The reference switches to the power series for |z| <= 1 (lib/fbem/src/numerical.f90:1258-1330) to
// keep E_m accurate RELATIVE TO ITSELF; what reaches the matrix is 1/r + E_2/r and 1/r^2 + E_3/r^2 (+ bounded terms), where the
// static 1 dominates E_m = O(|z|^m/m!) for |z| <= 1, so the direct subtraction e^z - 1 - z (- z^2/2), whose absolute error is a
// few ulp of 1, is as accurate for the kernel as the series is.  The series is kept only where it is also the cheaper branch:
// |z| <= 0.1, E_3 = z^3/3! (1 + z/4 (1 + z/5 (...))) with 9 terms (first neglected term |z|^9 3!/12! < 2e-17); otherwise the
// branch-free exp / sincos of bem_math.cuh (45 FP64 instructions against 7 per series term).
MFB_HD void pot_E23(cplx z, cplx& E2, cplx& E3) {
  const cplx z2 = z * z;
  if (z.re * z.re + z.im * z.im <= 0.01) {
    double tr = 1.0, ti = 0.0;
#pragma unroll
    for (int m = 12; m >= 4; m--) {
      const double inv = 1.0 / (double)m;
      const double ur = (z.re * tr - z.im * ti) * inv, ui = (z.re * ti + z.im * tr) * inv;
      tr = 1.0 + ur; ti = ui;
    }
    const cplx z3 = z2 * z;
    E3 = (z3 * mk(tr, ti)) * (1.0 / 6.0);
    E2 = cfmar(z2, 0.5, E3);
  } else {
    const double ex = mfb_exp(z.re); double sn, cs;
    mfb_sincos(z.im, sn, cs);
    const cplx E1 = mk(fma(ex, cs, -1.0), ex * sn);
    E2 = E1 - z;
    E3 = cfmar(z2, -0.5, E2);
  }
}

MFB_HD void pot_scalars(const PotParams& p, double r, double d1r1, cplx& fP, cplx& fQ) {
  const double d1r2 = d1r1 * d1r1;
  const cplx z = mk(p.k.im * r, -p.k.re * r);
  cplx E2, E3; pot_E23(z, E2, E3);
  const cplx e2 = E2 * d1r1, e3 = E3 * d1r2;
  fP = mk(d1r1 + p.P1.re + e2.re, p.P1.im + e2.im);
  fQ = cfma(p.Q2, e2, mk(d1r2 + p.Q1.re + e3.re, p.Q1.im + e3.im));
}

// accumulators of one (collocation point, element) pair: raw sums of fs_Q dr/dn phi_j J w (h) and fs_P phi_j J w (g)
template <int NN>
struct PAcc {
  double hr[NN], hi[NN], gr[NN], gi[NN];
  MFB_HD void zero() {
#pragma unroll
    for (int i = 0; i < NN; i++) { hr[i] = 0.0; hi[i] = 0.0; gr[i] = 0.0; gi[i] = 0.0; }
  }
};

// one quadrature point: x, n = point and unit normal, xc = collocation point, w[j] = phi_j * J * weight
template <int NN>
MFB_HD void pot_accumulate(PAcc<NN>& a, const PotParams& p, const double* x, const double* n, const double* xc, const double* w) {
  const double rv0 = x[0] - xc[0], rv1 = x[1] - xc[1], rv2 = x[2] - xc[2];
  const double r2 = fma(rv0, rv0, fma(rv1, rv1, rv2 * rv2));
  const double d1r1 = mfb_rsqrt(r2), r = r2 * d1r1;
  const double drdn = fma(rv0, n[0], fma(rv1, n[1], rv2 * n[2])) * d1r1;
  cplx fP, fQ; pot_scalars(p, r, d1r1, fP, fQ);
  const double qr = fQ.re * drdn, qi = fQ.im * drdn;
#pragma unroll
  for (int j = 0; j < NN; j++) {
    a.hr[j] = fma(qr, w[j], a.hr[j]); a.hi[j] = fma(qi, w[j], a.hi[j]);
    a.gr[j] = fma(fP.re, w[j], a.gr[j]); a.gi[j] = fma(fP.im, w[j], a.gi[j]);
  }
}

}  // namespace mfbd
