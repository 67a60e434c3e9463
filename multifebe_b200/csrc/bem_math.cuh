// bem_math.cuh -- per-Gauss-point arithmetic of the elastodynamic SBIE kernels (host+device inline).
//
// Evaluates the reference's *regularised* form of the full-space fundamental solution (lib/fbem/src/bem_harela3d.f90:
// 175-218 coefficients, :663-693 combination): psi, chi, T1, T2, T3 as static part + constant + sum coeff*E_m(z_j)/r^(m-1),
// E_m(z) = e^z - sum_{j<m} z^j/j!, z_j = -i k_j r.  E_m is formed GPU-style (branch on |z|<=1 like
// lib/fbem/src/numerical.f90:1258-1330, but a fixed-length Horner series instead of the reference's data-dependent term
// count, and only m = 2..5).  Everything here is continuous arithmetic (FMA contraction allowed); discrete decisions are
// taken on the host (plan_host.cpp) or by threshold comparison (classify kernel).
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define MFB_HD __host__ __device__ __forceinline__
#else
#define MFB_HD inline
#endif

namespace mfbd {

struct cplx { double re, im; };
MFB_HD cplx mk(double a, double b) { cplx c; c.re = a; c.im = b; return c; }
MFB_HD cplx operator+(cplx a, cplx b) { return mk(a.re + b.re, a.im + b.im); }
MFB_HD cplx operator-(cplx a, cplx b) { return mk(a.re - b.re, a.im - b.im); }
MFB_HD cplx operator*(cplx a, cplx b) { return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
MFB_HD cplx operator*(cplx a, double s) { return mk(a.re * s, a.im * s); }
MFB_HD cplx operator*(double s, cplx a) { return mk(a.re * s, a.im * s); }
MFB_HD cplx cfma(cplx a, cplx b, cplx c) { return mk(c.re + a.re * b.re - a.im * b.im, c.im + a.re * b.im + a.im * b.re); }  // a*b+c
MFB_HD cplx cfmar(cplx a, double s, cplx c) { return mk(c.re + a.re * s, c.im + a.im * s); }                                   // a*s+c

// Region/frequency parameters (fbem_bem_harela3d_parameters, SBIE subset), 1-based like the reference.
struct KParams {
  cplx k1, k2;
  cplx psi[7], chi[7], T1[11], T2[10], T3[10];
  cplx cte_u; double cte_t;
  cplx S1[12], S2[12], S3[13], S4[11], S5[12];   // hypersingular kernels d*, s* (bem_harela3d.f90:219-280)
  cplx cte_s; double cte_d;
};

// E_2..E_5 of z (returned already divided: e2 = E2/r, e3 = E3/r^2, e4 = E4/r^3, e5 = E5/r^4)
struct EnR { cplx e2, e3, e4, e5; };

// ---- branch-free exp / sincos for the kernel arguments ---------------------------------------------------------------
// z = -i k r has Re z = Im(k) r <= 0 (damped medium) and |Im z| = Re(k) r bounded by (wavenumber x model size), far inside
// the range where a three-constant Cody-Waite reduction is exact.  No slow path, no special cases: straight-line code
// that the scheduler can interleave for z1 and z2 (the library routines cost a call and several branches each).
MFB_HD double mfb_round_magic(double x, int& q) {   // nearest integer of |x| < 2^31 and its int value
  const double MAGIC = 6755399441055744.0;         // 1.5 * 2^52
  double t = x + MAGIC;
#if defined(__CUDA_ARCH__)
  q = __double2loint(t);
#else
  long long bits; memcpy(&bits, &t, 8); q = (int)(bits & 0xffffffffll);
#endif
  return t - MAGIC;
}
MFB_HD double mfb_exp(double a) {   // e^a, a in [-700, 700]; Taylor degree 13 on |f| <= ln2/2 (remainder 4e-18)
  a = fmax(a, -700.0);
  int n; const double nd = mfb_round_magic(a * 1.4426950408889634074, n);
  double f = fma(nd, -6.93147180369123816490e-01, a);
  f = fma(nd, -1.90821492927058770002e-10, f);
  double p = 1.0 / 6227020800.0;
  p = fma(p, f, 1.0 / 479001600.0); p = fma(p, f, 1.0 / 39916800.0); p = fma(p, f, 1.0 / 3628800.0); p = fma(p, f, 1.0 / 362880.0);
  p = fma(p, f, 1.0 / 40320.0); p = fma(p, f, 1.0 / 5040.0); p = fma(p, f, 1.0 / 720.0); p = fma(p, f, 1.0 / 120.0);
  p = fma(p, f, 1.0 / 24.0); p = fma(p, f, 1.0 / 6.0); p = fma(p, f, 0.5); p = fma(p, f, 1.0); p = fma(p, f, 1.0);
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
#else
  return ldexp(p, n);
#endif
}
MFB_HD void mfb_sincos(double b, double& sn, double& cs) {   // |b| < 1e5; kernels of fdlibm's __kernel_sin/__kernel_cos on |t| <= pi/4
  int q; const double nd = mfb_round_magic(b * 6.36619772367581382433e-01, q);
  double t = fma(nd, -1.57079632673412561417e+00, b);
  t = fma(nd, -6.07710050650619224932e-11, t);
  t = fma(nd, -2.02226624879595063154e-21, t);
  const double z = t * t;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08); ps = fma(ps, z, 2.75573137070700676789e-06); ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03); ps = fma(ps, z, -1.66666666666666324348e-01);
  const double s = fma(t * z, ps, t);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09); pc = fma(pc, z, -2.75573143513906633035e-07); pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03); pc = fma(pc, z, 4.16666666666666019037e-02);
  const double c = fma(z * z, pc, fma(z, -0.5, 1.0));
  const double a0 = (q & 1) ? c : s, a1 = (q & 1) ? s : c;
  sn = (q & 2) ? -a0 : a0;
  cs = ((q + 1) & 2) ? -a1 : a1;
}

// 1/(m+5)!, m = 0..16: E5(z) = z^5 * sum_m z^m/(m+5)!  (|z| <= 1: the first neglected term is 1/22! = 9e-22)
#define MFB_E5_COEFFS {1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0, 1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0, \
    1.0 / 479001600.0, 1.0 / 6227020800.0, 1.0 / 87178291200.0, 1.0 / 1307674368000.0, 1.0 / 20922789888000.0, \
    1.0 / 355687428096000.0, 1.0 / 6402373705728000.0, 1.0 / 121645100408832000.0, 1.0 / 2432902008176640000.0, \
    1.0 / 51090942171709440000.0}

#if defined(__CUDACC__)
static __device__ __constant__ double c_e5_coeffs[17] = MFB_E5_COEFFS;
#endif
MFB_HD double mfb_e5_coeff(int m) {
#if defined(__CUDA_ARCH__)
  return c_e5_coeffs[m];
#else
  const double c[17] = MFB_E5_COEFFS; return c[m];
#endif
}
// The series sum S(z) = sum_m z^m/(m+5)! is split into its even and odd parts, S = Se(w) + z So(w) with w = z^2: two
// independent Horner chains of 8 steps instead of one of 16 (the kernels run at low occupancy, so the length of the
// dependent chain is what this branch costs).  Real coefficients: 4 FMA-class operations per step.
MFB_HD void zexp_series(cplx z, cplx& E2, cplx& E3, cplx& E4, cplx& E5) {   // |z| <= 1
  const cplx z2 = z * z, z3 = z2 * z, z4 = z2 * z2;
  double er = mfb_e5_coeff(16), ei = 0.0, orr = 0.0, oi = 0.0;
#pragma unroll 1
  for (int m = 14; m >= 0; m -= 2) {
    const double tr = fma(er, z2.re, fma(-ei, z2.im, mfb_e5_coeff(m))), ur = fma(orr, z2.re, fma(-oi, z2.im, mfb_e5_coeff(m + 1)));
    ei = fma(er, z2.im, ei * z2.re); oi = fma(orr, z2.im, oi * z2.re);
    er = tr; orr = ur;
  }
  const cplx S = cfma(z, mk(orr, oi), mk(er, ei));
  E5 = (z4 * z) * S;
  E4 = cfmar(z4, 1.0 / 24.0, E5);
  E3 = cfmar(z3, 1.0 / 6.0, E4);
  E2 = cfmar(z2, 0.5, E3);
}
// two arguments in one rolled loop (four independent chains)
MFB_HD void zexp_series2(cplx za, cplx zb, cplx& A2, cplx& A3, cplx& A4, cplx& A5, cplx& B2, cplx& B3, cplx& B4, cplx& B5) {
  const cplx wa = za * za, wb = zb * zb;
  double aer = mfb_e5_coeff(16), aei = 0.0, aor = 0.0, aoi = 0.0, ber = aer, bei = 0.0, bor = 0.0, boi = 0.0;
#pragma unroll 1
  for (int m = 14; m >= 0; m -= 2) {
    const double ce = mfb_e5_coeff(m), co = mfb_e5_coeff(m + 1);
    const double t1 = fma(aer, wa.re, fma(-aei, wa.im, ce)), t2 = fma(aor, wa.re, fma(-aoi, wa.im, co));
    const double t3 = fma(ber, wb.re, fma(-bei, wb.im, ce)), t4 = fma(bor, wb.re, fma(-boi, wb.im, co));
    aei = fma(aer, wa.im, aei * wa.re); aoi = fma(aor, wa.im, aoi * wa.re);
    bei = fma(ber, wb.im, bei * wb.re); boi = fma(bor, wb.im, boi * wb.re);
    aer = t1; aor = t2; ber = t3; bor = t4;
  }
  { const cplx z3 = wa * za, z4 = wa * wa, S = cfma(za, mk(aor, aoi), mk(aer, aei));
    A5 = (z4 * za) * S; A4 = cfmar(z4, 1.0 / 24.0, A5); A3 = cfmar(z3, 1.0 / 6.0, A4); A2 = cfmar(wa, 0.5, A3); }
  { const cplx z3 = wb * zb, z4 = wb * wb, S = cfma(zb, mk(bor, boi), mk(ber, bei));
    B5 = (z4 * zb) * S; B4 = cfmar(z4, 1.0 / 24.0, B5); B3 = cfmar(z3, 1.0 / 6.0, B4); B2 = cfmar(wb, 0.5, B3); }
}
MFB_HD void zexp_direct(cplx z, cplx& E2, cplx& E3, cplx& E4, cplx& E5) {   // |z| > 1
  const cplx z2 = z * z, z3 = z2 * z, z4 = z2 * z2;
  double ex = mfb_exp(z.re), sn, cs;
  mfb_sincos(z.im, sn, cs);
  cplx E1 = mk(fma(ex, cs, -1.0), ex * sn);
  E2 = E1 - z;
  E3 = cfmar(z2, -0.5, E2);
  E4 = cfmar(z3, -1.0 / 6.0, E3);
  E5 = cfmar(z4, -1.0 / 24.0, E4);
}
MFB_HD void zexp_E2_5(cplx z, cplx& E2, cplx& E3, cplx& E4, cplx& E5) {
  if (z.re * z.re + z.im * z.im <= 1.0) zexp_series(z, E2, E3, E4, E5); else zexp_direct(z, E2, E3, E4, E5);
}
// both arguments at once: the common cases (both direct / both series) are straight-line code for the two arguments
// together, which the scheduler interleaves (the polynomial chains of one hide the latency of the other)
MFB_HD void zexp_pair(cplx z1, cplx z2, cplx& A2, cplx& A3, cplx& A4, cplx& A5, cplx& B2, cplx& B3, cplx& B4, cplx& B5) {
  const bool s1 = z1.re * z1.re + z1.im * z1.im <= 1.0, s2 = z2.re * z2.re + z2.im * z2.im <= 1.0;
  // |z1| <= |z2| on this path (k1 is the P wavenumber), but nothing here relies on it
  if (!s1 && !s2) { zexp_direct(z1, A2, A3, A4, A5); zexp_direct(z2, B2, B3, B4, B5); }
  else if (s1 && s2) zexp_series2(z1, z2, A2, A3, A4, A5, B2, B3, B4, B5);
  else { zexp_E2_5(z1, A2, A3, A4, A5); zexp_E2_5(z2, B2, B3, B4, B5); }
}

// E_2..E_6 for the hypersingular kernels (not a hot path: general form, both branches of numerical.f90:1258-1330)
MFB_HD void zexp_E2_6(cplx z, cplx& E2, cplx& E3, cplx& E4, cplx& E5, cplx& E6) {
  const cplx z2 = z * z, z3 = z2 * z, z4 = z2 * z2, z5 = z4 * z;
  if (z.re * z.re + z.im * z.im <= 1.0) {
    double c[18]; c[0] = 1.0 / 720.0;
    for (int m = 1; m < 18; m++) c[m] = c[m - 1] / (double)(m + 6);
    double sr = c[17], si = 0.0;
    for (int m = 16; m >= 0; m--) { const double tr = fma(sr, z.re, fma(-si, z.im, c[m])); si = fma(sr, z.im, si * z.re); sr = tr; }
    E6 = (z5 * z) * mk(sr, si);
    E5 = cfmar(z5, 1.0 / 120.0, E6); E4 = cfmar(z4, 1.0 / 24.0, E5); E3 = cfmar(z3, 1.0 / 6.0, E4); E2 = cfmar(z2, 0.5, E3);
  } else {
    zexp_direct(z, E2, E3, E4, E5);
    E6 = cfmar(z5, -1.0 / 120.0, E5);
  }
}

struct KScal { cplx psi, chi, T1, T2, T3; double d1r1, d1r2; };

// regular_only: drop the static 1/r^2 parts of T1..T3 (interior integration, bem_harela3d.f90:1394-1401)
template <bool REGULAR_ONLY>
MFB_HD void kernel_scalars(const KParams& p, double r, KScal& k) {
  double d1r1 = 1.0 / r, d1r2 = d1r1 * d1r1, d1r3 = d1r2 * d1r1, d1r4 = d1r2 * d1r2;
  // z = -i k r
  cplx z1 = mk(p.k1.im * r, -p.k1.re * r), z2 = mk(p.k2.im * r, -p.k2.re * r);
  cplx A2, A3, A4, A5, B2, B3, B4, B5;
  zexp_E2_5(z1, A2, A3, A4, A5);
  zexp_E2_5(z2, B2, B3, B4, B5);
  cplx E21 = A2 * d1r1, E22 = B2 * d1r1, E31 = A3 * d1r2, E32 = B3 * d1r2;
  cplx E41 = A4 * d1r3, E42 = B4 * d1r3, E51 = A5 * d1r4, E52 = B5 * d1r4;
  cplx t;
  t = p.psi[1] * d1r1 + p.psi[2] + E22;
  t = cfma(p.psi[3], E31, t); t = cfma(p.psi[4], E32, t); t = cfma(p.psi[5], E41, t); t = cfma(p.psi[6], E42, t);
  k.psi = t;
  t = p.chi[1] * d1r1 + E22;
  t = cfma(p.chi[2], E21, t); t = cfma(p.chi[3], E31, t); t = cfma(p.chi[4], E32, t); t = cfma(p.chi[5], E41, t); t = cfma(p.chi[6], E42, t);
  k.chi = t;
  t = REGULAR_ONLY ? p.T1[2] : (p.T1[1] * d1r2 + p.T1[2]);
  t = cfma(p.T1[3], E21, t); t = cfma(p.T1[4], E22, t); t = cfma(p.T1[5], E31, t); t = cfma(p.T1[6], E32, t);
  t = cfma(p.T1[7], E41, t); t = cfma(p.T1[8], E42, t); t = cfma(p.T1[9], E51, t); t = cfma(p.T1[10], E52, t);
  k.T1 = t;
  t = REGULAR_ONLY ? p.T2[2] : (p.T2[1] * d1r2 + p.T2[2]);
  t = cfma(p.T2[3], E22, t); t = cfma(p.T2[4], E31, t); t = cfma(p.T2[5], E32, t); t = cfma(p.T2[6], E41, t);
  t = cfma(p.T2[7], E42, t); t = cfma(p.T2[8], E51, t); t = cfma(p.T2[9], E52, t);
  k.T2 = t;
  t = REGULAR_ONLY ? p.T3[2] : (p.T3[1] * d1r2 + p.T3[2]);
  t = cfma(p.T3[3], E21, t); t = cfma(p.T3[4], E31, t); t = cfma(p.T3[5], E32, t); t = cfma(p.T3[6], E41, t);
  t = cfma(p.T3[7], E42, t); t = cfma(p.T3[8], E51, t); t = cfma(p.T3[9], E52, t);
  k.T3 = t;
  k.d1r1 = d1r1; k.d1r2 = d1r2;
}

// Kernel scalars for the regular (K1) kernel with PRE-SCALED parameters: q.psi/q.chi carry the factor -cte_u (their slot 0 is
// the coefficient of the bare E2(z2)/r term), q.T1..T3 carry cte_t, so that psi*delta - chi*r,l*r,k is directly the
// matrix contribution -g/(4 pi mu) of a u-known column and the T combination directly +h/(4 pi) of a t-known column.
// Takes r and 1/r (from rsqrt) instead of dividing.
// need_g / need_h: form psi, chi (displacement kernel) / T1..T3 (traction kernel) only when a column of the element uses them.
MFB_HD void kernel_scalars_scaled(const KParams& p, double r, double d1r1, KScal& k, bool need_g = true, bool need_h = true) {
  const double d1r2 = d1r1 * d1r1, d1r3 = d1r2 * d1r1, d1r4 = d1r2 * d1r2;
  const cplx z1 = mk(p.k1.im * r, -p.k1.re * r), z2 = mk(p.k2.im * r, -p.k2.re * r);
  cplx A2, A3, A4, A5, B2, B3, B4, B5;
  zexp_pair(z1, z2, A2, A3, A4, A5, B2, B3, B4, B5);
  const cplx E21 = A2 * d1r1, E22 = B2 * d1r1, E31 = A3 * d1r2, E32 = B3 * d1r2;
  const cplx E41 = A4 * d1r3, E42 = B4 * d1r3, E51 = A5 * d1r4, E52 = B5 * d1r4;
  cplx t;
  k.psi = mk(0.0, 0.0); k.chi = k.psi; k.T1 = k.psi; k.T2 = k.psi; k.T3 = k.psi;
  if (need_g) {
    t = cfmar(p.psi[1], d1r1, p.psi[2]);
    t = cfma(p.psi[0], E22, t); t = cfma(p.psi[3], E31, t); t = cfma(p.psi[4], E32, t); t = cfma(p.psi[5], E41, t); t = cfma(p.psi[6], E42, t);
    k.psi = t;
    t = p.chi[1] * d1r1;
    t = cfma(p.chi[0], E22, t); t = cfma(p.chi[2], E21, t); t = cfma(p.chi[3], E31, t); t = cfma(p.chi[4], E32, t); t = cfma(p.chi[5], E41, t); t = cfma(p.chi[6], E42, t);
    k.chi = t;
  }
  if (need_h) {
    t = cfmar(p.T1[1], d1r2, p.T1[2]);
    t = cfma(p.T1[3], E21, t); t = cfma(p.T1[4], E22, t); t = cfma(p.T1[5], E31, t); t = cfma(p.T1[6], E32, t);
    t = cfma(p.T1[7], E41, t); t = cfma(p.T1[8], E42, t); t = cfma(p.T1[9], E51, t); t = cfma(p.T1[10], E52, t);
    k.T1 = t;
    t = cfmar(p.T2[1], d1r2, p.T2[2]);
    t = cfma(p.T2[3], E22, t); t = cfma(p.T2[4], E31, t); t = cfma(p.T2[5], E32, t); t = cfma(p.T2[6], E41, t);
    t = cfma(p.T2[7], E42, t); t = cfma(p.T2[8], E51, t); t = cfma(p.T2[9], E52, t);
    k.T2 = t;
    t = cfmar(p.T3[1], d1r2, p.T3[2]);
    t = cfma(p.T3[3], E21, t); t = cfma(p.T3[4], E31, t); t = cfma(p.T3[5], E32, t); t = cfma(p.T3[6], E41, t);
    t = cfma(p.T3[7], E42, t); t = cfma(p.T3[8], E51, t); t = cfma(p.T3[9], E52, t);
    k.T3 = t;
  }
  k.d1r1 = d1r1; k.d1r2 = d1r2;
}

// Accumulators of one (collocation point, element) pair: h,g[l][k][j] for NL load directions l (3 = all, 1 = one
// direction per pass for elements with many nodes), split re/im so that every index is a compile-time constant.
template <int NN, int NL>
struct Acc {
  double hr[NL * 3 * NN], hi[NL * 3 * NN], gr[NL * 3 * NN], gi[NL * 3 * NN];
  MFB_HD void zero() {
#pragma unroll
    for (int i = 0; i < NL * 3 * NN; i++) { hr[i] = 0.0; hi[i] = 0.0; gr[i] = 0.0; gi[i] = 0.0; }
  }
};

// Exterior point (bem_harela3d.f90:663-693): x,n = integration point and unit normal, xc = collocation point,
// w[j] = phi_j*J*weight.  NL==3: all load directions; NL==1: only direction `il` (0..2).
template <int NN, int NL>
MFB_HD void accumulate_exterior(Acc<NN, NL>& a, const KParams& p, const double* x, const double* n, const double* xc,
                                const double* w, int il) {
  double rv0 = x[0] - xc[0], rv1 = x[1] - xc[1], rv2 = x[2] - xc[2];
  double r = sqrt(rv0 * rv0 + rv1 * rv1 + rv2 * rv2);
  KScal k; kernel_scalars<false>(p, r, k);
  double dx[3] = {rv0 * k.d1r1, rv1 * k.d1r1, rv2 * k.d1r1};
  double drdn = dx[0] * n[0] + dx[1] * n[1] + dx[2] * n[2];
#pragma unroll
  for (int ll = 0; ll < NL; ll++) {
    const int l = (NL == 3) ? ll : il;
    double dxl = (NL == 3) ? dx[ll] : (il == 0 ? dx[0] : (il == 1 ? dx[1] : dx[2]));
    double nl = (NL == 3) ? n[ll] : (il == 0 ? n[0] : (il == 1 ? n[1] : n[2]));
#pragma unroll
    for (int kk = 0; kk < 3; kk++) {
      double dl = (l == kk) ? 1.0 : 0.0;
      double dd = dxl * dx[kk];
      // fs_u = psi*delta - chi r,l r,k ;  fs_t = T1 r,l r,k drdn + T2 (drdn delta + r,k n_l) + T3 r,l n_k
      cplx fu = mk(k.psi.re * dl - k.chi.re * dd, k.psi.im * dl - k.chi.im * dd);
      double c1 = dd * drdn, c2 = drdn * dl + dx[kk] * nl, c3 = dxl * n[kk];
      cplx ft = mk(k.T1.re * c1 + k.T2.re * c2 + k.T3.re * c3, k.T1.im * c1 + k.T2.im * c2 + k.T3.im * c3);
#pragma unroll
      for (int j = 0; j < NN; j++) {
        const int q = (ll * 3 + kk) * NN + j;
        a.hr[q] += ft.re * w[j]; a.hi[q] += ft.im * w[j];
        a.gr[q] += fu.re * w[j]; a.gi[q] += fu.im * w[j];
      }
    }
  }
}

// Exterior point of the HYPERSINGULAR equation (fbem_bem_harela3d_hbie_ext_pre, bem_harela3d.f90:2606-2654): ni = unit normal at
// the collocation point; the s* combination goes to a.h (the reference's m), the d* combination to a.g (its l).
template <int NN, int NL>
MFB_HD void accumulate_exterior_hbie(Acc<NN, NL>& a, const KParams& p, const double* x, const double* n, const double* xc, const double* ni,
                                     const double* w, int il) {
  const double rv0 = x[0] - xc[0], rv1 = x[1] - xc[1], rv2 = x[2] - xc[2];
  const double r = sqrt(rv0 * rv0 + rv1 * rv1 + rv2 * rv2);
  const double d1r1 = 1.0 / r, d1r2 = d1r1 * d1r1, d1r3 = d1r2 * d1r1, d1r4 = d1r2 * d1r2, d1r5 = d1r4 * d1r1;
  const double dx[3] = {rv0 * d1r1, rv1 * d1r1, rv2 * d1r1};
  const double drdn = dx[0] * n[0] + dx[1] * n[1] + dx[2] * n[2], drdni = -(dx[0] * ni[0] + dx[1] * ni[1] + dx[2] * ni[2]);
  const double nni = n[0] * ni[0] + n[1] * ni[1] + n[2] * ni[2];
  const cplx z1 = mk(p.k1.im * r, -p.k1.re * r), z2 = mk(p.k2.im * r, -p.k2.re * r);
  cplx A2, A3, A4, A5, A6, B2, B3, B4, B5, B6;
  zexp_E2_6(z1, A2, A3, A4, A5, A6); zexp_E2_6(z2, B2, B3, B4, B5, B6);
  const cplx E21 = A2 * d1r1, E22 = B2 * d1r1, E31 = A3 * d1r2, E32 = B3 * d1r2, E41 = A4 * d1r3, E42 = B4 * d1r3;
  const cplx E51 = A5 * d1r4, E52 = B5 * d1r4, E61 = A6 * d1r5, E62 = B6 * d1r5;
  cplx t;
  t = cfmar(p.T1[1], d1r2, p.T1[2]);
  t = cfma(p.T1[3], E21, t); t = cfma(p.T1[4], E22, t); t = cfma(p.T1[5], E31, t); t = cfma(p.T1[6], E32, t);
  t = cfma(p.T1[7], E41, t); t = cfma(p.T1[8], E42, t); t = cfma(p.T1[9], E51, t); t = cfma(p.T1[10], E52, t);
  const cplx TT1 = t;
  t = cfmar(p.T2[1], d1r2, p.T2[2]);
  t = cfma(p.T2[3], E22, t); t = cfma(p.T2[4], E31, t); t = cfma(p.T2[5], E32, t); t = cfma(p.T2[6], E41, t);
  t = cfma(p.T2[7], E42, t); t = cfma(p.T2[8], E51, t); t = cfma(p.T2[9], E52, t);
  const cplx TT2 = t;
  t = cfmar(p.T3[1], d1r2, p.T3[2]);
  t = cfma(p.T3[3], E21, t); t = cfma(p.T3[4], E31, t); t = cfma(p.T3[5], E32, t); t = cfma(p.T3[6], E41, t);
  t = cfma(p.T3[7], E42, t); t = cfma(p.T3[8], E51, t); t = cfma(p.T3[9], E52, t);
  const cplx TT3 = t;
  t = p.S1[1] * d1r3 + p.S1[2] * d1r1;
  t = cfma(p.S1[3], E22, t); t = cfma(p.S1[4], E31, t); t = cfma(p.S1[5], E32, t); t = cfma(p.S1[6], E41, t); t = cfma(p.S1[7], E42, t);
  t = cfma(p.S1[8], E51, t); t = cfma(p.S1[9], E52, t); t = cfma(p.S1[10], E61, t); t = cfma(p.S1[11], E62, t);
  const cplx S1 = t;
  t = p.S2[1] * d1r3 + p.S2[2] * d1r1;
  t = cfma(p.S2[3], E21, t); t = cfma(p.S2[4], E31, t); t = cfma(p.S2[5], E32, t); t = cfma(p.S2[6], E41, t); t = cfma(p.S2[7], E42, t);
  t = cfma(p.S2[8], E51, t); t = cfma(p.S2[9], E52, t); t = cfma(p.S2[10], E61, t); t = cfma(p.S2[11], E62, t);
  const cplx S2 = t;
  t = p.S3[1] * d1r3 + p.S3[2] * d1r1;
  t = cfma(p.S3[3], E21, t); t = cfma(p.S3[4], E22, t); t = cfma(p.S3[5], E31, t); t = cfma(p.S3[6], E32, t); t = cfma(p.S3[7], E41, t);
  t = cfma(p.S3[8], E42, t); t = cfma(p.S3[9], E51, t); t = cfma(p.S3[10], E52, t); t = cfma(p.S3[11], E61, t); t = cfma(p.S3[12], E62, t);
  const cplx S3 = t;
  t = p.S4[1] * d1r3 + p.S4[2] * d1r1 + p.S4[3];
  t = cfma(p.S4[4], E32, t); t = cfma(p.S4[5], E41, t); t = cfma(p.S4[6], E42, t); t = cfma(p.S4[7], E51, t); t = cfma(p.S4[8], E52, t);
  t = cfma(p.S4[9], E61, t); t = cfma(p.S4[10], E62, t);
  const cplx S4 = t;
  t = p.S5[1] * d1r3 + p.S5[2] * d1r1 + p.S5[3];
  t = cfma(p.S5[4], E21, t); t = cfma(p.S5[5], E31, t); t = cfma(p.S5[6], E41, t); t = cfma(p.S5[7], E42, t); t = cfma(p.S5[8], E51, t);
  t = cfma(p.S5[9], E52, t); t = cfma(p.S5[10], E61, t); t = cfma(p.S5[11], E62, t);
  const cplx S5 = t;
#pragma unroll
  for (int ll = 0; ll < NL; ll++) {
    const int l = (NL == 3) ? ll : il;
    const double dxl = (NL == 3) ? dx[ll] : (il == 0 ? dx[0] : (il == 1 ? dx[1] : dx[2]));
    const double nl = (NL == 3) ? n[ll] : (il == 0 ? n[0] : (il == 1 ? n[1] : n[2]));
    const double nil = (NL == 3) ? ni[ll] : (il == 0 ? ni[0] : (il == 1 ? ni[1] : ni[2]));
#pragma unroll
    for (int kk = 0; kk < 3; kk++) {
      const double dl = (l == kk) ? 1.0 : 0.0;
      // fs_d = TT1 r,l r,k drdni - TT2 (-drdni delta + r,l ni_k) - TT3 r,k ni_l
      const double d1 = dxl * dx[kk] * drdni, d2 = -(-drdni * dl + dxl * ni[kk]), d3 = -dx[kk] * nil;
      const cplx fd = mk(TT1.re * d1 + TT2.re * d2 + TT3.re * d3, TT1.im * d1 + TT2.im * d2 + TT3.im * d3);
      const double s1 = dxl * ni[kk] * drdn - dx[kk] * nl * drdni - dl * drdn * drdni + dxl * dx[kk] * nni;
      const double s2 = dx[kk] * nil * drdn - dxl * n[kk] * drdni, s3 = dxl * dx[kk] * drdn * drdni, s4 = dl * nni + ni[kk] * nl, s5 = n[kk] * nil;
      const cplx fs = mk(S1.re * s1 + S2.re * s2 + S3.re * s3 + S4.re * s4 + S5.re * s5, S1.im * s1 + S2.im * s2 + S3.im * s3 + S4.im * s4 + S5.im * s5);
#pragma unroll
      for (int j = 0; j < NN; j++) {
        const int q = (ll * 3 + kk) * NN + j;
        a.hr[q] += fs.re * w[j]; a.hi[q] += fs.im * w[j];
        a.gr[q] += fd.re * w[j]; a.gi[q] += fd.im * w[j];
      }
    }
  }
}

// Interior (singular element) point (bem_harela3d.f90:1386-1415): regular parts + CPV-regularised part;
// phi[j] at the point, phi_i[j] at the collocation point, jw = J*rho*jthetap*w_ang*w_rad.
template <int NN, int NL>
MFB_HD void accumulate_interior(Acc<NN, NL>& a, const KParams& p, const double* x, const double* n, const double* xc,
                                const double* phi, const double* phi_i, double jw, int il) {
  double rv0 = x[0] - xc[0], rv1 = x[1] - xc[1], rv2 = x[2] - xc[2];
  double r = sqrt(rv0 * rv0 + rv1 * rv1 + rv2 * rv2);
  KScal k; kernel_scalars<true>(p, r, k);
  double dx[3] = {rv0 * k.d1r1, rv1 * k.d1r1, rv2 * k.d1r1};
  double drdn = dx[0] * n[0] + dx[1] * n[1] + dx[2] * n[2];
  cplx T1s = cfmar(p.T1[1], k.d1r2, k.T1), T2s = cfmar(p.T2[1], k.d1r2, k.T2), T21 = p.T2[1] * k.d1r2;
#pragma unroll
  for (int ll = 0; ll < NL; ll++) {
    const int l = (NL == 3) ? ll : il;
    double dxl = (NL == 3) ? dx[ll] : (il == 0 ? dx[0] : (il == 1 ? dx[1] : dx[2]));
    double nl = (NL == 3) ? n[ll] : (il == 0 ? n[0] : (il == 1 ? n[1] : n[2]));
#pragma unroll
    for (int kk = 0; kk < 3; kk++) {
      double dl = (l == kk) ? 1.0 : 0.0;
      double dd = dxl * dx[kk];
      cplx fu = mk(k.psi.re * dl - k.chi.re * dd, k.psi.im * dl - k.chi.im * dd);
      double c1 = dd * drdn, c2 = drdn * dl, c3 = dx[kk] * nl, c4 = dxl * n[kk];
      cplx ft = mk(T1s.re * c1 + T2s.re * c2 + k.T2.re * c3 + k.T3.re * c4, T1s.im * c1 + T2s.im * c2 + k.T2.im * c3 + k.T3.im * c4);
      double cc = nl * dx[kk] - n[kk] * dxl;
      cplx fc = T21 * cc;
#pragma unroll
      for (int j = 0; j < NN; j++) {
        const int q = (ll * 3 + kk) * NN + j;
        double wj = phi[j] * jw, wc = (phi[j] - phi_i[j]) * jw;
        a.hr[q] += ft.re * wj + fc.re * wc; a.hi[q] += ft.im * wj + fc.im * wc;
        a.gr[q] += fu.re * wj; a.gi[q] += fu.im * wj;
      }
    }
  }
}

// ---- element geometry at xi (isoparametric Lagrange elements, continuous) --------------------------------------
template <int ET> struct ElemTraits;
template <> struct ElemTraits<5> { static const int NN = 3, NV = 3; };
template <> struct ElemTraits<6> { static const int NN = 6, NV = 3; };
template <> struct ElemTraits<7> { static const int NN = 4, NV = 4; };
template <> struct ElemTraits<8> { static const int NN = 8, NV = 4; };
template <> struct ElemTraits<9> { static const int NN = 9, NV = 4; };

template <int ET>
MFB_HD void shape(const double xi1, const double xi2, double* phi, double* d1, double* d2) {
  if (ET == 5) {
    phi[0] = xi1; phi[1] = xi2; phi[2] = 1.0 - xi1 - xi2;
    d1[0] = 1.0; d1[1] = 0.0; d1[2] = -1.0; d2[0] = 0.0; d2[1] = 1.0; d2[2] = -1.0;
  } else if (ET == 6) {
    double a3 = 1.0 - xi1 - xi2;
    phi[0] = xi1 * (2.0 * xi1 - 1.0); phi[1] = xi2 * (2.0 * xi2 - 1.0); phi[2] = a3 * (2.0 * a3 - 1.0);
    phi[3] = 4.0 * xi1 * xi2; phi[4] = 4.0 * xi2 * a3; phi[5] = 4.0 * xi1 * a3;
    d1[0] = 4.0 * xi1 - 1.0; d1[1] = 0.0; d1[2] = 4.0 * xi1 + 4.0 * xi2 - 3.0; d1[3] = 4.0 * xi2; d1[4] = -4.0 * xi2; d1[5] = -4.0 * (xi2 + 2.0 * xi1 - 1.0);
    d2[0] = 0.0; d2[1] = 4.0 * xi2 - 1.0; d2[2] = 4.0 * xi1 + 4.0 * xi2 - 3.0; d2[3] = 4.0 * xi1; d2[4] = -4.0 * (2.0 * xi2 + xi1 - 1.0); d2[5] = -4.0 * xi1;
  } else if (ET == 7) {
    double a3 = 0.25 * (1.0 + xi1), a4 = 0.25 * (1.0 - xi1), a5 = 1.0 + xi2, a6 = 1.0 - xi2;
    phi[0] = a4 * a6; phi[1] = a3 * a6; phi[2] = a3 * a5; phi[3] = a4 * a5;
    d1[0] = -0.25 * a6; d1[1] = 0.25 * a6; d1[2] = 0.25 * a5; d1[3] = -0.25 * a5;
    d2[0] = -a4; d2[1] = -a3; d2[2] = a3; d2[3] = a4;
  } else if (ET == 8) {
    double a3 = 0.25 * (1.0 + xi1), a4 = 0.25 * (1.0 - xi1), a5 = 1.0 + xi2, a6 = 1.0 - xi2, a7 = 1.0 - xi1 * xi1, a8 = 1.0 - xi2 * xi2;
    phi[0] = a4 * a6 * (-xi1 - a5); phi[1] = a3 * a6 * (xi1 - a5); phi[2] = a3 * a5 * (xi1 - a6); phi[3] = a4 * a5 * (-xi1 - a6);
    phi[4] = 0.5 * a6 * a7; phi[5] = 2.0 * a3 * a8; phi[6] = 0.5 * a5 * a7; phi[7] = 2.0 * a4 * a8;
    double b3 = xi2 + 1.0, b4 = xi2 - 1.0, b5 = xi2 + 2.0 * xi1, b6 = xi2 - 2.0 * xi1;
    d1[0] = -0.25 * b4 * b5; d1[1] = 0.25 * b4 * b6; d1[2] = 0.25 * b3 * b5; d1[3] = -0.25 * b3 * b6;
    d1[4] = xi1 * b4; d1[5] = -0.5 * b3 * b4; d1[6] = -xi1 * b3; d1[7] = 0.5 * b3 * b4;
    double c3 = xi1 + 1.0, c4 = xi1 - 1.0, c5 = 2.0 * xi2 + xi1, c6 = 2.0 * xi2 - xi1;
    d2[0] = -0.25 * c4 * c5; d2[1] = 0.25 * c3 * c6; d2[2] = 0.25 * c3 * c5; d2[3] = -0.25 * c4 * c6;
    d2[4] = 0.5 * c3 * c4; d2[5] = -xi2 * c3; d2[6] = -0.5 * c3 * c4; d2[7] = xi2 * c4;
  } else {
    double a3 = 0.25 * xi1 * (xi1 + 1.0), a4 = 0.25 * xi1 * (xi1 - 1.0), a5 = xi2 * (xi2 + 1.0), a6 = xi2 * (xi2 - 1.0);
    double a7 = 1.0 - xi1 * xi1, a8 = 1.0 - xi2 * xi2;
    phi[0] = a4 * a6; phi[1] = a3 * a6; phi[2] = a3 * a5; phi[3] = a4 * a5;
    phi[4] = 0.5 * a6 * a7; phi[5] = 2.0 * a3 * a8; phi[6] = 0.5 * a5 * a7; phi[7] = 2.0 * a4 * a8; phi[8] = a7 * a8;
    double b3 = 2.0 * xi1 + 1.0, b4 = 2.0 * xi1 - 1.0, b5 = xi2 + 1.0, b6 = xi2 - 1.0, b7 = 0.25 * xi2, b8 = b5 * b6, b9 = -0.5 * b8, b10 = -xi1 * xi2;
    d1[0] = b7 * b4 * b6; d1[1] = b7 * b3 * b6; d1[2] = b7 * b3 * b5; d1[3] = b7 * b4 * b5;
    d1[4] = b10 * b6; d1[5] = b9 * b3; d1[6] = b10 * b5; d1[7] = b9 * b4; d1[8] = 2.0 * xi1 * b8;
    double c3 = 2.0 * xi2 + 1.0, c4 = 2.0 * xi2 - 1.0, c5 = xi1 + 1.0, c6 = xi1 - 1.0, c7 = 0.25 * xi1, c8 = c5 * c6, c9 = -0.5 * c8;
    d2[0] = c7 * c6 * c4; d2[1] = c7 * c5 * c4; d2[2] = c7 * c5 * c3; d2[3] = c7 * c6 * c3;
    d2[4] = c9 * c4; d2[5] = b10 * c5; d2[6] = c9 * c3; d2[7] = b10 * c6; d2[8] = 2.0 * xi2 * c8;
  }
}

// x(xi), unit normal and geometric jacobian |T1 x T2| (e.g. bem_harela3d.f90:815-831)
template <int ET>
MFB_HD void geometry_at(const double* xn, double xi1, double xi2, double* phi, double* x, double* n, double& jg) {
  const int NN = ElemTraits<ET>::NN;
  double d1[NN], d2[NN];
  shape<ET>(xi1, xi2, phi, d1, d2);
  double t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0};
  x[0] = x[1] = x[2] = 0.0;
#pragma unroll
  for (int k = 0; k < NN; k++) {
#pragma unroll
    for (int c = 0; c < 3; c++) { x[c] += phi[k] * xn[3 * k + c]; t1[c] += d1[k] * xn[3 * k + c]; t2[c] += d2[k] * xn[3 * k + c]; }
  }
  double N0 = t1[1] * t2[2] - t1[2] * t2[1], N1 = t1[2] * t2[0] - t1[0] * t2[2], N2 = t1[0] * t2[1] - t1[1] * t2[0];
  jg = sqrt(N0 * N0 + N1 * N1 + N2 * N2);
  double inv = 1.0 / jg;
  n[0] = N0 * inv; n[1] = N1 * inv; n[2] = N2 * inv;
}

// One point of a Telles/subdivision leaf (bem_harela3d.f90:764-863 quads, :908-1012 triangles):
// (g1,w1),(g2,w2) = 1D Gauss-Legendre abscissa/weight per direction ([-1,1] for quads, [0,1] for triangles);
// tp1,tp2 = Telles cubics; xi_s = sub-element corners in the parent xi space.  Outputs x, n, w[j] = phi_j*jw.
template <int ET>
MFB_HD void leaf_point(const double* xn, const double* xi_s, const double* tp1, const double* tp2, double g1, double w1,
                       double g2, double w2, double* x, double* n, double* w) {
  const int NN = ElemTraits<ET>::NN;
  const bool tri = (ElemTraits<ET>::NV == 3);
  double t1 = ((tp1[0] * g1 + tp1[1]) * g1 + tp1[2]) * g1 + tp1[3], j1 = (3.0 * tp1[0] * g1 + 2.0 * tp1[1]) * g1 + tp1[2];
  double t2 = ((tp2[0] * g2 + tp2[1]) * g2 + tp2[2]) * g2 + tp2[3], j2 = (3.0 * tp2[0] * g2 + 2.0 * tp2[1]) * g2 + tp2[2];
  double xi1, xi2, js;
  if (tri) {
    double p1 = (1.0 - t2) * t1, p2 = t2, jqt = 1.0 - t2;
    // sub-triangle map with vertex shape functions (p1, p2, 1-p1-p2)
    double p3 = 1.0 - p1 - p2;
    xi1 = p1 * xi_s[0] + p2 * xi_s[2] + p3 * xi_s[4];
    xi2 = p1 * xi_s[1] + p2 * xi_s[3] + p3 * xi_s[5];
    double a1 = xi_s[0] - xi_s[4], a2 = xi_s[1] - xi_s[5], b1 = xi_s[2] - xi_s[4], b2 = xi_s[3] - xi_s[5];
    js = (a1 * b2 - a2 * b1) * jqt;
  } else {
    double s3 = 0.25 * (1.0 + t1), s4 = 0.25 * (1.0 - t1), s5 = 1.0 + t2, s6 = 1.0 - t2;
    double f0 = s4 * s6, f1 = s3 * s6, f2 = s3 * s5, f3 = s4 * s5;
    xi1 = f0 * xi_s[0] + f1 * xi_s[2] + f2 * xi_s[4] + f3 * xi_s[6];
    xi2 = f0 * xi_s[1] + f1 * xi_s[3] + f2 * xi_s[5] + f3 * xi_s[7];
    double e0 = -0.25 * s6, e1 = 0.25 * s6, e2 = 0.25 * s5, e3 = -0.25 * s5;   // d/dt1
    double h0 = -s4, h1 = -s3, h2 = s3, h3 = s4;                               // d/dt2
    double a1 = e0 * xi_s[0] + e1 * xi_s[2] + e2 * xi_s[4] + e3 * xi_s[6], a2 = e0 * xi_s[1] + e1 * xi_s[3] + e2 * xi_s[5] + e3 * xi_s[7];
    double b1 = h0 * xi_s[0] + h1 * xi_s[2] + h2 * xi_s[4] + h3 * xi_s[6], b2 = h0 * xi_s[1] + h1 * xi_s[3] + h2 * xi_s[5] + h3 * xi_s[7];
    js = a1 * b2 - a2 * b1;
  }
  double phi[NN], jg;
  geometry_at<ET>(xn, xi1, xi2, phi, x, n, jg);
  double jw = jg * js * j1 * j2 * w1 * w2;
#pragma unroll
  for (int k = 0; k < NN; k++) w[k] = phi[k] * jw;
}

}  // namespace mfbd
