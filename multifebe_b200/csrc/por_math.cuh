// por_math.cuh -- per-Gauss-point arithmetic of the Biot poroelastic SBIE kernels (host+device inline, like bem_math.cuh / pot_math.cuh).
//
// Used by the kernels of poro.cu (validated on hardware in round 2: tests/test_gpu_poroelastic.py).  The header is also compiled for the HOST by
// tests/test_por_math_host.py and held to the oracle there, so the point formulas and parameter tables are checked without a GPU.
//
// Node variables: 0 = fluid phase (tau | Un), 1..3 = skeleton (u_k | t_k).  The fundamental solution is evaluated in the reference's
// regularised form (lib/fbem/src/bem_harpor3d.f90:944-996): twelve radial scalars, each a static part (1/r or 1/r^2) + a constant + a sum
// of coeff * E_m(z_j)/r^(m-1), z_j = -i k_j r for the two compressional (k1, k2) and the shear (k3) wavenumbers, m = 2..5:
//   u*_00 = eta            u*_0k = u*_k0 = vartheta r,k           u*_lk = psi delta_lk - chi r,l r,k
//   t*_00 = W0 dr/dn       t*_0k = T01 r,k dr/dn + T02 n_k        t*_l0 = W1 r,l dr/dn + W2 n_l
//   t*_lk = T1 r,l r,k dr/dn + T2 (dr/dn delta_lk + r,k n_l) + T3 r,l n_k
// times the constants cte_u(l,k), cte_t(l,k) (:545-556).
#pragma once
#include "bem_math.cuh"
#include <complex>

namespace mfbd {

// fbem_bem_harpor3d_parameters (bem_harpor3d.f90:91-134), SBIE subset, 1-based tables like the reference
struct PorParams {
  cplx k1, k2, k3, J, Z;
  cplx eta[4], vartheta[6], psi[9], chi[10], W0[7], T01[8], T02[9], W1[11], W2[10], T1[15], T2[13], T3[14];
  cplx cte_u[4][4], cte_t[4][4];
};

struct PorScal { cplx eta, vartheta, psi, chi, W0, T01, T02, W1, W2, T1, T2, T3; };

// The twelve radial scalars at distance r (d1r1 = 1/r).  REGULAR_ONLY drops the static 1/r^2 parts of W0, T1, T2, T3 (interior
// integration, bem_harpor3d.f90:1720-1750; the caller adds back the ones it integrates in full).
template <bool REGULAR_ONLY>
MFB_HD void por_scalars(const PorParams& p, double r, double d1r1, PorScal& s) {
  const double d1r2 = d1r1 * d1r1, d1r3 = d1r2 * d1r1, d1r4 = d1r2 * d1r2;
  cplx e2[3], e3[3], e4[3], e5[3];
  {
    const cplx k[3] = {p.k1, p.k2, p.k3};
#pragma unroll
    for (int j = 0; j < 3; j++) {
      cplx E2, E3, E4, E5;
      zexp_E2_5(mk(k[j].im * r, -k[j].re * r), E2, E3, E4, E5);
      e2[j] = E2 * d1r1; e3[j] = E3 * d1r2; e4[j] = E4 * d1r3; e5[j] = E5 * d1r4;
    }
  }
  cplx t;
  t = mk(d1r1 + p.eta[1].re, p.eta[1].im);
  s.eta = cfma(p.eta[2], e2[0], cfma(p.eta[3], e2[1], t));
  t = p.vartheta[1];
  t = cfma(p.vartheta[2], e2[0], t); t = cfma(p.vartheta[3], e2[1], t); t = cfma(p.vartheta[4], e3[0], t); t = cfma(p.vartheta[5], e3[1], t);
  s.vartheta = t;
  t = cfmar(p.psi[1], d1r1, p.psi[2]) + e2[2];
  t = cfma(p.psi[3], e3[0], t); t = cfma(p.psi[4], e3[1], t); t = cfma(p.psi[5], e3[2], t); t = cfma(p.psi[6], e4[0], t); t = cfma(p.psi[7], e4[1], t);
  t = cfma(p.psi[8], e4[2], t);
  s.psi = t;
  t = p.chi[1] * d1r1 + e2[2];
  t = cfma(p.chi[2], e2[0], t); t = cfma(p.chi[3], e2[1], t); t = cfma(p.chi[4], e3[0], t); t = cfma(p.chi[5], e3[1], t); t = cfma(p.chi[6], e3[2], t);
  t = cfma(p.chi[7], e4[0], t); t = cfma(p.chi[8], e4[1], t); t = cfma(p.chi[9], e4[2], t);
  s.chi = t;
  t = REGULAR_ONLY ? p.W0[2] : cfmar(p.W0[1], d1r2, p.W0[2]);
  t = cfma(p.W0[3], e2[0], t); t = cfma(p.W0[4], e2[1], t); t = cfma(p.W0[5], e3[0], t); t = cfma(p.W0[6], e3[1], t);
  s.W0 = t;
  t = p.T01[1] * d1r1;
  t = cfma(p.T01[2], e2[0], t); t = cfma(p.T01[3], e2[1], t); t = cfma(p.T01[4], e3[0], t); t = cfma(p.T01[5], e3[1], t); t = cfma(p.T01[6], e4[0], t);
  t = cfma(p.T01[7], e4[1], t);
  s.T01 = t;
  t = cfmar(p.T02[1], d1r1, p.T02[2]);
  t = cfma(p.T02[3], e2[0], t); t = cfma(p.T02[4], e2[1], t); t = cfma(p.T02[5], e3[0], t); t = cfma(p.T02[6], e3[1], t); t = cfma(p.T02[7], e4[0], t);
  t = cfma(p.T02[8], e4[1], t);
  s.T02 = t;
  t = p.W1[1] * d1r1;
  t = cfma(p.W1[2], e2[0], t); t = cfma(p.W1[3], e2[1], t); t = cfma(p.W1[4], e2[2], t); t = cfma(p.W1[5], e3[0], t); t = cfma(p.W1[6], e3[1], t);
  t = cfma(p.W1[7], e3[2], t); t = cfma(p.W1[8], e4[0], t); t = cfma(p.W1[9], e4[1], t); t = cfma(p.W1[10], e4[2], t);
  s.W1 = t;
  t = cfmar(p.W2[1], d1r1, p.W2[2]);
  t = cfma(p.W2[3], e2[2], t); t = cfma(p.W2[4], e3[0], t); t = cfma(p.W2[5], e3[1], t); t = cfma(p.W2[6], e3[2], t); t = cfma(p.W2[7], e4[0], t);
  t = cfma(p.W2[8], e4[1], t); t = cfma(p.W2[9], e4[2], t);
  s.W2 = t;
  t = REGULAR_ONLY ? p.T1[2] : cfmar(p.T1[1], d1r2, p.T1[2]);
  t = cfma(p.T1[3], e2[0], t); t = cfma(p.T1[4], e2[1], t); t = cfma(p.T1[5], e2[2], t); t = cfma(p.T1[6], e3[0], t); t = cfma(p.T1[7], e3[1], t);
  t = cfma(p.T1[8], e3[2], t); t = cfma(p.T1[9], e4[0], t); t = cfma(p.T1[10], e4[1], t); t = cfma(p.T1[11], e4[2], t); t = cfma(p.T1[12], e5[0], t);
  t = cfma(p.T1[13], e5[1], t); t = cfma(p.T1[14], e5[2], t);
  s.T1 = t;
  t = REGULAR_ONLY ? p.T2[2] : cfmar(p.T2[1], d1r2, p.T2[2]);
  t = cfma(p.T2[3], e2[2], t); t = cfma(p.T2[4], e3[0], t); t = cfma(p.T2[5], e3[1], t); t = cfma(p.T2[6], e3[2], t); t = cfma(p.T2[7], e4[0], t);
  t = cfma(p.T2[8], e4[1], t); t = cfma(p.T2[9], e4[2], t); t = cfma(p.T2[10], e5[0], t); t = cfma(p.T2[11], e5[1], t); t = cfma(p.T2[12], e5[2], t);
  s.T2 = t;
  t = REGULAR_ONLY ? p.T3[2] : cfmar(p.T3[1], d1r2, p.T3[2]);
  t = cfma(p.T3[3], e2[0], t); t = cfma(p.T3[4], e2[1], t); t = cfma(p.T3[5], e3[0], t); t = cfma(p.T3[6], e3[1], t); t = cfma(p.T3[7], e3[2], t);
  t = cfma(p.T3[8], e4[0], t); t = cfma(p.T3[9], e4[1], t); t = cfma(p.T3[10], e4[2], t); t = cfma(p.T3[11], e5[0], t); t = cfma(p.T3[12], e5[1], t);
  t = cfma(p.T3[13], e5[2], t);
  s.T3 = t;
}

// The 4 x 4 blocks fs_u, fs_t at one exterior point (UNSCALED: the caller multiplies by cte_u / cte_t once per pair, as the reference does).
// x, n = integration point and unit normal, xc = collocation point.
MFB_HD void por_exterior_blocks(const PorParams& p, const double* x, const double* n, const double* xc, cplx fu[4][4], cplx ft[4][4]) {
  const double rv0 = x[0] - xc[0], rv1 = x[1] - xc[1], rv2 = x[2] - xc[2];
  const double r = sqrt(rv0 * rv0 + rv1 * rv1 + rv2 * rv2), d1r1 = 1.0 / r;
  const double dx[3] = {rv0 * d1r1, rv1 * d1r1, rv2 * d1r1};
  const double drdn = dx[0] * n[0] + dx[1] * n[1] + dx[2] * n[2];
  PorScal s; por_scalars<false>(p, r, d1r1, s);
  fu[0][0] = s.eta; ft[0][0] = s.W0 * drdn;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    fu[0][c + 1] = s.vartheta * dx[c]; fu[c + 1][0] = fu[0][c + 1];
    ft[0][c + 1] = cfmar(s.T01, dx[c] * drdn, s.T02 * n[c]); ft[c + 1][0] = cfmar(s.W1, dx[c] * drdn, s.W2 * n[c]);
  }
#pragma unroll
  for (int l = 0; l < 3; l++)
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double dl = (l == k) ? 1.0 : 0.0, dd = dx[l] * dx[k];
      fu[l + 1][k + 1] = mk(s.psi.re * dl - s.chi.re * dd, s.psi.im * dl - s.chi.im * dd);
      const double c1 = dd * drdn, c2 = drdn * dl + dx[k] * n[l], c3 = dx[l] * n[k];
      ft[l + 1][k + 1] = mk(s.T1.re * c1 + s.T2.re * c2 + s.T3.re * c3, s.T1.im * c1 + s.T2.im * c2 + s.T3.im * c3);
    }
}

// The same at a point of the element that holds the collocation point (bem_harpor3d.f90:1752-1790): the static 1/r^2 parts of W0, of T1
// and of the dr/dn delta term of T2 are integrated in full; what is left of the static T2, T3 parts, T2(1)/r^2 (n_l r,k - n_k r,l), is
// returned separately (fc[l][k], skeleton block only): the caller integrates it against (phi_j - phi_j(xi_i)) and adds the line integrals.
MFB_HD void por_interior_blocks(const PorParams& p, const double* x, const double* n, const double* xc, cplx fu[4][4], cplx ft[4][4], cplx fc[3][3]) {
  const double rv0 = x[0] - xc[0], rv1 = x[1] - xc[1], rv2 = x[2] - xc[2];
  const double r = sqrt(rv0 * rv0 + rv1 * rv1 + rv2 * rv2), d1r1 = 1.0 / r, d1r2 = d1r1 * d1r1;
  const double dx[3] = {rv0 * d1r1, rv1 * d1r1, rv2 * d1r1};
  const double drdn = dx[0] * n[0] + dx[1] * n[1] + dx[2] * n[2];
  PorScal s; por_scalars<true>(p, r, d1r1, s);
  fu[0][0] = s.eta; ft[0][0] = cfmar(p.W0[1], d1r2, s.W0) * drdn;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    fu[0][c + 1] = s.vartheta * dx[c]; fu[c + 1][0] = fu[0][c + 1];
    ft[0][c + 1] = cfmar(s.T01, dx[c] * drdn, s.T02 * n[c]); ft[c + 1][0] = cfmar(s.W1, dx[c] * drdn, s.W2 * n[c]);
  }
  const cplx T1s = cfmar(p.T1[1], d1r2, s.T1), T2s = cfmar(p.T2[1], d1r2, s.T2), T21 = p.T2[1] * d1r2;
#pragma unroll
  for (int l = 0; l < 3; l++)
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double dl = (l == k) ? 1.0 : 0.0, dd = dx[l] * dx[k];
      fu[l + 1][k + 1] = mk(s.psi.re * dl - s.chi.re * dd, s.psi.im * dl - s.chi.im * dd);
      const double c1 = dd * drdn, c2 = drdn * dl, c3 = dx[k] * n[l], c4 = dx[l] * n[k];
      ft[l + 1][k + 1] = mk(T1s.re * c1 + T2s.re * c2 + s.T2.re * c3 + s.T3.re * c4, T1s.im * c1 + T2s.im * c2 + s.T2.im * c3 + s.T3.im * c4);
      fc[l][k] = T21 * (n[l] * dx[k] - n[k] * dx[l]);
    }
}

// fbem_bem_harpor3d_calculate_parameters (bem_harpor3d.f90:204-568, SBIE subset) on the host.  With a = (lambda + 2 mu), m = mu / a,
// v = (Q/R - Z) / a, D = k1^2 - k2^2, alpha_j = k_j^2 - m k3^2, beta_j = m k_j^2 - k1^2 k2^2 / k3^2, and the shorthands
// A_j = alpha_j / D, B_j = beta_j / D, V = v / D every table entry is a short product.
inline void por_params_host(std::complex<double> lambda, std::complex<double> mu, double rho1, double rho2, double rhoa, std::complex<double> R,
                            std::complex<double> Q, double b, double omega, PorParams& P) {
  typedef std::complex<double> cd;
  const cd I(0.0, 1.0);
  const cd r11 = rho1 + rhoa - I * b / omega, r12 = -rhoa + I * b / omega, r22 = rho2 + rhoa - I * b / omega;
  const cd J = 1.0 / (r22 * omega * omega), Z = r12 / r22, a = lambda + 2.0 * mu, QR = Q / R;
  cd k3 = std::sqrt((r11 / r22 - Z * Z) / (mu * J));
  if (k3.real() < 0.0) k3 = -k3;
  const cd cb = a / (J * R) + mu * k3 * k3 + (QR - Z) * (QR - Z) / J, cc = mu * k3 * k3 / (J * R), disc = std::sqrt(cb * cb - 4.0 * a * cc);
  cd k1 = std::sqrt(0.5 * (cb - disc) / a), k2 = std::sqrt(0.5 * (cb + disc) / a);
  if (k1.real() < 0.0) k1 = -k1;
  if (k2.real() < 0.0) k2 = -k2;
  if (k1.real() > k2.real()) std::swap(k1, k2);
  const cd q1 = k1 * k1, q2 = k2 * k2, q3 = k3 * k3, m = mu / a, v = (QR - Z) / a, D = q1 - q2;
  const cd al1 = q1 - m * q3, al2 = q2 - m * q3, be1 = m * q1 - q1 * q2 / q3, be2 = m * q2 - q1 * q2 / q3;
  const cd A1 = al1 / D, A2 = al2 / D, B1 = be1 / D, B2 = be2 / D, V = v / D;
  const cd i1 = I * k1, i2 = I * k2, i3 = I * k3;
  const cd sB = i1 * B1 - i2 * B2;            // (i k1 beta1 - i k2 beta2) / D
  const cd sK = (q1 * be1 - q2 * be2) / D;    // (k1^2 beta1 - k2^2 beta2) / D
  const cd c3 = I * (q1 * k1 - q2 * k2) / D;  // i (k1^3 - k2^3) / D
  const cd mv = mu * v, mV = mu * V, ZV1 = Z * v + J * al1, ZV2 = Z * v + J * al2, Z2m = Z * Z / mu;
  auto set = [](cplx& o, cd z) { o.re = z.real(); o.im = z.imag(); };
  set(P.k1, k1); set(P.k2, k2); set(P.k3, k3); set(P.J, J); set(P.Z, Z);
  cd eta[4] = {0, -(i1 * A1 - i2 * A2), A1, -A2};
  cd vt[6] = {0, 0.5 * v, V * i1, -V * i2, V, -V};
  cd psi[9] = {0, 0.5 * (lambda + 3.0 * mu) / a, -(sB + 2.0 * i3) / 3.0, -B1 / i1, B2 / i2, 1.0 / i3, B1 / q1, -B2 / q2, -1.0 / q3};
  cd chi[10] = {0, -0.5 * (lambda + mu) / a, -B1, B2, -3.0 * B1 / i1, 3.0 * B2 / i2, 3.0 / i3, 3.0 * B1 / q1, -3.0 * B2 / q2, -3.0 / q3};
  cd W0[7] = {0, J, 0.5 * (Z * v + J * (q1 * A1 - q2 * A2)), ZV1 * i1 / D, -ZV2 * i2 / D, ZV1 / D, -ZV2 / D};
  cd T01[8] = {0, mv, -2.0 * mV * q1, 2.0 * mV * q2, 6.0 * mV * i1, -6.0 * mV * i2, 6.0 * mV, -6.0 * mV};
  cd T02[9] = {0, mv + Z, (lambda + 2.0 / 3.0 * mu) * v * c3 - QR * (i1 * A1 - i2 * A2), QR * A1 - lambda * V * q1, -(QR * A2 - lambda * V * q2),
               -2.0 * mV * i1, 2.0 * mV * i2, -2.0 * mV, 2.0 * mV};
  cd W1[11] = {0, 0.5 * (QR * m - Z), -(mV * q1 + Z * B1), mV * q2 + Z * B2, Z, 3.0 * (mV * i1 - Z * B1 / i1), -3.0 * (mV * i2 - Z * B2 / i2), 3.0 * Z / i3,
               3.0 * (mV + Z * B1 / q1), -3.0 * (mV + Z * B2 / q2), -3.0 * Z / q3};
  cd W2[10] = {0, -0.5 * (QR * m + Z), (mv * c3 + Z * (sB + 2.0 * i3)) / 3.0, -Z, -(mV * i1 - Z * B1 / i1), mV * i2 - Z * B2 / i2, -Z / i3,
               -(mV + Z * B1 / q1), mV + Z * B2 / q2, Z / q3};
  cd T1[15] = {0, -3.0 * (lambda + mu) / a, 0.25 * (sK - q3), -2.0 * i1 * B1, 2.0 * i2 * B2, 2.0 * i3, -12.0 * B1, 12.0 * B2, 12.0, -30.0 * B1 / i1,
               30.0 * B2 / i2, 30.0 / i3, 30.0 * B1 / q1, -30.0 * B2 / q2, -30.0 / q3};
  cd T2[13] = {0, -m, -0.25 * (sK + q3), -i3, 2.0 * B1, -2.0 * B2, -3.0, 6.0 * B1 / i1, -6.0 * B2 / i2, -6.0 / i3, -6.0 * B1 / q1, 6.0 * B2 / q2, 6.0 / q3};
  const cd qv = QR * v / J, lm = lambda / mu;
  cd T3[14] = {0, m, 0.25 * (2.0 * qv - (2.0 * lm + 1.0) * sK + q3), (qv - lm * be1) * i1 / D, -(qv - lm * be2) * i2 / D, (qv - (lm - 2.0) * be1) / D,
               -(qv - (lm - 2.0) * be2) / D, -2.0, 6.0 * B1 / i1, -6.0 * B2 / i2, -6.0 / i3, -6.0 * B1 / q1, 6.0 * B2 / q2, 6.0 / q3};
  for (int i = 1; i < 4; i++) set(P.eta[i], eta[i]);
  for (int i = 1; i < 6; i++) set(P.vartheta[i], vt[i]);
  for (int i = 1; i < 9; i++) set(P.psi[i], psi[i]);
  for (int i = 1; i < 10; i++) set(P.chi[i], chi[i]);
  for (int i = 1; i < 7; i++) set(P.W0[i], W0[i]);
  for (int i = 1; i < 8; i++) set(P.T01[i], T01[i]);
  for (int i = 1; i < 9; i++) set(P.T02[i], T02[i]);
  for (int i = 1; i < 11; i++) set(P.W1[i], W1[i]);
  for (int i = 1; i < 10; i++) set(P.W2[i], W2[i]);
  for (int i = 1; i < 15; i++) set(P.T1[i], T1[i]);
  for (int i = 1; i < 13; i++) set(P.T2[i], T2[i]);
  for (int i = 1; i < 14; i++) set(P.T3[i], T3[i]);
  const double c4 = 0.07957747154594767280411105048;
  for (int l = 0; l < 4; l++)
    for (int k = 0; k < 4; k++) {
      set(P.cte_u[l][k], l == 0 ? cd(-c4) : (k == 0 ? -c4 / J : c4 / mu));
      set(P.cte_t[l][k], l == 0 ? cd(k == 0 ? -c4 : c4) : (k == 0 ? -c4 / mu : cd(c4)));
    }
}

}  // namespace mfbd
