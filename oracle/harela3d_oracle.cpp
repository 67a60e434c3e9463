// =====================================================================================
// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under
// multifebe_b200/ may include, link or call this file.  Only tests/, bench.py's
// cpu_baseline / --impl reference leg and __graft_entry__.smoke() use it, as the checker.
//
// CPU restatement (C++17, scalar FP64, __float128 where the reference uses real128) of
// MultiFEBE's time-harmonic 3D elastodynamic SBIE assembly path.  Every routine cites the
// reference file:line it follows (paths relative to /root/reference).
//
// PARITY STATUS: "parity unpinned" by the reference itself -- MultiFEBE ships no golden
// vectors / unit tests for this path and no Fortran compiler exists in the
// build container, so the reference binary cannot be run.  The oracle is pinned instead by
// (a) digit-for-digit quadrature tables parsed from the reference's .rc files,
// (b) the analytic solutions of the reference's own tutorials ME-TH-EL-001 / ME-ST-EL-002,
// (c) physics identities (static limit = Kelvin, rigid-body row sums) -- see tests/.
//
// Build: g++ -O2 -ffp-contract=off -fcx-fortran-rules -fopenmp (mimics gfortran
// -march=core2 -O3: no FMA contraction, Fortran complex mul/div rules), see oracle/Makefile.
// =====================================================================================
#include <complex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <quadmath.h>
#include <omp.h>
#include "../data/quad_tables.h"

typedef std::complex<double> cd;
typedef __float128 q128;

// element type ids: lib/fbem/src/shape_functions.f90:198-206
enum { LINE2 = 2, LINE3 = 3, TRI3 = 5, TRI6 = 6, QUAD4 = 7, QUAD8 = 8, QUAD9 = 9 };

// numerical constants: lib/fbem/src/numerical.f90:71-95
static const double c_pi = 3.14159265358979323846264338328;
static const double c_2pi = 6.28318530717958623199592693709;
static const double c_pi_2 = 1.57079632679489661923132169164;
static const double c_pi_4 = 0.78539816339744830961566084582;
static const double c_sqrt2 = 1.41421356237309504880168872421;
static const double c_1_4pi = 0.07957747154594767280411105048;
static const double check_xi_tol = 0.5e-12;  // shape_functions.f90:280

static inline int n_nodes_of(int et) {
  switch (et) { case LINE2: return 2; case LINE3: return 3; case TRI3: return 3; case TRI6: return 6;
                case QUAD4: return 4; case QUAD8: return 8; case QUAD9: return 9; }
  return 0;
}
static inline int n_vertices_of(int et) { return (et == TRI3 || et == TRI6) ? 3 : (et == LINE2 || et == LINE3) ? 2 : 4; }
static inline int n_edges_of(int et) { return (et == TRI3 || et == TRI6) ? 3 : (et == LINE2 || et == LINE3) ? 1 : 4; }
static inline int edge_type_of(int et) { return (et == TRI3 || et == QUAD4 || et == LINE2) ? LINE2 : LINE3; }
// fbem_edge_node(k,edge,etype): shape_functions.f90:841-905 (0-based here)
static inline int edge_node(int k, int edge, int et) {
  if (et == LINE2 || et == LINE3) return k;
  int nv = n_vertices_of(et);
  if (k == 0) return edge;
  if (k == 1) return (edge + 1) % nv;
  return nv + edge;  // mid-edge node
}

// quadrature table accessors (1-based rule n, 0-based point k): quad_rules.f90:88-138
static inline double gl11_x(int n, int k) { return QT_GL11_X[QT_GL11_OFF[n - 1] + k]; }
static inline double gl11_w(int n, int k) { return QT_GL11_W[QT_GL11_OFF[n - 1] + k]; }
static inline double gl01_x(int n, int k) { return QT_GL01_X[QT_GL01_OFF[n - 1] + k]; }
static inline double gl01_w(int n, int k) { return QT_GL01_W[QT_GL01_OFF[n - 1] + k]; }
static inline double gj01_x(int n, int k) { return QT_GJ01_X[QT_GJ01_OFF[n - 1] + k]; }
static inline double gj01_w(int n, int k) { return QT_GJ01_W[QT_GJ01_OFF[n - 1] + k]; }
static inline int wan_n(int order) { return QT_WAN_N[order - 1]; }
static inline double wan_x1(int order, int k) { return QT_WAN_X1[QT_WAN_OFF[order - 1] + k]; }
static inline double wan_x2(int order, int k) { return QT_WAN_X2[QT_WAN_OFF[order - 1] + k]; }
static inline double wan_w(int order, int k) { return QT_WAN_W[QT_WAN_OFF[order - 1] + k]; }

// -------------------------------------------------------------------------------------
// Shape functions (delta = 0, continuous).  T is the type of the "aux"/phi variables
// (double, or __float128 in the nearest-point Newton iteration where the reference
// declares them real128 while xi stays real64: geometry.f90:5118-5123) -- mixed-mode
// promotion then follows the same rules as in the Fortran expressions.
// lib/fbem/src/resources_shape_functions/{phi,dphidxi1,dphidxi2}_{tri3,tri6,quad4,quad8,quad9}.rc
// -------------------------------------------------------------------------------------
template <class T> static void phi2d(int et, const double* xi, T* phi) {
  T a1, a2, a3, a4, a5, a6, a7, a8;
  switch (et) {
    case TRI3:
      a1 = (T)xi[0]; a2 = (T)xi[1];
      phi[0] = a1; phi[1] = a2; phi[2] = (T)1.0 - a1 - a2; break;
    case TRI6:
      a1 = (T)xi[0]; a2 = (T)xi[1]; a3 = (T)1.0 - a1 - a2; a4 = (T)4.0 * a1;
      phi[0] = a1 * ((T)2.0 * a1 - (T)1.0); phi[1] = a2 * ((T)2.0 * a2 - (T)1.0); phi[2] = a3 * ((T)2.0 * a3 - (T)1.0);
      phi[3] = a4 * a2; phi[4] = (T)4.0 * a2 * a3; phi[5] = a4 * a3; break;
    case QUAD4:
      a1 = (T)xi[0]; a2 = (T)xi[1]; a3 = (T)0.25 * ((T)1.0 + a1); a4 = (T)0.25 * ((T)1.0 - a1); a5 = (T)1.0 + a2; a6 = (T)1.0 - a2;
      phi[0] = a4 * a6; phi[1] = a3 * a6; phi[2] = a3 * a5; phi[3] = a4 * a5; break;
    case QUAD8:
      a1 = (T)xi[0]; a2 = (T)xi[1]; a3 = (T)0.25 * ((T)1.0 + a1); a4 = (T)0.25 * ((T)1.0 - a1); a5 = (T)1.0 + a2; a6 = (T)1.0 - a2;
      a7 = (T)1.0 - a1 * a1; a8 = (T)1.0 - a2 * a2;
      phi[0] = a4 * a6 * (-a1 - a5); phi[1] = a3 * a6 * (a1 - a5); phi[2] = a3 * a5 * (a1 - a6); phi[3] = a4 * a5 * (-a1 - a6);
      phi[4] = (T)0.5 * a6 * a7; phi[5] = (T)2.0 * a3 * a8; phi[6] = (T)0.5 * a5 * a7; phi[7] = (T)2.0 * a4 * a8; break;
    case QUAD9:
      a1 = (T)xi[0]; a2 = (T)xi[1]; a3 = (T)0.25 * a1 * (a1 + (T)1.0); a4 = (T)0.25 * a1 * (a1 - (T)1.0);
      a5 = a2 * (a2 + (T)1.0); a6 = a2 * (a2 - (T)1.0); a7 = (T)1.0 - a1 * a1; a8 = (T)1.0 - a2 * a2;
      phi[0] = a4 * a6; phi[1] = a3 * a6; phi[2] = a3 * a5; phi[3] = a4 * a5;
      phi[4] = (T)0.5 * a6 * a7; phi[5] = (T)2.0 * a3 * a8; phi[6] = (T)0.5 * a5 * a7; phi[7] = (T)2.0 * a4 * a8; phi[8] = a7 * a8; break;
  }
}
template <class T> static void dphi2d(int et, const double* xi, T* d1, T* d2) {
  // real64-only sub-expressions (e.g. xi(2)+2.0d0*xi(1)) are evaluated in double before widening, as in Fortran.
  const double x1 = xi[0], x2 = xi[1];
  switch (et) {
    case TRI3:
      d1[0] = 1.0; d1[1] = 0.0; d1[2] = -1.0; d2[0] = 0.0; d2[1] = 1.0; d2[2] = -1.0; break;
    case TRI6: {
      T a1 = (T)(4.0 * x1), a2 = (T)(4.0 * x2), a3 = (T)1.0, a4 = (T)0.0;
      d1[0] = a3 * (a1 - a4 - (T)1.0); d1[1] = 0.0; d1[2] = a3 * (a1 + a2 + a4 - (T)3.0);
      d1[3] = (T)4.0 * a3 * ((T)x2 - a4); d1[4] = -d1[3]; d1[5] = (T)(-4.0) * a3 * (T)(x2 + 2.0 * x1 - 1.0);
      d2[0] = 0.0; d2[1] = a3 * (a2 - a4 - (T)1.0); d2[2] = a3 * (a1 + a2 + a4 - (T)3.0);
      d2[3] = (T)4.0 * a3 * ((T)x1 - a4); d2[4] = (T)(-4.0) * a3 * (T)(2.0 * x2 + x1 - 1.0); d2[5] = -d2[3]; break; }
    case QUAD4: {
      T a1 = (T)1.0, a2 = (T)(0.25 * x2) * a1 * a1, a3 = (T)0.25 * a1;
      d1[0] = a2 - a3; d1[1] = -d1[0]; d1[2] = a2 + a3; d1[3] = -d1[2];
      a2 = (T)(0.25 * x1) * a1 * a1;
      d2[0] = a2 - a3; d2[1] = -a2 - a3; d2[2] = -d2[1]; d2[3] = -d2[0]; break; }
    case QUAD8: {
      T a1 = (T)1.0, a2 = (T)(-1.0), a3 = (T)x2 + a1, a4 = (T)x2 - a1, a5 = (T)(x2 + 2.0 * x1), a6 = (T)(x2 - 2.0 * x1);
      T a7 = (T)0.25 * a2, a8 = a2 * (T)x1;
      d1[0] = a7 * a4 * a5; d1[1] = -a7 * a4 * a6; d1[2] = -a7 * a3 * a5; d1[3] = a7 * a3 * a6;
      d1[4] = -a8 * a4; d1[5] = (T)0.5 * a2 * a3 * a4; d1[6] = a8 * a3; d1[7] = -d1[5];
      a3 = (T)x1 + a1; a4 = (T)x1 - a1; a5 = (T)(2.0 * x2 + x1); a6 = (T)(2.0 * x2 - x1); a8 = a2 * (T)x2;
      d2[0] = a7 * a4 * a5; d2[1] = -a7 * a3 * a6; d2[2] = -a7 * a3 * a5; d2[3] = a7 * a4 * a6;
      d2[4] = (T)(-0.5) * a2 * a3 * a4; d2[5] = a8 * a3; d2[6] = -d2[4]; d2[7] = -a8 * a4; break; }
    case QUAD9: {
      T a1 = (T)1.0, a2 = (T)1.0, a3 = (T)(2.0 * x1) + a1, a4 = (T)(2.0 * x1) - a1, a5 = (T)x2 + a1, a6 = (T)x2 - a1;
      T a7 = (T)0.25 * a2 * (T)x2, a8 = a2 * a5 * a6, a9 = (T)(-0.5) * a8, a10 = -a2 * (T)x1 * (T)x2;
      d1[0] = a7 * a4 * a6; d1[1] = a7 * a3 * a6; d1[2] = a7 * a3 * a5; d1[3] = a7 * a4 * a5;
      d1[4] = a10 * a6; d1[5] = a9 * a3; d1[6] = a10 * a5; d1[7] = a9 * a4; d1[8] = (T)(2.0 * x1) * a8;
      a3 = (T)(2.0 * x2) + a1; a4 = (T)(2.0 * x2) - a1; a5 = (T)x1 + a1; a6 = (T)x1 - a1;
      a7 = (T)0.25 * a2 * (T)x1; a8 = a2 * a5 * a6; a9 = (T)(-0.5) * a8;
      d2[0] = a7 * a6 * a4; d2[1] = a7 * a5 * a4; d2[2] = a7 * a5 * a3; d2[3] = a7 * a6 * a3;
      d2[4] = a9 * a4; d2[5] = a10 * a5; d2[6] = a9 * a3; d2[7] = a10 * a6; d2[8] = (T)(2.0 * x2) * a8; break; }
  }
}
// phi_line2.rc, phi_line3.rc, dphidxi_line2.rc, dphidxi_line3.rc
template <class T> static void phi1d(int et, double xi, T* phi) {
  if (et == LINE2) { T a1 = (T)(0.5 * xi / 1.0); phi[0] = (T)0.5 - a1; phi[1] = (T)0.5 + a1; }
  else { T a1 = (T)(xi / 1.0), a2 = (T)0.5 * a1, a3 = a1 - (T)1.0, a4 = a1 + (T)1.0; phi[0] = a2 * a3; phi[1] = a2 * a4; phi[2] = -a3 * a4; }
}
template <class T> static void dphi1d(int et, double xi, T* d) {
  if (et == LINE2) { T a1 = (T)0.5; d[0] = -a1; d[1] = a1; }
  else { T a1 = (T)1.0, a2 = (T)xi * a1 * a1, a3 = (T)0.5 * a1; d[0] = a2 - a3; d[1] = a2 + a3; d[2] = (T)(-2.0) * a2; }
}
// xi_*_at_node.rc (delta=0)
static void xi_at_node(int et, int node, double* xi) {
  static const double tri[6][2] = {{1, 0}, {0, 1}, {0, 0}, {0.5, 0.5}, {0, 0.5}, {0.5, 0}};
  static const double quad[9][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}, {0, -1}, {1, 0}, {0, 1}, {-1, 0}, {0, 0}};
  if (et == TRI3 || et == TRI6) { xi[0] = tri[node][0]; xi[1] = tri[node][1]; }
  else if (et == LINE2 || et == LINE3) { static const double l[3] = {-1, 1, 0}; xi[0] = l[node]; }
  else { xi[0] = quad[node][0]; xi[1] = quad[node][1]; }
}
// fbem_check_xi1xi2: shape_functions.f90:1037-1062
static bool check_xi1xi2(int et, const double* xi) {
  if (et == TRI3 || et == TRI6) return !((xi[0] < 0.0 - check_xi_tol) || (xi[1] < 0.0 - check_xi_tol) || ((xi[0] + xi[1]) > 1.0 + check_xi_tol));
  return !((xi[0] < -1.0 - check_xi_tol) || (xi[0] > 1.0 + check_xi_tol) || (xi[1] < -1.0 - check_xi_tol) || (xi[1] > 1.0 + check_xi_tol));
}
// fbem_check_xi1xi2_edge: shape_functions.f90:1111-1147
static bool check_xi1xi2_edge(int et, const double* xi) {
  if (!check_xi1xi2(et, xi)) return false;
  if (et == TRI3 || et == TRI6) return (xi[0] < 0.0 + check_xi_tol) || (xi[1] < 0.0 + check_xi_tol) || ((xi[0] + xi[1]) > 1.0 - check_xi_tol);
  return (xi[0] < -1.0 + check_xi_tol) || (xi[0] > 1.0 - check_xi_tol) || (xi[1] < -1.0 + check_xi_tol) || (xi[1] > 1.0 - check_xi_tol);
}

static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// -------------------------------------------------------------------------------------
// fbem_zexp_decomposed: lib/fbem/src/numerical.f90:1258-1330
// -------------------------------------------------------------------------------------
static void zexp_decomposed(cd z, cd* E /*0..6*/) {
  double absz = std::abs(z);
  if (absz <= 1.0) {
    int n;
    if (absz <= 1.e-6) n = 1;
    else n = (int)lround(pow(10.0, 1.176 + 0.175 * log10(absz)));
    cd term[28];
    term[0] = cd(1.0, 0.0); term[1] = z;
    double factorial = 1.0; cd zpowk = z;
    for (int k = 2; k <= n + 6; k++) { factorial = factorial * (double)k; zpowk = zpowk * z; term[k] = zpowk / factorial; }
    E[6] = cd(0.0, 0.0);
    for (int k = n + 6; k >= 6; k--) E[6] = E[6] + term[k];
    E[5] = E[6] + term[5]; E[4] = E[5] + term[4]; E[3] = E[4] + term[3];
    E[2] = E[3] + term[2]; E[1] = E[2] + term[1]; E[0] = E[1] + term[0];
  } else {
    cd expz = std::exp(z), z2 = z * z, z3 = z2 * z, z4 = z3 * z, z5 = z4 * z;
    E[0] = expz; E[1] = expz - 1.0; E[2] = E[1] - z; E[3] = E[2] - 0.5 * z2;
    E[4] = E[3] - (1.0 / 6.0) * z3; E[5] = E[4] - (1.0 / 24.0) * z4; E[6] = E[5] - (1.0 / 120.0) * z5;
  }
}

// -------------------------------------------------------------------------------------
// fbem_bem_harela3d_parameters / _calculate_parameters: lib/fbem/src/bem_harela3d.f90:118-287
// (SBIE subset: psi, chi, T1..T3, cte_u, cte_t; 1-based arrays kept for readability)
// -------------------------------------------------------------------------------------
struct Params {
  cd lambda, mu; double rho, omega; cd c1, c2, k1, k2;
  cd psi[7], chi[7], T1[11], T2[10], T3[10]; cd cte_u, cte_t;
  cd S1[12], S2[12], S3[13], S4[11], S5[12]; cd cte_d, cte_s;   // hypersingular equation (d*, s*): bem_harela3d.f90:219-287
  // static elasticity (lib/fbem/src/bem_staela3d.f90): Kelvin solution, cte_u = cteu1, cte_t = ctet1 of :617-622
  bool statics = false; double cteu2 = 0.0, ctet2 = 0.0, nu_s = 0.0, ctes3 = 0.0;
  // scalar wave propagation (inviscid fluid / acoustics), lib/fbem/src/bem_harpot3d.f90:102-115: wavenumber k, P(1), Q(1:2).
  // h, g of a pair are vectors over the element nodes: stored in h[0..nn), g[0..nn) of the 9*nn containers (the rest stays zero);
  // cte_t = -1/(4 pi), cte_u = 1/(4 pi) carry the reference's `h=-h*c_1_4pi; g=g*c_1_4pi`.
  bool pot = false; cd kp, P1, Q1, Q2;
  // Biot poroelastic medium (PorParams below): h, g are (n, 4, 4)
  const struct PorParams* por = nullptr;
};
// fbem_bem_harpot3d_calculate_parameters: lib/fbem/src/bem_harpot3d.f90:123-165 (SBIE subset P, Q)
static void calculate_parameters_pot(double rho, cd c, double omega, Params& p) {
  const cd im(0.0, 1.0);
  cd k = omega / c;
  p.pot = true; p.omega = omega; p.rho = rho; p.c1 = c; p.kp = k;
  p.P1 = -im * k; p.Q1 = 0.5 * (k * k); p.Q2 = im * k;
  p.cte_u = c_1_4pi; p.cte_t = -c_1_4pi;
}
// -------------------------------------------------------------------------------------
// Biot poroelastic medium: fbem_bem_harpor3d_parameters / _calculate_parameters, lib/fbem/src/bem_harpor3d.f90:91-134, 204-568 (SBIE subset:
// eta, vartheta, psi, chi, W0, T01, T02, W1, W2, T1..T3, cte_u, cte_t; 1-based like the reference).  Variables of a node: 0 = fluid
// equivalent stress tau (primary) / fluid normal displacement Un (secondary), 1..3 = solid displacement u_k / solid traction t_k.
// h, g of a pair are (n, 4, 4): stored as h[(j*4 + il)*4 + ik].
// -------------------------------------------------------------------------------------
struct PorParams {
  cd lambda, mu, nu, R, Q; double rho1, rho2, rhoa, b, omega;
  cd rhohat11, rhohat22, rhohat12, Z, J, k1, k2, k3;
  cd eta[4], vartheta[6], psi[9], chi[10], W0[7], T01[8], T02[9], W1[11], W2[10], T1[15], T2[13], T3[14];
  cd cte_u[4][4], cte_t[4][4];
};
static void calculate_parameters_por(cd lambda, cd mu, double rho1, double rho2, double rhoa, cd R, cd Q, double b, double omega, PorParams& P) {
  const cd im(0.0, 1.0);
  cd rhohat11 = rho1 + rhoa - im * b / omega, rhohat12 = -rhoa + im * b / omega, rhohat22 = rho2 + rhoa - im * b / omega;
  cd J = 1.0 / (rhohat22 * (omega * omega)), Z = rhohat12 / rhohat22;
  cd k3 = std::sqrt(1.0 / (mu * J) * (rhohat11 / rhohat22 - Z * Z));
  if (k3.real() < 0.0) k3 = -k3;
  cd ca = lambda + 2.0 * mu;
  cd cb = (lambda + 2.0 * mu) / (J * R) + mu * (k3 * k3) + 1.0 / J * ((Q / R - Z) * (Q / R - Z));
  cd cc = mu / (J * R) * (k3 * k3);
  cd k1 = std::sqrt(0.5 * (cb - std::sqrt(cb * cb - 4.0 * ca * cc)) / ca), k2 = std::sqrt(0.5 * (cb + std::sqrt(cb * cb - 4.0 * ca * cc)) / ca);
  if (k1.real() < 0.0) k1 = -k1;
  if (k2.real() < 0.0) k2 = -k2;
  if (k1.real() > k2.real()) std::swap(k1, k2);
  P.omega = omega; P.lambda = lambda; P.mu = mu; P.nu = 0.5 * lambda / (lambda + mu); P.rho1 = rho1; P.rho2 = rho2; P.rhoa = rhoa; P.R = R; P.Q = Q; P.b = b;
  P.rhohat11 = rhohat11; P.rhohat22 = rhohat22; P.rhohat12 = rhohat12; P.Z = Z; P.J = J; P.k1 = k1; P.k2 = k2; P.k3 = k3;
  const cd l2m = lambda + 2.0 * mu, k1_2 = k1 * k1, k2_2 = k2 * k2, k3_2 = k3 * k3, k1_3 = k1_2 * k1, k2_3 = k2_2 * k2;
  cd alpha1 = k1_2 - mu / l2m * k3_2, alpha2 = k2_2 - mu / l2m * k3_2;
  cd beta1 = mu / l2m * k1_2 - k1_2 * k2_2 / k3_2, beta2 = mu / l2m * k2_2 - k1_2 * k2_2 / k3_2;
  cd vc = (Q / R - Z) / l2m, d12 = k1_2 - k2_2;
  const cd ik1 = im * k1, ik2 = im * k2, ik3 = im * k3, QR = Q / R;
  P.eta[1] = -(ik1 * alpha1 - ik2 * alpha2) / d12; P.eta[2] = alpha1 / d12; P.eta[3] = -alpha2 / d12;
  P.vartheta[1] = 0.5 * vc; P.vartheta[2] = vc * ik1 / d12; P.vartheta[3] = -vc * ik2 / d12; P.vartheta[4] = vc / d12; P.vartheta[5] = -vc / d12;
  P.psi[1] = 0.5 * (lambda + 3.0 * mu) / l2m; P.psi[2] = -1.0 / 3.0 * ((ik1 * beta1 - ik2 * beta2) / d12 + 2.0 * ik3);
  P.psi[3] = -beta1 / ik1 / d12; P.psi[4] = beta2 / ik2 / d12; P.psi[5] = 1.0 / ik3; P.psi[6] = beta1 / k1_2 / d12; P.psi[7] = -beta2 / k2_2 / d12; P.psi[8] = -1.0 / k3_2;
  P.chi[1] = -0.5 * (lambda + mu) / l2m; P.chi[2] = -beta1 / d12; P.chi[3] = beta2 / d12; P.chi[4] = -3.0 * beta1 / ik1 / d12; P.chi[5] = 3.0 * beta2 / ik2 / d12;
  P.chi[6] = 3.0 / ik3; P.chi[7] = 3.0 * beta1 / k1_2 / d12; P.chi[8] = -3.0 * beta2 / k2_2 / d12; P.chi[9] = -3.0 / k3_2;
  P.W0[1] = J; P.W0[2] = 0.5 * (Z * vc + J * (k1_2 * alpha1 - k2_2 * alpha2) / d12); P.W0[3] = (Z * vc + J * alpha1) * ik1 / d12;
  P.W0[4] = -(Z * vc + J * alpha2) * ik2 / d12; P.W0[5] = (Z * vc + J * alpha1) / d12; P.W0[6] = -(Z * vc + J * alpha2) / d12;
  P.T01[1] = mu * vc; P.T01[2] = -2.0 * mu * vc * k1_2 / d12; P.T01[3] = 2.0 * mu * vc * k2_2 / d12; P.T01[4] = 6.0 * mu * vc * ik1 / d12;
  P.T01[5] = -6.0 * mu * vc * ik2 / d12; P.T01[6] = 6.0 * mu * vc / d12; P.T01[7] = -6.0 * mu * vc / d12;
  P.T02[1] = mu * vc + Z; P.T02[2] = (lambda + 2.0 / 3.0 * mu) * vc * im * (k1_3 - k2_3) / d12 - QR * (ik1 * alpha1 - ik2 * alpha2) / d12;
  P.T02[3] = (QR * alpha1 - lambda * vc * k1_2) / d12; P.T02[4] = -(QR * alpha2 - lambda * vc * k2_2) / d12; P.T02[5] = -2.0 * mu * vc * ik1 / d12;
  P.T02[6] = 2.0 * mu * vc * ik2 / d12; P.T02[7] = -2.0 * mu * vc / d12; P.T02[8] = 2.0 * mu * vc / d12;
  P.W1[1] = 0.5 * (QR * mu / l2m - Z); P.W1[2] = -(mu * vc * k1_2 + Z * beta1) / d12; P.W1[3] = (mu * vc * k2_2 + Z * beta2) / d12; P.W1[4] = Z;
  P.W1[5] = 3.0 * (mu * vc * ik1 - Z * beta1 / ik1) / d12; P.W1[6] = -3.0 * (mu * vc * ik2 - Z * beta2 / ik2) / d12; P.W1[7] = 3.0 * Z / ik3;
  P.W1[8] = 3.0 * (mu * vc + Z * beta1 / k1_2) / d12; P.W1[9] = -3.0 * (mu * vc + Z * beta2 / k2_2) / d12; P.W1[10] = -3.0 * Z / k3_2;
  P.W2[1] = -0.5 * (QR * mu / l2m + Z); P.W2[2] = 1.0 / 3.0 * (mu * vc * im * (k1_3 - k2_3) / d12 + Z * ((ik1 * beta1 - ik2 * beta2) / d12 + 2.0 * ik3));
  P.W2[3] = -Z; P.W2[4] = -(mu * vc * ik1 - Z * beta1 / ik1) / d12; P.W2[5] = (mu * vc * ik2 - Z * beta2 / ik2) / d12; P.W2[6] = -Z / ik3;
  P.W2[7] = -(mu * vc + Z * beta1 / k1_2) / d12; P.W2[8] = (mu * vc + Z * beta2 / k2_2) / d12; P.W2[9] = Z / k3_2;
  const cd kb = (k1_2 * beta1 - k2_2 * beta2) / d12;
  P.T1[1] = -3.0 * (lambda + mu) / l2m; P.T1[2] = 0.25 * (kb - k3_2); P.T1[3] = -2.0 * ik1 * beta1 / d12; P.T1[4] = 2.0 * ik2 * beta2 / d12; P.T1[5] = 2.0 * ik3;
  P.T1[6] = -12.0 * beta1 / d12; P.T1[7] = 12.0 * beta2 / d12; P.T1[8] = 12.0; P.T1[9] = -30.0 * beta1 / ik1 / d12; P.T1[10] = 30.0 * beta2 / ik2 / d12;
  P.T1[11] = 30.0 / ik3; P.T1[12] = 30.0 * beta1 / k1_2 / d12; P.T1[13] = -30.0 * beta2 / k2_2 / d12; P.T1[14] = -30.0 / k3_2;
  P.T2[1] = -mu / l2m; P.T2[2] = -0.25 * (kb + k3_2); P.T2[3] = -ik3; P.T2[4] = 2.0 * beta1 / d12; P.T2[5] = -2.0 * beta2 / d12; P.T2[6] = -3.0;
  P.T2[7] = 6.0 * beta1 / ik1 / d12; P.T2[8] = -6.0 * beta2 / ik2 / d12; P.T2[9] = -6.0 / ik3; P.T2[10] = -6.0 * beta1 / k1_2 / d12; P.T2[11] = 6.0 * beta2 / k2_2 / d12;
  P.T2[12] = 6.0 / k3_2;
  const cd qv = QR * vc / J, lm = lambda / mu;
  P.T3[1] = mu / l2m; P.T3[2] = 0.25 * (2.0 * qv - (2.0 * lm + 1.0) * kb + k3_2); P.T3[3] = (qv - lm * beta1) * ik1 / d12; P.T3[4] = -(qv - lm * beta2) * ik2 / d12;
  P.T3[5] = (qv - (lm - 2.0) * beta1) / d12; P.T3[6] = -(qv - (lm - 2.0) * beta2) / d12; P.T3[7] = -2.0; P.T3[8] = 6.0 * beta1 / ik1 / d12;
  P.T3[9] = -6.0 * beta2 / ik2 / d12; P.T3[10] = -6.0 / ik3; P.T3[11] = -6.0 * beta1 / k1_2 / d12; P.T3[12] = 6.0 * beta2 / k2_2 / d12; P.T3[13] = 6.0 / k3_2;
  for (int a = 0; a < 4; a++) for (int c = 0; c < 4; c++) {
    P.cte_u[a][c] = (a == 0) ? cd(-c_1_4pi) : (c == 0 ? -c_1_4pi / J : c_1_4pi / mu);
    P.cte_t[a][c] = (a == 0) ? (c == 0 ? cd(-c_1_4pi) : cd(c_1_4pi)) : (c == 0 ? -c_1_4pi / mu : cd(c_1_4pi));
  }
}
// Kernel scalars of the poroelastic fundamental solution at distance r (bem_harpor3d.f90:944-975); regular_only drops the static 1/r^2 parts
// of W0 and TT1..TT3 as the interior integration does (:1720-1750), which adds them back where it integrates them in full.
struct PorScalars { cd eta, vartheta, psi, chi, W0, T01, T02, W1, W2, TT1, TT2, TT3; double d1r1, d1r2; };
static inline void por_scalars(const PorParams& p, double r, bool regular_only, PorScalars& k) {
  double d1r1 = 1.0 / r, d1r2 = d1r1 * d1r1, d1r3 = d1r2 * d1r1, d1r4 = d1r3 * d1r1;
  const cd mim(-0.0, -1.0);
  cd E[3][7];
  zexp_decomposed(mim * p.k1 * r, E[0]); zexp_decomposed(mim * p.k2 * r, E[1]); zexp_decomposed(mim * p.k3 * r, E[2]);
  cd E2[3], E3[3], E4[3], E5[3];
  for (int j = 0; j < 3; j++) { E2[j] = E[j][2] * d1r1; E3[j] = E[j][3] * d1r2; E4[j] = E[j][4] * d1r3; E5[j] = E[j][5] * d1r4; }
  k.eta = d1r1 + p.eta[1] + p.eta[2] * E2[0] + p.eta[3] * E2[1];
  k.vartheta = p.vartheta[1] + p.vartheta[2] * E2[0] + p.vartheta[3] * E2[1] + p.vartheta[4] * E3[0] + p.vartheta[5] * E3[1];
  k.psi = p.psi[1] * d1r1 + p.psi[2] + E2[2] + p.psi[3] * E3[0] + p.psi[4] * E3[1] + p.psi[5] * E3[2] + p.psi[6] * E4[0] + p.psi[7] * E4[1] + p.psi[8] * E4[2];
  k.chi = p.chi[1] * d1r1 + p.chi[2] * E2[0] + p.chi[3] * E2[1] + E2[2] + p.chi[4] * E3[0] + p.chi[5] * E3[1] + p.chi[6] * E3[2] + p.chi[7] * E4[0] + p.chi[8] * E4[1]
        + p.chi[9] * E4[2];
  cd W0r = p.W0[2] + p.W0[3] * E2[0] + p.W0[4] * E2[1] + p.W0[5] * E3[0] + p.W0[6] * E3[1];
  k.T01 = p.T01[1] * d1r1 + p.T01[2] * E2[0] + p.T01[3] * E2[1] + p.T01[4] * E3[0] + p.T01[5] * E3[1] + p.T01[6] * E4[0] + p.T01[7] * E4[1];
  k.T02 = p.T02[1] * d1r1 + p.T02[2] + p.T02[3] * E2[0] + p.T02[4] * E2[1] + p.T02[5] * E3[0] + p.T02[6] * E3[1] + p.T02[7] * E4[0] + p.T02[8] * E4[1];
  k.W1 = p.W1[1] * d1r1 + p.W1[2] * E2[0] + p.W1[3] * E2[1] + p.W1[4] * E2[2] + p.W1[5] * E3[0] + p.W1[6] * E3[1] + p.W1[7] * E3[2] + p.W1[8] * E4[0] + p.W1[9] * E4[1]
       + p.W1[10] * E4[2];
  k.W2 = p.W2[1] * d1r1 + p.W2[2] + p.W2[3] * E2[2] + p.W2[4] * E3[0] + p.W2[5] * E3[1] + p.W2[6] * E3[2] + p.W2[7] * E4[0] + p.W2[8] * E4[1] + p.W2[9] * E4[2];
  cd T1r = p.T1[2] + p.T1[3] * E2[0] + p.T1[4] * E2[1] + p.T1[5] * E2[2] + p.T1[6] * E3[0] + p.T1[7] * E3[1] + p.T1[8] * E3[2] + p.T1[9] * E4[0] + p.T1[10] * E4[1]
           + p.T1[11] * E4[2] + p.T1[12] * E5[0] + p.T1[13] * E5[1] + p.T1[14] * E5[2];
  cd T2r = p.T2[2] + p.T2[3] * E2[2] + p.T2[4] * E3[0] + p.T2[5] * E3[1] + p.T2[6] * E3[2] + p.T2[7] * E4[0] + p.T2[8] * E4[1] + p.T2[9] * E4[2] + p.T2[10] * E5[0]
           + p.T2[11] * E5[1] + p.T2[12] * E5[2];
  cd T3r = p.T3[2] + p.T3[3] * E2[0] + p.T3[4] * E2[1] + p.T3[5] * E3[0] + p.T3[6] * E3[1] + p.T3[7] * E3[2] + p.T3[8] * E4[0] + p.T3[9] * E4[1] + p.T3[10] * E4[2]
           + p.T3[11] * E5[0] + p.T3[12] * E5[1] + p.T3[13] * E5[2];
  if (regular_only) { k.W0 = W0r; k.TT1 = T1r; k.TT2 = T2r; k.TT3 = T3r; }
  else { k.W0 = p.W0[1] * d1r2 + W0r; k.TT1 = p.T1[1] * d1r2 + T1r; k.TT2 = p.T2[1] * d1r2 + T2r; k.TT3 = p.T3[1] * d1r2 + T3r; }
  k.d1r1 = d1r1; k.d1r2 = d1r2;
}
// order f of the estimator's model function 1/r^f: 5 for the elastic SBIE (bem_harela3d.f90:1522), 7 for the elastic HBIE (:3677),
// 3 for the scalar SBIE (bem_harpot3d.f90:1009, :693)
static inline int estimator_f(const Params& p, const double* n_i) { return n_i ? 7 : (p.pot ? 3 : 5); }   // poroelastic SBIE: 5 (bem_harpor3d.f90:1949)
static inline int nblk(const Params& p) { return p.por ? 16 : 9; }   // entries of one node block of h, g
// constants and orientation of a finished pair: `h = cte_t h, g = cte_u g; if (reverse) h = -h` of every integrator
static inline void finish_pair(const Params& p, bool reverse, const double* n_i, int nn, cd* h, cd* g) {
  if (p.por) {
    for (int j = 0; j < nn; j++) for (int a = 0; a < 4; a++) for (int c = 0; c < 4; c++) {
      h[(j * 4 + a) * 4 + c] = p.por->cte_t[a][c] * h[(j * 4 + a) * 4 + c]; g[(j * 4 + a) * 4 + c] = p.por->cte_u[a][c] * g[(j * 4 + a) * 4 + c]; }
  } else {
    for (int i = 0; i < 9 * nn; i++) { h[i] = (n_i ? p.cte_s : p.cte_t) * h[i]; g[i] = (n_i ? p.cte_d : p.cte_u) * g[i]; }
  }
  if (reverse) for (int i = 0; i < nblk(p) * nn; i++) h[i] = -h[i];
}
// fbem_decomposed_zexp: lib/fbem/src/numerical.f90:1138-1191 (E0..E4; used by fbem_bem_harpot3d_sbie_int)
static void decomposed_zexp(cd z, cd* E /*0..4*/) {
  double absz = std::abs(z);
  if (absz <= 1.0) {
    int n;
    if (absz <= 1.e-6) n = 1;
    else n = (int)lround(pow(10.0, 1.176 + 0.175 * log10(absz)));
    cd term[20];
    term[0] = cd(1.0, 0.0); term[1] = z;
    double factorial = 1.0; cd zpowk = z;
    for (int k = 2; k <= n + 4; k++) { factorial = factorial * (double)k; zpowk = zpowk * z; term[k] = zpowk / factorial; }
    E[4] = cd(0.0, 0.0);
    for (int k = n + 4; k >= 4; k--) E[4] = E[4] + term[k];
    E[3] = E[4] + term[3]; E[2] = E[3] + term[2]; E[1] = E[2] + term[1]; E[0] = E[1] + term[0];
  } else {
    cd expz = std::exp(z), z2 = z * z, z3 = z2 * z;
    E[0] = expz; E[1] = expz - 1.0; E[2] = E[1] - z; E[3] = E[2] - 0.5 * z2; E[4] = E[3] - 0.166666666666666667 * z3;
  }
}
// constants of fbem_bem_staela3d_sbie_ext_pre / _ext_st / _int: bem_staela3d.f90:617-622
static void calculate_parameters_static(double mu, double nu, Params& p) {
  p.statics = true; p.mu = mu; p.lambda = 2.0 * mu * nu / (1.0 - 2.0 * nu); p.rho = 0.0; p.omega = 0.0;
  p.cte_u = 1.0 / (16.0 * c_pi * mu * (1.0 - nu)); p.cteu2 = 3.0 - 4.0 * nu;
  p.cte_t = -1.0 / (8.0 * c_pi * (1.0 - nu)); p.ctet2 = 1.0 - 2.0 * nu;
  // hypersingular equation, fbem_bem_staela3d_hbie_ext_pre (bem_staela3d.f90:4408-4412): cted1 = cte_t, cted2 = ctes2 = ctet2
  p.cte_d = p.cte_t; p.cte_s = mu / (4.0 * c_pi * (1.0 - nu)); p.nu_s = nu; p.ctes3 = 1.0 - 4.0 * nu;
}
static void calculate_parameters(cd lambda, cd mu, double rho, double omega, Params& p) {
  const cd im(0.0, 1.0);
  cd c1 = std::sqrt((lambda + 2.0 * mu) / rho), c2 = std::sqrt(mu / rho);
  cd k1 = omega / c1, k2 = omega / c2;
  p.omega = omega; p.lambda = lambda; p.mu = mu; p.rho = rho; p.c1 = c1; p.c2 = c2; p.k1 = k1; p.k2 = k2;
  cd c1_2 = c1 * c1, c2_2 = c2 * c2, c1_3 = c1_2 * c1, c1_4 = c1_2 * c1_2, k2_2 = k2 * k2;
  cd ik1 = im * k1, ik2 = im * k2, ik1_2 = ik1 * ik1, ik2_2 = ik2 * ik2;
  cd r = c2_2 / c1_2;  // c2**2/c1**2
  double om2 = omega * omega;
  p.psi[1] = 0.5 * (1.0 + r);
  p.psi[2] = -1.0 / 3.0 * (2.0 / c2 + c2_2 / c1_3) * im * omega;
  p.psi[3] = im * k1 / k2_2; p.psi[4] = 1.0 / ik2; p.psi[5] = 1.0 / k2_2; p.psi[6] = -1.0 / k2_2;
  p.chi[1] = -0.5 * (1.0 - r); p.chi[2] = -r; p.chi[3] = -3.0 * c2_2 / c1_2 / ik1; p.chi[4] = 3.0 / ik2;
  p.chi[5] = -3.0 * c2_2 / c1_2 / ik1_2; p.chi[6] = 3.0 / ik2_2;
  p.T1[1] = 3.0 * (r - 1.0); p.T1[2] = -0.25 * (1.0 / c2_2 - c2_2 / c1_4) * om2;
  p.T1[3] = -2.0 * im * k1 * c2_2 / c1_2; p.T1[4] = 2.0 * im * k2; p.T1[5] = -12.0 * c2_2 / c1_2; p.T1[6] = 12.0;
  p.T1[7] = -r * 30.0 / ik1; p.T1[8] = 30.0 / ik2; p.T1[9] = -r * 30.0 / ik1_2; p.T1[10] = 30.0 / ik2_2;
  p.T2[1] = -r; p.T2[2] = -0.25 * (1.0 / c2_2 + c2_2 / c1_4) * om2; p.T2[3] = -im * k2; p.T2[4] = 2.0 * c2_2 / c1_2;
  p.T2[5] = -3.0; p.T2[6] = r * 6.0 / ik1; p.T2[7] = -6.0 / ik2; p.T2[8] = r * 6.0 / ik1_2; p.T2[9] = -6.0 / ik2_2;
  p.T3[1] = r; p.T3[2] = 0.25 * (3.0 * c2_2 / c1_4 - 2.0 / c1_2 + 1.0 / c2_2) * om2;
  p.T3[3] = (2.0 * c2_2 / c1_2 - 1.0) * im * k1; p.T3[4] = 4.0 * c2_2 / c1_2 - 1.0; p.T3[5] = -2.0;
  p.T3[6] = r * 6.0 / ik1; p.T3[7] = -6.0 / ik2; p.T3[8] = r * 6.0 / ik1_2; p.T3[9] = -6.0 / ik2_2;
  p.cte_u = c_1_4pi / mu; p.cte_t = c_1_4pi;
  // S1..S5, cte_d, cte_s: bem_harela3d.f90:219-287
  cd k1_2 = k1 * k1, c1_5 = c1_4 * c1, c2_3 = c2_2 * c2, om3 = om2 * omega;
  p.S1[1] = 3.0 * (1.0 - 2.0 * c2_2 / c1_2); p.S1[2] = -0.5 * c2_2 / c1_4 * om2; p.S1[3] = k2_2; p.S1[4] = 4.0 * im * k1 * k1_2 / k2_2;
  p.S1[5] = -7.0 * im * k2; p.S1[6] = 24.0 * k1_2 / k2_2; p.S1[7] = -27.0; p.S1[8] = 60.0 * k1_2 / k2_2 / ik1; p.S1[9] = -60.0 / ik2;
  p.S1[10] = 60.0 * k1_2 / k2_2 / ik1_2; p.S1[11] = -60.0 / ik2_2;
  p.S2[1] = 6.0 * c2_2 / c1_2; p.S2[2] = (0.5 / c2_2 + 1.5 * c2_2 / c1_4 - 1.0 / c1_2) * om2; p.S2[3] = 2.0 * c2_2 / c1_4 * (c1_2 / c2_2 - 2.0) * om2;
  p.S2[4] = 2.0 * im * c2_2 / c1_3 * (8.0 - 3.0 * c1_2 / c2_2) * omega; p.S2[5] = -4.0 * im * k2; p.S2[6] = 6.0 * c2_2 / c1_2 * (6.0 - c1_2 / c2_2);
  p.S2[7] = -24.0; p.S2[8] = c2_2 / c1_2 * 60.0 / ik1; p.S2[9] = -60.0 / ik2; p.S2[10] = 60.0 / ik2_2; p.S2[11] = -60.0 / ik2_2;
  p.S3[1] = 30.0 * (1.0 - c2_2 / c1_2); p.S3[2] = 1.5 * (1.0 / c2_2 - c2_2 / c1_4) * om2; p.S3[3] = -4.0 * c2_2 / c1_2 * k1_2; p.S3[4] = 4.0 * k2_2;
  p.S3[5] = 40.0 * c2_2 / c1_2 * im * k1; p.S3[6] = -40.0 * im * k2; p.S3[7] = 180.0 * c2_2 / c1_2; p.S3[8] = -180.0;
  p.S3[9] = 420.0 * c2_2 / c1_2 / ik1; p.S3[10] = -420.0 / ik2; p.S3[11] = 420.0 / ik2_2; p.S3[12] = -420.0 / ik2_2;
  p.S4[1] = 2.0 * c2_2 / c1_2; p.S4[2] = 0.5 * (c2_2 / c1_4 + 1.0 / c2_2) * om2; p.S4[3] = -2.0 / 5.0 * (1.0 / c2_3 + 2.0 / 3.0 * c2_2 / c1_5) * im * om3;
  p.S4[4] = 2.0 * im * k2; p.S4[5] = -4.0 * c2_2 / c1_2; p.S4[6] = 6.0; p.S4[7] = -12.0 * c2_2 / c1_2 / ik1; p.S4[8] = 12.0 / ik2;
  p.S4[9] = -12.0 / ik2_2; p.S4[10] = 12.0 / ik2_2;
  p.S5[1] = 2.0 * (1.0 - 3.0 * c2_2 / c1_2); p.S5[2] = (-2.0 / c1_2 + 0.5 / c2_2 + 0.5 * c2_2 / c1_4) * om2;
  p.S5[3] = (8.0 / 3.0 / c1_3 + 4.0 / 15.0 / c2_3 - 1.0 / c1 / c2_2 - 24.0 / 15.0 * c2_2 / c1_5) * im * om3;
  p.S5[4] = (-4.0 / c1_2 + 1.0 / c2_2 + 4.0 * c2_2 / c1_4) * om2; p.S5[5] = 4.0 * im * (1.0 / c1 - 2.0 * c2_2 / c1_3) * omega;
  p.S5[6] = 4.0 * (1.0 - 3.0 * c2_2 / c1_2); p.S5[7] = 4.0; p.S5[8] = 12.0 * im * k1 / k2_2; p.S5[9] = -12.0 * im / k2; p.S5[10] = 12.0 / k2_2; p.S5[11] = -12.0 / k2_2;
  p.cte_d = c_1_4pi; p.cte_s = c_1_4pi * mu;
}

// Kernel scalars psi, chi, TT1..3 at distance r (bem_harela3d.f90:663-685).  `regular_only` drops the
// static 1/r^2 parts of TT1..3 as in the interior integration (:1394-1401).
struct KernelScalars { cd psi, chi, TT1, TT2, TT3; double d1r1, d1r2; };
static inline void kernel_scalars(const Params& p, double r, bool regular_only, KernelScalars& k) {
  double d1r1 = 1.0 / r, d1r2 = d1r1 * d1r1, d1r3 = d1r2 * d1r1, d1r4 = d1r3 * d1r1;
  const cd mim(-0.0, -1.0);
  cd z[2] = {mim * p.k1 * r, mim * p.k2 * r};
  cd E1[7], E2[7];
  zexp_decomposed(z[0], E1); zexp_decomposed(z[1], E2);
  cd E21 = E1[2] * d1r1, E22 = E2[2] * d1r1, E31 = E1[3] * d1r2, E32 = E2[3] * d1r2;
  cd E41 = E1[4] * d1r3, E42 = E2[4] * d1r3, E51 = E1[5] * d1r4, E52 = E2[5] * d1r4;
  k.psi = p.psi[1] * d1r1 + p.psi[2] + E22 + p.psi[3] * E31 + p.psi[4] * E32 + p.psi[5] * E41 + p.psi[6] * E42;
  k.chi = p.chi[1] * d1r1 + p.chi[2] * E21 + E22 + p.chi[3] * E31 + p.chi[4] * E32 + p.chi[5] * E41 + p.chi[6] * E42;
  if (!regular_only) {
    k.TT1 = p.T1[1] * d1r2 + p.T1[2] + p.T1[3] * E21 + p.T1[4] * E22 + p.T1[5] * E31 + p.T1[6] * E32 + p.T1[7] * E41 + p.T1[8] * E42 + p.T1[9] * E51 + p.T1[10] * E52;
    k.TT2 = p.T2[1] * d1r2 + p.T2[2] + p.T2[3] * E22 + p.T2[4] * E31 + p.T2[5] * E32 + p.T2[6] * E41 + p.T2[7] * E42 + p.T2[8] * E51 + p.T2[9] * E52;
    k.TT3 = p.T3[1] * d1r2 + p.T3[2] + p.T3[3] * E21 + p.T3[4] * E31 + p.T3[5] * E32 + p.T3[6] * E41 + p.T3[7] * E42 + p.T3[8] * E51 + p.T3[9] * E52;
  } else {
    k.TT1 = p.T1[2] + p.T1[3] * E21 + p.T1[4] * E22 + p.T1[5] * E31 + p.T1[6] * E32 + p.T1[7] * E41 + p.T1[8] * E42 + p.T1[9] * E51 + p.T1[10] * E52;
    k.TT2 = p.T2[2] + p.T2[3] * E22 + p.T2[4] * E31 + p.T2[5] * E32 + p.T2[6] * E41 + p.T2[7] * E42 + p.T2[8] * E51 + p.T2[9] * E52;
    k.TT3 = p.T3[2] + p.T3[3] * E21 + p.T3[4] * E31 + p.T3[5] * E32 + p.T3[6] * E41 + p.T3[7] * E42 + p.T3[8] * E51 + p.T3[9] * E52;
  }
  k.d1r1 = d1r1; k.d1r2 = d1r2;
}
static const double dkr[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};

// One exterior Gauss point: h(:,l,k)+=fs_t*pphijw, g(:,l,k)+=fs_u*sphijw  (bem_harela3d.f90:663-693)
static inline void add_exterior_point(const Params& p, const double* x, const double* n, const double* x_i, int nn,
                                      const double* pphijw, const double* sphijw, cd* h, cd* g) {
  double rv[3] = {x[0] - x_i[0], x[1] - x_i[1], x[2] - x_i[2]};
  double r = sqrt(dot3(rv, rv));
  if (p.por) {   // fbem_bem_harpor3d_sbie_ext_pre: bem_harpor3d.f90:931-996 (the same point formula in _ext_st)
    PorScalars k; por_scalars(*p.por, r, false, k);
    double drdx[3] = {rv[0] * k.d1r1, rv[1] * k.d1r1, rv[2] * k.d1r1};
    double drdn = dot3(drdx, n);
    cd fs_u[4][4], fs_t[4][4];
    fs_u[0][0] = k.eta; fs_t[0][0] = k.W0 * drdn;
    for (int c = 0; c < 3; c++) {
      fs_u[0][c + 1] = k.vartheta * drdx[c]; fs_u[c + 1][0] = k.vartheta * drdx[c];
      fs_t[0][c + 1] = k.T01 * drdx[c] * drdn + k.T02 * n[c]; fs_t[c + 1][0] = k.W1 * drdx[c] * drdn + k.W2 * n[c];
    }
    for (int ik = 0; ik < 3; ik++) for (int il = 0; il < 3; il++) {
      fs_u[il + 1][ik + 1] = k.psi * dkr[il][ik] - k.chi * drdx[il] * drdx[ik];
      fs_t[il + 1][ik + 1] = k.TT1 * drdx[il] * drdx[ik] * drdn + k.TT2 * (drdn * dkr[il][ik] + drdx[ik] * n[il]) + k.TT3 * drdx[il] * n[ik];
    }
    for (int a = 0; a < 4; a++) for (int c = 0; c < 4; c++) for (int j = 0; j < nn; j++) {
      h[(j * 4 + a) * 4 + c] += fs_t[a][c] * pphijw[j]; g[(j * 4 + a) * 4 + c] += fs_u[a][c] * sphijw[j]; }
    return;
  }
  if (p.pot) {   // fbem_bem_harpot3d_sbie_ext_pre: bem_harpot3d.f90:305-320 (the same point formula in _ext_st :451-466, :605-622)
    double d1r1 = 1.0 / r, d1r2 = d1r1 * d1r1;
    double drdx[3] = {rv[0] * d1r1, rv[1] * d1r1, rv[2] * d1r1};
    double drdn = dot3(drdx, n);
    cd E[7]; zexp_decomposed(cd(-0.0, -1.0) * p.kp * r, E);
    cd E2r = E[2] * d1r1, E3r = E[3] * d1r2;
    cd fs_P = d1r1 + p.P1 + E2r;
    cd fs_Q = d1r2 + p.Q1 + p.Q2 * E2r + E3r;
    cd fq = fs_Q * drdn;
    for (int j = 0; j < nn; j++) { h[j] += fq * pphijw[j]; g[j] += fs_P * sphijw[j]; }
    return;
  }
  if (p.statics) {   // fbem_bem_staela3d_sbie_ext_pre: bem_staela3d.f90:629-645 (the same point formula in _ext_st :930-950)
    double d1r = 1.0 / r, d1r2 = d1r * d1r;
    double drdx[3] = {rv[0] * d1r, rv[1] * d1r, rv[2] * d1r};
    double drdn = dot3(drdx, n);
    for (int il = 0; il < 3; il++)
      for (int ik = 0; ik < 3; ik++) {
        double fs_u = d1r * (p.cteu2 * dkr[il][ik] + drdx[il] * drdx[ik]);
        double fs_t = d1r2 * ((p.ctet2 * dkr[il][ik] + 3.0 * drdx[il] * drdx[ik]) * drdn + p.ctet2 * (n[il] * drdx[ik] - n[ik] * drdx[il]));
        for (int j = 0; j < nn; j++) { h[(j * 3 + il) * 3 + ik] += fs_t * pphijw[j]; g[(j * 3 + il) * 3 + ik] += fs_u * sphijw[j]; }
      }
    return;
  }
  KernelScalars ks; kernel_scalars(p, r, false, ks);
  double drdx[3] = {rv[0] * ks.d1r1, rv[1] * ks.d1r1, rv[2] * ks.d1r1};
  double drdn = dot3(drdx, n);
  for (int il = 0; il < 3; il++)
    for (int ik = 0; ik < 3; ik++) {
      cd fs_u = ks.psi * dkr[il][ik] - ks.chi * drdx[il] * drdx[ik];
      cd fs_t = ks.TT1 * drdx[il] * drdx[ik] * drdn + ks.TT2 * (drdn * dkr[il][ik] + drdx[ik] * n[il]) + ks.TT3 * drdx[il] * n[ik];
      for (int j = 0; j < nn; j++) { h[(j * 3 + il) * 3 + ik] += fs_t * pphijw[j]; g[(j * 3 + il) * 3 + ik] += fs_u * sphijw[j]; }
    }
}

// One exterior Gauss point of the hypersingular equation: m(:,l,k) += s*_lk pphijw, l(:,l,k) += d*_lk sphijw with the unit normal
// n_i at the collocation point (fbem_bem_harela3d_hbie_ext_pre, bem_harela3d.f90:2606-2654; the same formulas in _ext_st)
static inline void add_exterior_point_hbie(const Params& p, const double* x, const double* n, const double* x_i, const double* n_i, int nn,
                                           const double* pphijw, const double* sphijw, cd* m, cd* l) {
  double rv[3] = {x[0] - x_i[0], x[1] - x_i[1], x[2] - x_i[2]};
  double r = sqrt(dot3(rv, rv));
  double d1r1 = 1.0 / r, d1r2 = d1r1 * d1r1, d1r3 = d1r2 * d1r1, d1r4 = d1r3 * d1r1, d1r5 = d1r4 * d1r1;
  double drdx[3] = {rv[0] * d1r1, rv[1] * d1r1, rv[2] * d1r1};
  double drdn = dot3(drdx, n), drdni = -dot3(drdx, n_i), n_dot_ni = dot3(n, n_i);
  if (p.statics) {   // fbem_bem_staela3d_hbie_ext_pre: bem_staela3d.f90:4428-4439
    const double cted2 = p.ctet2, ctes2 = p.ctet2, ctes3 = p.ctes3, nu = p.nu_s;
    for (int il = 0; il < 3; il++)
      for (int ik = 0; ik < 3; ik++) {
        double fs_d = d1r2 * ((cted2 * dkr[il][ik] + 3.0 * drdx[il] * drdx[ik]) * drdni + cted2 * (n_i[il] * drdx[ik] - n_i[ik] * drdx[il]));
        double fs_s = d1r3 * (3.0 * (5.0 * drdx[il] * drdx[ik] - nu * dkr[il][ik]) * drdn * drdni
                            + 3.0 * ctes2 * (drdx[ik] * n_i[il] * drdn - drdx[il] * n[ik] * drdni)
                            + 3.0 * nu * (drdx[il] * n_i[ik] * drdn - drdx[ik] * n[il] * drdni)
                            + (3.0 * nu * drdx[il] * drdx[ik] + ctes2 * dkr[il][ik]) * n_dot_ni
                            + ctes2 * n[il] * n_i[ik] - ctes3 * n[ik] * n_i[il]);
        for (int j = 0; j < nn; j++) { m[(j * 3 + il) * 3 + ik] += fs_s * pphijw[j]; l[(j * 3 + il) * 3 + ik] += fs_d * sphijw[j]; }
      }
    return;
  }
  const cd mim(-0.0, -1.0);
  cd z[2] = {mim * p.k1 * r, mim * p.k2 * r};
  cd E1[7], E2[7];
  zexp_decomposed(z[0], E1); zexp_decomposed(z[1], E2);
  cd E21 = E1[2] * d1r1, E22 = E2[2] * d1r1, E31 = E1[3] * d1r2, E32 = E2[3] * d1r2, E41 = E1[4] * d1r3, E42 = E2[4] * d1r3;
  cd E51 = E1[5] * d1r4, E52 = E2[5] * d1r4, E61 = E1[6] * d1r5, E62 = E2[6] * d1r5;
  cd TT1 = p.T1[1] * d1r2 + p.T1[2] + p.T1[3] * E21 + p.T1[4] * E22 + p.T1[5] * E31 + p.T1[6] * E32 + p.T1[7] * E41 + p.T1[8] * E42 + p.T1[9] * E51 + p.T1[10] * E52;
  cd TT2 = p.T2[1] * d1r2 + p.T2[2] + p.T2[3] * E22 + p.T2[4] * E31 + p.T2[5] * E32 + p.T2[6] * E41 + p.T2[7] * E42 + p.T2[8] * E51 + p.T2[9] * E52;
  cd TT3 = p.T3[1] * d1r2 + p.T3[2] + p.T3[3] * E21 + p.T3[4] * E31 + p.T3[5] * E32 + p.T3[6] * E41 + p.T3[7] * E42 + p.T3[8] * E51 + p.T3[9] * E52;
  cd S1 = p.S1[1] * d1r3 + p.S1[2] * d1r1 + p.S1[3] * E22 + p.S1[4] * E31 + p.S1[5] * E32 + p.S1[6] * E41 + p.S1[7] * E42 + p.S1[8] * E51 + p.S1[9] * E52 + p.S1[10] * E61 + p.S1[11] * E62;
  cd S2 = p.S2[1] * d1r3 + p.S2[2] * d1r1 + p.S2[3] * E21 + p.S2[4] * E31 + p.S2[5] * E32 + p.S2[6] * E41 + p.S2[7] * E42 + p.S2[8] * E51 + p.S2[9] * E52 + p.S2[10] * E61 + p.S2[11] * E62;
  cd S3 = p.S3[1] * d1r3 + p.S3[2] * d1r1 + p.S3[3] * E21 + p.S3[4] * E22 + p.S3[5] * E31 + p.S3[6] * E32 + p.S3[7] * E41 + p.S3[8] * E42 + p.S3[9] * E51 + p.S3[10] * E52 + p.S3[11] * E61 + p.S3[12] * E62;
  cd S4 = p.S4[1] * d1r3 + p.S4[2] * d1r1 + p.S4[3] + p.S4[4] * E32 + p.S4[5] * E41 + p.S4[6] * E42 + p.S4[7] * E51 + p.S4[8] * E52 + p.S4[9] * E61 + p.S4[10] * E62;
  cd S5 = p.S5[1] * d1r3 + p.S5[2] * d1r1 + p.S5[3] + p.S5[4] * E21 + p.S5[5] * E31 + p.S5[6] * E41 + p.S5[7] * E42 + p.S5[8] * E51 + p.S5[9] * E52 + p.S5[10] * E61 + p.S5[11] * E62;
  for (int il = 0; il < 3; il++)
    for (int ik = 0; ik < 3; ik++) {
      cd fs_d = TT1 * drdx[il] * drdx[ik] * drdni - TT2 * (-drdni * dkr[il][ik] + drdx[il] * n_i[ik]) - TT3 * drdx[ik] * n_i[il];
      cd fs_s = S1 * (drdx[il] * n_i[ik] * drdn - drdx[ik] * n[il] * drdni - dkr[il][ik] * drdn * drdni + drdx[il] * drdx[ik] * n_dot_ni)
              + S2 * (drdx[ik] * n_i[il] * drdn - drdx[il] * n[ik] * drdni) + S3 * drdx[il] * drdx[ik] * drdn * drdni
              + S4 * (dkr[il][ik] * n_dot_ni + n_i[ik] * n[il]) + S5 * n[ik] * n_i[il];
      for (int j = 0; j < nn; j++) { m[(j * 3 + il) * 3 + ik] += fs_s * pphijw[j]; l[(j * 3 + il) * 3 + ik] += fs_d * sphijw[j]; }
    }
}

// -------------------------------------------------------------------------------------
// Quasi-singular rule estimator: lib/fbem/src/quasisingular_integration.f90:181-287 (fit
// coefficients, numbers only), :318-398 (parameters), :402-554 (standard), :558-711 (Telles)
// -------------------------------------------------------------------------------------
// [coef][curve]
static const double lnr_alpha0[3][3] = {{2.511371433253553e-1, 1.588915580092415e-1, 1.303338290205561e-1}, {-8.318131267946077e-2, -9.469173675969373e-2, -9.041069352366727e-2}, {-1.973643360180158e-3, -2.500979779162619e-3, -2.302123487272636e-3}};
static const double lnr_alpha1[3][3] = {{-8.301769338383449e-1, -8.575272011921337e-1, -5.785210418300009e-1}, {-7.458331284021328e-2, -6.866562987114574e-2, -3.107089548772586e-2}, {-1.992535165956408e-3, -1.735469217262563e-3, -4.204818399993008e-4}};
static const double lnr_beta1[3][3] = {{-1.114783338380396e-1, -3.918234478283774e-1, -1.933844506192423e-1}, {-6.169581618278328e-2, -6.478096580939043e-2, -3.303326826061181e-2}, {-1.873698813959166e-3, -1.657459305149584e-3, -6.497990057394490e-4}};
static const double d1rn_alpha0[6][3] = {{2.443943976922257e-1, 1.855869272054077e-1, 1.964994120809251e-1}, {-7.845125775746743e-2, -8.441784317192081e-2, -7.533682025056084e-2}, {3.895326800087784e-2, 3.992519918023225e-2, 6.300967117113785e-2}, {8.783124822346978e-4, 1.003034542932314e-3, 1.453521059340246e-3}, {-1.719172638314362e-3, -2.012460287158036e-3, -1.653933461489553e-3}, {-1.575912106080523e-3, -1.596684136739970e-3, -2.613368441217468e-3}};
static const double d1rn_alpha1[6][3] = {{-8.421216556191472e-1, -8.012899104725332e-1, -5.891778953073477e-1}, {-8.778262996280104e-2, -7.060561516342915e-2, -4.515354064860566e-2}, {1.679496012553309e-2, 5.713299243166406e-3, 5.728946548799128e-3}, {6.522695414940627e-4, -1.952251227204736e-4, 1.106131193536092e-6}, {-2.649807213627071e-3, -1.984455335909311e-3, -1.179100966502172e-3}, {-8.472009602597869e-4, -6.680823196695776e-4, -5.355782068751523e-4}};
static const double d1rn_beta1[6][3] = {{-2.693032219216048e-2, -2.353515879579821e-1, -9.828561879516765e-2}, {-6.138311391493901e-2, -5.141856801187955e-2, -3.074365737570111e-2}, {2.273921394673689e-2, 1.693100431819354e-2, 1.306905121742232e-2}, {9.789745450155518e-4, -1.491381226735667e-5, 6.065437465834727e-5}, {-2.030377872751224e-3, -1.330369292483563e-3, -7.853234287066991e-4}, {-9.059289455187508e-4, -1.264704521827383e-3, -1.025889829482388e-3}};
static const double t_lnr_alpha0[3][3] = {{2.5834148732821033e-01, 5.1283007739426911e-01, 4.1061632629507161e-01}, {-8.8865167808209675e-02, -1.1479138136155583e-02, -3.0669024691563598e-02}, {-2.3640194116454241e-03, 1.7167160355773211e-03, 5.1920314153891956e-04}};
static const double t_lnr_alpha1[3][3] = {{-9.1721308433764848e-01, -1.0632414684959128e+00, -1.1740312565911180e+00}, {-6.2837979237008845e-02, -1.1633428696469847e-01, -1.3946434974037192e-01}, {-1.1580362636310136e-03, -4.4173578665990879e-03, -5.4264705520737193e-03}};
static const double t_lnr_beta1[3][3] = {{-4.9431525769444484e-01, -9.1742055835524161e-01, -1.1668939380105636e+00}, {-6.5628053690049357e-02, -1.4448614058834844e-01, -1.7686541251294283e-01}, {-1.2683068366630545e-03, -5.5849846784841173e-03, -6.9562836980345301e-03}};
static const double t_d1rn_alpha0[6][3] = {{2.9379042486857926e-01, 1.1155167367380846e-01, 9.5587534756324991e-02}, {-7.6518438657183094e-02, -9.0043220465500617e-02, -8.5841370957250140e-02}, {3.0011332045426935e-02, 3.9605615471752335e-02, 5.9724590614136340e-02}, {5.8900790667561719e-04, 1.0669688965738889e-03, 1.4101322853742803e-03}, {-1.7642732320416377e-03, -2.2428131063082419e-03, -2.0424217927592597e-03}, {-1.3634192833061319e-03, -1.6030309471606837e-03, -2.4264102214015559e-03}};
static const double t_d1rn_alpha1[6][3] = {{-9.8339088502232042e-01, -6.5856321275500784e-01, -2.4634605754204883e-01}, {-8.6761356304300027e-02, -4.7354561095619759e-02, 1.0051411705899942e-02}, {2.6028760963988956e-02, 1.9423558756522371e-03, -1.1994978255365170e-02}, {7.7677507397256337e-04, -2.9087437598165618e-04, -1.0022065157916291e-03}, {-2.3619290417702245e-03, -1.1859755938049318e-03, 7.9639034155470554e-04}, {-1.2891431252005461e-03, -4.5988734937654318e-04, -1.8103634920482468e-04}};
static const double t_d1rn_beta1[6][3] = {{-4.5453529921363467e-01, -2.6610069546799336e-01, 7.5605400079843565e-02}, {-7.3941529478921839e-02, -3.6048387668689889e-02, 2.4021121564322980e-02}, {3.7672240427426076e-02, 1.3288129149913631e-02, -9.3693044556669332e-03}, {9.8276802658567980e-04, -1.8099231920766130e-04, -1.3532246088221670e-03}, {-1.8897770518175436e-03, -7.6015457149682889e-04, 1.3564674989692248e-03}, {-1.8079562352345093e-03, -1.2247286683263207e-03, -7.4467072775070487e-04}};

struct QsParams {  // [curve 0..2][f 0..7]
  double re;
  double dmin[3][8], dmax[3][8], alpha0[3][8], alpha1[3][8], beta1[3][8];
  double t_dmin[3][8], t_dmax[3][8], t_alpha0[3][8], t_alpha1[3][8], t_beta1[3][8];
};
static inline double poly1(const double c[3][3], int cv, double x) { return c[0][cv] + c[1][cv] * x + c[2][cv] * (x * x); }
static inline double poly2(const double c[6][3], int cv, double x, double y) {
  return c[0][cv] + c[1][cv] * x + c[2][cv] * y + c[3][cv] * x * y + c[4][cv] * (x * x) + c[5][cv] * (y * y);
}
static void qs_calculate_parameters(double relative_error, QsParams& p) {
  double log10re = log10(fabs(relative_error));
  if (log10re > -3.0) log10re = -3.0;
  if (log10re < -15.0) log10re = -15.0;
  p.re = pow(10.0, log10re);
  // log10(30.), log10( 2.) are default-real (single precision) intrinsics promoted to double (:343-344)
  const double l30 = (double)log10f(30.f), l2 = (double)log10f(2.f);
  for (int c = 0; c < 3; c++) {
    p.alpha0[c][0] = poly1(lnr_alpha0, c, log10re); p.alpha1[c][0] = poly1(lnr_alpha1, c, log10re); p.beta1[c][0] = poly1(lnr_beta1, c, log10re);
    p.dmin[c][0] = pow(10.0, (l30 - p.alpha0[c][0]) / (p.alpha1[c][0] - l30 * p.beta1[c][0]));
    p.dmax[c][0] = pow(10.0, (l2 - p.alpha0[c][0]) / (p.alpha1[c][0] - l2 * p.beta1[c][0]));
    p.t_alpha0[c][0] = poly1(t_lnr_alpha0, c, log10re); p.t_alpha1[c][0] = poly1(t_lnr_alpha1, c, log10re); p.t_beta1[c][0] = poly1(t_lnr_beta1, c, log10re);
    p.t_dmin[c][0] = pow(10.0, (l30 - p.t_alpha0[c][0]) / (p.t_alpha1[c][0] - l30 * p.t_beta1[c][0]));
    p.t_dmax[c][0] = pow(10.0, (l2 - p.t_alpha0[c][0]) / (p.t_alpha1[c][0] - l2 * p.t_beta1[c][0]));
    if (p.t_dmin[c][0] < 1.e-6) p.t_dmin[c][0] = 1.e-6;
    if (p.t_dmin[c][0] > p.t_dmax[c][0]) p.t_dmin[c][0] = 1.e-6;
  }
  for (int i = 1; i <= 7; i++)
    for (int c = 0; c < 3; c++) {
      double y = (double)i;
      p.alpha0[c][i] = poly2(d1rn_alpha0, c, log10re, y); p.alpha1[c][i] = poly2(d1rn_alpha1, c, log10re, y); p.beta1[c][i] = poly2(d1rn_beta1, c, log10re, y);
      p.dmin[c][i] = pow(10.0, (l30 - p.alpha0[c][i]) / (p.alpha1[c][i] - l30 * p.beta1[c][i]));
      p.dmax[c][i] = pow(10.0, (l2 - p.alpha0[c][i]) / (p.alpha1[c][i] - l2 * p.beta1[c][i]));
      p.t_alpha0[c][i] = poly2(t_d1rn_alpha0, c, log10re, y); p.t_alpha1[c][i] = poly2(t_d1rn_alpha1, c, log10re, y); p.t_beta1[c][i] = poly2(t_d1rn_beta1, c, log10re, y);
      p.t_dmin[c][i] = pow(10.0, (l30 - p.t_alpha0[c][i]) / (p.t_alpha1[c][i] - l30 * p.t_beta1[c][i]));
      p.t_dmax[c][i] = pow(10.0, (l2 - p.t_alpha0[c][i]) / (p.t_alpha1[c][i] - l2 * p.t_beta1[c][i]));
      if (p.t_dmin[c][i] < 1.e-6) p.t_dmin[c][i] = 1.e-6;
      if (p.t_dmin[c][i] > p.t_dmax[c][i]) p.t_dmin[c][i] = 1.e-6;
    }
}
// N_c(d) for one curve; returns false if d<=dmin (estimation 0)
static inline bool qs_curve(const double dmin[3][8], const double dmax[3][8], const double a0[3][8], const double a1[3][8],
                            const double b1[3][8], int c, int f, double d, double log10d, double& N) {
  if (d <= dmin[c][f]) return false;
  if (d >= dmax[c][f]) N = 2.0;
  else N = pow(10.0, (a0[c][f] + a1[c][f] * log10d) / (1.0 + b1[c][f] * log10d));
  return true;
}
// telles=false: fbem_qs_n_estimation_standard (:402-554); telles=true: fbem_qs_n_estimation_telles (:558-711).
// Optional r,q arguments are never passed on this path.
static int qs_n_estimation(bool telles, int etype, int f, const QsParams& p, double d, const double* barxi) {
  const double(*dmin)[8] = telles ? p.t_dmin : p.dmin;
  const double(*dmax)[8] = telles ? p.t_dmax : p.dmax;
  const double(*a0)[8] = telles ? p.t_alpha0 : p.alpha0;
  const double(*a1)[8] = telles ? p.t_alpha1 : p.alpha1;
  const double(*b1)[8] = telles ? p.t_beta1 : p.beta1;
  int n;
  if (d <= 2.0) {
    double log10d = log10(d), N1, N2;
    if (etype == LINE2 || etype == LINE3) {
      if (fabs(barxi[0]) < 1.0) {
        if (!qs_curve(dmin, dmax, a0, a1, b1, 0, f, d, log10d, N1)) return 0;
        if (!qs_curve(dmin, dmax, a0, a1, b1, 1, f, d, log10d, N2)) return 0;
        n = (int)ceil(barxi[0] * barxi[0] * (N2 - N1) + N1);
      } else {
        if (!qs_curve(dmin, dmax, a0, a1, b1, 1, f, d, log10d, N2)) return 0;
        n = (int)ceil(N2);
      }
    } else {
      if (!qs_curve(dmin, dmax, a0, a1, b1, 0, f, d, log10d, N1)) return 0;
      if (!qs_curve(dmin, dmax, a0, a1, b1, 1, f, d, log10d, N2)) return 0;
      if (etype == TRI3 || etype == TRI6)
        n = (int)ceil(4.0 * (barxi[1] * (barxi[1] + barxi[0] - 1.0) + (barxi[0] - 1.0) * barxi[0]) * (N2 - N1) + N2);
      else
        n = (int)ceil((barxi[0] * barxi[0]) * (barxi[1] * barxi[1]) * (N2 - N1) + N1);
    }
  } else {
    // d>2 always uses the *standard* curve 3 with nint (also in the Telles variant, :700-706)
    if (d >= p.dmax[2][f]) n = 2;
    else { double log10d = log10(d); n = (int)lround(pow(10.0, (p.alpha0[2][f] + p.alpha1[2][f] * log10d) / (1.0 + p.beta1[2][f] * log10d))); }
  }
  if (n > 30) n = 0;
  return n;
}

// -------------------------------------------------------------------------------------
// Telles transformation: lib/fbem/src/telles_transformation.f90:77-232
// -------------------------------------------------------------------------------------
static double telles_barr_any(double d) {  // fbem_telles_barr(d, fbem_f_any) :77-127
  double b = (d < 3.0) ? d / (0.89039 * d + 0.32883) : 1.0;
  if (b > 1.0) b = 1.0;
  return b;
}
static inline double cbrt_signed(double v, double d13) { return v >= 0.0 ? pow(v, d13) : -pow(fabs(v), d13); }
static void telles11_parameters(double bar_xi, double bar_r, double* c) {  // :130-161
  double w = bar_xi / (1.0 + 2.0 * bar_r);
  double p = 1.0 / (3.0 * (1.0 + 2.0 * bar_r)) * (3.0 - 2.0 * bar_r - 3.0 * w * bar_xi);
  double q = w / 2.0 * ((3.0 - 2.0 * bar_r) / (1.0 + 2.0 * bar_r) - 2.0 * (w * w) - 1.0);
  double R2 = sqrt(q * q + p * p * p), d13 = 1.0 / 3.0;
  double R31 = cbrt_signed(-q + R2, d13), R32 = cbrt_signed(-q - R2, d13);
  double bg = R31 + R32 + w, Q = 1.0 + 3.0 * (bg * bg);
  c[0] = (1.0 - bar_r) / Q; c[1] = -3.0 * bg * c[0]; c[2] = (bar_r + 3.0 * (bg * bg)) / Q; c[3] = -c[1];
}
static void telles01_parameters(double bar_xi, double bar_r, double* c) {  // :164-196
  double waux = 1.0 + 2.0 * bar_r, w = (bar_xi + bar_r) / waux;
  double p = (3.0 * bar_xi + bar_r) / (3.0 * waux) - w * w;
  double q = w * ((3.0 * bar_xi + bar_r) / (2.0 * waux) - w * w) - bar_xi / (2.0 * waux);
  double R2 = sqrt(q * q + p * p * p), d13 = 1.0 / 3.0;
  double R31 = cbrt_signed(-q + R2, d13), R32 = cbrt_signed(-q - R2, d13);
  double bg = R31 + R32 + w, Q = 3.0 * bg * (bg - 1.0) + 1.0;
  c[0] = (1.0 - bar_r) / Q; c[1] = -3.0 * bg * c[0]; c[2] = (3.0 * bg * (bg - bar_r) + bar_r) / Q; c[3] = 0.0;
}
static inline void telles_xi_jac(const double* c, double g, double& xi, double& jac) {  // :221-232
  xi = c[0] * (g * g * g) + c[1] * (g * g) + c[2] * g + c[3];
  jac = 3.0 * c[0] * (g * g) + 2.0 * c[1] * g + c[2];
}

// -------------------------------------------------------------------------------------
// Geometry helpers: lib/fbem/src/geometry.f90
// -------------------------------------------------------------------------------------
static double jacobian3d_1d(int et, const double* x /*3 x nn*/, double xi) {  // :1469-1504
  double d[3]; dphi1d<double>(et, xi, d);
  double t[3] = {0, 0, 0};
  for (int i = 0; i < n_nodes_of(et); i++) for (int j = 0; j < 3; j++) t[j] = t[j] + d[i] * x[3 * i + j];
  return sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
}
static double length3d(int et, const double* x, double tol) {  // :2794-2882 (rule=0, step=0 -> 3)
  const int nr = 32, step = 3;
  double local_tol = tol <= 0.0 ? 1.e-6 : tol;
  double err = local_tol + 1.0;
  double length = jacobian3d_1d(et, x, gl11_x(1, 0)) * gl11_w(1, 0);
  int j = 1 + step;
  while (err >= local_tol) {
    double old = length; length = 0.0;
    for (int i = 0; i < j; i++) length = length + jacobian3d_1d(et, x, gl11_x(j, i)) * gl11_w(j, i);
    err = fabs((old - length) / length);
    j = j + step;
    if (j >= nr) {
      length = 0.0; for (int i = 0; i < nr - 1; i++) length = length + jacobian3d_1d(et, x, gl11_x(nr - 1, i)) * gl11_w(nr - 1, i);
      old = length;
      length = 0.0; for (int i = 0; i < nr; i++) length = length + jacobian3d_1d(et, x, gl11_x(nr, i)) * gl11_w(nr, i);
      err = fabs((old - length) / length);
      break;
    }
  }
  return length;
}
static double characteristic_length(int et, const double* x, double tol) {  // :3357-3389
  double best = 0.0; bool first = true;
  for (int e = 0; e < n_edges_of(et); e++) {
    int ety = edge_type_of(et); double xe[9];
    for (int n = 0; n < n_nodes_of(ety); n++) for (int c = 0; c < 3; c++) xe[3 * n + c] = x[3 * edge_node(n, e, et) + c];
    double l = length3d(ety, xe, tol);
    if (first || l > best) { best = l; first = false; }
  }
  return best;
}
// x, T1, T2, N, jg at xi (the idiom repeated throughout bem_harela3d.f90, e.g. :815-831)
static inline void geom_at(int et, int nn, const double* xn, const double* xi, double* gphi, double* x, double* N, double& jg) {
  double d1[9], d2[9];
  phi2d<double>(et, xi, gphi); dphi2d<double>(et, xi, d1, d2);
  double T1[3] = {0, 0, 0}, T2[3] = {0, 0, 0}; x[0] = x[1] = x[2] = 0.0;
  for (int k = 0; k < nn; k++) for (int c = 0; c < 3; c++) {
    x[c] = x[c] + gphi[k] * xn[3 * k + c]; T1[c] = T1[c] + d1[k] * xn[3 * k + c]; T2[c] = T2[c] + d2[k] * xn[3 * k + c]; }
  N[0] = T1[1] * T2[2] - T1[2] * T2[1]; N[1] = T1[2] * T2[0] - T1[0] * T2[2]; N[2] = T1[0] * T2[1] - T1[1] * T2[0];
  jg = sqrt(dot3(N, N));
}
// fbem_qs_phijac_ngp_2d: quasisingular_integration.f90:812-975
static int phijac_ngp_2d(int et, const double* xn, double error_height) {
  int nn = n_nodes_of(et);
  double tol = error_height <= 0.0 ? 1.e-6 : (error_height < 1.e-12 ? 1.e-12 : error_height);
  double integral[9] = {0}, old[9]; bool cont = true; int j = 1;
  while (cont) {
    for (int l = 0; l < nn; l++) { old[l] = integral[l]; integral[l] = 0.0; }
    bool tri = (et == TRI3 || et == TRI6);
    for (int k1 = 0; k1 < j; k1++) for (int k2 = 0; k2 < j; k2++) {
      double xi[2], w;
      if (tri) { xi[0] = (1.0 - gj01_x(j, k2)) * gl01_x(j, k1); xi[1] = gj01_x(j, k2); }
      else { xi[0] = gl11_x(j, k1); xi[1] = gl11_x(j, k2); }
      double phi[9], x[3], N[3], jac; geom_at(et, nn, xn, xi, phi, x, N, jac);
      double jw = tri ? jac * gl01_w(j, k1) * gj01_w(j, k2) : jac * gl11_w(j, k1) * gl11_w(j, k2);
      (void)w;
      for (int l = 0; l < nn; l++) integral[l] = integral[l] + (phi[l] + 1.0) * jw;
    }
    if (j == 1) j = j + 1;
    else {
      cont = false;
      for (int l = 0; l < nn; l++) if (fabs((old[l] - integral[l]) / integral[l]) > tol) { cont = true; j = j + 1; break; }
      if (j == 33) cont = false;
    }
  }
  return j - 1;
}
// fbem_geometry_element_ball: geometry.f90:3904-4235 (2D elements in R^3)
static void element_ball(int et, const double* xn, int glp, double* centre, double& radius) {
  int nn = n_nodes_of(et); bool tri = (et == TRI3 || et == TRI6);
  double esize = 0.0, xm[3] = {0, 0, 0};
  int npt = tri ? wan_n(2 * glp - 1) : glp * glp;
  auto point = [&](int k, double* xi, double& w) {
    if (tri) { xi[0] = wan_x1(2 * glp - 1, k); xi[1] = wan_x2(2 * glp - 1, k); w = wan_w(2 * glp - 1, k); }
    else { int k1 = k / glp, k2 = k % glp; xi[0] = gl11_x(glp, k1); xi[1] = gl11_x(glp, k2); w = -1.0; }
  };
  for (int k = 0; k < npt; k++) {
    double xi[2], w, phi[9], x[3], N[3], jg; point(k, xi, w); geom_at(et, nn, xn, xi, phi, x, N, jg);
    double jw = tri ? jg * w : jg * gl11_w(glp, k / glp) * gl11_w(glp, k % glp);
    esize = esize + jw; for (int c = 0; c < 3; c++) xm[c] = xm[c] + x[c] * jw;
  }
  for (int c = 0; c < 3; c++) centre[c] = xm[c] / esize;
  radius = 0.0;
  for (int k = 0; k < nn; k++) {
    double r[3] = {xn[3 * k] - centre[0], xn[3 * k + 1] - centre[1], xn[3 * k + 2] - centre[2]};
    double t = sqrt(dot3(r, r)); if (radius < t) radius = t;
  }
  for (int k = 0; k < npt; k++) {
    double xi[2], w, phi[9], x[3] = {0, 0, 0}; point(k, xi, w); phi2d<double>(et, xi, phi);
    for (int n = 0; n < nn; n++) for (int c = 0; c < 3; c++) x[c] = x[c] + phi[n] * xn[3 * n + c];
    double r[3] = {x[0] - centre[0], x[1] - centre[1], x[2] - centre[2]};
    double t = sqrt(dot3(r, r)); if (radius < t) radius = t;
  }
}
// fbem_obtain_element_subdivision_coordinates: geometry.f90:2591-2697
static void subdivision_coordinates(int et, const double* x, const double* xi_s /*2 x nv or 1 x 2*/, double* x_s) {
  int nn = n_nodes_of(et), nv = n_vertices_of(et);
  for (int k = 0; k < nn; k++) {
    double xis[2]; xi_at_node(et, k, xis);
    if (et == LINE2 || et == LINE3) {
      double phis[2]; phi1d<double>(LINE2, xis[0], phis);
      double xi = 0.0; for (int i = 0; i < 2; i++) xi = xi + phis[i] * xi_s[i];
      double phi[3]; phi1d<double>(et, xi, phi);
      for (int c = 0; c < 3; c++) { x_s[3 * k + c] = 0.0; for (int i = 0; i < nn; i++) x_s[3 * k + c] = x_s[3 * k + c] + phi[i] * x[3 * i + c]; }
    } else {
      double phis[4]; phi2d<double>(nv == 3 ? TRI3 : QUAD4, xis, phis);
      double xi[2] = {0, 0};
      for (int i = 0; i < nv; i++) { xi[0] = xi[0] + phis[i] * xi_s[2 * i]; xi[1] = xi[1] + phis[i] * xi_s[2 * i + 1]; }
      double phi[9]; phi2d<double>(et, xi, phi);
      for (int c = 0; c < 3; c++) { x_s[3 * k + c] = 0.0; for (int i = 0; i < nn; i++) x_s[3 * k + c] = x_s[3 * k + c] + phi[i] * x[3 * i + c]; }
    }
  }
}

// ---- nearest point -----------------------------------------------------------------
static void nearest_element_point_bem(int et, const double* x, double cl, const double* x_i, double* barxi, double& rmin, double& d, int& method);

// fbem_nearest_xi_nodes: geometry.f90:4263-4316
static void nearest_xi_nodes(int et, const double* x, const double* p, double* xi, double& rnear) {
  int nn = n_nodes_of(et), best = 0;
  double r[3] = {x[0] - p[0], x[1] - p[1], x[2] - p[2]};
  rnear = sqrt(dot3(r, r));
  for (int k = 1; k < nn; k++) {
    double rr[3] = {x[3 * k] - p[0], x[3 * k + 1] - p[1], x[3 * k + 2] - p[2]};
    double rm = sqrt(dot3(rr, rr));
    if (rm < rnear) { rnear = rm; best = k; }
  }
  xi_at_node(et, best, xi);
}
// fbem_nearest_xi (sampling, 1D): geometry.f90:4329-4441
static void nearest_xi_sampling_1d(int et, const double* x, const double* p, int ns, int nrs, int nit, double& xi_out, double& r_out) {
  int nn = n_nodes_of(et); double xi_min = 0, r_min = 0;
  auto dist = [&](double xi) { double phi[3], xx[3] = {0, 0, 0}; phi1d<double>(et, xi, phi);
    for (int j = 0; j < nn; j++) for (int c = 0; c < 3; c++) xx[c] = xx[c] + phi[j] * x[3 * j + c];
    double rv[3] = {xx[0] - p[0], xx[1] - p[1], xx[2] - p[2]}; return sqrt(dot3(rv, rv)); };
  for (int k = 0; k <= ns + 1; k++) {
    double xi = -1.0 + 2.0 / (double)(ns + 1) * (double)k, r = dist(xi);
    if (k == 0) { xi_min = xi; r_min = r; }
    if (r < r_min) { xi_min = xi; r_min = r; }
  }
  double xi_old = xi_min, width = 2.0 / (double)(ns + 1);
  for (int l = 1; l <= nit; l++) {
    for (int k = 1; k <= nrs; k++) {
      double xi = xi_old - width + (2.0 * width) / (double)(nrs + 1) * (double)k;
      if (xi > -1.0 && xi < 1.0) { double r = dist(xi); if (r < r_min) { xi_min = xi; r_min = r; } }
    }
    xi_old = xi_min; width = 2.0 * width / (double)(nrs + 1);
  }
  xi_out = xi_min; r_out = r_min;
}
// fbem_nearest_xi1xi2 (sampling, 2D): geometry.f90:4574-4795 (including the quad double-division of `width`)
static void nearest_xi_sampling_2d(int et, const double* x, const double* p, int ns, int nrs, int nit, double* xi_out, double& r_out) {
  int nn = n_nodes_of(et); bool tri = (et == TRI3 || et == TRI6);
  double xi_min[2] = {0, 0}, r_min = 0;
  auto dist = [&](const double* xi) { double phi[9], xx[3] = {0, 0, 0}; phi2d<double>(et, xi, phi);
    for (int j = 0; j < nn; j++) for (int c = 0; c < 3; c++) xx[c] = xx[c] + phi[j] * x[3 * j + c];
    double rv[3] = {xx[0] - p[0], xx[1] - p[1], xx[2] - p[2]}; return sqrt(dot3(rv, rv)); };
  for (int k1 = 0; k1 <= ns + 1; k1++) {
    int k2max = tri ? ns + 1 - k1 : ns + 1;
    for (int k2 = 0; k2 <= k2max; k2++) {
      double xi[2];
      if (tri) { xi[0] = (double)k1 / (double)(ns + 1); xi[1] = (double)k2 / (double)(ns + 1); }
      else { xi[0] = -1.0 + 2.0 * (double)k1 / (double)(ns + 1); xi[1] = -1.0 + 2.0 * (double)k2 / (double)(ns + 1); }
      double r = dist(xi);
      if (k1 == 0 && k2 == 0) { xi_min[0] = xi[0]; xi_min[1] = xi[1]; r_min = r; }
      if (r < r_min) { xi_min[0] = xi[0]; xi_min[1] = xi[1]; r_min = r; }
    }
  }
  double xo[2] = {xi_min[0], xi_min[1]};
  double width = tri ? 1.0 / (double)(ns + 1) : 2.0 / (double)(ns + 1);
  for (int l = 1; l <= nit; l++) {
    for (int k1 = 1; k1 <= nrs; k1++) {
      double xi[2];
      xi[0] = xo[0] + 2.0 * ((double)k1 - (double)(nrs + 1) / 2.0) / (double)(nrs + 1) * width;
      if (!tri && !(xi[0] > -1.0 && xi[0] < 1.0)) continue;
      for (int k2 = 1; k2 <= nrs; k2++) {
        xi[1] = xo[1] + 2.0 * ((double)k2 - (double)(nrs + 1) / 2.0) / (double)(nrs + 1) * width;
        bool in = tri ? (xi[0] > 0.0 && xi[1] > 0.0 && (xi[0] + xi[1]) < 1.0) : (xi[1] > -1.0 && xi[1] < 1.0);
        if (in) { double r = dist(xi); if (r < r_min) { xi_min[0] = xi[0]; xi_min[1] = xi[1]; r_min = r; } }
      }
    }
    xo[0] = xi_min[0]; xo[1] = xi_min[1];
    if (!tri) width = width / (double)(nrs + 1);
    width = 2.0 * width / (double)(nrs + 1);
  }
  xi_out[0] = xi_min[0]; xi_out[1] = xi_min[1]; r_out = r_min;
}
// fbem_nearest_minimization_1d (+ iteration in real128): geometry.f90:5071-5105, :5196-5287
static void nearest_minimization_1d(int et, const double* x, const double* x_i, double error, int nmax, double& barxi, double& rmin, int& info) {
  int nn = n_nodes_of(et);
  std::vector<double> bh(nmax + 2), eh(nmax + 1);
  bh[1] = (barxi < -1.0 || barxi > 1.0) ? 0.0 : barxi;
  int k = 1; info = 0;
  while (info == 0) {
    q128 phi[3], dphi[3], xb[3] = {0, 0, 0}, dx[3] = {0, 0, 0}, a[3], b[3];
    phi1d<q128>(et, bh[k], phi); dphi1d<q128>(et, bh[k], dphi);
    for (int i = 0; i < nn; i++) for (int c = 0; c < 3; c++) { xb[c] = xb[c] + phi[i] * (q128)x[3 * i + c]; dx[c] = dx[c] + dphi[i] * (q128)x[3 * i + c]; }
    for (int c = 0; c < 3; c++) { a[c] = xb[c] - dx[c] * (q128)bh[k] - (q128)x_i[c]; b[c] = dx[c]; }
    q128 ab = a[0] * b[0] + a[1] * b[1] + a[2] * b[2], bb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
    bh[k + 1] = (double)(-ab / bb);
    eh[k] = fabs(bh[k + 1] - bh[k]) * 0.5;
    if (eh[k] <= error) { barxi = bh[k + 1]; info = 1; }
    else if (k == nmax) { barxi = bh[k + 1]; info = 2; }
    else if (k > 5) { if (eh[k] > eh[k - 2]) { barxi = bh[k + 1]; info = 3; } else k = k + 1; }
    else k = k + 1;
  }
  if (info == 1) {
    if (barxi < -1.0 || barxi > 1.0) {
      barxi = -1.0;
      q128 r[3]; for (int c = 0; c < 3; c++) r[c] = (q128)(x[c] - x_i[c]);
      rmin = (double)sqrtq(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
      for (int c = 0; c < 3; c++) r[c] = (q128)(x[3 + c] - x_i[c]);
      if (sqrtq(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) < (q128)rmin) barxi = 1.0;
    }
    q128 phi[3], r[3] = {0, 0, 0}; phi1d<q128>(et, barxi, phi);
    for (int i = 0; i < nn; i++) for (int c = 0; c < 3; c++) r[c] = r[c] + phi[i] * (q128)x[3 * i + c];
    for (int c = 0; c < 3; c++) r[c] = r[c] - (q128)x_i[c];
    rmin = (double)sqrtq(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  }
}
// fbem_nearest_minimization_2d (+ nearest_minimization_iteration_2d in real128): geometry.f90:5107-5150, :5289-5408
static void nearest_minimization_2d(int et, const double* x, const double* x_i, double error, int nmax, double* barxi, double& rmin, int& info) {
  int nn = n_nodes_of(et), ne = n_edges_of(et);
  std::vector<double> b1(nmax + 2), b2(nmax + 2), eh(nmax + 1);
  if (!check_xi1xi2(et, barxi)) { if (ne == 3) { b1[1] = 1.0 / 3.0; b2[1] = 1.0 / 3.0; } else { b1[1] = 0.0; b2[1] = 0.0; } }
  else { b1[1] = barxi[0]; b2[1] = barxi[1]; }
  int k = 1; info = 0;
  while (info == 0) {
    double bx[2] = {b1[k], b2[k]};
    q128 phi[9], d1[9], d2[9], xb[3] = {0, 0, 0}, t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0}, a[3];
    phi2d<q128>(et, bx, phi); dphi2d<q128>(et, bx, d1, d2);
    for (int i = 0; i < nn; i++) for (int c = 0; c < 3; c++) {
      xb[c] = xb[c] + phi[i] * (q128)x[3 * i + c]; t1[c] = t1[c] + d1[i] * (q128)x[3 * i + c]; t2[c] = t2[c] + d2[i] * (q128)x[3 * i + c]; }
    for (int c = 0; c < 3; c++) a[c] = xb[c] - t1[c] * (q128)bx[0] - t2[c] * (q128)bx[1] - (q128)x_i[c];
    q128 bb = t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2], cc = t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2];
    q128 bc = t1[0] * t2[0] + t1[1] * t2[1] + t1[2] * t2[2];
    q128 ab = a[0] * t1[0] + a[1] * t1[1] + a[2] * t1[2], ac = a[0] * t2[0] + a[1] * t2[1] + a[2] * t2[2];
    q128 det = bb * cc - bc * bc;
    b1[k + 1] = (double)(-(ab * cc - bc * ac) / det);
    b2[k + 1] = (double)(-(bb * ac - ab * bc) / det);
    double e1 = b1[k + 1] - b1[k], e2 = b2[k + 1] - b2[k];
    eh[k] = sqrt(e1 * e1 + e2 * e2) * 0.5;
    if (eh[k] <= error) { barxi[0] = b1[k + 1]; barxi[1] = b2[k + 1]; info = 1; }
    else if (k == nmax) { barxi[0] = b1[k + 1]; barxi[1] = b2[k + 1]; info = 2; }
    else if (k > 5) { if (eh[k] >= eh[k - 2]) { barxi[0] = b1[k + 1]; barxi[1] = b2[k + 1]; info = 3; } else k = k + 1; }
    else k = k + 1;
  }
  if (info == 1) {
    if (!check_xi1xi2(et, barxi)) {
      double rmin_e[4], bxe[4]; int ety = edge_type_of(et), nne = n_nodes_of(ety);
      for (int e = 0; e < ne; e++) {
        double xe[9]; for (int n = 0; n < nne; n++) for (int c = 0; c < 3; c++) xe[3 * n + c] = x[3 * edge_node(n, e, et) + c];
        double cl = characteristic_length(ety, xe, 1.e-12), d; int m; double bb1[1];
        nearest_element_point_bem(ety, xe, cl, x_i, bb1, rmin_e[e], d, m); bxe[e] = bb1[0];
      }
      int ke = 0; for (int e = 1; e < ne; e++) if (rmin_e[e] < rmin_e[ke]) ke = e;  // minloc: first minimum
      q128 phie[3]; phi1d<q128>(ety, bxe[ke], phie);
      q128 s0 = 0, s1 = 0;
      for (int n = 0; n < nne; n++) { double xin[2]; xi_at_node(et, edge_node(n, ke, et), xin); s0 = s0 + phie[n] * (q128)xin[0]; s1 = s1 + phie[n] * (q128)xin[1]; }
      barxi[0] = (double)s0; barxi[1] = (double)s1;
    }
    q128 phi[9], r[3] = {0, 0, 0}; phi2d<q128>(et, barxi, phi);
    for (int i = 0; i < nn; i++) for (int c = 0; c < 3; c++) r[c] = r[c] + phi[i] * (q128)x[3 * i + c];
    for (int c = 0; c < 3; c++) r[c] = r[c] - (q128)x_i[c];
    rmin = (double)sqrtq(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  }
}
// fbem_nearest_element_point_bem: geometry.f90:5519-5566
static void nearest_element_point_bem(int et, const double* x, double cl, const double* x_i, double* barxi, double& rmin, double& d, int& method) {
  nearest_xi_nodes(et, x, x_i, barxi, rmin);
  d = rmin / cl;
  if (d < 1.0) {
    int info;
    if (et == LINE2 || et == LINE3) {
      nearest_minimization_1d(et, x, x_i, 1.e-14, 20, barxi[0], rmin, info);
      if (info != 1) { nearest_xi_sampling_1d(et, x, x_i, 25, 14, 2, barxi[0], rmin); method = 3; } else method = 2;
    } else {
      nearest_minimization_2d(et, x, x_i, 1.e-14, 20, barxi, rmin, info);
      if (info != 1) { nearest_xi_sampling_2d(et, x, x_i, 25, 14, 2, barxi, rmin); method = 3; } else method = 2;
    }
    d = rmin / cl;
  } else method = 1;
}

// -------------------------------------------------------------------------------------
// Calculation element + precalculated datasets: lib/fbem/src/bem_general.f90:51-107, :163-759
// -------------------------------------------------------------------------------------
struct PSet { int gln, ngp; std::vector<double> x, n, pphijw; };  // sphijw == pphijw (type_f1 == type_f2, delta = 0)
struct Element {
  int et, nn; double x[27]; double cl; int gln_far; double bc[3], br; bool reverse;
  std::vector<PSet> ps; int ps_gln_max;
};
static void init_precalculated_datasets(Element& e, int n_ps, const int* ps_gln) {
  bool tri = (e.et == TRI3 || e.et == TRI6);
  e.ps.resize(n_ps); e.ps_gln_max = 0;
  for (int i = 0; i < n_ps; i++) {
    PSet& s = e.ps[i]; s.gln = ps_gln[i]; if (s.gln > e.ps_gln_max) e.ps_gln_max = s.gln;
    bool wan = tri && s.gln <= 15;
    s.ngp = wan ? wan_n(2 * s.gln - 1) : s.gln * s.gln;
    s.x.resize(3 * s.ngp); s.n.resize(3 * s.ngp); s.pphijw.resize(e.nn * s.ngp);
    for (int kt = 0; kt < s.ngp; kt++) {
      double xi[2], w1, w2 = 1.0;
      if (wan) { xi[0] = wan_x1(2 * s.gln - 1, kt); xi[1] = wan_x2(2 * s.gln - 1, kt); w1 = wan_w(2 * s.gln - 1, kt); }
      else {
        int k1 = kt / s.gln, k2 = kt % s.gln;  // kt=k2+(k1-1)*gln  (:473)
        if (tri) { xi[0] = (1.0 - gj01_x(s.gln, k2)) * gl01_x(s.gln, k1); xi[1] = gj01_x(s.gln, k2); w1 = gl01_w(s.gln, k1); w2 = gj01_w(s.gln, k2); }
        else { xi[0] = gl11_x(s.gln, k1); xi[1] = gl11_x(s.gln, k2); w1 = gl11_w(s.gln, k1); w2 = gl11_w(s.gln, k2); }
      }
      double phi[9], xp[3], N[3], j; geom_at(e.et, e.nn, e.x, xi, phi, xp, N, j);
      for (int c = 0; c < 3; c++) { s.x[3 * kt + c] = xp[c]; s.n[3 * kt + c] = N[c] / j; }
      for (int k = 0; k < e.nn; k++) s.pphijw[e.nn * kt + k] = wan ? phi[k] * j * w1 : phi[k] * j * w1 * w2;
    }
  }
}

// plan statistics (for the algorithmic-work model of SURVEY.md section 8d)
struct Stats { long long pairs_regular[33], pts_regular, pairs_adaptive, leaves, pts_adaptive, pairs_singular, pts_singular, li_points; };
static void stats_add(Stats& a, const Stats& b) {
  for (int i = 0; i < 33; i++) a.pairs_regular[i] += b.pairs_regular[i];
  a.pts_regular += b.pts_regular; a.pairs_adaptive += b.pairs_adaptive; a.leaves += b.leaves; a.pts_adaptive += b.pts_adaptive;
  a.pairs_singular += b.pairs_singular; a.pts_singular += b.pts_singular; a.li_points += b.li_points;
}

// fbem_bem_harela3d_sbie_ext_pre: bem_harela3d.f90:628-700
// n_i != NULL: the hypersingular equation with the unit normal n_i at the collocation point (fbem_bem_harela3d_hbie_ext_pre
// :2573-2662, _ext_st :2664-3042, _ext_adp :3044-3167, _auto :3632-3697): the same traversal with the d*, s* point formulas,
// the estimator called with f = 7 instead of 5, h <- m (scaled by cte_s), g <- l (scaled by cte_d).
static void sbie_ext_pre(const PSet& s, const Element& e, const double* x_i, const Params& p, cd* h, cd* g, const double* n_i = nullptr) {
  int nn = e.nn;
  for (int i = 0; i < nblk(p) * nn; i++) { h[i] = 0.0; g[i] = 0.0; }
  for (int kip = 0; kip < s.ngp; kip++) {
    if (n_i) add_exterior_point_hbie(p, &s.x[3 * kip], &s.n[3 * kip], x_i, n_i, nn, &s.pphijw[nn * kip], &s.pphijw[nn * kip], h, g);
    else add_exterior_point(p, &s.x[3 * kip], &s.n[3 * kip], x_i, nn, &s.pphijw[nn * kip], &s.pphijw[nn * kip], h, g);
  }
  finish_pair(p, e.reverse, n_i, nn, h, g);
}
// fbem_bem_harela3d_sbie_ext_st: bem_harela3d.f90:702-1048
static void sbie_ext_st(const Element& e, const double* xi_s, const double* x_i, const double* barxip, double barr, const Params& p, int gln, cd* h, cd* g,
                        const double* n_i = nullptr) {
  int nn = e.nn; bool tri = (e.et == TRI3 || e.et == TRI6);
  for (int i = 0; i < nblk(p) * nn; i++) { h[i] = 0.0; g[i] = 0.0; }
  double tp1[4], tp2[4];
  if (!tri) { telles11_parameters(barxip[0], barr, tp1); telles11_parameters(barxip[1], barr, tp2); }
  else {
    double bpp[2];
    if (barxip[1] > 0.995) { bpp[0] = 0.5; bpp[1] = 1.0; } else { bpp[0] = barxip[0] / (1.0 - barxip[1]); bpp[1] = barxip[1]; }
    telles01_parameters(bpp[0], barr, tp1); telles01_parameters(bpp[1], barr, tp2);
  }
  for (int k1 = 0; k1 < gln; k1++) {
    double g1 = tri ? gl01_x(gln, k1) : gl11_x(gln, k1), w1 = tri ? gl01_w(gln, k1) : gl11_w(gln, k1), t1, jt1;
    telles_xi_jac(tp1, g1, t1, jt1);
    for (int k2 = 0; k2 < gln; k2++) {
      double g2 = tri ? gl01_x(gln, k2) : gl11_x(gln, k2), w2 = tri ? gl01_w(gln, k2) : gl11_w(gln, k2), t2, jt2;
      telles_xi_jac(tp2, g2, t2, jt2);
      double xip[2], jqt = 1.0;
      if (tri) { xip[0] = (1.0 - t2) * t1; xip[1] = t2; jqt = 1.0 - t2; } else { xip[0] = t1; xip[1] = t2; }
      // XIP->XI (sub-element map) and its jacobian js
      int nv = tri ? 3 : 4; double sp[4], sd1[4], sd2[4];
      phi2d<double>(tri ? TRI3 : QUAD4, xip, sp); dphi2d<double>(tri ? TRI3 : QUAD4, xip, sd1, sd2);
      double xi[2] = {0, 0}, dx1[2] = {0, 0}, dx2[2] = {0, 0};
      for (int k = 0; k < nv; k++) for (int c = 0; c < 2; c++) {
        xi[c] = xi[c] + sp[k] * xi_s[2 * k + c]; dx1[c] = dx1[c] + sd1[k] * xi_s[2 * k + c]; dx2[c] = dx2[c] + sd2[k] * xi_s[2 * k + c]; }
      double js = dx1[0] * dx2[1] - dx1[1] * dx2[0];
      double phi[9], x[3], N[3], jg; geom_at(e.et, nn, e.x, xi, phi, x, N, jg);
      double n[3] = {N[0] / jg, N[1] / jg, N[2] / jg};
      double jw = tri ? jg * js * jqt * jt1 * jt2 * w1 * w2 : jg * js * jt1 * jt2 * w1 * w2;
      double pj[9]; for (int k = 0; k < nn; k++) pj[k] = phi[k] * jw;
      if (n_i) add_exterior_point_hbie(p, x, n, x_i, n_i, nn, pj, pj, h, g);
      else add_exterior_point(p, x, n, x_i, nn, pj, pj, h, g);
    }
  }
  finish_pair(p, e.reverse, n_i, nn, h, g);
}
// fbem_bem_harela3d_sbie_ext_adp: bem_harela3d.f90:1050-1172
static void sbie_ext_adp(const Element& e, double* xi_s, const double* x_i, const Params& p, const QsParams& qsp, int ks, int ns, cd* h, cd* g, Stats& st,
                         const double* n_i = nullptr) {
  int nn = e.nn, nv = n_vertices_of(e.et);
  double barxip[2], rmin, d; int method;
  if (ks == 1) {
    for (int i = 0; i < nblk(p) * nn; i++) { h[i] = 0.0; g[i] = 0.0; }
    if (nv == 3) { xi_s[0] = 1; xi_s[1] = 0; xi_s[2] = 0; xi_s[3] = 1; xi_s[4] = 0; xi_s[5] = 0; }
    else { xi_s[0] = -1; xi_s[1] = -1; xi_s[2] = 1; xi_s[3] = -1; xi_s[4] = 1; xi_s[5] = 1; xi_s[6] = -1; xi_s[7] = 1; }
    nearest_element_point_bem(e.et, e.x, e.cl, x_i, barxip, rmin, d, method);
  } else {
    double x_s[27]; subdivision_coordinates(e.et, e.x, xi_s, x_s);
    double cl = characteristic_length(e.et, x_s, 1.e-12);
    nearest_element_point_bem(e.et, x_s, cl, x_i, barxip, rmin, d, method);
  }
  int gln_near = qs_n_estimation(true, e.et, estimator_f(p, n_i), qsp, d, barxip);
  bool subdivide = false;
  if (ks == ns) { if (gln_near == 0) gln_near = 30; } else if (gln_near == 0) subdivide = true;
  if (subdivide) {
    double t[8];
    auto mid = [&](int a, int b, double* o) { o[0] = 0.50 * (xi_s[2 * a] + xi_s[2 * b]); o[1] = 0.50 * (xi_s[2 * a + 1] + xi_s[2 * b + 1]); };
    auto cpy = [&](int a, double* o) { o[0] = xi_s[2 * a]; o[1] = xi_s[2 * a + 1]; };
    if (nv == 3) {
      cpy(0, t); mid(0, 1, t + 2); mid(0, 2, t + 4); sbie_ext_adp(e, t, x_i, p, qsp, ks + 1, ns, h, g, st, n_i);
      cpy(1, t); mid(1, 2, t + 2); mid(0, 1, t + 4); sbie_ext_adp(e, t, x_i, p, qsp, ks + 1, ns, h, g, st, n_i);
      cpy(2, t); mid(0, 2, t + 2); mid(1, 2, t + 4); sbie_ext_adp(e, t, x_i, p, qsp, ks + 1, ns, h, g, st, n_i);
      mid(0, 1, t); mid(1, 2, t + 2); mid(0, 2, t + 4); sbie_ext_adp(e, t, x_i, p, qsp, ks + 1, ns, h, g, st, n_i);
    } else {
      auto ctr = [&](double* o) { o[0] = 0.25 * (xi_s[0] + xi_s[2] + xi_s[4] + xi_s[6]); o[1] = 0.25 * (xi_s[1] + xi_s[3] + xi_s[5] + xi_s[7]); };
      cpy(0, t); mid(0, 1, t + 2); ctr(t + 4); mid(0, 3, t + 6); sbie_ext_adp(e, t, x_i, p, qsp, ks + 1, ns, h, g, st, n_i);
      mid(0, 1, t); cpy(1, t + 2); mid(1, 2, t + 4); ctr(t + 6); sbie_ext_adp(e, t, x_i, p, qsp, ks + 1, ns, h, g, st, n_i);
      ctr(t); mid(1, 2, t + 2); cpy(2, t + 4); mid(2, 3, t + 6); sbie_ext_adp(e, t, x_i, p, qsp, ks + 1, ns, h, g, st, n_i);
      mid(0, 3, t); ctr(t + 2); mid(2, 3, t + 4); cpy(3, t + 6); sbie_ext_adp(e, t, x_i, p, qsp, ks + 1, ns, h, g, st, n_i);
    }
  } else {
    double barr = telles_barr_any(d);
    int gln = std::max(gln_near, e.gln_far);
    cd ht[144], gt[144];
    sbie_ext_st(e, xi_s, x_i, barxip, barr, p, gln, ht, gt, n_i);
    for (int i = 0; i < nblk(p) * nn; i++) { h[i] = h[i] + ht[i]; g[i] = g[i] + gt[i]; }
    st.leaves++; st.pts_adaptive += (long long)gln * gln;
  }
}

// fbem_polar_transformation_setup: polar_transformation.f90:251-498
static void polar_setup(int et, const double* xi_i, int& nsub, int* sub, double th[8][2], double thp[8][2]) {
  bool tri = (et == TRI3 || et == TRI6); const double tol = check_xi_tol;
  bool in_edge = check_xi1xi2_edge(et, xi_i);
  nsub = 0;
  if (!in_edge) { nsub = tri ? 6 : 8; for (int k = 0; k < nsub; k++) sub[k] = k + 1; }
  else if (!tri) {
    double a = xi_i[0], b = xi_i[1];
    if (a <= -1.0 + tol && b <= -1.0 + tol) { nsub = 2; sub[0] = 4; sub[1] = 5; }
    if (a >= 1.0 - tol && b <= -1.0 + tol) { nsub = 2; sub[0] = 6; sub[1] = 7; }
    if (a >= 1.0 - tol && b >= 1.0 - tol) { nsub = 2; sub[0] = 8; sub[1] = 1; }
    if (a <= -1.0 + tol && b >= 1.0 - tol) { nsub = 2; sub[0] = 2; sub[1] = 3; }
    if (nsub == 0) {
      if (b <= -1.0 + tol) { nsub = 4; sub[0] = 4; sub[1] = 5; sub[2] = 6; sub[3] = 7; }
      if (a >= 1.0 - tol) { nsub = 4; sub[0] = 6; sub[1] = 7; sub[2] = 8; sub[3] = 1; }
      if (b >= 1.0 - tol) { nsub = 4; sub[0] = 8; sub[1] = 1; sub[2] = 2; sub[3] = 3; }
      if (a <= -1.0 + tol) { nsub = 4; sub[0] = 2; sub[1] = 3; sub[2] = 4; sub[3] = 5; }
    }
  } else {
    double a = xi_i[0], b = xi_i[1];
    if (a >= 1.0 - tol) { nsub = 1; sub[0] = 3; }
    if (b >= 1.0 - tol) { nsub = 1; sub[0] = 6; }
    if (a <= tol && b <= tol) { nsub = 2; sub[0] = 1; sub[1] = 2; }
    if (nsub == 0) {
      if ((a + b) >= 1.0 - tol) { nsub = 4; sub[0] = 3; sub[1] = 4; sub[2] = 5; sub[3] = 6; }
      if (a <= tol) { nsub = 3; sub[0] = 6; sub[1] = 1; sub[2] = 2; }
      if (b <= tol) { nsub = 3; sub[0] = 1; sub[1] = 2; sub[2] = 3; }
    }
  }
  for (int k = 0; k < nsub; k++) {
    double a = xi_i[0], b = xi_i[1], t;
    if (!tri) {
      switch (sub[k]) {
        case 1: t = c_pi - asin((-1.0 - b) / sqrt((-1.0 - a) * (-1.0 - a) + (-1.0 - b) * (-1.0 - b)));
          th[k][0] = t; th[k][1] = 1.5 * c_pi;
          thp[k][0] = (1.0 + b) * log(tan(0.5 * (th[k][0] - c_pi))); thp[k][1] = (1.0 + b) * log(tan(0.5 * (th[k][1] - c_pi))); break;
        case 2: t = c_2pi + asin((-1.0 - b) / sqrt((1.0 - a) * (1.0 - a) + (-1.0 - b) * (-1.0 - b)));
          th[k][0] = 1.5 * c_pi; th[k][1] = t;
          thp[k][0] = (1.0 + b) * log(tan(0.5 * (th[k][0] - c_pi))); thp[k][1] = (1.0 + b) * log(tan(0.5 * (th[k][1] - c_pi))); break;
        case 3: t = c_2pi + asin((-1.0 - b) / sqrt((1.0 - a) * (1.0 - a) + (-1.0 - b) * (-1.0 - b)));
          th[k][0] = t; th[k][1] = c_2pi;
          thp[k][0] = (1.0 - a) * log(tan(0.5 * (th[k][0] + c_pi_2))); thp[k][1] = (1.0 - a) * log(tan(0.5 * (th[k][1] + c_pi_2))); break;
        case 4: t = asin((1.0 - b) / sqrt((1.0 - a) * (1.0 - a) + (1.0 - b) * (1.0 - b)));
          th[k][0] = 0.0; th[k][1] = t;
          thp[k][0] = (1.0 - a) * log(tan(0.5 * (th[k][0] + c_pi_2))); thp[k][1] = (1.0 - a) * log(tan(0.5 * (th[k][1] + c_pi_2))); break;
        case 5: t = asin((1.0 - b) / sqrt((1.0 - a) * (1.0 - a) + (1.0 - b) * (1.0 - b)));
          th[k][0] = t; th[k][1] = c_pi_2;
          thp[k][0] = (1.0 - b) * log(tan(0.5 * th[k][0])); thp[k][1] = (1.0 - b) * log(tan(0.5 * th[k][1])); break;
        case 6: t = c_pi - asin((1.0 - b) / sqrt((-1.0 - a) * (-1.0 - a) + (1.0 - b) * (1.0 - b)));
          th[k][0] = c_pi_2; th[k][1] = t;
          thp[k][0] = (1.0 - b) * log(tan(0.5 * th[k][0])); thp[k][1] = (1.0 - b) * log(tan(0.5 * th[k][1])); break;
        case 7: t = c_pi - asin((1.0 - b) / sqrt((-1.0 - a) * (-1.0 - a) + (1.0 - b) * (1.0 - b)));
          th[k][0] = t; th[k][1] = c_pi;
          thp[k][0] = (1.0 + a) * log(tan(0.5 * (th[k][0] - c_pi_2))); thp[k][1] = (1.0 + a) * log(tan(0.5 * (th[k][1] - c_pi_2))); break;
        case 8: t = c_pi - asin((-1.0 - b) / sqrt((-1.0 - a) * (-1.0 - a) + (-1.0 - b) * (-1.0 - b)));
          th[k][0] = c_pi; th[k][1] = t;
          thp[k][0] = (1.0 + a) * log(tan(0.5 * (th[k][0] - c_pi_2))); thp[k][1] = (1.0 + a) * log(tan(0.5 * (th[k][1] - c_pi_2))); break;
      }
    } else {
      switch (sub[k]) {
        case 1: t = c_2pi - asin(b / sqrt((1.0 - a) * (1.0 - a) + b * b));
          th[k][0] = t; th[k][1] = c_pi_4 + c_2pi;
          thp[k][0] = (1.0 - a - b) / c_sqrt2 * log(tan(0.5 * (th[k][0] + c_pi_4))); thp[k][1] = (1.0 - a - b) / c_sqrt2 * log(tan(0.5 * (th[k][1] + c_pi_4))); break;
        case 2: t = c_pi - asin((1.0 - b) / sqrt(a * a + (1.0 - b) * (1.0 - b)));
          th[k][0] = c_pi_4; th[k][1] = t;
          thp[k][0] = (1.0 - a - b) / c_sqrt2 * log(tan(0.5 * (th[k][0] + c_pi_4))); thp[k][1] = (1.0 - a - b) / c_sqrt2 * log(tan(0.5 * (th[k][1] + c_pi_4))); break;
        case 3: t = c_pi - asin((1.0 - b) / sqrt(a * a + (1.0 - b) * (1.0 - b)));
          th[k][0] = t; th[k][1] = c_pi;
          thp[k][0] = a * log(tan(0.5 * (th[k][0] - c_pi_2))); thp[k][1] = a * log(tan(0.5 * (th[k][1] - c_pi_2))); break;
        case 4: t = c_pi + asin(b / sqrt(a * a + b * b));
          th[k][0] = c_pi; th[k][1] = t;
          thp[k][0] = a * log(tan(0.5 * (th[k][0] - c_pi_2))); thp[k][1] = a * log(tan(0.5 * (th[k][1] - c_pi_2))); break;
        case 5: t = c_pi + asin(b / sqrt(a * a + b * b));
          th[k][0] = t; th[k][1] = 1.5 * c_pi;
          thp[k][0] = b * log(tan(0.5 * (th[k][0] - c_pi))); thp[k][1] = b * log(tan(0.5 * (th[k][1] - c_pi))); break;
        case 6: t = c_2pi - asin(b / sqrt((1.0 - a) * (1.0 - a) + b * b));
          th[k][0] = 1.5 * c_pi; th[k][1] = t;
          thp[k][0] = b * log(tan(0.5 * (th[k][0] - c_pi))); thp[k][1] = b * log(tan(0.5 * (th[k][1] - c_pi))); break;
      }
    }
  }
}
// fbem_polar_transformation_angular: polar_transformation.f90:501-544
static void polar_angular(int et, const double* xi_i, int sub, double thetap, double& theta, double& rhoij) {
  bool tri = (et == TRI3 || et == TRI6); double a = xi_i[0], b = xi_i[1];
  if (!tri) {
    switch (sub) {
      case 1: case 2: theta = 2.0 * atan(exp(thetap / (1.0 + b))) + c_pi; rhoij = (-1.0 - b) / sin(theta); break;
      case 3: case 4: theta = 2.0 * atan(exp(thetap / (1.0 - a))) - c_pi_2; rhoij = (1.0 - a) / cos(theta); break;
      case 5: case 6: theta = 2.0 * atan(exp(thetap / (1.0 - b))); rhoij = (1.0 - b) / sin(theta); break;
      default: theta = 2.0 * atan(exp(thetap / (1.0 + a))) + c_pi_2; rhoij = (-1.0 - a) / cos(theta); break;
    }
  } else {
    switch (sub) {
      case 1: case 2: theta = 2.0 * atan(exp(c_sqrt2 * thetap / (1.0 - a - b))) - c_pi_4; rhoij = (1.0 - a - b) / (cos(theta) + sin(theta)); break;
      case 3: case 4: theta = 2.0 * atan(exp(thetap / a)) + c_pi_2; rhoij = -a / cos(theta); break;
      default: theta = 2.0 * atan(exp(thetap / b)) + c_pi; rhoij = -b / sin(theta); break;
    }
  }
}
// fbem_bem_staela3d_sbie_int_li: lib/fbem/src/bem_staela3d.f90:2245-2376
static void staela3d_sbie_int_li(int et, const double* xn, double* xi_s, const double* x_i, int ngp_min, const QsParams& qsp, int ks, int ns, double hli[3][3], Stats& st) {
  int nn = n_nodes_of(et); double x_s[9];
  if (ks == 1) { xi_s[0] = -1.0; xi_s[1] = 1.0; for (int i = 0; i < 3 * nn; i++) x_s[i] = xn[i]; }
  else subdivision_coordinates(et, xn, xi_s, x_s);
  double cl = characteristic_length(et, x_s, 1.e-12), barxip[1], rmin, d; int method;
  nearest_element_point_bem(et, x_s, cl, x_i, barxip, rmin, d, method);
  int gln_near = qs_n_estimation(true, et, 1, qsp, d, barxip);
  bool subdivide = false;
  if (ks == ns) { if (gln_near == 0) gln_near = 30; } else if (gln_near == 0) subdivide = true;
  if (subdivide) {
    double t[2];
    t[0] = xi_s[0]; t[1] = 0.5 * (xi_s[0] + xi_s[1]); staela3d_sbie_int_li(et, xn, t, x_i, ngp_min, qsp, ks + 1, ns, hli, st);
    t[0] = 0.5 * (xi_s[0] + xi_s[1]); t[1] = xi_s[1]; staela3d_sbie_int_li(et, xn, t, x_i, ngp_min, qsp, ks + 1, ns, hli, st);
  } else {
    double ht[3][3] = {{0}};
    int gln = std::max(gln_near, ngp_min);
    double barr = telles_barr_any(d), tp[4]; telles11_parameters(barxip[0], barr, tp);
    for (int kip = 0; kip < gln; kip++) {
      double gam = gl11_x(gln, kip), w = gl11_w(gln, kip), xip, jt; telles_xi_jac(tp, gam, xip, jt);
      double xi = 0.5 * (1.0 - xip) * xi_s[0] + 0.5 * (1.0 + xip) * xi_s[1], js = 0.5 * (xi_s[1] - xi_s[0]);
      double gphi[3], dg[3], x[3] = {0, 0, 0}, T[3] = {0, 0, 0}; phi1d<double>(et, xi, gphi); dphi1d<double>(et, xi, dg);
      for (int k = 0; k < nn; k++) for (int c = 0; c < 3; c++) { x[c] = x[c] + gphi[k] * xn[3 * k + c]; T[c] = T[c] + dg[k] * xn[3 * k + c]; }
      double jg = sqrt(dot3(T, T)); double t[3] = {T[0] / jg, T[1] / jg, T[2] / jg};
      double rv[3] = {x[0] - x_i[0], x[1] - x_i[1], x[2] - x_i[2]}; double r = sqrt(dot3(rv, rv)), dr1 = 1.0 / r;
      double jw = jg * js * jt * w;
      ht[0][1] = ht[0][1] - dr1 * t[2] * jw; ht[0][2] = ht[0][2] + dr1 * t[1] * jw; ht[1][2] = ht[1][2] - dr1 * t[0] * jw;
    }
    ht[1][0] = -ht[0][1]; ht[2][0] = -ht[0][2]; ht[2][1] = -ht[1][2];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) hli[a][b] = hli[a][b] + ht[a][b];
    st.li_points += gln;
  }
}
// fbem_bem_harela3d_sbie_int: bem_harela3d.f90:1174-1472
static void sbie_int(const Element& e, const double* xi_i, const Params& p, cd* h, cd* g, Stats& st) {
  int nn = e.nn, et = e.et;
  for (int i = 0; i < nblk(p) * nn; i++) { h[i] = 0.0; g[i] = 0.0; }
  double phi_g[9], x_i[3] = {0, 0, 0}, phi_i[9];
  phi2d<double>(et, xi_i, phi_g);
  for (int k = 0; k < nn; k++) for (int c = 0; c < 3; c++) x_i[c] = x_i[c] + phi_g[k] * e.x[3 * k + c];
  phi2d<double>(et, xi_i, phi_i);
  int nsub, sub[8]; double th[8][2], thp[8][2];
  polar_setup(et, xi_i, nsub, sub, th, thp);
  for (int ks = 0; ks < nsub; ks++) {
    double thetai = th[ks][0], thetaf = th[ks][1], thetapi = thp[ks][0], thetapf = thp[ks][1];
    int ngp_rho = 15, ngp_theta = 5 + (int)lround(25.0 * (thetaf - thetai) / c_pi_2);
    for (int kt = 0; kt < ngp_theta; kt++) {
      double thetapp = gl01_x(ngp_theta, kt), w_ang = gl01_w(ngp_theta, kt), jthetap = thetapf - thetapi;
      double thetap = jthetap * thetapp + thetapi, theta, rhoij;
      polar_angular(et, xi_i, sub[ks], thetap, theta, rhoij);
      double ct = cos(theta), sn = sin(theta);
      for (int kr = 0; kr < ngp_rho; kr++) {
        double rhop = gl01_x(ngp_rho, kr), w_rad = gl01_w(ngp_rho, kr), rho = rhoij * rhop;
        double xi[2] = {xi_i[0] + rho * ct, xi_i[1] + rho * sn};
        double phi[9], x[3], N[3], jg; geom_at(et, nn, e.x, xi, phi, x, N, jg);
        double n[3] = {N[0] / jg, N[1] / jg, N[2] / jg};
        double rv[3] = {x[0] - x_i[0], x[1] - x_i[1], x[2] - x_i[2]};
        double r = sqrt(dot3(rv, rv));
        if (p.por) {   // fbem_bem_harpor3d_sbie_int: bem_harpor3d.f90:1690-1790 -- the solid block as in the elastic integrator (static 1/r^2 parts of
                       // T1 and of the dr/dn delta term integrated in full, the remaining T2(1)/r^2 part through the CPV term and the line integrals);
                       // the fluid and coupling blocks are weakly singular and integrated as they are
          PorScalars k; por_scalars(*p.por, r, true, k);
          const PorParams& P = *p.por;
          double drdx[3] = {rv[0] * k.d1r1, rv[1] * k.d1r1, rv[2] * k.d1r1}, drdn = dot3(drdx, n);
          double jw = jg * rho * jthetap * w_ang * w_rad;
          cd fs_u[4][4], fs_t[4][4];
          fs_u[0][0] = k.eta; fs_t[0][0] = (P.W0[1] * k.d1r2 + k.W0) * drdn;
          for (int c = 0; c < 3; c++) {
            fs_u[0][c + 1] = k.vartheta * drdx[c]; fs_u[c + 1][0] = k.vartheta * drdx[c];
            fs_t[0][c + 1] = k.T01 * drdx[c] * drdn + k.T02 * n[c]; fs_t[c + 1][0] = k.W1 * drdx[c] * drdn + k.W2 * n[c];
          }
          for (int ik = 0; ik < 3; ik++) for (int il = 0; il < 3; il++) {
            fs_u[il + 1][ik + 1] = k.psi * dkr[il][ik] - k.chi * drdx[il] * drdx[ik];
            fs_t[il + 1][ik + 1] = (P.T1[1] * k.d1r2 + k.TT1) * drdx[il] * drdx[ik] * drdn + (P.T2[1] * k.d1r2 + k.TT2) * drdn * dkr[il][ik] + k.TT2 * drdx[ik] * n[il]
                                 + k.TT3 * drdx[il] * n[ik];
          }
          for (int j = 0; j < nn; j++) {
            double fjw = phi[j] * jw;
            for (int a = 0; a < 4; a++) for (int c = 0; c < 4; c++) { h[(j * 4 + a) * 4 + c] += fs_t[a][c] * fjw; g[(j * 4 + a) * 4 + c] += fs_u[a][c] * fjw; }
            for (int ik = 0; ik < 3; ik++) for (int il = 0; il < 3; il++)
              h[(j * 4 + il + 1) * 4 + ik + 1] += P.T2[1] * k.d1r2 * (n[il] * drdx[ik] - n[ik] * drdx[il]) * (phi[j] - phi_i[j]) * jw;
          }
          st.pts_singular++;
          continue;
        }
        if (p.pot) {   // fbem_bem_harpot3d_sbie_int: bem_harpot3d.f90:905-957 (weakly singular: no CPV part, no line integrals)
          double d1r = 1.0 / r, d1r2 = d1r * d1r;
          double drdn = dot3(rv, N) * d1r / jg;
          double jw = jg * rho * jthetap * w_ang * w_rad;
          cd E[5]; decomposed_zexp(cd(-0.0, -1.0) * p.kp * r, E);
          cd fs_P = d1r + p.P1 + d1r * E[2];
          cd fs_Q = d1r2 + p.Q1 + p.Q2 * d1r * E[2] + d1r2 * E[3];
          cd fq = fs_Q * drdn;
          for (int j = 0; j < nn; j++) { double fjw = phi[j] * jw; h[j] += fq * fjw; g[j] += fs_P * fjw; }
          st.pts_singular++;
          continue;
        }
        if (p.statics) {   // fbem_bem_staela3d_sbie_int: bem_staela3d.f90:1284-1320
          double d1r = 1.0 / r, d1r2 = d1r * d1r;
          double drdx[3] = {rv[0] * d1r, rv[1] * d1r, rv[2] * d1r}, drdn = dot3(drdx, n);
          double jw = jg * rho * jthetap * w_ang * w_rad;
          double fjw[9]; for (int j = 0; j < nn; j++) fjw[j] = phi[j] * jw;
          for (int il = 0; il < 3; il++) for (int ik = 0; ik < 3; ik++) {
            double fs_u = d1r * (p.cteu2 * dkr[il][ik] + drdx[il] * drdx[ik]);
            double fs_t = d1r2 * drdn * (p.ctet2 * dkr[il][ik] + 3.0 * drdx[il] * drdx[ik]);
            double fs_c = d1r2 * p.ctet2 * (n[il] * drdx[ik] - n[ik] * drdx[il]);
            for (int j = 0; j < nn; j++) {
              h[(j * 3 + il) * 3 + ik] += fs_t * fjw[j];
              g[(j * 3 + il) * 3 + ik] += fs_u * fjw[j];
              h[(j * 3 + il) * 3 + ik] += fs_c * (phi[j] - phi_i[j]) * jw;
            }
          }
          st.pts_singular++;
          continue;
        }
        KernelScalars k; kernel_scalars(p, r, true, k);
        double dr1 = k.d1r1, dr2 = k.d1r2;
        double drdx[3] = {rv[0] * dr1, rv[1] * dr1, rv[2] * dr1}, drdn = dot3(drdx, n);
        double jw = jg * rho * jthetap * w_ang * w_rad;
        double fjw[9]; for (int j = 0; j < nn; j++) fjw[j] = phi[j] * jw;
        for (int il = 0; il < 3; il++) for (int ik = 0; ik < 3; ik++) {
          cd fs_u = k.psi * dkr[il][ik] - k.chi * drdx[il] * drdx[ik];
          cd fs_t = (k.TT1 + p.T1[1] * dr2) * drdx[il] * drdx[ik] * drdn + (k.TT2 + p.T2[1] * dr2) * drdn * dkr[il][ik] + k.TT2 * drdx[ik] * n[il] + k.TT3 * drdx[il] * n[ik];
          cd fs_c = p.T2[1] * dr2 * (n[il] * drdx[ik] - n[ik] * drdx[il]);
          for (int j = 0; j < nn; j++) {
            h[(j * 3 + il) * 3 + ik] += fs_t * fjw[j];
            g[(j * 3 + il) * 3 + ik] += fs_u * fjw[j];
            h[(j * 3 + il) * 3 + ik] += fs_c * (phi[j] - phi_i[j]) * jw;
          }
        }
        st.pts_singular++;
      }
    }
  }
  if (p.pot) {
    for (int i = 0; i < nn; i++) { h[i] = p.cte_t * h[i]; g[i] = p.cte_u * g[i]; }
    if (e.reverse) for (int i = 0; i < nn; i++) h[i] = -h[i];
    return;
  }
  // line integrals
  double hli[3][3] = {{0}}; QsParams qsl; qs_calculate_parameters(1.e-15, qsl);
  int nedges = n_edges_of(et), ety = edge_type_of(et), nne = n_nodes_of(ety);
  for (int ke = 1; ke <= nedges; ke++) {
    bool integrate = false;
    for (int ks = 0; ks < nsub; ks++) if (sub[ks] == 2 * ke - 1 || sub[ks] == 2 * ke) { integrate = true; break; }
    if (integrate) {
      double xe[9]; for (int k = 0; k < nne; k++) for (int c = 0; c < 3; c++) xe[3 * k + c] = e.x[3 * edge_node(k, ke - 1, et) + c];
      double xi_s[2]; staela3d_sbie_int_li(ety, xe, xi_s, x_i, 5, qsl, 1, 16, hli, st);
    }
  }
  if (p.por) {   // bem_harpor3d.f90:1876-1880: the line-integral term of the solid block
    for (int il = 0; il < 3; il++) for (int ik = 0; ik < 3; ik++) for (int j = 0; j < nn; j++) h[(j * 4 + il + 1) * 4 + ik + 1] += phi_i[j] * p.por->T2[1] * hli[il][ik];
    finish_pair(p, e.reverse, nullptr, nn, h, g);
    return;
  }
  const cd c_li = p.statics ? cd(p.ctet2) : p.T2[1];   // bem_staela3d.f90:1373 / bem_harela3d.f90:1462-1466
  for (int il = 0; il < 3; il++) for (int ik = 0; ik < 3; ik++) for (int j = 0; j < nn; j++) h[(j * 3 + il) * 3 + ik] += phi_i[j] * c_li * hli[il][ik];
  finish_pair(p, e.reverse, nullptr, nn, h, g);
}
// fbem_bem_harela3d_sbie_auto: bem_harela3d.f90:1474-1538.  Returns mode: 1..30 regular gln (ps), 100 adaptive, 200 singular.
static int sbie_auto(const Element& e, const double* x_i, const Params& p, const QsParams& qsp, int ns, cd* h, cd* g, Stats& st, const double* n_i = nullptr) {
  double r[3] = {e.bc[0] - x_i[0], e.bc[1] - x_i[1], e.bc[2] - x_i[2]};
  double rmin = sqrt(dot3(r, r)) - e.br, barxi[2], d; int delta, method;
  if (rmin > (4.0 * e.br)) { delta = 0; barxi[0] = 0.0; barxi[1] = 0.0; d = rmin / e.cl; }
  else { nearest_element_point_bem(e.et, e.x, e.cl, x_i, barxi, rmin, d, method); delta = (d <= 1.e-12) ? 1 : 0; }
  if (delta == 1 && n_i) return -1;   // fbem_bem_harela3d_hbie_int (collocation point ON the element) is not restated: interior points never get here
  if (delta == 1) { st.pairs_singular++; sbie_int(e, barxi, p, h, g, st); return 200; }
  int gln_near = qs_n_estimation(false, e.et, estimator_f(p, n_i), qsp, d, barxi);
  int gln = std::max(e.gln_far, gln_near);
  if (gln <= e.ps_gln_max && gln_near > 0) {
    int ps = 0; for (size_t i = 0; i < e.ps.size(); i++) if (e.ps[i].gln >= gln) { ps = (int)i; break; }
    sbie_ext_pre(e.ps[ps], e, x_i, p, h, g, n_i);
    st.pairs_regular[e.ps[ps].gln]++; st.pts_regular += e.ps[ps].ngp;
    return e.ps[ps].gln;
  }
  double xi_s[8]; st.pairs_adaptive++;
  sbie_ext_adp(e, xi_s, x_i, p, qsp, 1, ns, h, g, st, n_i);
  return 100;
}

// fbem_bem_harela3d_sbie_freeterm (Mantic C-matrix): bem_harela3d.f90:365-542
// cp_out: the scalar free term c = (2 pi + sum_a)/(4 pi) alone == fbem_bem_pot3d_sbie_freeterm (lib/fbem/src/bem_stapot3d.f90:155-296:
// the same sorting of the tangents and the same sum of dihedral angles, without the tensor part)
static int sbie_freeterm(int ne, const double* n_in, const double* t_in, double tol, cd nu, cd c[3][3], double* cp_out = nullptr) {
  double ltol = (tol < 1.0e-12 || tol > 1.0e-3) ? 1.0e-6 : tol;
  std::vector<double> ln(3 * (ne + 2)), lt(3 * (ne + 2)), lti(3 * (ne + 1)), theta(ne + 1);
  std::vector<int> tc(ne + 1);
  for (int i = 1; i <= ne; i++) for (int k = 0; k < 3; k++) { ln[3 * i + k] = n_in[3 * (i - 1) + k]; lt[3 * i + k] = t_in[3 * (i - 1) + k]; }
  for (int ki = 1; ki <= ne - 1; ki++) {
    double e1[3], e2[3], e3[3];
    for (int k = 0; k < 3; k++) { e1[k] = lt[3 * ki + k]; e3[k] = ln[3 * ki + k]; }
    e2[0] = e3[1] * e1[2] - e3[2] * e1[1]; e2[1] = e3[2] * e1[0] - e3[0] * e1[2]; e2[2] = e3[0] * e1[1] - e3[1] * e1[0];
    for (int kj = 1; kj <= ne; kj++) {
      lti[3 * kj + 0] = e1[0] * lt[3 * kj] + e1[1] * lt[3 * kj + 1] + e1[2] * lt[3 * kj + 2];
      lti[3 * kj + 1] = e2[0] * lt[3 * kj] + e2[1] * lt[3 * kj + 1] + e2[2] * lt[3 * kj + 2];
      lti[3 * kj + 2] = e3[0] * lt[3 * kj] + e3[1] * lt[3 * kj + 1] + e3[2] * lt[3 * kj + 2];
    }
    int ntc = 0;
    for (int kj = ki + 1; kj <= ne; kj++) if (fabs(lti[3 * kj + 2]) <= ltol) tc[ntc++] = kj;
    if (ntc == 0) return 1;  // 'the normals/tangents configuration is not valid'
    for (int kj = 0; kj < ntc; kj++) { theta[tc[kj]] = atan2(lti[3 * tc[kj] + 1], lti[3 * tc[kj]]); if (theta[tc[kj]] < 0.0) theta[tc[kj]] = theta[tc[kj]] + 2.0 * c_pi; }
    double mint = theta[tc[0]]; int minkj = tc[0];
    for (int kj = 1; kj < ntc; kj++) if (theta[tc[kj]] < mint) { mint = theta[tc[kj]]; minkj = tc[kj]; }
    for (int k = 0; k < 3; k++) { std::swap(lt[3 * (ki + 1) + k], lt[3 * minkj + k]); std::swap(ln[3 * (ki + 1) + k], ln[3 * minkj + k]); }
  }
  for (int k = 0; k < 3; k++) { ln[k] = ln[3 * ne + k]; lt[k] = lt[3 * ne + k]; ln[3 * (ne + 1) + k] = ln[3 + k]; lt[3 * (ne + 1) + k] = lt[3 + k]; }
  double sum_a = 0.0;
  for (int ki = 1; ki <= ne; ki++) {
    const double *a = &ln[3 * (ki - 1)], *b = &ln[3 * ki];
    double nxn[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    double nxndr = nxn[0] * lt[3 * ki] + nxn[1] * lt[3 * ki + 1] + nxn[2] * lt[3 * ki + 2];
    if (nxndr < 0.0) nxndr = -1.0;
    if (nxndr > 0.0) nxndr = 1.0;
    double ndn = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    if (ndn > 1.0) ndn = 1.0;
    sum_a = sum_a + nxndr * acos(ndn);
  }
  double cp = 1.0 / (4.0 * c_pi) * (2.0 * c_pi + sum_a);
  if (cp_out) *cp_out = cp;
  double sum_b[3][3] = {{0}};
  for (int ki = 1; ki <= ne; ki++) {
    double rmr[3]; for (int k = 0; k < 3; k++) rmr[k] = lt[3 * (ki + 1) + k] - lt[3 * ki + k];
    const double* nn = &ln[3 * ki];
    double v[3] = {rmr[1] * nn[2] - rmr[2] * nn[1], rmr[2] * nn[0] - rmr[0] * nn[2], rmr[0] * nn[1] - rmr[1] * nn[0]};
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) sum_b[a][b] = sum_b[a][b] + v[a] * nn[b];
  }
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
    c[a][b] = -1.0 / (8.0 * c_pi * (1.0 - nu)) * sum_b[a][b];
    if (a == b) c[a][b] = c[a][b] + cp;
  }
  return 0;
}
// unit normal and element-boundary tangents at an element node: src/build_data_at_geometrical_nodes.f90:198-222,
// lib/fbem/src/geometry.f90:657-914 (fbem_utangents_at_boundary)
static void node_normal_tangents(int et, const double* xn, int node, double* n, double* tbp, double* tbm) {
  int nn = n_nodes_of(et); double xi[2]; xi_at_node(et, node, xi);
  double d1[9], d2[9]; dphi2d<double>(et, xi, d1, d2);
  double T1[3] = {0, 0, 0}, T2[3] = {0, 0, 0};
  for (int k = 0; k < nn; k++) for (int c = 0; c < 3; c++) { T1[c] = T1[c] + d1[k] * xn[3 * k + c]; T2[c] = T2[c] + d2[k] * xn[3 * k + c]; }
  double N[3] = {T1[1] * T2[2] - T1[2] * T2[1], T1[2] * T2[0] - T1[0] * T2[2], T1[0] * T2[1] - T1[1] * T2[0]};
  double jn = sqrt(dot3(N, N)); for (int c = 0; c < 3; c++) n[c] = N[c] / jn;
  double n1 = sqrt(T1[0] * T1[0] + T1[1] * T1[1] + T1[2] * T1[2]), n2 = sqrt(T2[0] * T2[0] + T2[1] * T2[1] + T2[2] * T2[2]);
  double t1[3] = {T1[0] / n1, T1[1] / n1, T1[2] / n1}, t2[3] = {T2[0] / n2, T2[1] / n2, T2[2] / n2};
  auto set = [&](double* o, const double* v, double s) { for (int c = 0; c < 3; c++) o[c] = s * v[c]; };
  if (et == TRI3 || et == TRI6) {
    double d3[6] = {0, 0, 0, 0, 0, 0};
    if (et == TRI3) { d3[0] = 1.0; d3[1] = -1.0; }
    else { d3[0] = 4.0 * xi[0] - 1.0; d3[1] = 4.0 * xi[0] - 3.0; d3[3] = 4.0 * (1.0 - 2.0 * xi[0]); }
    double T3[3] = {0, 0, 0}; for (int k = 0; k < nn; k++) for (int c = 0; c < 3; c++) T3[c] = T3[c] + d3[k] * xn[3 * k + c];
    double n3 = sqrt(T3[0] * T3[0] + T3[1] * T3[1] + T3[2] * T3[2]); double t3[3] = {T3[0] / n3, T3[1] / n3, T3[2] / n3};
    switch (node) {
      case 0: set(tbp, t3, -1); set(tbm, t1, -1); break; case 1: set(tbp, t2, -1); set(tbm, t3, 1); break;
      case 2: set(tbp, t1, 1); set(tbm, t2, 1); break;   case 3: set(tbp, t3, -1); set(tbm, t3, 1); break;
      case 4: set(tbp, t2, -1); set(tbm, t2, 1); break;  default: set(tbp, t1, 1); set(tbm, t1, -1); break;
    }
  } else {
    switch (node) {
      case 0: set(tbp, t1, 1); set(tbm, t2, 1); break;   case 1: set(tbp, t2, 1); set(tbm, t1, -1); break;
      case 2: set(tbp, t1, -1); set(tbm, t2, -1); break; case 3: set(tbp, t2, -1); set(tbm, t1, 1); break;
      case 4: set(tbp, t1, 1); set(tbm, t1, -1); break;  case 5: set(tbp, t2, 1); set(tbm, t2, -1); break;
      case 6: set(tbp, t1, -1); set(tbm, t1, 1); break;  case 7: set(tbp, t2, -1); set(tbm, t2, 1); break;
      default: set(tbp, t1, 0); set(tbm, t1, 0); break;
    }
  }
}

// =====================================================================================
// Model handle + assembly driver
//   src/build_data_of_be_elements.f90:62-110 (csize, n_phi, bounding ball)
//   src/build_lse_mechanics_bem_harela.f90:227-238 (element loop), :273-747 (free terms), :972-1324 (collocation loop)
//   src/assemble_bem_harela_equation.f90:78-113 (ordinary BE boundary scatter)
// =====================================================================================
struct Model {
  int n_node, n_elem, n_colloc, n_dof;
  std::vector<Element> elem; std::vector<int> eptr, enode;
  std::vector<double> cx; std::vector<int> cnode, celem, ckn; std::vector<double> cxi;
  std::vector<int> row, col_u, col_t, ctype;
  int nd = 3;   // equations / unknowns per node: 3 (elastic solid), 1 (inviscid fluid: row(1), col(1) = p, col(2) = Un)
  QsParams qsp; int ns_max; double geometric_tolerance;
  std::vector<int> n2e_ptr, n2e_elem, n2e_kn;  // node -> (element, local node) incidences
  Stats last;
  // symmetry planes (lib/fbem/src/symmetry.f90:60-171, src/read_symmetry_planes.f90:228-283): image ks of every element, ks = 1..n_sym-1
  std::vector<int> ps_gln; std::vector<double> node_x;
  int n_planes = 0, n_sym = 1, plane_eid[3] = {0, 0, 0};
  double plane_m[3][3], conf_m[8][3], conf_t[8][3], conf_s[8]; bool conf_rev[8];
  std::vector<std::vector<Element>> img;
  // incident field at the nodes of every element (element()%incident_c): [(eptr[e] + kn) * 3 + ik], empty = none
  std::vector<cd> u_inc, t_inc;
  std::vector<double> n_fn;   // nodal unit normals node()%n_fn (ctype 10), empty = none
};

extern "C" {

void* orc_setup(int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr, const int* elem_node,
                const unsigned char* elem_reversed, int n_colloc, const double* colloc_x, const int* colloc_node,
                const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                const int* row, const int* col_u, const int* col_t, const int* ctype, int n_dof,
                double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln, double geometric_tolerance) {
  Model* m = new Model();
  m->n_node = n_node; m->n_elem = n_elem; m->n_colloc = n_colloc; m->n_dof = n_dof;
  m->eptr.assign(elem_ptr, elem_ptr + n_elem + 1); m->enode.assign(elem_node, elem_node + elem_ptr[n_elem]);
  m->cx.assign(colloc_x, colloc_x + 3 * n_colloc); m->cnode.assign(colloc_node, colloc_node + n_colloc);
  m->celem.assign(colloc_elem, colloc_elem + n_colloc); m->ckn.assign(colloc_kn, colloc_kn + n_colloc);
  m->cxi.assign(colloc_xi, colloc_xi + 2 * n_colloc);
  m->row.assign(row, row + 3 * n_node); m->col_u.assign(col_u, col_u + 3 * n_node); m->col_t.assign(col_t, col_t + 3 * n_node);
  m->ctype.assign(ctype, ctype + 3 * n_node);
  qs_calculate_parameters(qsi_relative_error, m->qsp); m->ns_max = qsi_ns_max; m->geometric_tolerance = geometric_tolerance;
  m->ps_gln.assign(precalset_gln, precalset_gln + n_precalsets); m->node_x.assign(node_x, node_x + 3 * n_node);
  for (int ks = 0; ks < 8; ks++) { m->conf_s[ks] = 1.0; m->conf_rev[ks] = false; for (int c = 0; c < 3; c++) { m->conf_m[ks][c] = 1.0; m->conf_t[ks][c] = 1.0; } }
  m->elem.resize(n_elem);
  for (int e = 0; e < n_elem; e++) {
    Element& el = m->elem[e]; el.et = etype[e]; el.nn = n_nodes_of(el.et); el.reverse = elem_reversed[e] != 0;
    for (int k = 0; k < el.nn; k++) for (int c = 0; c < 3; c++) el.x[3 * k + c] = node_x[3 * elem_node[elem_ptr[e] + k] + c];
    el.cl = characteristic_length(el.et, el.x, 1.e-9);
    el.gln_far = phijac_ngp_2d(el.et, el.x, qsi_relative_error);
    element_ball(el.et, el.x, el.gln_far, el.bc, el.br);
    init_precalculated_datasets(el, n_precalsets, precalset_gln);
  }
  std::vector<int> cnt(n_node + 1, 0);
  for (int e = 0; e < n_elem; e++) for (int k = elem_ptr[e]; k < elem_ptr[e + 1]; k++) cnt[elem_node[k] + 1]++;
  for (int i = 0; i < n_node; i++) cnt[i + 1] += cnt[i];
  m->n2e_ptr = cnt; m->n2e_elem.resize(cnt[n_node]); m->n2e_kn.resize(cnt[n_node]);
  std::vector<int> pos(cnt.begin(), cnt.end() - 1);
  for (int e = 0; e < n_elem; e++) for (int k = elem_ptr[e]; k < elem_ptr[e + 1]; k++) { int nd = elem_node[k]; m->n2e_elem[pos[nd]] = e; m->n2e_kn[pos[nd]] = k - elem_ptr[e]; pos[nd]++; }
  memset(&m->last, 0, sizeof(Stats));
  return m;
}
void orc_free(void* h) { delete (Model*)h; }

// Symmetry planes through the origin, normal to axis eid[i] (1..3, ascending), with the translation multipliers t[3*i..] of
// src/read_symmetry_planes.f90:76-228 (symmetry: -1 on the normal axis, +1 elsewhere; antisymmetry: the opposite signs).
// The images follow the step table of fbem_symmetry_multipliers (lib/fbem/src/symmetry.f90:81-170): 1 root, 2 SP1, 3 SP1+SP2,
// 4 SP2, 5 SP3, 6 SP1+SP3, 7 SP1+SP2+SP3, 8 SP2+SP3; an odd number of reflections reverses the orientation.  As in
// build_lse_mechanics_bem_harela.f90:1052-1107 only the nodal coordinates of the calculation element are reflected: csize, n_phi
// and the bounding ball (centre included) stay those of the root element.
static int set_symmetry_impl(void* h, int n_planes, const int* eid, const double* t, const double* sc);
int orc_set_symmetry(void* h, int n_planes, const int* eid, const double* t) { return set_symmetry_impl(h, n_planes, eid, t, nullptr); }
// the same with the scalar multipliers symplane_s (fluid pressure / fluid-phase variables: +1 symmetry, -1 antisymmetry)
int orc_set_symmetry_s(void* h, int n_planes, const int* eid, const double* t, const double* sc) { return set_symmetry_impl(h, n_planes, eid, t, sc); }
static int set_symmetry_impl(void* h, int n_planes, const int* eid, const double* t, const double* sc) {
  Model* m = (Model*)h;
  if (n_planes < 0 || n_planes > 3) return 1;
  m->n_planes = n_planes; m->n_sym = 1 << n_planes; m->img.clear();
  double pt[3][3];
  for (int i = 0; i < n_planes; i++) {
    if (eid[i] < 1 || eid[i] > 3 || (i > 0 && eid[i] <= eid[i - 1])) return 1;
    m->plane_eid[i] = eid[i];
    for (int c = 0; c < 3; c++) { m->plane_m[i][c] = (c == eid[i] - 1) ? -1.0 : 1.0; pt[i][c] = t[3 * i + c]; }
  }
  static const int steps[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
  for (int ks = 0; ks < m->n_sym; ks++) {
    int cnt = 0;
    for (int c = 0; c < 3; c++) { m->conf_m[ks][c] = 1.0; m->conf_t[ks][c] = 1.0; }
    m->conf_s[ks] = 1.0;
    for (int i = 0; i < n_planes; i++) if (steps[ks][i]) {
      cnt++; for (int c = 0; c < 3; c++) { m->conf_m[ks][c] *= m->plane_m[i][c]; m->conf_t[ks][c] *= pt[i][c]; }
      // default: the multiplier of a tangential translation (symmetry +1, antisymmetry -1), as read_symmetry_planes.f90:160-228 pairs them
      m->conf_s[ks] *= sc ? sc[i] : pt[i][eid[i] % 3];
    }
    m->conf_rev[ks] = (cnt & 1) != 0;
  }
  m->img.resize(m->n_sym - 1);
  for (int ks = 1; ks < m->n_sym; ks++) {
    m->img[ks - 1] = m->elem;
    for (int e = 0; e < m->n_elem; e++) {
      Element& el = m->img[ks - 1][e];
      for (int k = 0; k < el.nn; k++) for (int c = 0; c < 3; c++) el.x[3 * k + c] = m->conf_m[ks][c] * m->elem[e].x[3 * k + c];
      el.reverse = m->elem[e].reverse != m->conf_rev[ks];
      init_precalculated_datasets(el, (int)m->ps_gln.size(), m->ps_gln.data());
    }
  }
  return 0;
}
void orc_set_node_normals(void* h, const double* n_fn) { Model* m = (Model*)h; m->n_fn.assign(n_fn, n_fn + 3 * (size_t)m->n_node); }
// incident field of the region: u_inc, t_inc [(elem_ptr[e] + kn) * 3 + ik] interleaved complex, or NULL to clear
void orc_set_incident(void* h, const double* u_ri, const double* t_ri) {
  Model* m = (Model*)h; m->u_inc.clear(); m->t_inc.clear();
  if (!u_ri || !t_ri) return;
  size_t n = (size_t)m->eptr[m->n_elem] * (size_t)m->nd;   // three values per element node for a solid, one (p_inc | Un_inc) for a fluid
  m->u_inc.assign((const cd*)u_ri, (const cd*)u_ri + n); m->t_inc.assign((const cd*)t_ri, (const cd*)t_ri + n);
}
// planes (indices into plane_eid) that contain the node: fbem_node_symplanes_connectivity (lib/fbem/src/data_structures.f90:1116-1153)
static int node_planes(const Model* m, int sn, int* planes) {
  int n = 0;
  for (int i = 0; i < m->n_planes; i++) if (fabs(m->node_x[3 * sn + m->plane_eid[i] - 1]) <= m->geometric_tolerance) planes[n++] = i;
  return n;
}
static inline const Element& image_of(const Model* m, int e, int ks) { return ks == 0 ? m->elem[e] : m->img[ks - 1][e]; }
// normals / tangents of the fan of elements around node sn as seen from the region (reverse: the boundary is reversed in it), completed with the
// mirror images when the node lies in one or two symmetry planes (build_lse_mechanics_bem_harela.f90:430-555); returns the fan size, -1 if unsupported
static int node_fan(const Model* m, int sn, bool reverse, std::vector<double>& ns, std::vector<double>& ts) {
  int b0 = m->n2e_ptr[sn], ne = m->n2e_ptr[sn + 1] - b0;
  int planes[3]; const int npl = node_planes(m, sn, planes);
  if (npl > 2) return -1;
  const int fan = ne << npl;
  ns.assign(3 * fan, 0.0); ts.assign(3 * fan, 0.0); std::vector<double> tr(3 * ne);
  for (int k = 0; k < ne; k++) {
    const Element& ee = m->elem[m->n2e_elem[b0 + k]]; double n[3], tbp[3], tbm[3];
    node_normal_tangents(ee.et, ee.x, m->n2e_kn[b0 + k], n, tbp, tbm);
    for (int cc = 0; cc < 3; cc++) { ns[3 * k + cc] = reverse ? -n[cc] : n[cc]; ts[3 * k + cc] = reverse ? tbm[cc] : tbp[cc]; tr[3 * k + cc] = reverse ? tbp[cc] : tbm[cc]; }
  }
  if (npl >= 1) {
    const double* m1 = m->plane_m[planes[0]]; const double* m2 = (npl == 2) ? m->plane_m[planes[1]] : nullptr;
    for (int k = 0; k < ne; k++) for (int cc = 0; cc < 3; cc++) {
      ns[3 * (k + ne) + cc] = m1[cc] * ns[3 * k + cc]; ts[3 * (k + ne) + cc] = m1[cc] * tr[3 * k + cc];
      if (npl == 2) {
        ns[3 * (k + 2 * ne) + cc] = m1[cc] * m2[cc] * ns[3 * k + cc]; ts[3 * (k + 2 * ne) + cc] = m1[cc] * m2[cc] * ts[3 * k + cc];
        ns[3 * (k + 3 * ne) + cc] = m2[cc] * ns[3 * k + cc]; ts[3 * (k + 3 * ne) + cc] = m2[cc] * tr[3 * k + cc];
      }
    }
  }
  return fan;
}
static inline void apply_symconf(const Model* m, int ks, int nn, cd* h, cd* g) {   // build_lse_mechanics_bem_harela.f90:1203-1206
  if (ks == 0) return;
  for (int kn = 0; kn < nn; kn++) for (int il = 0; il < 3; il++) for (int ik = 0; ik < 3; ik++) { h[(kn * 3 + il) * 3 + ik] *= m->conf_t[ks][ik]; g[(kn * 3 + il) * 3 + ik] *= m->conf_t[ks][ik]; }
}

void orc_element_data(void* h, int e, double* cl, int* gln_far, double* bc, double* br) {
  Model* m = (Model*)h; *cl = m->elem[e].cl; *gln_far = m->elem[e].gln_far; for (int c = 0; c < 3; c++) bc[c] = m->elem[e].bc[c]; *br = m->elem[e].br;
}

// scatter of one (collocation node, element) block: assemble_bem_harela_equation.f90:78-113 (coupling be / class ordinary)
static void scatter(const Model* m, int e, int sn_col, const cd* hp, const cd* gp, const cd* cvalue, cd* A, cd* b) {
  int nn = m->elem[e].nn; long long nd = m->n_dof;
  for (int il = 0; il < 3; il++) {
    long long row = m->row[3 * sn_col + il];
    for (int ik = 0; ik < 3; ik++)
      for (int kn = 0; kn < nn; kn++) {
        int sn = m->enode[m->eptr[e] + kn];
        cd hh = hp[(kn * 3 + il) * 3 + ik], gg = gp[(kn * 3 + il) * 3 + ik];
        switch (m->ctype[3 * sn + ik]) {
          case 0: { long long col = m->col_t[3 * sn + ik]; A[row + nd * col] = A[row + nd * col] - gg; b[row] = b[row] - hh * cvalue[3 * sn + ik]; break; }
          case 1: { long long col = m->col_u[3 * sn + ik]; A[row + nd * col] = A[row + nd * col] + hh; b[row] = b[row] + gg * cvalue[3 * sn + ik]; break; }
          case 2: case 3: {   // u_k unknown, t_k unknown (local-axes conditions; their rows are the host's): assemble_bem_harela_equation.f90:107-112
            long long col = m->col_u[3 * sn + ik]; A[row + nd * col] = A[row + nd * col] + hh;
            col = m->col_t[3 * sn + ik]; A[row + nd * col] = A[row + nd * col] - gg;
            break; }
          case 10: {   // p known (normal pressure), u_k unknown: assemble_bem_harela_equation.f90:97-106
            long long col = m->col_u[3 * sn + ik]; A[row + nd * col] = A[row + nd * col] + hh;
            if (!m->elem[e].reverse) b[row] = b[row] + gg * cvalue[3 * sn + ik] * m->n_fn[3 * sn + ik];
            else b[row] = b[row] - gg * cvalue[3 * sn + ik] * m->n_fn[3 * sn + ik];
            break; }
        }
        // incident wave field: assemble_bem_harela_equation.f90:651-666 (ordinary boundary)
        if (!m->u_inc.empty()) b[row] = b[row] + hh * m->u_inc[(size_t)(m->eptr[e] + kn) * 3 + ik] - gg * m->t_inc[(size_t)(m->eptr[e] + kn) * 3 + ik];
      }
  }
}

// One frequency: accumulates (+=) this region's BIE rows into A (col-major n_dof x n_dof) and b.
static int assemble_impl(Model* m, const Params& p, cd nu, const cd* cvalue, cd* A, cd* b, int nthreads, long long* stats_out);
int orc_assemble(void* h, double omega, const double* lambda_ri, const double* mu_ri, double rho, const double* nu_ri,
                 const double* cvalue_ri, double* A_ri, double* b_ri, int nthreads, long long* stats_out /*44*/) {
  Params p; calculate_parameters(cd(lambda_ri[0], lambda_ri[1]), cd(mu_ri[0], mu_ri[1]), rho, omega, p);
  return assemble_impl((Model*)h, p, cd(nu_ri[0], nu_ri[1]), (const cd*)cvalue_ri, (cd*)A_ri, (cd*)b_ri, nthreads, stats_out);
}
// Static elasticity: build_lse_mechanics_bem_staela (src/build_lse_mechanics_bem_staela.f90; same element / collocation loops,
// free-term pass :273-520 with fbem_bem_staela3d_sbie_freeterm, scatter assemble_bem_staela_equation.f90) with the Kelvin
// kernels of lib/fbem/src/bem_staela3d.f90:524-1381.  Real in, real out (A col-major n_dof x n_dof, b); the traversal is the
// harmonic one (the reference's static and harmonic integrators differ only in the point formulas), evaluated in real
// arithmetic and carried in complex containers whose imaginary parts stay exactly zero.
int orc_assemble_static(void* h, double mu, double nu, const double* cvalue_r, double* A_r, double* b_r, int nthreads, long long* stats_out) {
  Model* m = (Model*)h; const long long n = m->n_dof;
  Params p; calculate_parameters_static(mu, nu, p);
  std::vector<cd> A((size_t)n * n, cd(0.0, 0.0)), b((size_t)n, cd(0.0, 0.0)), cv((size_t)3 * m->n_node);
  for (size_t i = 0; i < cv.size(); i++) cv[i] = cvalue_r[i];
  int err = assemble_impl(m, p, cd(nu, 0.0), cv.data(), A.data(), b.data(), nthreads, stats_out);
  for (size_t i = 0; i < A.size(); i++) { if (A[i].imag() != 0.0) err = 7; A_r[i] += A[i].real(); }
  for (size_t i = 0; i < b.size(); i++) { if (b[i].imag() != 0.0) err = 7; b_r[i] += b[i].real(); }
  return err;
}
// u*, t* of Kelvin (fbem_bem_staela3d_sbie_u / _t, bem_staela3d.f90:408-461), [l][k]
void orc_fundamental_solutions_static(const double* x, const double* n, const double* x_i, double mu, double nu, double* u, double* t) {
  Params p; calculate_parameters_static(mu, nu, p);
  double one = 1.0; cd h[9], g[9]; for (int i = 0; i < 9; i++) { h[i] = 0; g[i] = 0; }
  add_exterior_point(p, x, n, x_i, 1, &one, &one, h, g);
  for (int i = 0; i < 9; i++) { u[i] = (p.cte_u * g[i]).real(); t[i] = (p.cte_t * h[i]).real(); }
}
int orc_pair_static(void* h, int e, const double* x_i, double mu, double nu, double* h_r, double* g_r) {
  Model* m = (Model*)h; Params p; calculate_parameters_static(mu, nu, p);
  Stats st; memset(&st, 0, sizeof(st)); cd hh[81], gg[81];
  int mode = sbie_auto(m->elem[e], x_i, p, m->qsp, m->ns_max, hh, gg, st);
  for (int i = 0; i < 9 * m->elem[e].nn; i++) { h_r[i] = hh[i].real(); g_r[i] = gg[i].real(); }
  return mode;
}
static int assemble_impl(Model* m, const Params& p, cd nu, const cd* cvalue, cd* A, cd* b, int nthreads, long long* stats_out) {
  Stats total; memset(&total, 0, sizeof(total));
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
  {
    Stats st; memset(&st, 0, sizeof(st));
#pragma omp for schedule(dynamic)
    for (int e = 0; e < m->n_elem; e++) {
      cd hh[81], gg[81];
      for (int ks = 0; ks < m->n_sym; ks++) {
        const Element& el = image_of(m, e, ks);
        for (int c = 0; c < m->n_colloc; c++) {
          sbie_auto(el, &m->cx[3 * c], p, m->qsp, m->ns_max, hh, gg, st);
          apply_symconf(m, ks, el.nn, hh, gg);
#pragma omp critical
          scatter(m, e, m->cnode[c], hh, gg, cvalue, A, b);
        }
      }
    }
#pragma omp critical
    stats_add(total, st);
  }
  // free terms (serial): build_lse_mechanics_bem_harela.f90:273-747
  int err = 0;
  for (int c = 0; c < m->n_colloc; c++) {
    if (m->celem[c] < 0) continue;   // a point off the boundary (interior point of the region): no free term
    int e = m->celem[c], kn = m->ckn[c], sn = m->cnode[c]; const Element& el = m->elem[e];
    cd hp[81], gp[81]; for (int i = 0; i < 9 * el.nn; i++) { hp[i] = 0.0; gp[i] = 0.0; }
    if (kn >= 0 && m->cxi[2 * c] == -9.0) {  // marker: nodal SBIE (xi not used)
    }
    bool mca = (m->cxi[2 * c] != -9.0);
    if (!mca) {
      double xi_i[2]; xi_at_node(el.et, kn, xi_i);
      cd cplus[3][3];
      if (check_xi1xi2_edge(el.et, xi_i)) {
        std::vector<double> ns, ts;
        const int fan = node_fan(m, sn, el.reverse, ns, ts);
        if (fan < 0) { err = 1; continue; }   // the reference builds fans for nodes in one or two planes only (:500-555)
        if (sbie_freeterm(fan, ns.data(), ts.data(), m->geometric_tolerance, nu, cplus)) err = 1;
      } else {
        for (int a = 0; a < 3; a++) for (int bb = 0; bb < 3; bb++) cplus[a][bb] = (a == bb) ? 0.5 : 0.0;
      }
      for (int il = 0; il < 3; il++) for (int ik = 0; ik < 3; ik++) hp[(kn * 3 + il) * 3 + ik] = hp[(kn * 3 + il) * 3 + ik] + cplus[il][ik];
    } else {
      double phi[9]; phi2d<double>(el.et, &m->cxi[2 * c], phi);
      for (int il = 0; il < 3; il++) for (int j = 0; j < el.nn; j++) hp[(j * 3 + il) * 3 + il] = hp[(j * 3 + il) * 3 + il] + 0.5 * phi[j];
    }
    scatter(m, e, sn, hp, gp, cvalue, A, b);
  }
  m->last = total;
  if (stats_out) {
    for (int i = 0; i < 33; i++) stats_out[i] = total.pairs_regular[i];
    stats_out[33] = total.pts_regular; stats_out[34] = total.pairs_adaptive; stats_out[35] = total.leaves; stats_out[36] = total.pts_adaptive;
    stats_out[37] = total.pairs_singular; stats_out[38] = total.pts_singular; stats_out[39] = total.li_points;
  }
  return err;
}

// -------------------------------------------------------------------------------------------------------------------------
// Inviscid fluid (acoustic) BE region: build_lse_mechanics_bem_harpot (src/build_lse_mechanics_bem_harpot.f90: element loop
// :211-217, free-term pass :243-660 with fbem_bem_pot3d_sbie_freeterm :533, collocation loop :722-1133 with
// fbem_bem_harpot3d_sbie_auto :923) and the scatter of assemble_bem_harpot_equation.f90:78-96 (be boundary, ordinary class,
// ctype 0: p known / Un unknown, ctype 1: Un known / p unknown).  One equation and one unknown per node:
// row[n_node], col_p[n_node] = node%col(1,1), col_q[n_node] = node%col(2,1), ctype[n_node] = node%ctype(1,1).
// -------------------------------------------------------------------------------------------------------------------------
void* orc_setup_pot(int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr, const int* elem_node,
                    const unsigned char* elem_reversed, int n_colloc, const double* colloc_x, const int* colloc_node,
                    const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                    const int* row, const int* col_p, const int* col_q, const int* ctype, int n_dof,
                    double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln, double geometric_tolerance) {
  std::vector<int> r3(3 * n_node, -1), cu3(3 * n_node, -1), ct3(3 * n_node, -1), ty3(3 * n_node, 0);
  Model* m = (Model*)orc_setup(n_node, node_x, n_elem, etype, elem_ptr, elem_node, elem_reversed, n_colloc, colloc_x, colloc_node, colloc_elem, colloc_kn,
                               colloc_xi, r3.data(), cu3.data(), ct3.data(), ty3.data(), n_dof, qsi_relative_error, qsi_ns_max, n_precalsets, precalset_gln,
                               geometric_tolerance);
  m->nd = 1;
  m->row.assign(row, row + n_node); m->col_u.assign(col_p, col_p + n_node); m->col_t.assign(col_q, col_q + n_node); m->ctype.assign(ctype, ctype + n_node);
  return m;
}
static void scatter_pot(const Model* m, int e, int sn_col, const cd* hp, const cd* gp, const cd* cvalue, cd* A, cd* b) {
  int nn = m->elem[e].nn; long long nd = m->n_dof;
  long long row = m->row[sn_col];
  for (int kn = 0; kn < nn; kn++) {
    int sn = m->enode[m->eptr[e] + kn];
    switch (m->ctype[sn]) {
      case 0: { long long col = m->col_t[sn]; A[row + nd * col] = A[row + nd * col] - gp[kn]; b[row] = b[row] - hp[kn] * cvalue[sn]; break; }
      case 1: { long long col = m->col_u[sn]; A[row + nd * col] = A[row + nd * col] + hp[kn]; b[row] = b[row] + gp[kn] * cvalue[sn]; break; }
    }
    // incident wave field: assemble_bem_harpot_equation.f90:471-481 (ordinary boundary)
    if (!m->u_inc.empty()) b[row] = b[row] + hp[kn] * m->u_inc[(size_t)(m->eptr[e] + kn)] - gp[kn] * m->t_inc[(size_t)(m->eptr[e] + kn)];
  }
}
int orc_assemble_pot(void* hd, double omega, double rho, const double* c_ri, const double* cvalue_ri, double* A_ri, double* b_ri, int nthreads,
                     long long* stats_out /*44*/) {
  Model* m = (Model*)hd; if (m->nd != 1) return 9;
  Params p; calculate_parameters_pot(rho, cd(c_ri[0], c_ri[1]), omega, p);
  const cd* cvalue = (const cd*)cvalue_ri; cd* A = (cd*)A_ri; cd* b = (cd*)b_ri;
  Stats total; memset(&total, 0, sizeof(total));
  const double d1J = rho * (omega * omega);   // rho*omega**2
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
  {
    Stats st; memset(&st, 0, sizeof(st));
#pragma omp for schedule(dynamic)
    for (int e = 0; e < m->n_elem; e++)
    for (int ks = 0; ks < m->n_sym; ks++) {   // symmetry images: build_lse_mechanics_bem_harpot.f90:790-800, h, g times symconf_s (:947-948)
      const Element& el = image_of(m, e, ks);
      cd hh[81], gg[81];
      for (int c = 0; c < m->n_colloc; c++) {
        sbie_auto(el, &m->cx[3 * c], p, m->qsp, m->ns_max, hh, gg, st);
        // the flux variable is the normal displacement Un = 1/(rho omega^2) dp/dn: gp=gp*d1J (build_lse_mechanics_bem_harpot.f90:751,1104)
        for (int j = 0; j < el.nn; j++) { gg[j] = gg[j] * d1J * m->conf_s[ks]; hh[j] = hh[j] * m->conf_s[ks]; }
#pragma omp critical
        scatter_pot(m, e, m->cnode[c], hh, gg, cvalue, A, b);
      }
    }
#pragma omp critical
    stats_add(total, st);
  }
  int err = 0;
  for (int c = 0; c < m->n_colloc; c++) {
    if (m->celem[c] < 0) continue;   // a point off the boundary (interior point of the region): no free term
    int e = m->celem[c], kn = m->ckn[c], sn = m->cnode[c]; const Element& el = m->elem[e];
    cd hp[9], gp[9]; for (int i = 0; i < el.nn; i++) { hp[i] = 0.0; gp[i] = 0.0; }
    bool mca = (m->cxi[2 * c] != -9.0);
    if (!mca) {
      double xi_i[2]; xi_at_node(el.et, kn, xi_i);
      double c_plus = 0.5;
      if (check_xi1xi2_edge(el.et, xi_i)) {
        std::vector<double> ns, ts;
        const int ne = node_fan(m, sn, el.reverse, ns, ts);
        if (ne < 0) { err = 1; continue; }
        cd dummy[3][3];
        if (sbie_freeterm(ne, ns.data(), ts.data(), m->geometric_tolerance, cd(0.0, 0.0), dummy, &c_plus)) err = 1;
      }
      hp[kn] = hp[kn] + c_plus;
    } else {
      double phi[9]; phi2d<double>(el.et, &m->cxi[2 * c], phi);
      for (int j = 0; j < el.nn; j++) hp[j] = hp[j] + 0.5 * phi[j];
    }
    scatter_pot(m, e, sn, hp, gp, cvalue, A, b);
  }
  m->last = total;
  if (stats_out) {
    for (int i = 0; i < 33; i++) stats_out[i] = total.pairs_regular[i];
    stats_out[33] = total.pts_regular; stats_out[34] = total.pairs_adaptive; stats_out[35] = total.leaves; stats_out[36] = total.pts_adaptive;
    stats_out[37] = total.pairs_singular; stats_out[38] = total.pts_singular; stats_out[39] = total.li_points;
  }
  return err;
}
// h, g (n nodes each) of one (collocation point, element) pair through fbem_bem_harpot3d_sbie_auto; returns the mode
int orc_pair_pot(void* hd, int e, const double* x_i, double omega, double rho, const double* c_ri, double* h_ri, double* g_ri) {
  Model* m = (Model*)hd; Params p; calculate_parameters_pot(rho, cd(c_ri[0], c_ri[1]), omega, p);
  Stats st; memset(&st, 0, sizeof(st)); cd hh[81], gg[81];
  const int ks = e / m->n_elem; const Element& el = image_of(m, e % m->n_elem, ks);   // e >= n_elem: a symmetry image, h and g times symconf_s
  int mode = sbie_auto(el, x_i, p, m->qsp, m->ns_max, hh, gg, st);
  for (int j = 0; j < el.nn; j++) { hh[j] *= m->conf_s[ks]; gg[j] *= m->conf_s[ks]; }
  memcpy(h_ri, hh, sizeof(cd) * el.nn); memcpy(g_ri, gg, sizeof(cd) * el.nn);
  return mode;
}
// p*, q* (fbem_bem_harpot3d_sbie_p / _q, bem_harpot3d.f90:229-273)
void orc_fundamental_solutions_pot(const double* x, const double* n, const double* x_i, double omega, double rho, const double* c_ri, double* p_ri, double* q_ri) {
  Params p; calculate_parameters_pot(rho, cd(c_ri[0], c_ri[1]), omega, p);
  double one = 1.0; cd h[9], g[9]; for (int i = 0; i < 9; i++) { h[i] = 0; g[i] = 0; }
  add_exterior_point(p, x, n, x_i, 1, &one, &one, h, g);
  cd po = p.cte_u * g[0], qo = p.cte_t * h[0];
  p_ri[0] = po.real(); p_ri[1] = po.imag(); q_ri[0] = qo.real(); q_ri[1] = qo.imag();
}
// -------------------------------------------------------------------------------------------------------------------------
// Biot poroelastic BE region: build_lse_mechanics_bem_harpor (src/build_lse_mechanics_bem_harpor.f90: element loop, free-term pass :295-700
// with c(0,0) = J c_pot, c(1:3,1:3) = Mantic's matrix of the drained skeleton, MCA points J phi/2 and phi/2; collocation loop with
// fbem_bem_harpor3d_sbie_auto) and the scatter of assemble_bem_harpor_equation.f90:78-110, :140-170 for an ordinary `be` boundary with open-pore
// conditions: fluid phase ctype(0) 0: tau known / Un unknown, 1: Un known / tau unknown; skeleton ctype(k) 0: u_k known, 1: t_k known.
// Four equations and four unknowns per node: row[4*n_node], col_p (tau, u1, u2, u3), col_s (Un, t1, t2, t3), ctype[4*n_node].
// props = lambda(2), mu(2), rho1, rho2, rhoa, R(2), Q(2), b.
// -------------------------------------------------------------------------------------------------------------------------
static void por_params_from(const double* props, double omega, PorParams& P) {
  calculate_parameters_por(cd(props[0], props[1]), cd(props[2], props[3]), props[4], props[5], props[6], cd(props[7], props[8]), cd(props[9], props[10]), props[11], omega, P);
}
void* orc_setup_por(int n_node, const double* node_x, int n_elem, const int* etype, const int* elem_ptr, const int* elem_node,
                    const unsigned char* elem_reversed, int n_colloc, const double* colloc_x, const int* colloc_node,
                    const int* colloc_elem, const int* colloc_kn, const double* colloc_xi,
                    const int* row, const int* col_p, const int* col_s, const int* ctype, int n_dof,
                    double qsi_relative_error, int qsi_ns_max, int n_precalsets, const int* precalset_gln, double geometric_tolerance) {
  std::vector<int> r3(3 * n_node, -1), ty3(3 * n_node, 0);
  Model* m = (Model*)orc_setup(n_node, node_x, n_elem, etype, elem_ptr, elem_node, elem_reversed, n_colloc, colloc_x, colloc_node, colloc_elem, colloc_kn,
                               colloc_xi, r3.data(), r3.data(), r3.data(), ty3.data(), n_dof, qsi_relative_error, qsi_ns_max, n_precalsets, precalset_gln,
                               geometric_tolerance);
  m->nd = 4;
  m->row.assign(row, row + 4 * n_node); m->col_u.assign(col_p, col_p + 4 * n_node); m->col_t.assign(col_s, col_s + 4 * n_node); m->ctype.assign(ctype, ctype + 4 * n_node);
  return m;
}
static void scatter_por(const Model* m, int e, int sn_col, const cd* hp, const cd* gp, const cd* cvalue, cd* A, cd* b) {
  int nn = m->elem[e].nn; long long nd = m->n_dof;
  for (int il = 0; il < 4; il++) {
    long long row = m->row[4 * sn_col + il];
    for (int ik = 0; ik < 4; ik++)
      for (int kn = 0; kn < nn; kn++) {
        int sn = m->enode[m->eptr[e] + kn];
        cd hh = hp[(kn * 4 + il) * 4 + ik], gg = gp[(kn * 4 + il) * 4 + ik];
        switch (m->ctype[4 * sn + ik]) {
          case 0: { long long col = m->col_t[4 * sn + ik]; A[row + nd * col] = A[row + nd * col] - gg; b[row] = b[row] - hh * cvalue[4 * sn + ik]; break; }
          case 1: { long long col = m->col_u[4 * sn + ik]; A[row + nd * col] = A[row + nd * col] + hh; b[row] = b[row] + gg * cvalue[4 * sn + ik]; break; }
        }
        // incident wave field: assemble_bem_harpor_equation.f90:1277-1289 (ordinary boundary), (tau | u_k)_inc and (Un | t_k)_inc per element node
        if (!m->u_inc.empty()) b[row] = b[row] + hh * m->u_inc[(size_t)(m->eptr[e] + kn) * 4 + ik] - gg * m->t_inc[(size_t)(m->eptr[e] + kn) * 4 + ik];
      }
  }
}
int orc_assemble_por(void* hd, double omega, const double* props, const double* cvalue_ri, double* A_ri, double* b_ri, int nthreads, long long* stats_out /*44*/) {
  Model* m = (Model*)hd; if (m->nd != 4) return 9;
  PorParams P; por_params_from(props, omega, P);
  Params p; p.por = &P;
  const cd* cvalue = (const cd*)cvalue_ri; cd* A = (cd*)A_ri; cd* b = (cd*)b_ri;
  Stats total; memset(&total, 0, sizeof(total));
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
  {
    Stats st; memset(&st, 0, sizeof(st));
#pragma omp for schedule(dynamic)
    for (int e = 0; e < m->n_elem; e++)
    for (int ks = 0; ks < m->n_sym; ks++) {   // symmetry images: build_lse_mechanics_bem_harpor.f90:855-865
      const Element& el = image_of(m, e, ks);
      cd hh[144], gg[144];
      for (int c = 0; c < m->n_colloc; c++) {
        sbie_auto(el, &m->cx[3 * c], p, m->qsp, m->ns_max, hh, gg, st);
        if (ks > 0)   // h(:,:,0), g(:,:,0) times symconf_s, h(:,:,ik), g(:,:,ik) times symconf_t(ik) (:971-975)
          for (int kn = 0; kn < el.nn; kn++) for (int il = 0; il < 4; il++) for (int ik = 0; ik < 4; ik++) {
            const double f = (ik == 0) ? m->conf_s[ks] : m->conf_t[ks][ik - 1];
            hh[(kn * 4 + il) * 4 + ik] *= f; gg[(kn * 4 + il) * 4 + ik] *= f;
          }
#pragma omp critical
        scatter_por(m, e, m->cnode[c], hh, gg, cvalue, A, b);
      }
    }
#pragma omp critical
    stats_add(total, st);
  }
  int err = 0;
  for (int c = 0; c < m->n_colloc; c++) {
    if (m->celem[c] < 0) continue;
    int e = m->celem[c], kn = m->ckn[c], sn = m->cnode[c]; const Element& el = m->elem[e];
    cd hp[144], gp[144]; for (int i = 0; i < 16 * el.nn; i++) { hp[i] = 0.0; gp[i] = 0.0; }
    bool mca = (m->cxi[2 * c] != -9.0);
    if (!mca) {
      double xi_i[2]; xi_at_node(el.et, kn, xi_i);
      double cpot = 0.5; cd cela[3][3];
      for (int a = 0; a < 3; a++) for (int bb = 0; bb < 3; bb++) cela[a][bb] = (a == bb) ? 0.5 : 0.0;
      if (check_xi1xi2_edge(el.et, xi_i)) {
        std::vector<double> ns, ts;
        const int ne = node_fan(m, sn, el.reverse, ns, ts);
        if (ne < 0) { err = 1; continue; }
        if (sbie_freeterm(ne, ns.data(), ts.data(), m->geometric_tolerance, P.nu, cela, &cpot)) err = 1;
      }
      hp[(kn * 4 + 0) * 4 + 0] += P.J * cpot;
      for (int il = 0; il < 3; il++) for (int ik = 0; ik < 3; ik++) hp[(kn * 4 + il + 1) * 4 + ik + 1] += cela[il][ik];
    } else {
      double phi[9]; phi2d<double>(el.et, &m->cxi[2 * c], phi);
      for (int j = 0; j < el.nn; j++) {
        hp[(j * 4 + 0) * 4 + 0] += P.J * 0.5 * phi[j];
        for (int il = 1; il < 4; il++) hp[(j * 4 + il) * 4 + il] += 0.5 * phi[j];
      }
    }
    scatter_por(m, e, sn, hp, gp, cvalue, A, b);
  }
  m->last = total;
  if (stats_out) {
    for (int i = 0; i < 33; i++) stats_out[i] = total.pairs_regular[i];
    stats_out[33] = total.pts_regular; stats_out[34] = total.pairs_adaptive; stats_out[35] = total.leaves; stats_out[36] = total.pts_adaptive;
    stats_out[37] = total.pairs_singular; stats_out[38] = total.pts_singular; stats_out[39] = total.li_points;
  }
  return err;
}
// h, g (n, 4, 4) of one (collocation point, element) pair through fbem_bem_harpor3d_sbie_auto; returns the mode
int orc_pair_por(void* hd, int e, const double* x_i, double omega, const double* props, double* h_ri, double* g_ri) {
  Model* m = (Model*)hd; PorParams P; por_params_from(props, omega, P); Params p; p.por = &P;
  Stats st; memset(&st, 0, sizeof(st)); cd hh[144], gg[144];
  const int ks = e / m->n_elem; const Element& el = image_of(m, e % m->n_elem, ks);   // e >= n_elem: a symmetry image (dof 0 times symconf_s, dofs 1..3 times symconf_t)
  int mode = sbie_auto(el, x_i, p, m->qsp, m->ns_max, hh, gg, st);
  if (ks > 0)
    for (int kn = 0; kn < el.nn; kn++) for (int il = 0; il < 4; il++) for (int ik = 0; ik < 4; ik++) {
      const double f = (ik == 0) ? m->conf_s[ks] : m->conf_t[ks][ik - 1];
      hh[(kn * 4 + il) * 4 + ik] *= f; gg[(kn * 4 + il) * 4 + ik] *= f;
    }
  memcpy(h_ri, hh, sizeof(cd) * 16 * el.nn); memcpy(g_ri, gg, sizeof(cd) * 16 * el.nn);
  return mode;
}
// u*, t* (4 x 4, [l][k]) and the wavenumbers k1, k2, k3, Z, J of the poroelastic fundamental solution
void orc_fundamental_solutions_por(const double* x, const double* n, const double* x_i, double omega, const double* props, double* u_ri, double* t_ri, double* k_ri /*5 complex*/) {
  PorParams P; por_params_from(props, omega, P); Params p; p.por = &P;
  double one = 1.0; cd h[16], g[16]; for (int i = 0; i < 16; i++) { h[i] = 0; g[i] = 0; }
  add_exterior_point(p, x, n, x_i, 1, &one, &one, h, g);
  finish_pair(p, false, nullptr, 1, h, g);
  memcpy(u_ri, g, sizeof(g)); memcpy(t_ri, h, sizeof(h));
  cd kk[5] = {P.k1, P.k2, P.k3, P.Z, P.J}; memcpy(k_ri, kk, sizeof(kk));
}
// pieces of the free-term pass for the multi-region driver (oracle/multiregion.py): unit normal and the two element-boundary tangents at an
// element node, and the scalar free term of fbem_bem_pot3d_sbie_freeterm
void orc_node_normal_tangents(int et, const double* xn, int node, double* n, double* tbp, double* tbm) { node_normal_tangents(et, xn, node, n, tbp, tbm); }
int orc_freeterm_pot(int ne, const double* n, const double* t, double tol, double* cp) {
  cd dummy[3][3]; return sbie_freeterm(ne, n, t, tol, cd(0.0, 0.0), dummy, cp);
}
void orc_decomposed_zexp(const double* z_ri, double* E_ri /*5 complex*/) { cd E[5]; decomposed_zexp(cd(z_ri[0], z_ri[1]), E); memcpy(E_ri, E, sizeof(E)); }

// Bounded sample of one frequency's assembly for the CPU baseline: every element against the collocation points
// c = c_offset, c_offset + c_stride, ... with the reference's parallel structure (OpenMP dynamic over integration elements,
// critical scatter).  Each sampled collocation point gets its own three rows in the compact matrix
// A_s (3*n_sample x n_dof, column-major).  Returns the number of sampled collocation points.
static int sample_impl(Model* m, const Params& p, const cd* cvalue, int c_offset, int c_stride, cd* A, cd* b, int nthreads, long long* points_out);
int orc_assemble_colloc_sample(void* h, double omega, const double* lambda_ri, const double* mu_ri, double rho, const double* cvalue_ri,
                               int c_offset, int c_stride, double* A_ri, double* b_ri, int nthreads, long long* points_out) {
  Params p; calculate_parameters(cd(lambda_ri[0], lambda_ri[1]), cd(mu_ri[0], mu_ri[1]), rho, omega, p);
  return sample_impl((Model*)h, p, (const cd*)cvalue_ri, c_offset, c_stride, (cd*)A_ri, (cd*)b_ri, nthreads, points_out);
}
// the same bounded sample with the static (Kelvin) kernels; complex containers, imaginary parts stay zero
int orc_assemble_colloc_sample_static(void* h, double mu, double nu, const double* cvalue_ri, int c_offset, int c_stride, double* A_ri, double* b_ri,
                                      int nthreads, long long* points_out) {
  Params p; calculate_parameters_static(mu, nu, p);
  return sample_impl((Model*)h, p, (const cd*)cvalue_ri, c_offset, c_stride, (cd*)A_ri, (cd*)b_ri, nthreads, points_out);
}
static int sample_impl(Model* m, const Params& p, const cd* cvalue, int c_offset, int c_stride, cd* A, cd* b, int nthreads, long long* points_out) {
  std::vector<int> cs; for (int c = c_offset; c < m->n_colloc; c += c_stride) cs.push_back(c);
  const long long ns = (long long)cs.size(), ld = 3 * ns;
  Stats total; memset(&total, 0, sizeof(total));
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
  {
    Stats st; memset(&st, 0, sizeof(st));
#pragma omp for schedule(dynamic)
    for (int e = 0; e < m->n_elem; e++)
    for (int ks = 0; ks < m->n_sym; ks++) {
      const Element& el = image_of(m, e, ks);
      cd hh[81], gg[81];
      for (long long q = 0; q < ns; q++) {
        sbie_auto(el, &m->cx[3 * cs[q]], p, m->qsp, m->ns_max, hh, gg, st);
        apply_symconf(m, ks, el.nn, hh, gg);
#pragma omp critical
        for (int il = 0; il < 3; il++) {
          long long row = 3 * q + il;
          for (int ik = 0; ik < 3; ik++)
            for (int kn = 0; kn < el.nn; kn++) {
              int sn = m->enode[m->eptr[e] + kn];
              cd hv = hh[(kn * 3 + il) * 3 + ik], gv = gg[(kn * 3 + il) * 3 + ik];
              if (m->ctype[3 * sn + ik] == 0) { long long col = m->col_t[3 * sn + ik]; A[row + ld * col] = A[row + ld * col] - gv; b[row] = b[row] - hv * cvalue[3 * sn + ik]; }
              else { long long col = m->col_u[3 * sn + ik]; A[row + ld * col] = A[row + ld * col] + hv; b[row] = b[row] + gv * cvalue[3 * sn + ik]; }
            }
        }
      }
    }
#pragma omp critical
    stats_add(total, st);
  }
  if (points_out) *points_out = total.pts_regular + total.pts_adaptive + total.pts_singular;
  return (int)ns;
}

// h,g (n x 3 x 3, [j][l][k] interleaved complex) of one (collocation point, element) pair; returns the integration mode.
int orc_pair(void* h, int e, const double* x_i, double omega, const double* lambda_ri, const double* mu_ri, double rho, double* h_ri, double* g_ri, long long* stats_out) {
  Model* m = (Model*)h; Params p; calculate_parameters(cd(lambda_ri[0], lambda_ri[1]), cd(mu_ri[0], mu_ri[1]), rho, omega, p);
  Stats st; memset(&st, 0, sizeof(st));
  const int ks = e / m->n_elem; const Element& el = image_of(m, e % m->n_elem, ks);   // e >= n_elem: image ks of a model with symmetry planes, signs applied
  int mode = sbie_auto(el, x_i, p, m->qsp, m->ns_max, (cd*)h_ri, (cd*)g_ri, st);
  apply_symconf(m, ks, el.nn, (cd*)h_ri, (cd*)g_ri);
  if (stats_out) { stats_out[0] = st.pts_regular; stats_out[1] = st.leaves; stats_out[2] = st.pts_adaptive; stats_out[3] = st.pts_singular; stats_out[4] = st.li_points; }
  return mode;
}
// m, l (n x 3 x 3) of the hypersingular equation for a point OFF the element with unit normal n_i (fbem_bem_harela3d_hbie_auto)
int orc_pair_hbie(void* h, int e, const double* x_i, const double* n_i, double omega, const double* lambda_ri, const double* mu_ri, double rho, double* m_ri, double* l_ri) {
  Model* md = (Model*)h; Params p; calculate_parameters(cd(lambda_ri[0], lambda_ri[1]), cd(mu_ri[0], mu_ri[1]), rho, omega, p);
  Stats st; memset(&st, 0, sizeof(st));
  const int ks = e / md->n_elem; const Element& el = image_of(md, e % md->n_elem, ks);
  int mode = sbie_auto(el, x_i, p, md->qsp, md->ns_max, (cd*)m_ri, (cd*)l_ri, st, n_i);
  apply_symconf(md, ks, el.nn, (cd*)m_ri, (cd*)l_ri);   // m(:,:,ik), l(:,:,ik) times symconf_t(ik): build_lse_mechanics_bem_harela.f90:1159-1164
  return mode;
}
// d*, s* (fbem_bem_harela3d_hbie_d / _s, bem_harela3d.f90:2472-2568), [l][k] interleaved complex
void orc_fundamental_solutions_hbie(const double* x, const double* n, const double* x_i, const double* n_i, double omega, const double* lambda_ri, const double* mu_ri, double rho,
                                    double* d_ri, double* s_ri) {
  Params p; calculate_parameters(cd(lambda_ri[0], lambda_ri[1]), cd(mu_ri[0], mu_ri[1]), rho, omega, p);
  double one = 1.0; cd mm[9], ll[9]; for (int i = 0; i < 9; i++) { mm[i] = 0; ll[i] = 0; }
  add_exterior_point_hbie(p, x, n, x_i, n_i, 1, &one, &one, mm, ll);
  cd* d = (cd*)d_ri; cd* s = (cd*)s_ri;
  for (int i = 0; i < 9; i++) { d[i] = p.cte_d * ll[i]; s[i] = p.cte_s * mm[i]; }
}
int orc_pair_hbie_static(void* h, int e, const double* x_i, const double* n_i, double mu, double nu, double* m_r, double* l_r) {
  Model* md = (Model*)h; Params p; calculate_parameters_static(mu, nu, p);
  Stats st; memset(&st, 0, sizeof(st)); cd mm[81], ll[81];
  int mode = sbie_auto(md->elem[e], x_i, p, md->qsp, md->ns_max, mm, ll, st, n_i);
  for (int i = 0; i < 9 * md->elem[e].nn; i++) { m_r[i] = mm[i].real(); l_r[i] = ll[i].real(); }
  return mode;
}
// plan only (mode per pair) -- for comparing discrete decisions with the product's planner
// (element index ks * n_elem + r addresses image ks of root element r of a model with symmetry planes)
int orc_pair_mode(void* h, int e, const double* x_i, double* d_out, double* barxi_out) {
  Model* m = (Model*)h; const Element& el = image_of(m, e % m->n_elem, e / m->n_elem);
  double r[3] = {el.bc[0] - x_i[0], el.bc[1] - x_i[1], el.bc[2] - x_i[2]};
  double rmin = sqrt(dot3(r, r)) - el.br, barxi[2], d; int method;
  if (rmin > (4.0 * el.br)) { barxi[0] = 0.0; barxi[1] = 0.0; d = rmin / el.cl; }
  else { nearest_element_point_bem(el.et, el.x, el.cl, x_i, barxi, rmin, d, method); if (d <= 1.e-12) { *d_out = d; barxi_out[0] = barxi[0]; barxi_out[1] = barxi[1]; return 200; } }
  *d_out = d; barxi_out[0] = barxi[0]; barxi_out[1] = barxi[1];
  int gln_near = qs_n_estimation(false, el.et, 5, m->qsp, d, barxi);
  int gln = std::max(el.gln_far, gln_near);
  if (gln <= el.ps_gln_max && gln_near > 0) { for (size_t i = 0; i < el.ps.size(); i++) if (el.ps[i].gln >= gln) return el.ps[i].gln; }
  return 100;
}

// ---- small entry points for unit tests ----
void orc_zexp_decomposed(const double* z_ri, double* E_ri /*7 complex*/) { cd E[7]; zexp_decomposed(cd(z_ri[0], z_ri[1]), E); memcpy(E_ri, E, sizeof(E)); }
// fundamental solutions u*, t* (bem_harela3d.f90:545-622), [l][k] interleaved complex
void orc_fundamental_solutions(const double* x, const double* n, const double* x_i, double omega, const double* lambda_ri, const double* mu_ri, double rho, double* u_ri, double* t_ri) {
  Params p; calculate_parameters(cd(lambda_ri[0], lambda_ri[1]), cd(mu_ri[0], mu_ri[1]), rho, omega, p);
  double one = 1.0; cd h[9], g[9]; for (int i = 0; i < 9; i++) { h[i] = 0; g[i] = 0; }
  add_exterior_point(p, x, n, x_i, 1, &one, &one, h, g);
  for (int i = 0; i < 9; i++) { ((cd*)u_ri)[i] = p.cte_u * g[i]; ((cd*)t_ri)[i] = p.cte_t * h[i]; }
}
int orc_qs_n(int telles, int etype, int f, double re, double d, const double* barxi) { QsParams q; qs_calculate_parameters(re, q); return qs_n_estimation(telles != 0, etype, f, q, d, barxi); }
void orc_qs_params(double re, double* out /*dmin[3], dmax[3] for f=5 standard; then telles*/) {
  QsParams q; qs_calculate_parameters(re, q);
  for (int c = 0; c < 3; c++) { out[c] = q.dmin[c][5]; out[3 + c] = q.dmax[c][5]; out[6 + c] = q.t_dmin[c][5]; out[9 + c] = q.t_dmax[c][5]; }
}
void orc_telles(int dom01, double bar_xi, double bar_r, double* c) { if (dom01) telles01_parameters(bar_xi, bar_r, c); else telles11_parameters(bar_xi, bar_r, c); }
double orc_telles_barr(double d) { return telles_barr_any(d); }
void orc_nearest(int etype, const double* x, const double* x_i, double* barxi, double* rmin, double* d, int* method) {
  double cl = characteristic_length(etype, x, 1.e-9); nearest_element_point_bem(etype, x, cl, x_i, barxi, *rmin, *d, *method);
}
void orc_phi(int etype, const double* xi, double* phi, double* d1, double* d2) { phi2d<double>(etype, xi, phi); dphi2d<double>(etype, xi, d1, d2); }
int orc_freeterm(int ne, const double* n, const double* t, double tol, const double* nu_ri, double* c_ri) {
  cd c[3][3]; int e = sbie_freeterm(ne, n, t, tol, cd(nu_ri[0], nu_ri[1]), c); memcpy(c_ri, c, sizeof(c)); return e;
}
void orc_tables(int family, int n, double* x, double* w) {  // 0 gl11, 1 gl01, 2 gj01, 3 wantri(order n: x=[xi1..,xi2..])
  if (family == 3) { int np = wan_n(n); for (int k = 0; k < np; k++) { x[k] = wan_x1(n, k); x[np + k] = wan_x2(n, k); w[k] = wan_w(n, k); } return; }
  for (int k = 0; k < n; k++) { x[k] = family == 0 ? gl11_x(n, k) : family == 1 ? gl01_x(n, k) : gj01_x(n, k); w[k] = family == 0 ? gl11_w(n, k) : family == 1 ? gl01_w(n, k) : gj01_w(n, k); }
}
int orc_wantri_n(int order) { return wan_n(order); }

}  // extern "C"
