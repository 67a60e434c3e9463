// plan_values.cpp -- the VALUE-producing host geometry of the quadrature plans (product code, once per mesh).
//
// plan_host.cpp keeps the discrete decision procedure of the reference (which rule, subdivide or not, how many angular points): those
// decisions must come out bit-identical, so that file follows the reference's arithmetic operation by operation.  Everything in THIS file
// produces continuous values (points, weights, rays, line integrals, free-term geometry, transformation coefficients) and is written from
// the mathematics, not from the reference's code: the formulas are derived in the comments, in forms that differ from the reference's
// (and from the oracle's restatement of them), so that agreement of the assembled matrices with the oracle to 1e-11 is a test of two
// independent derivations.  Each block names the reference routine whose RESULT it must reproduce.
#include "plan_host.h"
#include "plan_values.h"
#include "../../data/quad_tables.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace mfbh {

static const double PI = 3.14159265358979323846264338328;

// ----------------------------------------------------------------------------------------------------------------------------------
// Shape functions from the element definitions (results of lib/fbem/src/resources_shape_functions/*.rc, continuous elements).
// Node order: triangles (1,0) (0,1) (0,0) | mid-sides (1/2,1/2) (0,1/2) (1/2,0); quadrilaterals (-1,-1) (1,-1) (1,1) (-1,1) | mid-sides
// (0,-1) (1,0) (0,1) (-1,0) | centre.  Triangles: barycentric L = (xi1, xi2, 1 - xi1 - xi2), corner L(2L - 1), mid-side 4 L_a L_b.
// Quadrilaterals: node (s1, s2), s in {-1, 0, 1}: quad9 is the tensor product of the 1-D quadratic Lagrange basis on {-1, 0, 1}; quad8 the
// serendipity family (corner (1 + s1 x)(1 + s2 y)(s1 x + s2 y - 1)/4, mid-side (1 - x^2)(1 + s2 y)/2 or (1 + s1 x)(1 - y^2)/2); quad4 bilinear.
// ----------------------------------------------------------------------------------------------------------------------------------
static const int QSGN[9][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}, {0, -1}, {1, 0}, {0, 1}, {-1, 0}, {0, 0}};
static inline void lag3(int s, double t, double& v, double& d) {   // 1-D quadratic Lagrange polynomial of node s on {-1, 0, 1} and its derivative
  if (s == 0) { v = 1.0 - t * t; d = -2.0 * t; }
  else { v = 0.5 * t * (t + s); d = t + 0.5 * s; }
}
void shape_all(int et, const double* xi, double* phi, double* d1, double* d2) {
  const double x = xi[0], y = xi[1];
  if (et == TRI3 || et == TRI6) {
    const double L[3] = {x, y, 1.0 - x - y};
    static const double dL1[3] = {1.0, 0.0, -1.0}, dL2[3] = {0.0, 1.0, -1.0};
    if (et == TRI3) { for (int k = 0; k < 3; k++) { phi[k] = L[k]; d1[k] = dL1[k]; d2[k] = dL2[k]; } return; }
    static const int MID[3][2] = {{0, 1}, {1, 2}, {2, 0}};
    for (int k = 0; k < 3; k++) { phi[k] = L[k] * (2.0 * L[k] - 1.0); d1[k] = (4.0 * L[k] - 1.0) * dL1[k]; d2[k] = (4.0 * L[k] - 1.0) * dL2[k]; }
    for (int k = 0; k < 3; k++) {
      const int a = MID[k][0], b = MID[k][1];
      phi[3 + k] = 4.0 * L[a] * L[b]; d1[3 + k] = 4.0 * (dL1[a] * L[b] + L[a] * dL1[b]); d2[3 + k] = 4.0 * (dL2[a] * L[b] + L[a] * dL2[b]);
    }
    return;
  }
  const int nn = (et == QUAD4) ? 4 : (et == QUAD8 ? 8 : 9);
  for (int k = 0; k < nn; k++) {
    const int s1 = QSGN[k][0], s2 = QSGN[k][1];
    if (et == QUAD9) {
      double a, da, b, db; lag3(s1, x, a, da); lag3(s2, y, b, db);
      phi[k] = a * b; d1[k] = da * b; d2[k] = a * db;
    } else if (et == QUAD4 || k < 4) {
      const double a = 1.0 + s1 * x, b = 1.0 + s2 * y;
      if (et == QUAD4) { phi[k] = 0.25 * a * b; d1[k] = 0.25 * s1 * b; d2[k] = 0.25 * s2 * a; }
      else { const double c = s1 * x + s2 * y - 1.0; phi[k] = 0.25 * a * b * c; d1[k] = 0.25 * s1 * b * (c + a); d2[k] = 0.25 * s2 * a * (c + b); }
    } else if (s1 == 0) { const double b = 1.0 + s2 * y; phi[k] = 0.5 * (1.0 - x * x) * b; d1[k] = -x * b; d2[k] = 0.5 * s2 * (1.0 - x * x); }
    else { const double a = 1.0 + s1 * x; phi[k] = 0.5 * a * (1.0 - y * y); d1[k] = 0.5 * s1 * (1.0 - y * y); d2[k] = -y * a; }
  }
}
static inline int nn_of(int et) { return et == TRI3 ? 3 : et == TRI6 ? 6 : et == QUAD4 ? 4 : et == QUAD8 ? 8 : et == QUAD9 ? 9 : et; }
void shape_values(int et, const double* xi, double* phi) { double d1[9], d2[9]; shape_all(et, xi, phi, d1, d2); }

struct SurfPoint { double x[3], a1[3], a2[3], N[3], J; };
static void surface_at(int et, const double* xn, const double* xi, double* phi, SurfPoint& p) {
  double d1[9], d2[9]; shape_all(et, xi, phi, d1, d2);
  const int nn = nn_of(et);
  for (int c = 0; c < 3; c++) {
    double x = 0.0, a = 0.0, b = 0.0;
    for (int k = 0; k < nn; k++) { x += phi[k] * xn[3 * k + c]; a += d1[k] * xn[3 * k + c]; b += d2[k] * xn[3 * k + c]; }
    p.x[c] = x; p.a1[c] = a; p.a2[c] = b;
  }
  p.N[0] = p.a1[1] * p.a2[2] - p.a1[2] * p.a2[1]; p.N[1] = p.a1[2] * p.a2[0] - p.a1[0] * p.a2[2]; p.N[2] = p.a1[0] * p.a2[1] - p.a1[1] * p.a2[0];
  p.J = std::sqrt(p.N[0] * p.N[0] + p.N[1] * p.N[1] + p.N[2] * p.N[2]);
}
void element_point(int et, const double* xn, const double* xi, double* x) {
  double phi[9]; SurfPoint p; surface_at(et, xn, xi, phi, p);
  for (int c = 0; c < 3; c++) x[c] = p.x[c];
}

// ----------------------------------------------------------------------------------------------------------------------------------
// Precalculated point set of one element and one rule (result of fbem_bem_element%init_precalculated_datasets, lib/fbem/src/bem_general.f90:450-755):
// per point x, unit normal, phi_j * J * w.  Rules: quadrilaterals gln x gln Gauss-Legendre; triangles Wandzura of order 2 gln - 1 up to gln = 15,
// beyond that Gauss-Legendre x Gauss-Jacobi on the collapsed square (xi1 = (1 - t) s, xi2 = t; the Jacobi weight carries the factor 1 - t).
// ----------------------------------------------------------------------------------------------------------------------------------
int pointset_size(int et, int gln) { return ((et == TRI3 || et == TRI6) && gln <= 15) ? QT_WAN_N[2 * gln - 2] : gln * gln; }
void build_pointset(const Elem& e, int gln, double* out) {
  const bool tri = (e.et == TRI3 || e.et == TRI6);
  const int ngp = pointset_size(e.et, gln), nn = nn_of(e.et);
  for (int q = 0; q < ngp; q++) {
    double xi[2], w;
    if (tri && gln <= 15) { const int o = QT_WAN_OFF[2 * gln - 2] + q; xi[0] = QT_WAN_X1[o]; xi[1] = QT_WAN_X2[o]; w = QT_WAN_W[o]; }
    else {
      const int i = q / gln, j = q % gln;      // point order of the reference: second index fastest
      if (tri) {
        const double s = QT_GL01_X[QT_GL01_OFF[gln - 1] + i], t = QT_GJ01_X[QT_GJ01_OFF[gln - 1] + j];
        xi[0] = (1.0 - t) * s; xi[1] = t; w = QT_GL01_W[QT_GL01_OFF[gln - 1] + i] * QT_GJ01_W[QT_GJ01_OFF[gln - 1] + j];
      } else {
        xi[0] = QT_GL11_X[QT_GL11_OFF[gln - 1] + i]; xi[1] = QT_GL11_X[QT_GL11_OFF[gln - 1] + j];
        w = QT_GL11_W[QT_GL11_OFF[gln - 1] + i] * QT_GL11_W[QT_GL11_OFF[gln - 1] + j];
      }
    }
    double phi[9]; SurfPoint p; surface_at(e.et, e.x, xi, phi, p);
    double* o = out + (size_t)q * (6 + nn);
    const double jw = p.J * w;
    for (int c = 0; c < 3; c++) { o[c] = p.x[c]; o[3 + c] = p.N[c] / p.J; }
    for (int k = 0; k < nn; k++) o[6 + k] = phi[k] * jw;
  }
}

// ----------------------------------------------------------------------------------------------------------------------------------
// Telles' cubic xi(g) = a g^3 + b g^2 + c g + d (results of fbem_telles11/01_calculate_parameters, lib/fbem/src/telles_transformation.f90:130-196).
// Defining conditions: the map fixes the interval ends, its second derivative vanishes at the image gbar of the nearest point xibar and its
// first derivative there is rbar.  On [-1, 1]: b = -3 a gbar, d = -b, c = 1 - a, a = (1 - rbar) / (1 + 3 gbar^2), and gbar is the root of
//   F(g) = g + 2 (1 - rbar) g (1 - g^2) / (1 + 3 g^2) - xibar,            F(-1) <= 0 <= F(1),  F' >= rbar > 0.
// On [0, 1]: d = 0, b = -3 a gbar, c = rbar + 3 a gbar^2, a = (1 - rbar) / (3 gbar^2 - 3 gbar + 1), and
//   F(g) = (1 - rbar) g^3 / (3 g^2 - 3 g + 1) + rbar g - xibar,           F(0) <= 0 <= F(1).
// The reference evaluates Cardano's closed form of the same root; here it is found by a bracketed Newton iteration to the last bit.
// ----------------------------------------------------------------------------------------------------------------------------------
void telles_cubic(bool unit_interval, double xibar, double rbar, double* c) {
  double lo = unit_interval ? 0.0 : -1.0, hi = 1.0;
  auto F = [&](double g, double& dF) {
    if (unit_interval) {
      const double q = 3.0 * g * g - 3.0 * g + 1.0, dq = 6.0 * g - 3.0;
      dF = (1.0 - rbar) * (3.0 * g * g * q - g * g * g * dq) / (q * q) + rbar;
      return (1.0 - rbar) * g * g * g / q + rbar * g - xibar;
    }
    const double q = 1.0 + 3.0 * g * g, u = g * (1.0 - g * g);
    dF = 1.0 + 2.0 * (1.0 - rbar) * ((1.0 - 3.0 * g * g) * q - u * 6.0 * g) / (q * q);
    return g + 2.0 * (1.0 - rbar) * u / q - xibar;
  };
  double g = std::min(std::max(xibar, lo), hi);
  for (int it = 0; it < 200; it++) {
    double dF; const double f = F(g, dF);
    if (f == 0.0) break;
    if (f > 0.0) hi = g; else lo = g;
    double gn = (dF > 0.0) ? g - f / dF : 0.5 * (lo + hi);
    if (!(gn > lo && gn < hi)) gn = 0.5 * (lo + hi);
    if (gn == g || hi - lo <= 0.0) break;
    if (std::fabs(gn - g) <= 2.3e-16 * std::max(1.0, std::fabs(g))) { g = gn; break; }
    g = gn;
  }
  if (unit_interval) {
    const double a = (1.0 - rbar) / (3.0 * g * g - 3.0 * g + 1.0);
    c[0] = a; c[1] = -3.0 * a * g; c[2] = rbar + 3.0 * a * g * g; c[3] = 0.0;
  } else {
    const double a = (1.0 - rbar) / (1.0 + 3.0 * g * g);
    c[0] = a; c[1] = -3.0 * a * g; c[2] = 1.0 - a; c[3] = 3.0 * a * g;
  }
}

// ----------------------------------------------------------------------------------------------------------------------------------
// Rays of the polar integration around a collocation point xi_i ON the element (results of fbem_polar_transformation_setup / _angular,
// lib/fbem/src/polar_transformation.f90:251-544, as used by fbem_bem_harela3d_sbie_int).
// Reference polygon with counter-clockwise vertices V.  For the edge V_a -> V_b: unit direction e, outward normal nu = (e_y, -e_x),
// h = (V_a - xi_i).nu >= 0 the distance of xi_i to the edge line, s = (xi - xi_i).e the abscissa along the edge measured from the foot of
// the perpendicular.  The ray of polar angle theta meets the edge at s = h tan(theta - beta) (beta = direction of nu) after
// rho_max = h / cos(theta - beta).  The reference integrates in theta' with d theta' = rho_max d theta, i.e.
//   theta' = h ln tan((theta - beta)/2 + pi/4) = h asinh(s / h)            (inverse Gudermannian),
// so a quadrature node theta'_q gives directly   s_q = h sinh(theta'_q / h),   rho_max = sqrt(h^2 + s_q^2),   (cos, sin)(theta) = (h nu + s_q e) / rho_max
// without any angle ever being formed.  Every edge is split at the foot of the perpendicular into the two sub-triangles (2k - 1, 2k) of the
// reference's numbering; counts[] (from the decision core: presence and number of angular points of each sub-triangle) selects them.
// ----------------------------------------------------------------------------------------------------------------------------------
static int polygon_of(int et, double V[4][2]) {
  if (et == TRI3 || et == TRI6) { V[0][0] = 1; V[0][1] = 0; V[1][0] = 0; V[1][1] = 1; V[2][0] = 0; V[2][1] = 0; return 3; }
  V[0][0] = -1; V[0][1] = -1; V[1][0] = 1; V[1][1] = -1; V[2][0] = 1; V[2][1] = 1; V[3][0] = -1; V[3][1] = 1; return 4;
}
void polar_rays(int et, const double* xi_i, const int* counts /* [2 * n_edges], 0 = sub-triangle absent */, std::vector<Ray>& rays) {
  double V[4][2]; const int nv = polygon_of(et, V);
  for (int k = 0; k < nv; k++) {
    const double* A = V[k]; const double* B = V[(k + 1) % nv];
    const double ex = B[0] - A[0], ey = B[1] - A[1], L = std::sqrt(ex * ex + ey * ey);
    const double e[2] = {ex / L, ey / L}, nu[2] = {e[1], -e[0]};
    const double h = (A[0] - xi_i[0]) * nu[0] + (A[1] - xi_i[1]) * nu[1];
    const double sa = (A[0] - xi_i[0]) * e[0] + (A[1] - xi_i[1]) * e[1], sb = sa + L;
    for (int half = 0; half < 2; half++) {
      const int ng = counts[2 * k + half];
      if (ng <= 0) continue;
      const double s0 = half ? 0.0 : sa, s1 = half ? sb : 0.0;          // [V_a, foot] then [foot, V_b]
      const double t0 = h * std::asinh(s0 / h), t1 = h * std::asinh(s1 / h), dt = t1 - t0;
      for (int q = 0; q < ng; q++) {
        const double tq = t0 + dt * QT_GL01_X[QT_GL01_OFF[ng - 1] + q];
        const double s = h * std::sinh(tq / h), rho = std::sqrt(h * h + s * s);
        Ray r; r.ct = (h * nu[0] + s * e[0]) / rho; r.st = (h * nu[1] + s * e[1]) / rho; r.rhoij = rho; r.w = dt * QT_GL01_W[QT_GL01_OFF[ng - 1] + q];
        rays.push_back(r);
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------------------------
// Edge line integrals of the singular element integral (result of fbem_bem_staela3d_sbie_int_li, lib/fbem/src/bem_staela3d.f90:2245-2376,
// summed over the edges that do not contain the collocation point): I = sum_edges int t / r ds = int x'(u) / |x(u) - x_i| du (t ds = x' du),
// delivered as hli[l][k] = -eps_lkm I_m.  Straight edges (2 nodes): with a = x_A - x_i, t the unit direction, s_a = a.t, p^2 = |a|^2 - s_a^2,
//   int ds / sqrt(s^2 + p^2) = asinh(s_b / p) - asinh(s_a / p)   in closed form.
// Curved edges (3 nodes): x(u) = M + u (B - A)/2 + u^2 (A + B - 2M)/2 on [-1, 1], integrated with 20-point Gauss-Legendre panels graded
// geometrically away from the nearest point (the reference uses its Telles + subdivision machinery with a 1e-15 target).
// ----------------------------------------------------------------------------------------------------------------------------------
static const int EDGE_T[3][3] = {{0, 1, 3}, {1, 2, 4}, {2, 0, 5}}, EDGE_Q[4][3] = {{0, 1, 4}, {1, 2, 5}, {2, 3, 6}, {3, 0, 7}};
struct CurvedEdge { double M[3], D[3], Q[3], xi[3]; };
static inline double ce_dist2(const CurvedEdge& E, double u) {
  double r2 = 0.0;
  for (int c = 0; c < 3; c++) { const double x = E.M[c] + u * (E.D[c] + u * E.Q[c]) - E.xi[c]; r2 += x * x; }
  return r2;
}
// 20-point Gauss-Legendre on [a, b] of f(u) = x'(u) / |x(u) - x_i|
static void ce_panel(const CurvedEdge& E, double a, double b, double* I) {
  const double m = 0.5 * (a + b), hw = 0.5 * (b - a);
  for (int q = 0; q < 20; q++) {
    const double u = m + hw * QT_GL11_X[QT_GL11_OFF[19] + q], w = hw * QT_GL11_W[QT_GL11_OFF[19] + q];
    const double ir = 1.0 / std::sqrt(ce_dist2(E, u));
    for (int c = 0; c < 3; c++) I[c] += w * (E.D[c] + 2.0 * u * E.Q[c]) * ir;
  }
}
// The integrand is analytic on [-1, 1]; its complex singularities sit at a distance ~ rmin / |x'| from the parameter u0 of the nearest point.
// Panels graded geometrically away from u0 (first panel as wide as that distance, every next one twice as wide) keep the ratio
// panel width / distance to the singularity bounded, where 20 Gauss points reach machine precision.  Bounded work, no error-driven recursion.
static void ce_integrate(const CurvedEdge& E, double* I) {
  double u0 = 0.0, best = ce_dist2(E, 0.0);
  for (int k = 0; k <= 256; k++) { const double u = -1.0 + k / 128.0, d2 = ce_dist2(E, u); if (d2 < best) { best = d2; u0 = u; } }
  double lo = std::max(-1.0, u0 - 1.0 / 128.0), hi = std::min(1.0, u0 + 1.0 / 128.0);
  for (int it = 0; it < 60; it++) {   // golden-section refinement of the minimiser (approximate is enough: it only places the panels)
    const double g = 0.381966011250105, a = lo + g * (hi - lo), b = hi - g * (hi - lo);
    if (ce_dist2(E, a) < ce_dist2(E, b)) hi = b; else lo = a;
  }
  u0 = 0.5 * (lo + hi);
  double sp = 0.0;
  for (int c = 0; c < 3; c++) { const double d = E.D[c] + 2.0 * u0 * E.Q[c]; sp += d * d; }
  const double w0 = std::min(std::max(std::sqrt(ce_dist2(E, u0) / std::max(sp, 1e-300)), 1e-9), 2.0);
  for (int side = -1; side <= 1; side += 2) {
    double a = u0, w = w0;
    for (int k = 0; k < 64; k++) {
      double b = a + side * w;
      const bool last = side > 0 ? b >= 1.0 : b <= -1.0;
      if (last) b = side;
      if (b != a) { if (side > 0) ce_panel(E, a, b, I); else ce_panel(E, b, a, I); }
      if (last) break;
      a = b; w *= 2.0;
    }
  }
}
void edge_integrals(int et, const double* xn, const double* x_i, const bool* edge_on /* [n_edges] */, double* hli /* 9, accumulated */) {
  const bool tri = (et == TRI3 || et == TRI6), curved = (et == TRI6 || et == QUAD8 || et == QUAD9);
  const int ne = tri ? 3 : 4;
  double I[3] = {0, 0, 0};
  for (int k = 0; k < ne; k++) {
    if (!edge_on[k]) continue;
    const int* en = tri ? EDGE_T[k] : EDGE_Q[k];
    const double *A = xn + 3 * en[0], *B = xn + 3 * en[1];
    bool straight = !curved;
    if (curved) {   // a quadratic edge whose middle node sits exactly at the midpoint of the chord is the straight segment
      const double* M = xn + 3 * en[2]; double dev = 0.0, len = 0.0;
      for (int c = 0; c < 3; c++) { dev = std::max(dev, std::fabs(A[c] + B[c] - 2.0 * M[c])); len = std::max(len, std::fabs(B[c] - A[c])); }
      straight = dev <= 1e-13 * len;
    }
    if (straight) {
      double t[3], a[3], L = 0.0, sa = 0.0, a2 = 0.0;
      for (int c = 0; c < 3; c++) { t[c] = B[c] - A[c]; a[c] = A[c] - x_i[c]; L += t[c] * t[c]; a2 += a[c] * a[c]; }
      L = std::sqrt(L);
      for (int c = 0; c < 3; c++) { t[c] /= L; sa += a[c] * t[c]; }
      const double sb = sa + L, p2 = std::max(a2 - sa * sa, 0.0), p = std::sqrt(p2);
      const double v = (p > 1e-14 * L) ? std::asinh(sb / p) - std::asinh(sa / p) : std::log(std::fabs(sb / sa));
      for (int c = 0; c < 3; c++) I[c] += t[c] * v;
    } else {
      CurvedEdge E; const double* M = xn + 3 * en[2];
      for (int c = 0; c < 3; c++) { E.M[c] = M[c]; E.D[c] = 0.5 * (B[c] - A[c]); E.Q[c] = 0.5 * (A[c] + B[c] - 2.0 * M[c]); E.xi[c] = x_i[c]; }
      ce_integrate(E, I);
    }
  }
  // hli[l][k] = -eps_lkm I_m
  hli[1] += -I[2]; hli[2] += I[1]; hli[5] += -I[0];
  hli[3] += I[2];  hli[6] += -I[1]; hli[7] += I[0];
}

// ----------------------------------------------------------------------------------------------------------------------------------
// Unit normal and the unit tangent of the element boundary that LEAVES a node counter-clockwise (seen against the normal), or, for the
// reversed orientation, the other way round (results of fbem_unormal3d / fbem_utangents_at_boundary with the reversion rule of
// src/build_lse_mechanics_bem_harela.f90:428-447).  The boundary of the reference polygon is traversed counter-clockwise; at a node with
// parametric position xi on that boundary the traversal direction is a 2-vector d (for a corner: the edge that starts there, and for the
// reversed orientation minus the edge that ends there), and the tangent is the push-forward a1 d1 + a2 d2, normalised.
// ----------------------------------------------------------------------------------------------------------------------------------
void node_xi(int et, int node, double* xi) {
  static const double T[6][2] = {{1, 0}, {0, 1}, {0, 0}, {0.5, 0.5}, {0, 0.5}, {0.5, 0}};
  if (et == TRI3 || et == TRI6) { xi[0] = T[node][0]; xi[1] = T[node][1]; }
  else { xi[0] = QSGN[node][0]; xi[1] = QSGN[node][1]; }
}
void node_normal_tangent(int et, const double* xn, int node, bool reversed, double* n, double* t) {
  double V[4][2]; const int nv = polygon_of(et, V);
  double xi[2]; node_xi(et, node, xi);
  double d[2] = {0.0, 0.0};
  if (node < nv) {            // corner: leaving edge (node -> node + 1), or minus the arriving edge (node - 1 -> node)
    const int a = reversed ? (node + nv - 1) % nv : node, b = reversed ? node : (node + 1) % nv;
    d[0] = V[b][0] - V[a][0]; d[1] = V[b][1] - V[a][1];
    if (reversed) { d[0] = -d[0]; d[1] = -d[1]; }
  } else if (node < 2 * nv) { // mid-side node of edge (node - nv)
    const int a = node - nv, b = (a + 1) % nv;
    d[0] = V[b][0] - V[a][0]; d[1] = V[b][1] - V[a][1];
    if (reversed) { d[0] = -d[0]; d[1] = -d[1]; }
  }
  double phi[9]; SurfPoint p; surface_at(et, xn, xi, phi, p);
  double tt[3], tn = 0.0;
  for (int c = 0; c < 3; c++) { tt[c] = p.a1[c] * d[0] + p.a2[c] * d[1]; tn += tt[c] * tt[c]; }
  tn = std::sqrt(tn);
  for (int c = 0; c < 3; c++) { n[c] = (reversed ? -1.0 : 1.0) * p.N[c] / p.J; t[c] = tn > 0.0 ? tt[c] / tn : 0.0; }
}

// ----------------------------------------------------------------------------------------------------------------------------------
// Geometry of the free term at an edge / vertex node (result of fbem_bem_harela3d_sbie_freeterm, lib/fbem/src/bem_harela3d.f90:365-542:
// c = cp I - S / (8 pi (1 - nu))).  Derivation.  The free term of the Somigliana identity is the integral of the Kelvin traction kernel over the
// part of a vanishing sphere inside the body:  c_ij = 1 / (8 pi (1 - nu)) int_w [ (1 - 2 nu) delta_ij + 3 e_i e_j ] dw, w the interior solid angle,
// e the unit vector.  With K the interior cone cut by the unit ball, the divergence theorem for the field x_j e_i on K gives
//   delta_ij W / 3 = int_w e_i e_j dw + sum_faces n_i int_face x_j dS,       int_face x_j dS = (1/3) int_arc u_j dphi = (1/3) (n x (u_a - u_b))_j,
// (W = int_w dw; faces are the tangent planes of the m elements at the node, outward normal n, bounded by the unit edge tangents u_a -> u_b
// met counter-clockwise about n), hence
//   c_ij = (W / 4 pi) delta_ij - (1 / (8 pi (1 - nu))) sum_faces n_i (n x (u_a - u_b))_j :   cp = W / 4 pi,   S = that sum (symmetric as a whole).
// W follows from Gauss-Bonnet on the unit sphere: the boundary of w is a geodesic polygon, so W = 2 pi - sum of the exterior angles, and the
// exterior angle at the edge shared by two consecutive faces is the signed angle between their normals about that edge.
// Input: per element its outward unit normal and the unit tangent of the boundary edge that leaves the node counter-clockwise.  The faces are
// chained by following, in the plane of the current face, the smallest counter-clockwise rotation to another element's tangent.
// ----------------------------------------------------------------------------------------------------------------------------------
int mantic_terms(int m, const double* normals, const double* tangents, double tol, double* cp, double* sum_b) {
  if (m < 2) return 1;
  const double ptol = (tol < 1.0e-12 || tol > 1.0e-3) ? 1.0e-6 : tol;
  std::vector<int> order(1, 0); std::vector<char> used(m, 0); used[0] = 1;
  auto N = [&](int f) { return normals + 3 * f; };
  auto U = [&](int f) { return tangents + 3 * f; };
  for (int step = 1; step < m; step++) {
    const int f = order.back();
    const double *n = N(f), *u = U(f);
    const double v[3] = {n[1] * u[2] - n[2] * u[1], n[2] * u[0] - n[0] * u[2], n[0] * u[1] - n[1] * u[0]};   // n x u: a quarter turn counter-clockwise of u in the face
    int best = -1; double best_ang = 0.0;
    for (int g = 0; g < m; g++) {
      if (used[g]) continue;
      const double* w = U(g);
      if (std::fabs(w[0] * n[0] + w[1] * n[1] + w[2] * n[2]) > ptol) continue;       // not in the plane of face f
      double ang = std::atan2(w[0] * v[0] + w[1] * v[1] + w[2] * v[2], w[0] * u[0] + w[1] * u[1] + w[2] * u[2]);
      if (ang < 0.0) ang += 2.0 * PI;
      if (best < 0 || ang < best_ang) { best = g; best_ang = ang; }
    }
    if (best < 0) return 1;
    used[best] = 1; order.push_back(best);
  }
  double W = 2.0 * PI, S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int q = 0; q < m; q++) {
    const int f = order[q], g = order[(q + 1) % m], fp = order[(q + m - 1) % m];
    const double *n = N(f), *ua = U(f), *ub = U(g), *np = N(fp);
    // exterior angle at the edge ua between the previous face and this one: signed angle from np to n about ua, counted positive when the
    // boundary of w turns towards its inside (convex corner): W = 2 pi - sum(exterior) with exterior = -atan2((np x n).ua, np.n)
    const double cx[3] = {np[1] * n[2] - np[2] * n[1], np[2] * n[0] - np[0] * n[2], np[0] * n[1] - np[1] * n[0]};
    W += std::atan2(cx[0] * ua[0] + cx[1] * ua[1] + cx[2] * ua[2], np[0] * n[0] + np[1] * n[1] + np[2] * n[2]);
    const double du[3] = {ua[0] - ub[0], ua[1] - ub[1], ua[2] - ub[2]};
    const double k[3] = {n[1] * du[2] - n[2] * du[1], n[2] * du[0] - n[0] * du[2], n[0] * du[1] - n[1] * du[0]};   // n x (u_a - u_b)
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) S[i][j] += n[i] * k[j];
  }
  *cp = W / (4.0 * PI);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) sum_b[3 * i + j] = 0.5 * (S[i][j] + S[j][i]);
  return 0;
}

}  // namespace mfbh
