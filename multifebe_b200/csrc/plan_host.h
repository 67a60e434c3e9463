// plan_host.h -- host-side (omega-independent) quadrature planning of the B200 assembly path.
//
// Product code (not the oracle).  Runs once per mesh inside mfb_harela3d_setup: per-element data, precalculated
// point sets, rule thresholds for the GPU classifier, and -- for the O(N) near pairs the GPU classifier hands
// back -- the discrete decisions of fbem_bem_harela3d_sbie_auto (lib/fbem/src/bem_harela3d.f90:1474-1538):
// regular rule, adaptive Telles/subdivision leaf list (:1050-1172) or singular polar data (:1174-1472).
// The nearest-point Newton iteration runs in __float128 like the reference's real128 (geometry.f90:5107-5150).
// Compiled with -ffp-contract=off so that every value feeding a discrete decision is formed as in the reference.
#pragma once
#include <vector>
#include <complex>

namespace mfbh {

enum { LINE2 = 2, LINE3 = 3, TRI3 = 5, TRI6 = 6, QUAD4 = 7, QUAD8 = 8, QUAD9 = 9 };
int nodes_of(int et);

struct QsTable {  // N(d) fit parameters, [curve][f], f = 1 and 5 only are used on this path
  double dmin[3][8], dmax[3][8], a0[3][8], a1[3][8], b1[3][8];
  double t_dmin[3][8], t_dmax[3][8], t_a0[3][8], t_a1[3][8], t_b1[3][8];
};
void qs_table(double relative_error, QsTable& q);

struct Settings {
  double qsi_relative_error; int qsi_ns_max; std::vector<int> ps_gln; double geometric_tolerance;
  QsTable qs;       // for qsi_relative_error
  QsTable qs_li;    // for the 1e-15 line integrals of the singular path
  double far_thr[32];  // far_thr[n], n=2..30: smallest d>2 with N_far(d) <= n   (GPU classifier compares d against it)
  double far_dmax;     // d >= far_dmax -> gln_near = 2
  int f = 5;           // order of the estimator's model function: 5 for the SBIE kernels, 7 for the hypersingular ones
                       // (fbem_bem_harela3d_sbie_auto :1522 / _hbie_auto :3677)
};
void init_settings(Settings& s);

struct Elem {
  int et, nn; double x[27]; bool reversed;
  double cl; int gln_far; double bc[3], br;
};
void element_data(Elem& e, const Settings& s);
int pointset_size(int et, int gln);
// packed point set of one element and one rule: ngp records of (x[3], n[3], phi_j*J*w [nn])
void build_pointset(const Elem& e, int gln, double* out);

struct Leaf { double xi_s[8]; double tp1[4], tp2[4]; int gln; };
struct Ray { double ct, st, rhoij, w; };  // cos(theta), sin(theta), rho_max(theta), jthetap*w_angular
struct NearPlan {
  int mode;   // 0 regular (set index in `set`), 1 adaptive, 2 singular
  int set; int gln;
  std::vector<Leaf> leaves;
  double xi_i[2]; double x_i[3]; double hli[9]; std::vector<Ray> rays;
  long long points;
};
void plan_near_pair(const Elem& e, const double* x_i, const Settings& s, NearPlan& out);

// Geometry of the Mantic free-term matrix at an edge/vertex node (bem_harela3d.f90:365-542):
// c(l,k) = cp*delta_lk - sum_b(l,k) / (8 pi (1-nu)).  Returns nonzero on an invalid normals/tangents configuration.
int mantic_terms(int n_elements, const double* normals, const double* tangents, double tol, double* cp, double* sum_b);
void node_normal_tangent(int et, const double* xn, int node, bool reversed, double* n, double* t);
bool xi_on_element_boundary(int et, const double* xi);
void shape_values(int et, const double* xi, double* phi);
void node_xi(int et, int node, double* xi);

}  // namespace mfbh
