// plan_host.cpp -- see plan_host.h.  Host-side quadrature planning (product code, runs once per mesh).
//
// The geometric primitives here must take bit-identical discrete decisions to the reference
// (rule order from ceiling()/nint() of a fit in d, nearest point in real128, subdivision depth), so they follow the
// reference formulas operation by operation; file:line citations are relative to /root/reference.
#include "plan_host.h"
#include "plan_values.h"
#include <cmath>
#include <cstring>
#include <cstdint>
#include <algorithm>
#include <quadmath.h>
#include "../../data/quad_tables.h"

namespace mfbh {
typedef __float128 q128;
typedef std::complex<double> cd;

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
// Shape functions (continuous, delta = 0); T = double or __float128 (the reference mixes real64 xi with real128
// aux variables in the nearest-point iteration, lib/fbem/src/geometry.f90:5118-5123).
// lib/fbem/src/resources_shape_functions/{phi,dphidxi1,dphidxi2}_{tri3,tri6,quad4,quad8,quad9}.rc, phi_line{2,3}.rc
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


// ---- rule-order estimator N(d): lib/fbem/src/quasisingular_integration.f90:181-287 (fit data), :318-398, :402-711 ----
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

static inline double poly1(const double c[3][3], int cv, double x) { return c[0][cv] + c[1][cv] * x + c[2][cv] * (x * x); }
static inline double poly2(const double c[6][3], int cv, double x, double y) {
  return c[0][cv] + c[1][cv] * x + c[2][cv] * y + c[3][cv] * x * y + c[4][cv] * (x * x) + c[5][cv] * (y * y);
}
void qs_table(double relative_error, QsTable& p) {
  double log10re = log10(fabs(relative_error));
  if (log10re > -3.0) log10re = -3.0;
  if (log10re < -15.0) log10re = -15.0;
  // log10(30.), log10( 2.) are default-real (single precision) intrinsics promoted to double (:343-344)
  const double l30 = (double)log10f(30.f), l2 = (double)log10f(2.f);
  for (int c = 0; c < 3; c++) {
    p.a0[c][0] = poly1(lnr_alpha0, c, log10re); p.a1[c][0] = poly1(lnr_alpha1, c, log10re); p.b1[c][0] = poly1(lnr_beta1, c, log10re);
    p.dmin[c][0] = pow(10.0, (l30 - p.a0[c][0]) / (p.a1[c][0] - l30 * p.b1[c][0]));
    p.dmax[c][0] = pow(10.0, (l2 - p.a0[c][0]) / (p.a1[c][0] - l2 * p.b1[c][0]));
    p.t_a0[c][0] = poly1(t_lnr_alpha0, c, log10re); p.t_a1[c][0] = poly1(t_lnr_alpha1, c, log10re); p.t_b1[c][0] = poly1(t_lnr_beta1, c, log10re);
    p.t_dmin[c][0] = pow(10.0, (l30 - p.t_a0[c][0]) / (p.t_a1[c][0] - l30 * p.t_b1[c][0]));
    p.t_dmax[c][0] = pow(10.0, (l2 - p.t_a0[c][0]) / (p.t_a1[c][0] - l2 * p.t_b1[c][0]));
    if (p.t_dmin[c][0] < 1.e-6) p.t_dmin[c][0] = 1.e-6;
    if (p.t_dmin[c][0] > p.t_dmax[c][0]) p.t_dmin[c][0] = 1.e-6;
  }
  for (int i = 1; i <= 7; i++)
    for (int c = 0; c < 3; c++) {
      double y = (double)i;
      p.a0[c][i] = poly2(d1rn_alpha0, c, log10re, y); p.a1[c][i] = poly2(d1rn_alpha1, c, log10re, y); p.b1[c][i] = poly2(d1rn_beta1, c, log10re, y);
      p.dmin[c][i] = pow(10.0, (l30 - p.a0[c][i]) / (p.a1[c][i] - l30 * p.b1[c][i]));
      p.dmax[c][i] = pow(10.0, (l2 - p.a0[c][i]) / (p.a1[c][i] - l2 * p.b1[c][i]));
      p.t_a0[c][i] = poly2(t_d1rn_alpha0, c, log10re, y); p.t_a1[c][i] = poly2(t_d1rn_alpha1, c, log10re, y); p.t_b1[c][i] = poly2(t_d1rn_beta1, c, log10re, y);
      p.t_dmin[c][i] = pow(10.0, (l30 - p.t_a0[c][i]) / (p.t_a1[c][i] - l30 * p.t_b1[c][i]));
      p.t_dmax[c][i] = pow(10.0, (l2 - p.t_a0[c][i]) / (p.t_a1[c][i] - l2 * p.t_b1[c][i]));
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
static int qs_n_estimation(bool telles, int etype, int f, const QsTable& p, double d, const double* barxi) {
  const double(*dmin)[8] = telles ? p.t_dmin : p.dmin;
  const double(*dmax)[8] = telles ? p.t_dmax : p.dmax;
  const double(*a0)[8] = telles ? p.t_a0 : p.a0;
  const double(*a1)[8] = telles ? p.t_a1 : p.a1;
  const double(*b1)[8] = telles ? p.t_b1 : p.b1;
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
    else { double log10d = log10(d); n = (int)lround(pow(10.0, (p.a0[2][f] + p.a1[2][f] * log10d) / (1.0 + p.b1[2][f] * log10d))); }
  }
  if (n > 30) n = 0;
  return n;
}


// ---- Telles transformation: lib/fbem/src/telles_transformation.f90:77-232 ----
static double telles_barr_any(double d) {  // fbem_telles_barr(d, fbem_f_any) :77-127
  double b = (d < 3.0) ? d / (0.89039 * d + 0.32883) : 1.0;
  if (b > 1.0) b = 1.0;
  return b;
}


// ---- geometry: lib/fbem/src/geometry.f90 ----
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


// DECISION part of fbem_polar_transformation_setup (polar_transformation.f90:251-498) + fbem_bem_harela3d_sbie_int :1298-1299: which of the 8 (6)
// sub-triangles around xi_i exist and how many angular Gauss points each one gets, ngp = 5 + nint(25 (theta2 - theta1) / (pi/2)).  nint of a
// continuous quantity (25 dtheta / (pi/2) = 12.5 for the very common dtheta = pi/4) is a discrete decision, so the angles are formed exactly as
// the reference forms them.  The rays themselves (values) are computed independently in plan_values.cpp.  counts[sub - 1], 0 = absent.
static void polar_theta_counts(int et, const double* xi_i, int* counts) {
  int nsub; int sub[8]; double th[8][2];
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
          th[k][0] = t; th[k][1] = 1.5 * c_pi; break;
        case 2: t = c_2pi + asin((-1.0 - b) / sqrt((1.0 - a) * (1.0 - a) + (-1.0 - b) * (-1.0 - b)));
          th[k][0] = 1.5 * c_pi; th[k][1] = t; break;
        case 3: t = c_2pi + asin((-1.0 - b) / sqrt((1.0 - a) * (1.0 - a) + (-1.0 - b) * (-1.0 - b)));
          th[k][0] = t; th[k][1] = c_2pi; break;
        case 4: t = asin((1.0 - b) / sqrt((1.0 - a) * (1.0 - a) + (1.0 - b) * (1.0 - b)));
          th[k][0] = 0.0; th[k][1] = t; break;
        case 5: t = asin((1.0 - b) / sqrt((1.0 - a) * (1.0 - a) + (1.0 - b) * (1.0 - b)));
          th[k][0] = t; th[k][1] = c_pi_2; break;
        case 6: t = c_pi - asin((1.0 - b) / sqrt((-1.0 - a) * (-1.0 - a) + (1.0 - b) * (1.0 - b)));
          th[k][0] = c_pi_2; th[k][1] = t; break;
        case 7: t = c_pi - asin((1.0 - b) / sqrt((-1.0 - a) * (-1.0 - a) + (1.0 - b) * (1.0 - b)));
          th[k][0] = t; th[k][1] = c_pi; break;
        case 8: t = c_pi - asin((-1.0 - b) / sqrt((-1.0 - a) * (-1.0 - a) + (-1.0 - b) * (-1.0 - b)));
          th[k][0] = c_pi; th[k][1] = t; break;
      }
    } else {
      switch (sub[k]) {
        case 1: t = c_2pi - asin(b / sqrt((1.0 - a) * (1.0 - a) + b * b));
          th[k][0] = t; th[k][1] = c_pi_4 + c_2pi; break;
        case 2: t = c_pi - asin((1.0 - b) / sqrt(a * a + (1.0 - b) * (1.0 - b)));
          th[k][0] = c_pi_4; th[k][1] = t; break;
        case 3: t = c_pi - asin((1.0 - b) / sqrt(a * a + (1.0 - b) * (1.0 - b)));
          th[k][0] = t; th[k][1] = c_pi; break;
        case 4: t = c_pi + asin(b / sqrt(a * a + b * b));
          th[k][0] = c_pi; th[k][1] = t; break;
        case 5: t = c_pi + asin(b / sqrt(a * a + b * b));
          th[k][0] = t; th[k][1] = 1.5 * c_pi; break;
        case 6: t = c_2pi - asin(b / sqrt((1.0 - a) * (1.0 - a) + b * b));
          th[k][0] = 1.5 * c_pi; th[k][1] = t; break;
      }
    }
  }
  for (int k = 0; k < 8; k++) counts[k] = 0;
  for (int k = 0; k < nsub; k++) {
    int ngp_theta = 5 + (int)lround(25.0 * (th[k][1] - th[k][0]) / c_pi_2);
    if (ngp_theta > 32) ngp_theta = 32;
    counts[sub[k] - 1] = ngp_theta;
  }
}

// =====================================================================================
// Product-specific planning on top of the primitives above
// =====================================================================================
int nodes_of(int et) { return n_nodes_of(et); }
bool xi_on_element_boundary(int et, const double* xi) { return check_xi1xi2_edge(et, xi); }

static inline int qs_n(bool telles, int et, int f, const QsTable& q, double d, const double* barxi) { return qs_n_estimation(telles, et, f, q, d, barxi); }

// N_far(d), d>2: curve 3 with nint (quasisingular_integration.f90:543-549); 31 stands for ">30".
static int n_far(const QsTable& q, double d, int f) {
  const double bx[2] = {0.0, 0.0};
  int n = qs_n(false, TRI3, f, q, d, bx);
  return n == 0 ? 31 : n;
}
static inline double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t to_bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

void init_settings(Settings& s) {
  qs_table(s.qsi_relative_error, s.qs);
  qs_table(1.e-15, s.qs_li);
  s.far_dmax = s.qs.dmax[2][s.f];
  // far_thr[n] = smallest double d > 2 with N_far(d) <= n, found by bisection over the ordered bit patterns of
  // positive doubles with the exact host estimator, so that the GPU only compares d against thresholds.
  double dlo = nextafter(2.0, 3.0), dhi = std::max(s.far_dmax, dlo);
  for (int n = 2; n <= 30; n++) {
    if (n_far(s.qs, dlo, s.f) <= n) { s.far_thr[n] = dlo; continue; }
    uint64_t lo = to_bits(dlo), hi = to_bits(dhi);  // N(lo) > n, N(hi) = 2 <= n
    while (hi - lo > 1) { uint64_t mid = lo + (hi - lo) / 2; if (n_far(s.qs, from_bits(mid), s.f) <= n) hi = mid; else lo = mid; }
    s.far_thr[n] = from_bits(hi);
  }
  s.far_thr[0] = s.far_thr[1] = s.far_thr[31] = 0.0;
}

// csize, n_phi, bounding ball: src/build_data_of_be_elements.f90:62-110
void element_data(Elem& e, const Settings& s) {
  e.nn = n_nodes_of(e.et);
  e.cl = characteristic_length(e.et, e.x, 1.e-9);
  e.gln_far = phijac_ngp_2d(e.et, e.x, s.qsi_relative_error);
  element_ball(e.et, e.x, e.gln_far, e.bc, e.br);
}

// Leaf list of the adaptive quasi-singular integration (the decisions of fbem_bem_harela3d_sbie_ext_adp, bem_harela3d.f90:1050-1172): a
// sub-element is integrated when the Telles estimator returns a rule for its normalised distance (or the depth limit is reached), otherwise it is
// split into four at its edge midpoints.  The recursion of the reference is run as an explicit work list (depth-first, children pushed in reverse
// so that leaves come out in the reference's order); the DECISIONS (nearest point and distance of each sub-element, rule order, split or not) use
// the decision core above, the leaf's transformation coefficients (values) come from plan_values.cpp.
struct SubElem { double xi[8]; int depth; };
static void collect_leaves(const Elem& e, const double* x_i, const Settings& s, NearPlan& out) {
  const int nv = n_vertices_of(e.et);
  std::vector<SubElem> work(1);
  {
    static const double T0[6] = {1, 0, 0, 1, 0, 0}, Q0[8] = {-1, -1, 1, -1, 1, 1, -1, 1};
    memcpy(work[0].xi, nv == 3 ? T0 : Q0, sizeof(double) * 2 * nv); work[0].depth = 1;
  }
  while (!work.empty()) {
    const SubElem cur = work.back(); work.pop_back();
    double barxip[2], rmin, d; int method;
    if (cur.depth == 1) nearest_element_point_bem(e.et, e.x, e.cl, x_i, barxip, rmin, d, method);
    else {
      double x_s[27]; subdivision_coordinates(e.et, e.x, cur.xi, x_s);
      const double cl = characteristic_length(e.et, x_s, 1.e-12);
      nearest_element_point_bem(e.et, x_s, cl, x_i, barxip, rmin, d, method);
    }
    int gln_near = qs_n(true, e.et, s.f, s.qs, d, barxip);
    if (gln_near == 0 && cur.depth == s.qsi_ns_max) gln_near = 30;
    if (gln_near > 0) {
      Leaf lf; memset(&lf, 0, sizeof(lf));
      memcpy(lf.xi_s, cur.xi, sizeof(double) * 2 * nv);
      const double barr = telles_barr_any(d);
      if (nv == 4) { telles_cubic(false, barxip[0], barr, lf.tp1); telles_cubic(false, barxip[1], barr, lf.tp2); }
      else {   // collapsed square of the triangle (bem_harela3d.f90:897-924): (xi1 / (1 - xi2), xi2), the apex mapped to the middle of its side
        const bool apex = barxip[1] > 0.995;
        telles_cubic(true, apex ? 0.5 : barxip[0] / (1.0 - barxip[1]), barr, lf.tp1); telles_cubic(true, apex ? 1.0 : barxip[1], barr, lf.tp2);
      }
      lf.gln = std::max(gln_near, e.gln_far);
      out.leaves.push_back(lf); out.points += (long long)lf.gln * lf.gln;
      continue;
    }
    // split: corner c keeps its vertex, the others become the midpoints of its two edges (and the centroid for a quadrilateral); a triangle has a
    // fourth, central child made of the three midpoints
    const double* v = cur.xi;
    auto put = [&](double* o, double a, double b) { o[0] = a; o[1] = b; };
    auto midp = [&](int i, int j, double* o) { put(o, 0.50 * (v[2 * i] + v[2 * j]), 0.50 * (v[2 * i + 1] + v[2 * j + 1])); };
    SubElem ch[4];
    for (int c = 0; c < 4; c++) ch[c].depth = cur.depth + 1;
    if (nv == 3) {
      for (int c = 0; c < 3; c++) { put(ch[c].xi, v[2 * c], v[2 * c + 1]); }
      midp(0, 1, ch[0].xi + 2); midp(0, 2, ch[0].xi + 4);
      midp(1, 2, ch[1].xi + 2); midp(0, 1, ch[1].xi + 4);
      midp(0, 2, ch[2].xi + 2); midp(1, 2, ch[2].xi + 4);
      midp(0, 1, ch[3].xi); midp(1, 2, ch[3].xi + 2); midp(0, 2, ch[3].xi + 4);
    } else {
      double ctr[2]; put(ctr, 0.25 * (v[0] + v[2] + v[4] + v[6]), 0.25 * (v[1] + v[3] + v[5] + v[7]));
      for (int c = 0; c < 4; c++) {     // child c: vertex c stays in slot c, slot c + 2 is the centroid, the other two slots are edge midpoints
        const int nx = (c + 1) & 3, pv = (c + 3) & 3;
        put(ch[c].xi + 2 * c, v[2 * c], v[2 * c + 1]);
        midp(c, nx, ch[c].xi + 2 * nx); midp(c, pv, ch[c].xi + 2 * pv);
        put(ch[c].xi + 2 * ((c + 2) & 3), ctr[0], ctr[1]);
      }
    }
    for (int c = 3; c >= 0; c--) work.push_back(ch[c]);
  }
}

// omega-independent data of the singular element integral (fbem_bem_harela3d_sbie_int, bem_harela3d.f90:1174-1472): the decision core says which
// sub-triangles exist and how many angular points each gets; rays and edge line integrals are values (plan_values.cpp).
static void plan_singular(const Elem& e, const double* xi_i, const Settings& s, NearPlan& out) {
  (void)s;
  out.xi_i[0] = xi_i[0]; out.xi_i[1] = xi_i[1];
  element_point(e.et, e.x, xi_i, out.x_i);
  int counts[8]; polar_theta_counts(e.et, xi_i, counts);
  polar_rays(e.et, xi_i, counts, out.rays);
  out.points = (long long)out.rays.size() * 15;
  for (int i = 0; i < 9; i++) out.hli[i] = 0.0;
  bool edge_on[4];
  for (int k = 0; k < n_edges_of(e.et); k++) edge_on[k] = counts[2 * k] > 0 || counts[2 * k + 1] > 0;   // edges that do not contain the collocation point
  edge_integrals(e.et, e.x, out.x_i, edge_on, out.hli);
}

// fbem_bem_harela3d_sbie_auto decisions for one pair the GPU classifier could not settle with the ball test
// (lib/fbem/src/bem_harela3d.f90:1501-1537).
void plan_near_pair(const Elem& e, const double* x_i, const Settings& s, NearPlan& out) {
  out.leaves.clear(); out.rays.clear(); out.points = 0; out.set = -1; out.gln = 0;
  double r[3] = {e.bc[0] - x_i[0], e.bc[1] - x_i[1], e.bc[2] - x_i[2]};
  double rmin = sqrt(dot3(r, r)) - e.br, barxi[2], d; int method;
  if (rmin > (4.0 * e.br)) { barxi[0] = 0.0; barxi[1] = 0.0; d = rmin / e.cl; }
  else {
    nearest_element_point_bem(e.et, e.x, e.cl, x_i, barxi, rmin, d, method);
    if (d <= 1.e-12) { out.mode = 2; plan_singular(e, barxi, s, out); return; }
  }
  int gln_near = qs_n(false, e.et, s.f, s.qs, d, barxi);
  int gln = std::max(e.gln_far, gln_near);
  int ps_gln_max = 0; for (int g : s.ps_gln) ps_gln_max = std::max(ps_gln_max, g);
  if (gln <= ps_gln_max && gln_near > 0) {
    for (size_t i = 0; i < s.ps_gln.size(); i++) if (s.ps_gln[i] >= gln) { out.mode = 0; out.set = (int)i; out.gln = s.ps_gln[i]; out.points = pointset_size(e.et, s.ps_gln[i]); return; }
  }
  out.mode = 1; collect_leaves(e, x_i, s, out);
}

}  // namespace mfbh
