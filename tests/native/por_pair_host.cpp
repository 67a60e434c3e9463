// Host build of the per-pair bodies of the poroelastic kernels (multifebe_b200/csrc/por_pair.cuh: what a lane of R1 / R2 / R3 does at an
// integration point, the line-integral terms, the constants and orientation of a finished pair) driven lane-serially by the product's own
// host planner (plan_host.cpp: precalculated point sets, Telles / subdivision leaves, polar rays), so that tests/test_por_pair_host.py can
// hold the whole pair integral -- everything of poro.cu except the lane mapping, the warp reduction and the scatter -- to the CPU oracle
// without a GPU.  Test infrastructure only.
#include "../../multifebe_b200/csrc/por_pair.cuh"
#include "../../multifebe_b200/csrc/plan_host.h"
#include "../../data/quad_tables.h"
#include <vector>
using namespace mfbd;
typedef std::complex<double> cd;

template <int ET>
static void pair_t(const mfbh::Elem& e, const mfbh::NearPlan& pl, const double* x_i, const PorParams& P, cd* h, cd* g) {
  constexpr int NN = ElemTraits<ET>::NN, RECN = 6 + NN;
  constexpr bool tri = (ElemTraits<ET>::NV == 3);
  std::vector<double> pts;
  int npts = 0;
  if (pl.mode == 0) { npts = mfbh::pointset_size(e.et, pl.gln); pts.resize((size_t)npts * RECN); mfbh::build_pointset(e, pl.gln, pts.data()); }
  double phi_i[NN];
  if (pl.mode == 2) { double d1[NN], d2[NN]; shape<ET>(pl.xi_i[0], pl.xi_i[1], phi_i, d1, d2); }
  for (int l = 0; l < 4; l++) {
    RAcc<NN> acc; acc.zero();
    if (pl.mode == 0) {                                                   // k_por_regular
      for (int kp = 0; kp < npts; kp++) { const double* q = pts.data() + (size_t)kp * RECN; por_regular_point<NN>(acc, P, q, q + 3, q + 6, x_i, l); }
    } else if (pl.mode == 1) {                                            // k_por_adaptive
      const double* gx = tri ? QT_GL01_X : QT_GL11_X; const double* gw = tri ? QT_GL01_W : QT_GL11_W;
      for (const mfbh::Leaf& lf : pl.leaves) {
        const int gln = lf.gln, off = gln * (gln - 1) / 2;
        for (int idx = 0; idx < gln * gln; idx++) {
          const int k1 = idx / gln, k2 = idx - k1 * gln;
          por_leaf_point<ET>(acc, P, e.x, lf.xi_s, lf.tp1, lf.tp2, gx[off + k1], gw[off + k1], gx[off + k2], gw[off + k2], x_i, l);
        }
      }
    } else {                                                              // k_por_singular
      const double* gx = QT_GL01_X + 15 * 14 / 2; const double* gw = QT_GL01_W + 15 * 14 / 2;
      for (const mfbh::Ray& r : pl.rays)
        for (int kk = 0; kk < 15; kk++)
          por_singular_point<ET>(acc, P, e.x, pl.xi_i[0], pl.xi_i[1], phi_i, r.ct, r.st, r.rhoij * gx[kk], r.w, gw[kk], pl.x_i, l);
      por_singular_line_terms<NN>(acc, P, phi_i, pl.hli, l);
    }
    for (int k = 0; k < 4; k++) for (int j = 0; j < NN; j++) {
      double hr, hi, gr, gi;
      por_finished_entry<NN>(acc, P, l, k, j, e.reversed, hr, hi, gr, gi);
      h[(j * 4 + l) * 4 + k] = cd(hr, hi); g[(j * 4 + l) * 4 + k] = cd(gr, gi);
    }
  }
}

extern "C" int pph_pair(int et, const double* xn, int reversed, const double* x_i, double omega, const double* pr, double qsi_relative_error, int qsi_ns_max,
                        int n_sets, const int* set_gln, double geometric_tolerance, cd* h, cd* g) {
  mfbh::Settings S;
  S.qsi_relative_error = qsi_relative_error; S.qsi_ns_max = qsi_ns_max; S.geometric_tolerance = geometric_tolerance;
  S.ps_gln.assign(set_gln, set_gln + n_sets); S.f = 5;
  mfbh::init_settings(S);
  mfbh::Elem e; e.et = et; e.nn = mfbh::nodes_of(et); e.reversed = reversed != 0;
  for (int i = 0; i < 3 * e.nn; i++) e.x[i] = xn[i];
  mfbh::element_data(e, S);
  mfbh::NearPlan pl; mfbh::plan_near_pair(e, x_i, S, pl);
  PorParams P; por_params_host(cd(pr[0], pr[1]), cd(pr[2], pr[3]), pr[4], pr[5], pr[6], cd(pr[7], pr[8]), cd(pr[9], pr[10]), pr[11], omega, P);
  switch (et) {
    case 5: pair_t<5>(e, pl, x_i, P, h, g); break;
    case 6: pair_t<6>(e, pl, x_i, P, h, g); break;
    case 7: pair_t<7>(e, pl, x_i, P, h, g); break;
    case 8: pair_t<8>(e, pl, x_i, P, h, g); break;
    case 9: pair_t<9>(e, pl, x_i, P, h, g); break;
    default: return -1;
  }
  return pl.mode;
}
