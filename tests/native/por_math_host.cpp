// Host build of the product's poroelastic point arithmetic (multifebe_b200/csrc/por_math.cuh is host+device inline) so that
// tests/test_por_math_host.py can compare it with the CPU oracle without a GPU.  Test infrastructure only.
#include "../../multifebe_b200/csrc/por_math.cuh"
using namespace mfbd;
typedef std::complex<double> cd;
static void params(const double* pr, double omega, PorParams& P) {
  por_params_host(cd(pr[0], pr[1]), cd(pr[2], pr[3]), pr[4], pr[5], pr[6], cd(pr[7], pr[8]), cd(pr[9], pr[10]), pr[11], omega, P);
}
static cd C(cplx z) { return cd(z.re, z.im); }
extern "C" {
// u*, t* (4 x 4, [l][k], constants applied) at one exterior point, and k1, k2, k3, Z, J
void pmh_por_exterior(double omega, const double* props, const double* x, const double* n, const double* xc, cd* u, cd* t, cd* k5) {
  PorParams P; params(props, omega, P);
  cplx fu[4][4], ft[4][4];
  por_exterior_blocks(P, x, n, xc, fu, ft);
  for (int l = 0; l < 4; l++) for (int k = 0; k < 4; k++) { u[4 * l + k] = C(P.cte_u[l][k]) * C(fu[l][k]); t[4 * l + k] = C(P.cte_t[l][k]) * C(ft[l][k]); }
  k5[0] = C(P.k1); k5[1] = C(P.k2); k5[2] = C(P.k3); k5[3] = C(P.Z); k5[4] = C(P.J);
}
// interior form at the same point: u*, t* with the CPV kernel fc added back (must equal the exterior form), and fc alone (3 x 3)
void pmh_por_interior(double omega, const double* props, const double* x, const double* n, const double* xc, cd* u, cd* t, cd* fc9) {
  PorParams P; params(props, omega, P);
  cplx fu[4][4], ft[4][4], fc[3][3];
  por_interior_blocks(P, x, n, xc, fu, ft, fc);
  for (int l = 0; l < 4; l++) for (int k = 0; k < 4; k++) {
    cd tt = C(ft[l][k]);
    if (l > 0 && k > 0) tt += C(fc[l - 1][k - 1]);
    u[4 * l + k] = C(P.cte_u[l][k]) * C(fu[l][k]); t[4 * l + k] = C(P.cte_t[l][k]) * tt;
  }
  for (int l = 0; l < 3; l++) for (int k = 0; k < 3; k++) fc9[3 * l + k] = C(fc[l][k]);
}
}
