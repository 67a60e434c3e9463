// Host build of the product's scalar-wave point arithmetic (multifebe_b200/csrc/pot_math.cuh is host+device inline) so that
// tests/test_pot_math_host.py can compare it with the CPU oracle without a GPU.  Test infrastructure only.
#include "../../multifebe_b200/csrc/pot_math.cuh"
#include <complex>
using namespace mfbd;
extern "C" {
void pmh_E23(const double* z_ri, double* out4) {
  cplx E2, E3; pot_E23(mk(z_ri[0], z_ri[1]), E2, E3);
  out4[0] = E2.re; out4[1] = E2.im; out4[2] = E3.re; out4[3] = E3.im;
}
// one quadrature point with unit weight: out = (sum_h re, im, sum_g re, im) of pot_accumulate<1>, i.e. fs_Q dr/dn and fs_P
void pmh_point(double omega, double rho, const double* c_ri, const double* x, const double* n, const double* xc, double* out4) {
  typedef std::complex<double> cd;
  const cd im(0.0, 1.0), k = omega / cd(c_ri[0], c_ri[1]);
  PotParams pp;
  pp.k = mk(k.real(), k.imag());
  { cd v = -im * k; pp.P1 = mk(v.real(), v.imag()); }
  { cd v = 0.5 * (k * k); pp.Q1 = mk(v.real(), v.imag()); }
  { cd v = im * k; pp.Q2 = mk(v.real(), v.imag()); }
  pp.c4pi = 0.07957747154594767280411105048; pp.d1J = rho * omega * omega;
  PAcc<1> a; a.zero();
  const double w = 1.0;
  pot_accumulate<1>(a, pp, x, n, xc, &w);
  out4[0] = a.hr[0]; out4[1] = a.hi[0]; out4[2] = a.gr[0]; out4[3] = a.gi[0];
}
}
