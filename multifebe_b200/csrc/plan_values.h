// plan_values.h -- value-producing host geometry of the quadrature plans (plan_values.cpp): independent derivations, see that file's header.
#pragma once
#include "plan_host.h"
#include <vector>

namespace mfbh {

// shape functions and their parametric derivatives at xi (2-D elements TRI3 / TRI6 / QUAD4 / QUAD8 / QUAD9)
void shape_all(int et, const double* xi, double* phi, double* d1, double* d2);
// x(xi) on the element
void element_point(int et, const double* xn, const double* xi, double* x);
// coefficients (a, b, c, d) of Telles' cubic on [-1, 1] (unit_interval = false) or [0, 1] for the nearest point xibar and end slope rbar
void telles_cubic(bool unit_interval, double xibar, double rbar, double* c);
// rays of the polar integration around xi_i; counts[2 * edge + half] = angular points of the sub-triangle (0: absent), from the decision core
void polar_rays(int et, const double* xi_i, const int* counts, std::vector<Ray>& rays);
// hli[l][k] += -eps_lkm sum_edges int t_m / r ds over the edges with edge_on[k]
void edge_integrals(int et, const double* xn, const double* x_i, const bool* edge_on, double* hli);

}  // namespace mfbh
