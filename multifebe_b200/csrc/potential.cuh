// potential.cuh -- device-side parameters and launchers of the scalar-wave (inviscid fluid / acoustic) assembly kernels
// (first brick: one fluid BE region with ordinary `be` boundaries).  The problem object, the
// collocation tiles, the quadrature plan (estimator order f = 3) and the leaf / ray lists are the ones of assembly.cuh; a
// fluid node has ONE equation and ONE unknown, so every per-(element, node) descriptor (ecol, ekind, ecv) has stride 1.
#pragma once
#include "assembly.cuh"
#include "pot_math.cuh"

namespace mfbd {

void set_pot_params(const PotParams& pp, cudaStream_t st);
void launch_pot_regular(const DevGroup& g, const DevColloc& c, const DevSystem& s, const unsigned char* plan, cudaStream_t st);
void launch_pot_adaptive(const DevGroup& g, const DevColloc& c, const DevSystem& s, const DevAdaptive& a, const DevTables& t, cudaStream_t st);
void launch_pot_singular(const DevGroup& g, const DevColloc& c, const DevSystem& s, const DevSingular& a, const DevTables& t, cudaStream_t st);

}  // namespace mfbd
