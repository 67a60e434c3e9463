// poro.cuh -- launchers of the Biot poroelastic assembly kernels (poro.cu); data layout of assembly.cuh with four components per node.
#pragma once
#include "assembly.cuh"
#include "por_math.cuh"

namespace mfbd {

void set_por_params(const PorParams& pp, cudaStream_t st);
void launch_por_regular(const DevGroup& g, const DevColloc& c, const DevSystem& s, const unsigned char* plan, cudaStream_t st);
void launch_por_adaptive(const DevGroup& g, const DevColloc& c, const DevSystem& s, const DevAdaptive& a, const DevTables& t, cudaStream_t st);
void launch_por_singular(const DevGroup& g, const DevColloc& c, const DevSystem& s, const DevSingular& a, const DevTables& t, cudaStream_t st);

}  // namespace mfbd
