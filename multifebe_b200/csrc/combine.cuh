// combine.cuh -- launchers of the resident column combination (combine.cu)
#pragma once
#include "assembly.cuh"

namespace mfbd {
void launch_combine(const DevSystem& src, const DevSystem& dst, int n_rows, const int* src_row, const int* dst_row, int n_terms, const int* src_col,
                    const int* dst_col, const double* coef, cudaStream_t st);
void launch_add_entries(const DevSystem& dst, int n, const int* rows, const int* cols, const double* v, cudaStream_t st);
}  // namespace mfbd
