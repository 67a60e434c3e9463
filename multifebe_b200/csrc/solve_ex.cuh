// solve_ex.cuh -- optional stages of solve_lse_c (equilibration, condition estimate, iterative refinement) on planar device storage; see solve_ex.cu
#pragma once
#include <cuda_runtime.h>
#include <complex>
#include <functional>
#include <vector>

namespace mfbd {

// zgeequ + zlaqge: computes r, c (host copies and device arrays d_r, d_c), scales A in place as LAPACK would, sets equed in {N, R, C, B};
// info > 0: row info (or column info - n) is exactly zero.  Returns a cudaError as int.
int zequilibrate(double* Are, double* Aim, long long lda, int n, double* d_r, double* d_c, std::vector<double>& r, std::vector<double>& c, double* rowcnd, double* colcnd,
                 double* amax, char* equed, int* info, cudaStream_t st);
void scale_vector(double* re, double* im, int n, const double* d_f, cudaStream_t st);
// max_j sum_i |a_ij| (d_tmp: n doubles of device scratch)
double matrix_norm1(const double* re, const double* im, long long ld, int n, double* d_tmp, cudaStream_t st);
// x := inv(A)^H x with the factors P A = L U and the diagonal-block inverses of the factorisation; tmp: 4 n + 128 doubles of device scratch
int zgetrs_conjtrans_planar(const double* Are, const double* Aim, long long lda, int n, const int* d_perm, const double* inv, double* xre, double* xim, double* tmp, cudaStream_t st);
// 1-norm estimate of an operator given by v := B v (false) / v := B^H v (true): Hager's method in Higham's form (zlacn2)
double norm1_estimate(int n, const std::function<void(std::vector<std::complex<double>>&, bool)>& apply);

}  // namespace mfbd
