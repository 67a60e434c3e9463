// lu.cuh -- blocked right-looking complex LU (partial pivoting) + triangular solves on planar (split re/im) storage.
// Replaces OpenBLAS zgetrf/zgetrs behind solve_lse_c (src/solve_lse_c.f90:124,176).
#pragma once
#include <cuda_runtime.h>

namespace mfbd {

struct LuWork {
  int nb;                 // outer block size
  int ib;                 // sub-panel width (columns kept in shared memory by the cooperative panel kernel)
  int n_sm;
  double *cand_val;       // [2][grid]         pivot candidates (|re|+|im|)
  int *cand_row;          // [2][grid]
  double *cand_data;      // [2][grid][2*32]   candidate rows (re, im)
  double *diag_data;      // [2][2*32]         current diagonal row
  int *info;              // device flag: first zero pivot (1-based), 0 = ok
  float ms_panel, ms_swap, ms_trsm, ms_gemm; long long launches; long long gemm_launches; double gemm_flops, gemm_exec_flops;   // algorithmic (8mnk) and executed (6mnk with the 3M kernel) flops of the trailing updates
  cudaEvent_t* evs; int n_evs, n_steps_timed;   // 5 events per block step, recorded without synchronising
  cudaStream_t panel_stream;                    // high-priority stream of the look-ahead panel factorisation
  cudaEvent_t ev_next_cols, ev_panel_done;      // next panel's columns updated / next panel factorised
  cudaEvent_t* pevs;                            // 2 events per block step on the panel stream (panel timing under look-ahead)
  int lookahead, panel_ctas;
};
// sums the per-phase event times of the last timed factorisation (call after the stream has been synchronised)
void lu_collect_times(LuWork& w);

int lu_work_alloc(LuWork& w, int n, int nb);
void lu_work_free(LuWork& w);

// In-place LU of the n x n planar matrix; ipiv (device, 1-based, LAPACK convention).  Returns cudaError as int (0 ok).
int zgetrf_planar(double* Are, double* Aim, long long lda, int n, int* ipiv, LuWork& w, cudaStream_t st, bool timing);
// Solve with the factors: b (planar, n x nrhs, ldb) overwritten by the solution.
int zgetrs_planar(const double* Are, const double* Aim, long long lda, int n, const int* ipiv, double* bre, double* bim, long long ldb,
                  int nrhs, cudaStream_t st);
// C -= A*B on planar storage (the trailing-matrix update; FP64 tensor pipe, mma.sync m8n8k4).  k must be a multiple of 4.
void zgemm_minus_planar(int m, int n, int k, const double* Are, const double* Aim, long long lda, const double* Bre, const double* Bim,
                        long long ldb, double* Cre, double* Cim, long long ldc, cudaStream_t st);
// micro-benchmarks (TFLOP/s, GB/s)
double bench_dfma(cudaStream_t st);
double bench_dmma(cudaStream_t st);
double bench_copy(cudaStream_t st);

}  // namespace mfbd
