// lu.cuh -- blocked right-looking complex LU (partial pivoting) + triangular solves on planar (split re/im) storage.
// Replaces OpenBLAS zgetrf/zgetrs behind solve_lse_c (src/solve_lse_c.f90:124,176).
#pragma once
#include <cuda_runtime.h>
#include <vector>

namespace mfbd {

// Tensor maps of the operand planes of the TMA trailing-update kernel (gemm_tma.cu): m[0..1] = planes of the matrix the A operand is cut from,
// m[2..3] = planes of the matrix the B operand is cut from (whole planes; the operands are addressed by row / column offsets at launch).
struct GemmTmaMaps { alignas(64) unsigned char m[4][128]; int ok; int pad_; };
int gemm_tma_make_maps(GemmTmaMaps& t, const double* Are, const double* Aim, long long lda, int a_rows, int a_cols, const double* Bre, const double* Bim, long long ldb,
                       int b_rows, int b_cols);
bool gemm_tma_usable(const GemmTmaMaps& t, int m, int n, int k);      // maps valid and k a positive multiple of 16
// C[0:m, 0:n] -= A[a_row0 : a_row0 + m, a_col0 : a_col0 + k] * B[b_row0 : b_row0 + k, b_col0 : b_col0 + n]  (3M product on the FP64 tensor pipe)
void zgemm_minus_planar_tma(const GemmTmaMaps& t, int m, int n, int k, int a_row0, int a_col0, int b_row0, int b_col0, double* Cre, double* Cim, long long ldc, cudaStream_t st);

struct LuWork {
  int nb;                 // outer block size
  int ib;                 // sub-panel width (columns kept in shared memory by the cooperative panel kernel)
  int n_sm;
  double *cand_val;       // [2][grid]         pivot candidates (|re|+|im|)
  int *cand_row;          // [2][grid]
  double *cand_data;      // [2][grid][2*32]   candidate rows (re, im)
  double *diag_data;      // [2][2*32]         current diagonal row
  int *info;              // device flag: first zero pivot (1-based), 0 = ok
  double *asum[2];        // Ar + Ai of the L21 panel of the current / next step (pre-summed operand of the 3M trailing update)
  double *solve_ws;       // 2 n doubles: scratch of zgetrs_planar
  double *inv;            // inverses of the 64 x 64 diagonal blocks of L and U of the last factorisation (NULL: substitution kernels)
  int *pu_arrive;         // [64] arrival counters of k_panel_update (one per column block; self re-arming)
  float ms_panel, ms_swap, ms_trsm, ms_gemm; long long launches; long long gemm_launches; double gemm_flops, gemm_exec_flops;   // algorithmic (8mnk) and executed (6mnk with the 3M kernel) flops of the trailing updates
  cudaEvent_t* evs; int n_evs, n_steps_timed;   // 5 events per block step, recorded without synchronising
  cudaStream_t panel_stream;                    // high-priority stream of the look-ahead panel factorisation
  cudaEvent_t ev_next_cols, ev_panel_done;      // next panel's columns updated / next panel factorised
  cudaEvent_t* pevs;                            // 2 events per block step on the panel stream (panel timing under look-ahead)
  int lookahead, panel_ctas;
  int fused_panel_update;                      // 1: TRSM + update of the panel columns right of a sub-panel in one kernel (k_panel_update)
  int cluster_ib;                              // preferred sub-panel width of the cluster kernel (if the slab fits)
  int cluster_min_rows;                        // rows per CTA below which a sub-panel takes a smaller cluster
  int cluster, cluster_max_rows;               // CTAs of the cluster-resident panel kernel (0 = grid-wide kernel only); tallest panel it takes
  GemmTmaMaps tma; const double* tma_key;      // tensor maps of the matrix being factorised (rebuilt when the matrix pointer changes)
};
// sums the per-phase event times of the last timed factorisation (call after the stream has been synchronised)
void lu_collect_times(LuWork& w);

int lu_work_alloc(LuWork& w, int n, int nb);
void lu_work_free(LuWork& w);

// In-place LU of the n x n planar matrix; ipiv (device, 1-based, LAPACK convention).  Returns cudaError as int (0 ok).
int zgetrf_planar(double* Are, double* Aim, long long lda, int n, int* ipiv, LuWork& w, cudaStream_t st, bool timing);
// Solve with the factors: b (planar, n x nrhs, ldb) overwritten by the solution.
// inv = LuWork::inv of the factorisation (or NULL: serial substitution on the diagonal blocks).
int zgetrs_planar(const double* Are, const double* Aim, long long lda, int n, const int* ipiv, double* bre, double* bim, long long ldb,
                  int nrhs, cudaStream_t st, const double* inv = nullptr, double* ws = nullptr);
// perm (device) from ipiv (device, 1-based) without the host; bad (device int) is set when a pivot is out of range.  Returns -1 for n > 12000.
int launch_perm_from_ipiv(const int* ipiv, int* perm, int n, int* bad, cudaStream_t st);
// C -= A*B on planar storage (the trailing-matrix update; FP64 tensor pipe, mma.sync m8n8k4).  k must be a multiple of 4.
void zgemm_minus_planar(int m, int n, int k, const double* Are, const double* Aim, long long lda, const double* Bre, const double* Bim,
                        long long ldb, double* Cre, double* Cim, long long ldc, cudaStream_t st);

// ------------------------------------------------------------------------------------------------------------------
// Distributed LU of ONE system over P ranks (one GPU each): 1-D block-cyclic COLUMN layout, block = nb columns, block j
// lives on rank j % P as its local block j / P.  Every rank holds all n rows of its columns plus a replicated copy of the
// right-hand side as one extra local column, so that the forward substitution happens inside the factorisation.
// Per step the owner factorises the panel and the panel (L11, L21, pivots) is BROADCAST; everybody then interchanges,
// solves (TRSM) and updates (GEMM on the FP64 tensor pipe) its own columns.  One panel of look-ahead hides the broadcast.
// The collectives are behind DistComm: NCCL between processes, or plain device copies between the virtual ranks of one
// process (single-GPU self test of the index logic).
// ------------------------------------------------------------------------------------------------------------------
struct DistRank {
  int rank, ncl;                 // global rank; local columns (without the right-hand-side column)
  double *Lre, *Lim;             // planar lda x (ncl + 1), column ncl = right-hand side / solution workspace
  double* pbuf[3];               // packed panel: [2 planes][nbw][mp] doubles + nbw pivots (int) behind them
  double* xfin;                  // [2][lda] solution blocks of the columns this rank owns (zero elsewhere)
  int* ipiv;                     // [n] device, 1-based global rows
  LuWork w;                      // panel workspace of the cooperative kernel
  cudaStream_t main, comm;       // trailing updates / panel factorisation + broadcast (the same stream in loopback mode)
  cudaEvent_t ev_panel, ev_cols, ev_free[2];
  double gemm_flops;
};
struct DistComm {
  int P;                                             // world size
  virtual ~DistComm() {}
  // every call is made once per phase with the buffers / streams of ALL ranks local to this process (global ranks in `ranks`)
  virtual int bcast_bytes(int root, const int* ranks, void* const* bufs, size_t bytes, cudaStream_t const* st, int n_local) = 0;
  virtual int reduce_sum2(int root, const int* ranks, double* const* a, double* const* b, size_t count, cudaStream_t const* st, int n_local) = 0;   // two vectors, in place at the root
  virtual int allreduce_sum(const int* ranks, double* const* bufs, size_t count, cudaStream_t const* st, int n_local) = 0;
  // point-to-point exchange of the redistribution (processes only): send[q] (ns[q] doubles) to rank q, recv[q] (nr[q]) from rank q, q != own rank
  virtual int exchange(double* const* send, const size_t* ns, double* const* recv, const size_t* nr, cudaStream_t st) { (void)send; (void)ns; (void)recv; (void)nr; (void)st; return -1; }
  virtual const char* last_error() { return ""; }
};
struct DistLU {
  int n, nb, nblk, P; long long lda;
  std::vector<DistRank> r;       // the ranks local to this process
  DistComm* comm;
  float ms_lu, ms_solve;
  int owner_waits_for_panel;     // 1 (default): the owner of the next panel runs its trailing update after that panel, not beside it
};
inline int dist_owner(int blk, int P) { return blk % P; }
inline int dist_local_col(int blk, int P, int nb) { return (blk / P) * nb; }
// number of local columns of rank r (blocks j = r, r+P, ... < nblk; only the globally last block may be narrower)
inline int dist_ncols_local(int n, int nb, int P, int r) {
  const int nblk = (n + nb - 1) / nb; int c = 0;
  for (int j = r; j < nblk; j += P) c += (n - j * nb < nb) ? (n - j * nb) : nb;
  return c;
}
// main_stream: the stream the caller assembles on; separate_comm: panel factorisation + broadcasts on their own high-priority stream
int dist_rank_alloc(DistRank& R, int rank, int n, long long lda, int nb, int P, cudaStream_t main_stream, bool separate_comm);
void dist_rank_free(DistRank& R);
// factorise (the right-hand side column rides along: it ends as y = inv(L) P b on every rank)
int zgetrf_dist(DistLU& D);
// back substitution U x = y; the solution (planar, internal column order) ends in xfin of every rank
int zgetrs_dist(DistLU& D);

// micro-benchmarks (TFLOP/s, GB/s)
double bench_dfma(cudaStream_t st);
double bench_dmma(cudaStream_t st);
double bench_copy(cudaStream_t st);

}  // namespace mfbd
