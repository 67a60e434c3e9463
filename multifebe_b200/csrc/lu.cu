// lu.cu -- blocked right-looking complex LU with partial pivoting on planar (split re/im) FP64 storage, sm_100a.
//
// Replaces OpenBLAS zgetrf + zgetrs behind solve_lse_c (src/solve_lse_c.f90:124,176).  Structure per block column
// (LAPACK zgetrf right-looking variant):
//   k_panel      cooperative persistent kernel: pivot search (izamax semantics: max |re|+|im|, first occurrence),
//                row interchange, reciprocal scaling and rank-1 update of the panel, ONE grid barrier per column
//                (pivot candidates travel together with a copy of their row, so no second barrier is needed)
//   k_laswp      row interchanges outside the panel
//   k_trtri_neg  N = -inv(L11) (unit lower), then U12 = inv(L11) A12 runs as a GEMM on the tensor pipe
//   k_zgemm      trailing update A22 -= A21*U12: the one true dense contraction.  Complex product as 4 real
//                products on the FP64 tensor pipe: mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4; tcgen05 has no
//                FP64 kind, so the accumulators stay in registers), operands staged by cp.async (LDGSTS) through a
//                3-stage shared-memory ring, padded so that fragment loads are bank-conflict free.
#include "lu.cuh"
#include <cooperative_groups.h>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;

namespace mfbd {

// ------------------------------------------------------------------------------------------------------------------
// ZGEMM (C -= A*B), planar complex, column-major
// ------------------------------------------------------------------------------------------------------------------
const int BM = 128, BN = 64, BK = 16, STAGES = 3;
const int SA_LD = BM + 4;   // doubles; (4k + m) mod 16 distinct for the 16 lanes of a half warp
const int SB_LD = BK + 4;
const int SA_STAGE = 2 * BK * SA_LD;  // doubles (re plane, im plane)
const int SB_STAGE = 2 * BN * SB_LD;
const int GEMM_SMEM = STAGES * (SA_STAGE + SB_STAGE) * 8;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 1) k_zgemm_minus(int M, int N, int K, const double* __restrict__ Are, const double* __restrict__ Aim,
                                                        long long lda, const double* __restrict__ Bre, const double* __restrict__ Bim, long long ldb,
                                                        double* __restrict__ Cre, double* __restrict__ Cim, long long ldc) {
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + STAGES * SA_STAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;          // 4 warps along M (32 rows each), 2 along N (32 cols each)
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int gid = lane >> 2, tig = lane & 3;
  const int KT = (K + BK - 1) / BK;

  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * BK;
    double* a = sA + stage * SA_STAGE;
    double* b = sB + stage * SB_STAGE;
#pragma unroll
    for (int i = 0; i < (2 * BK * (BM / 2)) / 256; i++) {
      int idx = tid + 256 * i;
      int p = idx / (BK * (BM / 2)), rem = idx % (BK * (BM / 2)), k = rem / (BM / 2), c2 = rem % (BM / 2);
      int m = m0 + 2 * c2, kk = k0 + k;
      const double* src = (p ? Aim : Are) + (long long)kk * lda + m;
      int bytes = (kk < K) ? max(0, min(16, (M - m) * 8)) : 0;
      if (bytes == 0) src = (p ? Aim : Are);
      cp_async16(a + (p * BK + k) * SA_LD + 2 * c2, src, bytes);
    }
#pragma unroll
    for (int i = 0; i < (2 * BN * (BK / 2)) / 256; i++) {
      int idx = tid + 256 * i;
      int p = idx / (BN * (BK / 2)), rem = idx % (BN * (BK / 2)), nn = rem / (BK / 2), c2 = rem % (BK / 2);
      int n = n0 + nn, kk = k0 + 2 * c2;
      const double* src = (p ? Bim : Bre) + (long long)n * ldb + kk;
      int bytes = (n < N) ? max(0, min(16, (K - kk) * 8)) : 0;
      if (bytes == 0) src = (p ? Bim : Bre);
      cp_async16(b + (p * BN + nn) * SB_LD + 2 * c2, src, bytes);
    }
  };

  // accumulators start from C (C -= A*B is computed as C += A*(-B))
  double cr[4][4][2], ci[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < 4; ni++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int m = m0 + wm * 32 + mi * 8 + gid, n = n0 + wn * 32 + ni * 8 + 2 * tig + h;
        bool ok = (m < M) && (n < N);
        cr[mi][ni][h] = ok ? Cre[(long long)n * ldc + m] : 0.0;
        ci[mi][ni][h] = ok ? Cim[(long long)n * ldc + m] : 0.0;
      }

  for (int s = 0; s < STAGES - 1; s++) { if (s < KT) load_stage(s, s); cp_async_commit(); }
  for (int kt = 0; kt < KT; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    { int nk = kt + STAGES - 1; if (nk < KT) load_stage(nk % STAGES, nk); cp_async_commit(); }
    const double* a = sA + (kt % STAGES) * SA_STAGE;
    const double* b = sB + (kt % STAGES) * SB_STAGE;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; k4++) {
      double ar[4], ai[4], br[4], bi[4], nbr[4], nbi[4];
#pragma unroll
      for (int mi = 0; mi < 4; mi++) {
        int off = (k4 * 4 + tig) * SA_LD + wm * 32 + mi * 8 + gid;
        ar[mi] = a[off]; ai[mi] = a[BK * SA_LD + off];
      }
#pragma unroll
      for (int ni = 0; ni < 4; ni++) {
        int off = (wn * 32 + ni * 8 + gid) * SB_LD + k4 * 4 + tig;
        br[ni] = b[off]; bi[ni] = b[BN * SB_LD + off];
        nbr[ni] = -br[ni]; nbi[ni] = -bi[ni];
      }
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) { dmma(cr[mi][ni][0], cr[mi][ni][1], ar[mi], nbr[ni]); dmma(ci[mi][ni][0], ci[mi][ni][1], ar[mi], nbi[ni]); }
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) { dmma(cr[mi][ni][0], cr[mi][ni][1], ai[mi], bi[ni]); dmma(ci[mi][ni][0], ci[mi][ni][1], ai[mi], nbr[ni]); }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < 4; ni++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int m = m0 + wm * 32 + mi * 8 + gid, n = n0 + wn * 32 + ni * 8 + 2 * tig + h;
        if (m < M && n < N) { Cre[(long long)n * ldc + m] = cr[mi][ni][h]; Cim[(long long)n * ldc + m] = ci[mi][ni][h]; }
      }
}

void zgemm_minus_planar(int m, int n, int k, const double* Are, const double* Aim, long long lda, const double* Bre, const double* Bim,
                        long long ldb, double* Cre, double* Cim, long long ldc, cudaStream_t st) {
  if (m <= 0 || n <= 0 || k <= 0) return;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_zgemm_minus, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM); attr = true; }
  dim3 grid((m + BM - 1) / BM, (n + BN - 1) / BN);
  k_zgemm_minus<<<grid, 256, GEMM_SMEM, st>>>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc);
}

// ------------------------------------------------------------------------------------------------------------------
// Panel factorisation (cooperative, one grid barrier per column)
// ------------------------------------------------------------------------------------------------------------------
struct PanelArgs {
  double *Are, *Aim; long long lda; int n, k0, nbw, rpc;
  int* ipiv; double* cand_val; int* cand_row; double* cand_data; double* diag_data; int* info; int nb;
};

__device__ __forceinline__ void block_argmax(double v, int row, double* s_val, int* s_row, double& best, int& brow) {
  const int tid = threadIdx.x;
  s_val[tid] = v; s_row[tid] = row;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (tid < o) {
      double v2 = s_val[tid + o]; int r2 = s_row[tid + o];
      if (v2 > s_val[tid] || (v2 == s_val[tid] && r2 < s_row[tid])) { s_val[tid] = v2; s_row[tid] = r2; }
    }
    __syncthreads();
  }
  best = s_val[0]; brow = s_row[0];
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_panel(PanelArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double s_val[256]; __shared__ int s_row[256];
  __shared__ double s_ure[256], s_uim[256];
  const int tid = threadIdx.x, c = blockIdx.x, G = gridDim.x;
  const int rs = a.k0 + c * a.rpc, re = min(rs + a.rpc, a.n);
  double* Are = a.Are; double* Aim = a.Aim; const long long lda = a.lda;
  const int nbw = a.nbw, k0 = a.k0;
  const int BIG = 0x7fffffff;

  // candidate for column 0
  {
    double v = -1.0; int r = BIG;
    for (int i = rs + tid; i < re; i += blockDim.x) {
      double t = fabs(Are[(long long)k0 * lda + i]) + fabs(Aim[(long long)k0 * lda + i]);
      if (t > v) { v = t; r = i; }
    }
    double best; int brow; block_argmax(v, r, s_val, s_row, best, brow);
    if (tid == 0) { a.cand_val[c] = best; a.cand_row[c] = brow; }
    if (brow != BIG) for (int jj = tid; jj < nbw; jj += blockDim.x) {
      a.cand_data[((size_t)c) * 2 * a.nb + jj] = Are[(long long)(k0 + jj) * lda + brow];
      a.cand_data[((size_t)c) * 2 * a.nb + a.nb + jj] = Aim[(long long)(k0 + jj) * lda + brow];
    }
    if (k0 >= rs && k0 < re) for (int jj = tid; jj < nbw; jj += blockDim.x) {
      a.diag_data[jj] = Are[(long long)(k0 + jj) * lda + k0]; a.diag_data[a.nb + jj] = Aim[(long long)(k0 + jj) * lda + k0];
    }
  }
  for (int j = 0; j < nbw; j++) {
    const int buf = j & 1, nbuf = buf ^ 1;
    const int dj = k0 + j;   // diagonal row / global column
    __threadfence();
    grid.sync();
    // ---- reduce the G candidates (every CTA does it redundantly) ----
    double v = -1.0; int r = BIG;
    for (int i = tid; i < G; i += blockDim.x) {
      double t = a.cand_val[buf * G + i]; int rr = a.cand_row[buf * G + i];
      if (t > v || (t == v && rr < r)) { v = t; r = rr; }
    }
    double best; int p; block_argmax(v, r, s_val, s_row, best, p);
    const int cstar = (p - k0) / a.rpc;
    const double* cd = a.cand_data + ((size_t)buf * G + cstar) * 2 * a.nb;
    for (int jj = tid; jj < nbw; jj += blockDim.x) { s_ure[jj] = cd[jj]; s_uim[jj] = cd[a.nb + jj]; }
    __syncthreads();
    const double pr = s_ure[j], pi = s_uim[j];
    const bool zero_pivot = (pr == 0.0 && pi == 0.0);
    if (c == 0 && tid == 0) { a.ipiv[dj] = p + 1; if (zero_pivot) atomicCAS(a.info, 0, dj + 1); }
    // ---- row interchange inside the panel ----
    if (p != dj) {
      if (p >= rs && p < re) {   // row p receives the old diagonal row
        const double* dd = a.diag_data + (size_t)buf * 2 * a.nb;
        for (int jj = tid; jj < nbw; jj += blockDim.x) { Are[(long long)(k0 + jj) * lda + p] = dd[jj]; Aim[(long long)(k0 + jj) * lda + p] = dd[a.nb + jj]; }
      }
      if (dj >= rs && dj < re) { // diagonal row receives the pivot row
        for (int jj = tid; jj < nbw; jj += blockDim.x) { Are[(long long)(k0 + jj) * lda + dj] = s_ure[jj]; Aim[(long long)(k0 + jj) * lda + dj] = s_uim[jj]; }
      }
    }
    __syncthreads();
    // ---- scale column j and rank-1 update of the rest of the panel; fused pivot search for column j+1 ----
    double ir = 0.0, ii = 0.0;
    if (!zero_pivot) {  // reciprocal 1/pivot (zgetf2 scales by the reciprocal), Smith's algorithm
      if (fabs(pr) >= fabs(pi)) { double t = pi / pr, d = pr + pi * t; ir = 1.0 / d; ii = -t / d; }
      else { double t = pr / pi, d = pr * t + pi; ir = t / d; ii = -1.0 / d; }
    }
    double nv = -1.0; int nr = BIG;
    for (int i = rs + tid; i < re; i += blockDim.x) {
      if (i <= dj) continue;
      double lr = Are[(long long)dj * lda + i], li = Aim[(long long)dj * lda + i];
      if (!zero_pivot) { double t = lr * ir - li * ii; li = lr * ii + li * ir; lr = t; Are[(long long)dj * lda + i] = lr; Aim[(long long)dj * lda + i] = li; }
      for (int jj = j + 1; jj < nbw; jj++) {
        long long o = (long long)(k0 + jj) * lda + i;
        double xr = Are[o], xi = Aim[o];
        xr -= lr * s_ure[jj] - li * s_uim[jj];
        xi -= lr * s_uim[jj] + li * s_ure[jj];
        Are[o] = xr; Aim[o] = xi;
        if (jj == j + 1) { double t = fabs(xr) + fabs(xi); if (t > nv) { nv = t; nr = i; } }
      }
    }
    if (j + 1 < nbw) {
      double nbest; int nbrow; block_argmax(nv, nr, s_val, s_row, nbest, nbrow);   // includes the __syncthreads that orders the updates
      if (tid == 0) { a.cand_val[nbuf * G + c] = nbest; a.cand_row[nbuf * G + c] = nbrow; }
      if (nbrow != BIG) for (int jj = tid; jj < nbw; jj += blockDim.x) {
        a.cand_data[((size_t)nbuf * G + c) * 2 * a.nb + jj] = Are[(long long)(k0 + jj) * lda + nbrow];
        a.cand_data[((size_t)nbuf * G + c) * 2 * a.nb + a.nb + jj] = Aim[(long long)(k0 + jj) * lda + nbrow];
      }
      const int nd = dj + 1;
      if (nd >= rs && nd < re) for (int jj = tid; jj < nbw; jj += blockDim.x) {
        a.diag_data[(size_t)nbuf * 2 * a.nb + jj] = Are[(long long)(k0 + jj) * lda + nd];
        a.diag_data[(size_t)nbuf * 2 * a.nb + a.nb + jj] = Aim[(long long)(k0 + jj) * lda + nd];
      }
    }
  }
}

// row interchanges of one block step applied to columns [c0,c1) (outside the panel)
__global__ void k_laswp(double* Are, double* Aim, long long lda, int c0, int c1, int k0, int nbw, const int* __restrict__ ipiv) {
  int col = c0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= c1) return;
  double* ar = Are + (long long)col * lda; double* ai = Aim + (long long)col * lda;
  for (int j = 0; j < nbw; j++) {
    int p = ipiv[k0 + j] - 1, d = k0 + j;
    if (p != d) { double t = ar[d]; ar[d] = ar[p]; ar[p] = t; t = ai[d]; ai[d] = ai[p]; ai[p] = t; }
  }
}

// N = -inv(L11), L11 = unit lower nbw x nbw block at (k0,k0).  Thread c owns column c of the inverse.
__global__ void k_trtri_neg(const double* __restrict__ Are, const double* __restrict__ Aim, long long lda, int k0, int nbw,
                            double* xt_re, double* xt_im, double* nre, double* nim, int ldn) {
  int c = threadIdx.x;
  if (c < nbw) {
    // xt[k*ldn + c] = X(k,c) (coalesced scratch)
    for (int i = 0; i < nbw; i++) { xt_re[i * ldn + c] = (i == c) ? 1.0 : 0.0; xt_im[i * ldn + c] = 0.0; }
    for (int i = c + 1; i < nbw; i++) {
      double sr = 0.0, si = 0.0;
      for (int k = c; k < i; k++) {
        double lr = Are[(long long)(k0 + k) * lda + k0 + i], li = Aim[(long long)(k0 + k) * lda + k0 + i];
        double xr = xt_re[k * ldn + c], xi = xt_im[k * ldn + c];
        sr += lr * xr - li * xi; si += lr * xi + li * xr;
      }
      xt_re[i * ldn + c] = -sr; xt_im[i * ldn + c] = -si;
    }
    for (int i = 0; i < nbw; i++) { nre[(long long)c * ldn + i] = -xt_re[i * ldn + c]; nim[(long long)c * ldn + i] = -xt_im[i * ldn + c]; }
  }
}

int lu_work_alloc(LuWork& w, int n, int nb) {
  w.nb = nb;
  cudaDeviceProp prop; int dev; cudaGetDevice(&dev); cudaGetDeviceProperties(&prop, dev);
  w.n_sm = prop.multiProcessorCount;
  size_t G = (size_t)w.n_sm;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
  A((void**)&w.cand_val, 2 * G * sizeof(double)); A((void**)&w.cand_row, 2 * G * sizeof(int));
  A((void**)&w.cand_data, 2 * G * 2 * nb * sizeof(double)); A((void**)&w.diag_data, 2 * 2 * nb * sizeof(double));
  A((void**)&w.ninv_re, (size_t)4 * nb * nb * sizeof(double));   // N re, N im, scratch re, scratch im
  w.ninv_im = w.ninv_re + (size_t)nb * nb;
  w.ldt = nb;
  A((void**)&w.t_re, (size_t)2 * nb * (size_t)n * sizeof(double));
  w.t_im = w.t_re + (size_t)nb * n;
  A((void**)&w.info, sizeof(int));
  w.n_evs = 5 * ((n + nb - 1) / nb);
  w.evs = new cudaEvent_t[w.n_evs];
  for (int i = 0; i < w.n_evs; i++) cudaEventCreate(&w.evs[i]);
  w.ms_panel = w.ms_swap = w.ms_trsm = w.ms_gemm = 0.f; w.n_steps_timed = 0; w.gemm_launches = 0; w.gemm_flops = 0.0;
  return (int)e;
}
void lu_work_free(LuWork& w) {
  cudaFree(w.cand_val); cudaFree(w.cand_row); cudaFree(w.cand_data); cudaFree(w.diag_data); cudaFree(w.ninv_re); cudaFree(w.t_re); cudaFree(w.info);
  for (int i = 0; i < w.n_evs; i++) cudaEventDestroy(w.evs[i]);
  delete[] w.evs;
}
void lu_collect_times(LuWork& w) {
  w.ms_panel = w.ms_swap = w.ms_trsm = w.ms_gemm = 0.f;
  for (int s = 0; s < w.n_steps_timed; s++) {
    cudaEvent_t* ev = w.evs + 5 * s; float t;
    cudaEventElapsedTime(&t, ev[0], ev[1]); w.ms_panel += t;
    cudaEventElapsedTime(&t, ev[1], ev[2]); w.ms_swap += t;
    cudaEventElapsedTime(&t, ev[2], ev[3]); w.ms_trsm += t;
    cudaEventElapsedTime(&t, ev[3], ev[4]); w.ms_gemm += t;
  }
}

int zgetrf_planar(double* Are, double* Aim, long long lda, int n, int* ipiv, LuWork& w, cudaStream_t st, bool timing) {
  const int nb = w.nb;
  cudaMemsetAsync(w.info, 0, sizeof(int), st);
  w.launches = 0; w.gemm_launches = 0; w.gemm_flops = 0.0; w.n_steps_timed = 0;
  for (int k0 = 0; k0 < n; k0 += nb) {
    cudaEvent_t* ev = w.evs + 5 * (k0 / nb);
    const int nbw = (n - k0 < nb) ? (n - k0) : nb;
    const int m = n - k0;
    if (timing) cudaEventRecord(ev[0], st);
    // ---- panel ----
    int G = w.n_sm;
    int rpc = (m + G - 1) / G; if (rpc < 32) rpc = 32;
    G = (m + rpc - 1) / rpc;
    PanelArgs pa; pa.Are = Are; pa.Aim = Aim; pa.lda = lda; pa.n = n; pa.k0 = k0; pa.nbw = nbw; pa.rpc = rpc; pa.ipiv = ipiv;
    pa.cand_val = w.cand_val; pa.cand_row = w.cand_row; pa.cand_data = w.cand_data; pa.diag_data = w.diag_data; pa.info = w.info; pa.nb = nb;
    void* args[] = {&pa};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_panel, dim3(G), dim3(256), args, 0, st);
    if (e != cudaSuccess) return (int)e;
    w.launches += 1 + (k0 > 0) + (n - k0 - nbw > 0) * 4;
    if (timing) cudaEventRecord(ev[1], st);
    // ---- interchanges outside the panel ----
    if (k0 > 0) k_laswp<<<(k0 + 127) / 128, 128, 0, st>>>(Are, Aim, lda, 0, k0, k0, nbw, ipiv);
    const int nrest = n - k0 - nbw;
    if (nrest > 0) k_laswp<<<(nrest + 127) / 128, 128, 0, st>>>(Are, Aim, lda, k0 + nbw, n, k0, nbw, ipiv);
    if (timing) cudaEventRecord(ev[2], st);
    if (nrest > 0) {
      // ---- U12 = inv(L11) * A12 as T = 0 - N*A12 with N = -inv(L11) ----
      double* sc_re = w.ninv_re + (size_t)2 * nb * nb; double* sc_im = sc_re + (size_t)nb * nb;
      k_trtri_neg<<<1, nb, 0, st>>>(Are, Aim, lda, k0, nbw, sc_re, sc_im, w.ninv_re, w.ninv_im, nb);
      cudaMemsetAsync(w.t_re, 0, (size_t)nb * (size_t)nrest * sizeof(double), st);
      cudaMemsetAsync(w.t_im, 0, (size_t)nb * (size_t)nrest * sizeof(double), st);
      const double* B_re = Are + (long long)(k0 + nbw) * lda + k0; const double* B_im = Aim + (long long)(k0 + nbw) * lda + k0;
      zgemm_minus_planar(nbw, nrest, nbw, w.ninv_re, w.ninv_im, nb, B_re, B_im, lda, w.t_re, w.t_im, w.ldt, st);
      cudaMemcpy2DAsync((void*)B_re, lda * 8, w.t_re, w.ldt * 8, (size_t)nbw * 8, nrest, cudaMemcpyDeviceToDevice, st);
      cudaMemcpy2DAsync((void*)B_im, lda * 8, w.t_im, w.ldt * 8, (size_t)nbw * 8, nrest, cudaMemcpyDeviceToDevice, st);
      if (timing) cudaEventRecord(ev[3], st);
      // ---- trailing update A22 -= A21 * U12 ----
      zgemm_minus_planar(nrest, nrest, nbw, Are + (long long)k0 * lda + k0 + nbw, Aim + (long long)k0 * lda + k0 + nbw, lda, B_re, B_im, lda,
                         Are + (long long)(k0 + nbw) * lda + k0 + nbw, Aim + (long long)(k0 + nbw) * lda + k0 + nbw, lda, st);
      w.gemm_launches += 1; w.gemm_flops += 8.0 * (double)nrest * (double)nrest * (double)nbw;
    } else if (timing) cudaEventRecord(ev[3], st);
    if (timing) { cudaEventRecord(ev[4], st); w.n_steps_timed++; }
  }
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// Triangular solves (zgetrs, 'N'): b := P b ; L y = b (unit lower) ; U x = y.  Blocked by TS rows.
// ------------------------------------------------------------------------------------------------------------------
const int TS = 128;
__global__ void k_permute(const double* __restrict__ sre, const double* __restrict__ sim, double* dre, double* dim_, const int* __restrict__ perm, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { dre[i] = sre[perm[i]]; dim_[i] = sim[perm[i]]; }
}
// diagonal block solve, one CTA of TS threads, one right-hand side
__global__ void k_trsv_diag(const double* __restrict__ Are, const double* __restrict__ Aim, long long lda, int kb, int nbw, double* bre, double* bim, int lower) {
  __shared__ double yr[TS], yi[TS];
  int i = threadIdx.x;
  double vr = 0.0, vi = 0.0;
  if (i < nbw) { vr = bre[kb + i]; vi = bim[kb + i]; }
  if (lower) {
    for (int j = 0; j < nbw; j++) {
      if (i == j) { yr[j] = vr; yi[j] = vi; }
      __syncthreads();
      if (i > j && i < nbw) {
        double lr = Are[(long long)(kb + j) * lda + kb + i], li = Aim[(long long)(kb + j) * lda + kb + i];
        vr -= lr * yr[j] - li * yi[j]; vi -= lr * yi[j] + li * yr[j];
      }
    }
  } else {
    for (int j = nbw - 1; j >= 0; j--) {
      if (i == j) {
        double ur = Are[(long long)(kb + j) * lda + kb + j], ui = Aim[(long long)(kb + j) * lda + kb + j];
        double qr, qi;   // (vr + i vi)/(ur + i ui), Smith
        if (fabs(ur) >= fabs(ui)) { double t = ui / ur, d = ur + ui * t; qr = (vr + vi * t) / d; qi = (vi - vr * t) / d; }
        else { double t = ur / ui, d = ur * t + ui; qr = (vr * t + vi) / d; qi = (vi * t - vr) / d; }
        vr = qr; vi = qi; yr[j] = vr; yi[j] = vi;
      }
      __syncthreads();
      if (i < j) {
        double ur = Are[(long long)(kb + j) * lda + kb + i], ui = Aim[(long long)(kb + j) * lda + kb + i];
        vr -= ur * yr[j] - ui * yi[j]; vi -= ur * yi[j] + ui * yr[j];
      }
    }
  }
  if (i < nbw) { bre[kb + i] = vr; bim[kb + i] = vi; }
}
// b[r0:r1) -= A[r0:r1, kb:kb+nbw) * x[kb:kb+nbw)
__global__ void k_gemv_update(const double* __restrict__ Are, const double* __restrict__ Aim, long long lda, int r0, int r1, int kb, int nbw, double* bre, double* bim) {
  __shared__ double xr[TS], xi[TS];
  if (threadIdx.x < nbw) { xr[threadIdx.x] = bre[kb + threadIdx.x]; xi[threadIdx.x] = bim[kb + threadIdx.x]; }
  __syncthreads();
  int i = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r1) return;
  double sr = 0.0, si = 0.0;
#pragma unroll 4
  for (int j = 0; j < nbw; j++) {
    double ar = Are[(long long)(kb + j) * lda + i], ai = Aim[(long long)(kb + j) * lda + i];
    sr += ar * xr[j] - ai * xi[j]; si += ar * xi[j] + ai * xr[j];
  }
  bre[i] -= sr; bim[i] -= si;
}

int zgetrs_planar(const double* Are, const double* Aim, long long lda, int n, const int* ipiv_host_perm_dev, double* bre, double* bim, long long ldb,
                  int nrhs, cudaStream_t st) {
  // ipiv_host_perm_dev: device array perm[i] = source row of row i after all interchanges (built on the host from ipiv)
  double* tmp = nullptr;
  if (cudaMalloc((void**)&tmp, (size_t)2 * n * sizeof(double)) != cudaSuccess) return (int)cudaGetLastError();
  for (int c = 0; c < nrhs; c++) {
    double* br = bre + (long long)c * ldb; double* bi = bim + (long long)c * ldb;
    k_permute<<<(n + 255) / 256, 256, 0, st>>>(br, bi, tmp, tmp + n, ipiv_host_perm_dev, n);
    cudaMemcpyAsync(br, tmp, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(bi, tmp + n, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
    for (int kb = 0; kb < n; kb += TS) {
      int nbw = (n - kb < TS) ? n - kb : TS;
      k_trsv_diag<<<1, TS, 0, st>>>(Are, Aim, lda, kb, nbw, br, bi, 1);
      int r0 = kb + nbw;
      if (r0 < n) k_gemv_update<<<(n - r0 + TS - 1) / TS, TS, 0, st>>>(Are, Aim, lda, r0, n, kb, nbw, br, bi);
    }
    int nblk = (n + TS - 1) / TS;
    for (int b = nblk - 1; b >= 0; b--) {
      int kb = b * TS, nbw = (n - kb < TS) ? n - kb : TS;
      k_trsv_diag<<<1, TS, 0, st>>>(Are, Aim, lda, kb, nbw, br, bi, 0);
      if (kb > 0) k_gemv_update<<<(kb + TS - 1) / TS, TS, 0, st>>>(Are, Aim, lda, 0, kb, kb, nbw, br, bi);
    }
  }
  cudaStreamSynchronize(st);
  cudaFree(tmp);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// Roofline denominators measured on the box: FP64 FMA pipe, FP64 tensor pipe (DMMA), device copy
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_bench_dfma(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_bench_dmma(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
static double time_kernel(cudaStream_t st, void (*launch)(cudaStream_t)) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(st); cudaStreamSynchronize(st);
  float best = 1e30f;
  for (int r = 0; r < 3; r++) { cudaEventRecord(e0, st); launch(st); cudaEventRecord(e1, st); cudaEventSynchronize(e1); float t; cudaEventElapsedTime(&t, e0, e1); if (t < best) best = t; }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return best;
}
static double* g_bench_buf = nullptr;
static void launch_dfma(cudaStream_t st) { k_bench_dfma<<<148 * 8, 256, 0, st>>>(g_bench_buf, 20000); }
static void launch_dmma(cudaStream_t st) { k_bench_dmma<<<148 * 8, 256, 0, st>>>(g_bench_buf, 4000); }
double bench_dfma(cudaStream_t st) {
  cudaMalloc((void**)&g_bench_buf, (size_t)148 * 8 * 256 * 8);
  double ms = time_kernel(st, launch_dfma);
  cudaFree(g_bench_buf);
  return (148.0 * 8 * 256 * 20000.0 * 8 * 2) / (ms * 1e-3) / 1e12;
}
double bench_dmma(cudaStream_t st) {
  cudaMalloc((void**)&g_bench_buf, (size_t)148 * 8 * 256 * 8);
  double ms = time_kernel(st, launch_dmma);
  cudaFree(g_bench_buf);
  return (148.0 * 8 * 8 /*warps*/ * 4000.0 * 8 * (8 * 8 * 4 * 2)) / (ms * 1e-3) / 1e12;
}
double bench_copy(cudaStream_t st) {
  const size_t bytes = (size_t)2 << 30;
  char *a = nullptr, *b = nullptr;
  if (cudaMalloc((void**)&a, bytes) != cudaSuccess || cudaMalloc((void**)&b, bytes) != cudaSuccess) { cudaFree(a); return 0.0; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, st);
  float best = 1e30f;
  for (int r = 0; r < 5; r++) { cudaEventRecord(e0, st); cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, st); cudaEventRecord(e1, st); cudaEventSynchronize(e1); float t; cudaEventElapsedTime(&t, e0, e1); if (t < best) best = t; }
  cudaFree(a); cudaFree(b); cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 2.0 * bytes / (best * 1e-3) / 1e9;
}

}  // namespace mfbd
