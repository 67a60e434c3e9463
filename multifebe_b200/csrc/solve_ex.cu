// solve_ex.cu -- the OPTIONAL stages of solve_lse_c (src/solve_lse_c.f90:81-117 scaling = zgeequ + zlaqge, :140-165 condition = zgecon,
// :191-206 refine = zgerfs) on planar device storage, around the LU of lu.cu.  The reference delegates them to LAPACK; here they are written
// from LAPACK's published definitions:
//   equilibration  r_i = 1 / max_j |a_ij|_1, c_j = 1 / max_i r_i |a_ij|_1 (|z|_1 = |re| + |im|), applied when the spread of the factors is
//                  below 0.1 or the largest entry is outside [small, large]                                    (zgeequ / zlaqge)
//   condition      rcond = 1 / (|A|_1 |inv(A)|_1), |inv(A)|_1 by Hager's estimator in Higham's form (zlacn2): solves with A and A^H
//   refinement     x += inv(A) (b - A x) while the componentwise backward error  max_i |r_i| / (|A||x| + |b|)_i  halves, at most 5 times;
//                  forward error bound  | inv(A) (|r| + (n + 1) eps (|A||x| + |b|)) |_inf / |x|_inf  with the same estimator     (zgerfs)
// Solves with A^H use the factors of P A = L U in dot-product (left-looking) form: U^H y = v forward, L^H z = y backward, x = P^T z.  Column i
// of a factor is contiguous in memory, so the dot products stream it with one warp per column.
#include "solve_ex.cuh"
#include <algorithm>
#include <cmath>
#include <cfloat>
#include <cstdio>
#include <vector>

namespace mfbd {

static const int TSX = 64;   // diagonal block of the triangular solves (== TS of lu.cu: layout of LuWork::inv)

// ---- equilibration -----------------------------------------------------------------------------------------------------------
__global__ void k_row_amax(const double* __restrict__ re, const double* __restrict__ im, long long ld, int n, int cchunk, const double* __restrict__ rs, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c0 = blockIdx.y * cchunk, c1 = min(c0 + cchunk, n);
  double m = 0.0;
  for (int j = c0; j < c1; j++) m = fmax(m, fabs(re[(size_t)j * ld + i]) + fabs(im[(size_t)j * ld + i]));
  // non-negative doubles order like their bit patterns: atomicMax on the integer view
  atomicMax(reinterpret_cast<unsigned long long*>(out) + i, (unsigned long long)__double_as_longlong(m));
  (void)rs;
}
__global__ void k_col_amax(const double* __restrict__ re, const double* __restrict__ im, long long ld, int n, const double* __restrict__ r, double* out) {
  // one warp per column: max_i r_i |a_ij|_1
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= n) return;
  double m = 0.0;
  for (int i = lane; i < n; i += 32) m = fmax(m, r[i] * (fabs(re[(size_t)j * ld + i]) + fabs(im[(size_t)j * ld + i])));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) out[j] = m;
}
__global__ void k_scale_rc(double* re, double* im, long long ld, int n, const double* __restrict__ r, const double* __restrict__ c) {
  const long long total = (long long)n * n;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t / n), i = (int)(t - (long long)j * n);
    const double f = (r ? r[i] : 1.0) * (c ? c[j] : 1.0);
    re[(size_t)j * ld + i] *= f; im[(size_t)j * ld + i] *= f;
  }
}
__global__ void k_scale_vec(double* re, double* im, int n, const double* __restrict__ f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { re[i] *= f[i]; im[i] *= f[i]; }
}

int zequilibrate(double* Are, double* Aim, long long lda, int n, double* d_r, double* d_c, std::vector<double>& r, std::vector<double>& c, double* rowcnd, double* colcnd,
                 double* amax, char* equed, int* info, cudaStream_t st) {
  const double smlnum = DBL_MIN, bignum = 1.0 / smlnum;
  *info = 0; *equed = 'N';
  r.assign(n, 0.0); c.assign(n, 0.0);
  cudaMemsetAsync(d_r, 0, (size_t)n * 8, st);
  const int cchunk = 1024;
  k_row_amax<<<dim3((n + 255) / 256, (n + cchunk - 1) / cchunk), 256, 0, st>>>(Are, Aim, lda, n, cchunk, nullptr, d_r);
  cudaMemcpyAsync(r.data(), d_r, (size_t)n * 8, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
  double rcmin = bignum, rcmax = 0.0;
  for (int i = 0; i < n; i++) { rcmax = std::max(rcmax, r[i]); rcmin = std::min(rcmin, r[i]); }
  *amax = rcmax;
  if (rcmin == 0.0) { for (int i = 0; i < n; i++) if (r[i] == 0.0) { *info = i + 1; return 0; } }
  for (int i = 0; i < n; i++) r[i] = 1.0 / std::min(std::max(r[i], smlnum), bignum);
  *rowcnd = std::max(rcmin, smlnum) / std::min(rcmax, bignum);
  cudaMemcpyAsync(d_r, r.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st);
  k_col_amax<<<(n + 7) / 8, 256, 0, st>>>(Are, Aim, lda, n, d_r, d_c);
  cudaMemcpyAsync(c.data(), d_c, (size_t)n * 8, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
  rcmin = bignum; rcmax = 0.0;
  for (int j = 0; j < n; j++) { rcmax = std::max(rcmax, c[j]); rcmin = std::min(rcmin, c[j]); }
  if (rcmin == 0.0) { for (int j = 0; j < n; j++) if (c[j] == 0.0) { *info = n + j + 1; return 0; } }
  for (int j = 0; j < n; j++) c[j] = 1.0 / std::min(std::max(c[j], smlnum), bignum);
  *colcnd = std::max(rcmin, smlnum) / std::min(rcmax, bignum);
  cudaMemcpyAsync(d_c, c.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st);
  // zlaqge: thresholds
  const double thresh = 0.1, small = DBL_MIN / (DBL_EPSILON * 0.5), large = 1.0 / small;
  const bool rows = !(*rowcnd >= thresh && *amax >= small && *amax <= large), cols = !(*colcnd >= thresh);
  if (rows || cols) k_scale_rc<<<2368, 256, 0, st>>>(Are, Aim, lda, n, rows ? d_r : nullptr, cols ? d_c : nullptr);
  *equed = rows ? (cols ? 'B' : 'R') : (cols ? 'C' : 'N');
  return (int)cudaGetLastError();
}
void scale_vector(double* re, double* im, int n, const double* d_f, cudaStream_t st) { k_scale_vec<<<(n + 255) / 256, 256, 0, st>>>(re, im, n, d_f); }

// ---- 1-norm of the (planar) matrix: max_j sum_i |a_ij| (true modulus, as solve_lse_c.f90:143-150) ----------------------------------
__global__ void k_col_sum_abs(const double* __restrict__ re, const double* __restrict__ im, long long ld, int n, double* out) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= n) return;
  double s = 0.0;
  for (int i = lane; i < n; i += 32) s += hypot(re[(size_t)j * ld + i], im[(size_t)j * ld + i]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[j] = s;
}
double matrix_norm1(const double* re, const double* im, long long ld, int n, double* d_tmp, cudaStream_t st) {
  k_col_sum_abs<<<(n + 7) / 8, 256, 0, st>>>(re, im, ld, n, d_tmp);
  std::vector<double> h(n); cudaMemcpyAsync(h.data(), d_tmp, (size_t)n * 8, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
  double m = 0.0; for (double v : h) m = std::max(m, v);
  return m;
}

// ---- conjugate-transposed solves with the factors -----------------------------------------------------------------------------------
// w_c = v_c - sum_{j in [j0, j1)} conj(A[j, c]) y_j for the columns c of one diagonal block; one warp per column
__global__ void __launch_bounds__(256) k_tdot(const double* __restrict__ Are, const double* __restrict__ Aim, long long lda, int c0, int nc, int j0, int j1,
                                               const double* __restrict__ yre, const double* __restrict__ yim, const double* __restrict__ vre, const double* __restrict__ vim,
                                               double* wre, double* wim) {
  const int cc = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (cc >= nc) return;
  const double* ar = Are + (size_t)(c0 + cc) * lda; const double* ai = Aim + (size_t)(c0 + cc) * lda;
  double sr = 0.0, si = 0.0;
  for (int j = j0 + lane; j < j1; j += 32) { const double a = ar[j], b = -ai[j], yr = yre[j], yi = yim[j]; sr += a * yr - b * yi; si += a * yi + b * yr; }
  for (int o = 16; o > 0; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); si += __shfl_xor_sync(0xffffffffu, si, o); }
  if (lane == 0) { wre[cc] = vre[c0 + cc] - sr; wim[cc] = vim[c0 + cc] - si; }
}
// y[kb .. kb + nbw) = inv(D)^H w with the precomputed inverse of the diagonal block (layout of lu_invert_diagonal_blocks: o[col * TS + row])
__global__ void __launch_bounds__(64) k_tdiag(const double* __restrict__ invb, int kb, int nbw, const double* __restrict__ wre, const double* __restrict__ wim, double* yre, double* yim) {
  const int j = threadIdx.x;
  if (j >= nbw) return;
  const double* xr = invb; const double* xi = invb + TSX * TSX;
  double sr = 0.0, si = 0.0;
  for (int c = 0; c < nbw; c++) {   // (inv^H)[j][c] = conj(inv[c][j]) = conj(o[j * TS + c])
    const double a = xr[j * TSX + c], b = -xi[j * TSX + c];
    sr += a * wre[c] - b * wim[c]; si += a * wim[c] + b * wre[c];
  }
  yre[kb + j] = sr; yim[kb + j] = si;
}
__global__ void k_unpermute(const double* __restrict__ sre, const double* __restrict__ sim, double* dre, double* dim_, const int* __restrict__ perm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // forward: d[i] = s[perm[i]]; inverse: d[perm[i]] = s[i]
  if (i < n) { dre[perm[i]] = sre[i]; dim_[perm[i]] = sim[i]; }
}
// x := inv(A)^H x with P A = L U (perm[i] = source row of row i, as zgetrs_planar takes it); tmp: 4 n + 4 * 64 doubles of device scratch
int zgetrs_conjtrans_planar(const double* Are, const double* Aim, long long lda, int n, const int* d_perm, const double* inv, double* xre, double* xim, double* tmp, cudaStream_t st) {
  if (!inv) return -1;
  double *yre = tmp, *yim = tmp + n, *zre = tmp + 2 * (size_t)n, *zim = tmp + 3 * (size_t)n, *wre = tmp + 4 * (size_t)n, *wim = wre + TSX;
  const int nblk = (n + TSX - 1) / TSX;
  for (int b = 0; b < nblk; b++) {              // U^H y = x
    const int kb = b * TSX, nbw = std::min(TSX, n - kb);
    k_tdot<<<(nbw + 7) / 8, 256, 0, st>>>(Are, Aim, lda, kb, nbw, 0, kb, yre, yim, xre, xim, wre, wim);
    k_tdiag<<<1, 64, 0, st>>>(inv + ((size_t)b * 2 + 1) * 2 * TSX * TSX, kb, nbw, wre, wim, yre, yim);
  }
  for (int b = nblk - 1; b >= 0; b--) {         // L^H z = y
    const int kb = b * TSX, nbw = std::min(TSX, n - kb);
    k_tdot<<<(nbw + 7) / 8, 256, 0, st>>>(Are, Aim, lda, kb, nbw, kb + nbw, n, zre, zim, yre, yim, wre, wim);
    k_tdiag<<<1, 64, 0, st>>>(inv + ((size_t)b * 2 + 0) * 2 * TSX * TSX, kb, nbw, wre, wim, zre, zim);
  }
  k_unpermute<<<(n + 255) / 256, 256, 0, st>>>(zre, zim, xre, xim, d_perm, n);   // x = P^T z
  return (int)cudaGetLastError();
}

// ---- Hager / Higham 1-norm estimator (the algorithm of LAPACK's zlacn2) driven from the host --------------------------------------------
// apply(v, conj_transposed): v := B v or B^H v for the operator whose 1-norm is wanted
double norm1_estimate(int n, const std::function<void(std::vector<std::complex<double>>&, bool)>& apply) {
  typedef std::complex<double> cd;
  const double safmin = DBL_MIN;
  std::vector<cd> x(n, cd(1.0 / n, 0.0));
  apply(x, false);
  if (n == 1) return std::abs(x[0]);
  double est = 0.0; for (auto& v : x) est += std::abs(v);
  for (auto& v : x) { const double a = std::abs(v); v = a > safmin ? v / a : cd(1.0, 0.0); }
  apply(x, true);
  auto imax = [&]() { int j = 0; double m = -1.0; for (int i = 0; i < n; i++) { const double a = std::abs(x[i]); if (a > m) { m = a; j = i; } } return j; };   // izmax1: true modulus
  int j = imax();
  for (int iter = 2; iter <= 5; iter++) {
    std::fill(x.begin(), x.end(), cd(0.0, 0.0)); x[j] = cd(1.0, 0.0);
    apply(x, false);
    const double estold = est;
    est = 0.0; for (auto& v : x) est += std::abs(v);
    if (est <= estold) break;
    for (auto& v : x) { const double a = std::abs(v); v = a > safmin ? v / a : cd(1.0, 0.0); }
    apply(x, true);
    const int jlast = j; j = imax();
    if (std::abs(x[jlast]) == std::abs(x[j])) break;
  }
  // alternating-sign test vector
  double altsgn = 1.0;
  for (int i = 0; i < n; i++) { x[i] = cd(altsgn * (1.0 + (double)i / (double)(n - 1)), 0.0); altsgn = -altsgn; }
  apply(x, false);
  double temp = 0.0; for (auto& v : x) temp += std::abs(v);
  temp = 2.0 * (temp / (3.0 * n));
  return std::max(est, temp);
}

}  // namespace mfbd
