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
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
namespace cg = cooperative_groups;

namespace mfbd {

// A NULL imaginary plane means a REAL matrix (static path, solve_lse_r): offsets must keep it NULL.
__host__ __device__ inline double* poff(double* p, long long o) { return p ? p + o : nullptr; }
__host__ __device__ inline const double* poff(const double* p, long long o) { return p ? p + o : nullptr; }

// ------------------------------------------------------------------------------------------------------------------
// ZGEMM (C -= A*B), planar complex, column-major.  CTA tile (32*WM) x (32*WN), one 32x32 warp tile per warp
// (4 x 4 DMMA.8x8x4 tiles, re and im accumulators in registers), BK-deep k-tiles through a cp.async ring.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int WM, int WN, int BK, int STAGES>
struct GemmCfg {
  static const int BM = 32 * WM, BN = 32 * WN, T = 32 * WM * WN;
  static const int SA_LD = BM + 4;   // doubles; (4k + m) mod 16 distinct for the 16 lanes of a half warp
  static const int SB_LD = BK + 4;
  static const int SA_STAGE = 2 * BK * SA_LD;  // doubles (re plane, im plane)
  static const int SB_STAGE = 2 * BN * SB_LD;
  static const int SMEM = STAGES * (SA_STAGE + SB_STAGE) * 8;
};

template <int WM, int WN, int BK, int STAGES, int MINB>
__global__ void __launch_bounds__(32 * WM * WN, MINB)
k_zgemm_minus(int M, int N, int K, const double* __restrict__ Are, const double* __restrict__ Aim, long long lda, const double* __restrict__ Bre,
              const double* __restrict__ Bim, long long ldb, double* __restrict__ Cre, double* __restrict__ Cim, long long ldc) {
  typedef GemmCfg<WM, WN, BK, STAGES> C;
  constexpr int BM = C::BM, BN = C::BN, T = C::T, SA_LD = C::SA_LD, SB_LD = C::SB_LD, SA_STAGE = C::SA_STAGE, SB_STAGE = C::SB_STAGE;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + STAGES * SA_STAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int gid = lane >> 2, tig = lane & 3;
  const int KT = (K + BK - 1) / BK;

  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * BK;
    double* a = sA + stage * SA_STAGE;
    double* b = sB + stage * SB_STAGE;
#pragma unroll
    for (int i = 0; i < (2 * BK * (BM / 2)) / T; i++) {
      int idx = tid + T * i;
      int p = idx / (BK * (BM / 2)), rem = idx % (BK * (BM / 2)), k = rem / (BM / 2), c2 = rem % (BM / 2);
      int m = m0 + 2 * c2, kk = k0 + k;
      const double* src = (p ? Aim : Are) + (long long)kk * lda + m;
      int bytes = (kk < K) ? max(0, min(16, (M - m) * 8)) : 0;
      if (bytes == 0) src = (p ? Aim : Are);
      cp_async16(a + (p * BK + k) * SA_LD + 2 * c2, src, bytes);
    }
#pragma unroll
    for (int i = 0; i < (2 * BN * (BK / 2)) / T; i++) {
      int idx = tid + T * i;
      int p = idx / (BN * (BK / 2)), rem = idx % (BN * (BK / 2)), nn = rem / (BK / 2), c2 = rem % (BK / 2);
      int n = n0 + nn, kk = k0 + 2 * c2;
      const double* src = (p ? Bim : Bre) + (long long)n * ldb + kk;
      int bytes = (n < N) ? max(0, min(16, (K - kk) * 8)) : 0;
      if (bytes == 0) src = (p ? Bim : Bre);
      cp_async16(b + (p * BN + nn) * SB_LD + 2 * c2, src, bytes);
    }
  };

  for (int s = 0; s < STAGES - 1; s++) { if (s < KT) load_stage(s, s); cp_async_commit(); }

  // accumulators start from C (C -= A*B is computed as C += A*(-B)); these loads overlap the first k-tiles in flight
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

// ------------------------------------------------------------------------------------------------------------------
// ZGEMM, 3M form (C -= A*B with three real products instead of four, the ZGEMM3M of the BLAS):
//   X = Ar*Br, Y = Ai*Bi, Z = (Ar+Ai)*(Br+Bi);  Cr -= X - Y;  Ci -= Z - X - Y.
// 25 % fewer DMMA per complex product; the sums Ar+Ai, Br+Bi are formed on the fragments (one DADD per fragment element).
// Normwise as accurate as the 4-product form (the BLAS offers it for exactly this use); the imaginary part loses the
// componentwise bound, which an LU with partial pivoting does not rely on.  Warp tile 32 x (8*NI), accumulators X, Y, Z in
// registers, C read in the epilogue only.  Same shared-memory staging as k_zgemm_minus.
// ------------------------------------------------------------------------------------------------------------------
#ifndef MFB_GEMM3M_PRELOAD
#define MFB_GEMM3M_PRELOAD 1
#endif
template <int WM, int WN, int NI>
struct Gemm3Cfg {
  static const int BK = 16, STAGES = 2;
  static const int BM = 32 * WM, BN = 8 * NI * WN, T = 32 * WM * WN;
  static const int SA_LD = BM + 4, SB_LD = BK + 4;
  static const int SA_STAGE = 2 * BK * SA_LD, SB_STAGE = 2 * BN * SB_LD;
  static const int SMEM = STAGES * (SA_STAGE + SB_STAGE) * 8;
};
// PSA: the operand sum Ar + Ai comes PRE-SUMMED from a third plane Asum (same layout as A) instead of being formed on the
// fragments: four of the six DADDs per k-step leave the FP64 pipe they share with the DMMAs (the stall samples of the ncu
// source page sit on exactly these DADDs); costs a third shared-memory plane for A (72.7 KB per CTA).  MEASURED SLOWER
// (LU 2160 instead of 1982 ms at 30258 DOF: the extra plane's staging and fragment loads cost more than the DADDs saved), so it
// is an opt-in experiment (MFB_GEMM_PRESUM=1), not the default.
template <int WM, int WN, int NI, int MINB, bool PSA>
__global__ void __launch_bounds__(32 * WM * WN, MINB)
k_zgemm3m_minus(int M, int N, int K, const double* __restrict__ Are, const double* __restrict__ Aim, const double* __restrict__ Asum, long long lda,
                const double* __restrict__ Bre, const double* __restrict__ Bim, long long ldb, double* __restrict__ Cre, double* __restrict__ Cim, long long ldc) {
  typedef Gemm3Cfg<WM, WN, NI> C;
  constexpr int NPA = PSA ? 3 : 2;
  constexpr int BM = C::BM, BN = C::BN, BK = C::BK, STAGES = C::STAGES, T = C::T, SA_LD = C::SA_LD, SB_LD = C::SB_LD, SA_STAGE = NPA * BK * C::SA_LD, SB_STAGE = C::SB_STAGE;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + STAGES * SA_STAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int gid = lane >> 2, tig = lane & 3;
  const int KT = (K + BK - 1) / BK;
  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * BK;
    double* a = sA + stage * SA_STAGE;
    double* b = sB + stage * SB_STAGE;
#pragma unroll
    for (int i = 0; i < (NPA * BK * (BM / 2) + T - 1) / T; i++) {
      int idx = tid + T * i;
      if ((NPA * BK * (BM / 2)) % T != 0 && idx >= NPA * BK * (BM / 2)) break;
      int p = idx / (BK * (BM / 2)), rem = idx % (BK * (BM / 2)), k = rem / (BM / 2), c2 = rem % (BM / 2);
      int m = m0 + 2 * c2, kk = k0 + k;
      const double* base = (p == 0) ? Are : (p == 1 ? Aim : Asum);
      const double* src = base + (long long)kk * lda + m;
      int bytes = (kk < K) ? max(0, min(16, (M - m) * 8)) : 0;
      if (bytes == 0) src = base;
      cp_async16(a + (p * BK + k) * SA_LD + 2 * c2, src, bytes);
    }
#pragma unroll
    for (int i = 0; i < (2 * BN * (BK / 2) + T - 1) / T; i++) {
      int idx = tid + T * i;
      if ((2 * BN * (BK / 2)) % T != 0 && idx >= 2 * BN * (BK / 2)) break;
      int p = idx / (BN * (BK / 2)), rem = idx % (BN * (BK / 2)), nn = rem / (BK / 2), c2 = rem % (BK / 2);
      int n = n0 + nn, kk = k0 + 2 * c2;
      const double* src = (p ? Bim : Bre) + (long long)n * ldb + kk;
      int bytes = (n < N) ? max(0, min(16, (K - kk) * 8)) : 0;
      if (bytes == 0) src = (p ? Bim : Bre);
      cp_async16(b + (p * BN + nn) * SB_LD + 2 * c2, src, bytes);
    }
  };
  for (int s = 0; s < STAGES - 1; s++) { if (s < KT) load_stage(s, s); cp_async_commit(); }
  // C enters through the accumulators (its loads overlap the first k-tiles in flight, the epilogue only stores):
  // x = -Cr + X, y = Y, z = -(Ci + Cr) + Z  =>  Cr' = y - x,  Ci' = x + y - z
  double x[4][NI][2], y[4][NI][2], z[4][NI][2];
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < NI; ni++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int m = m0 + wm * 32 + mi * 8 + gid, n = n0 + wn * 8 * NI + ni * 8 + 2 * tig + h;
        const bool ok = (m < M) && (n < N);
#if MFB_GEMM3M_PRELOAD
        const double cr = ok ? Cre[(long long)n * ldc + m] : 0.0, ci = ok ? Cim[(long long)n * ldc + m] : 0.0;
        x[mi][ni][h] = -cr; y[mi][ni][h] = 0.0; z[mi][ni][h] = -(ci + cr);
#else
        if (ok) {   // pull the C tile into L2 now; it is read in the epilogue
          asm volatile("prefetch.global.L2 [%0];" ::"l"(Cre + (long long)n * ldc + m));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(Cim + (long long)n * ldc + m));
        }
        x[mi][ni][h] = 0.0; y[mi][ni][h] = 0.0; z[mi][ni][h] = 0.0;
#endif
      }
  for (int kt = 0; kt < KT; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    { int nk = kt + STAGES - 1; if (nk < KT) load_stage(nk % STAGES, nk); cp_async_commit(); }
    const double* a = sA + (kt % STAGES) * SA_STAGE;
    const double* b = sB + (kt % STAGES) * SB_STAGE;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; k4++) {
      double ar[4], ai[4], sa[4];
#pragma unroll
      for (int mi = 0; mi < 4; mi++) {
        int off = (k4 * 4 + tig) * SA_LD + wm * 32 + mi * 8 + gid;
        ar[mi] = a[off]; ai[mi] = a[BK * SA_LD + off]; sa[mi] = PSA ? a[2 * BK * SA_LD + off] : ar[mi] + ai[mi];
      }
#pragma unroll
      for (int ni = 0; ni < NI; ni++) {
        int off = (wn * 8 * NI + ni * 8 + gid) * SB_LD + k4 * 4 + tig;
        const double br = b[off], bi = b[BN * SB_LD + off], sb = br + bi;
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
          dmma(x[mi][ni][0], x[mi][ni][1], ar[mi], br);
          dmma(y[mi][ni][0], y[mi][ni][1], ai[mi], bi);
          dmma(z[mi][ni][0], z[mi][ni][1], sa[mi], sb);
        }
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < NI; ni++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int m = m0 + wm * 32 + mi * 8 + gid, n = n0 + wn * 8 * NI + ni * 8 + 2 * tig + h;
        if (m < M && n < N) {
          const long long o = (long long)n * ldc + m;
#if MFB_GEMM3M_PRELOAD
          Cre[o] = y[mi][ni][h] - x[mi][ni][h];
          Cim[o] = (x[mi][ni][h] + y[mi][ni][h]) - z[mi][ni][h];
#else
          Cre[o] -= x[mi][ni][h] - y[mi][ni][h];
          Cim[o] -= z[mi][ni][h] - x[mi][ni][h] - y[mi][ni][h];
#endif
        }
      }
}
// Real GEMM (C -= A*B) for the real LU of the static path (solve_lse_r -> dgetrf): the same tiling and staging as the 3M
// kernel with one plane and one accumulator set; C enters through the accumulators (x = -C + A*B, C' = -x).
template <int WM, int WN, int NI, int MINB>
__global__ void __launch_bounds__(32 * WM * WN, MINB)
k_dgemm_minus(int M, int N, int K, const double* __restrict__ A, long long lda, const double* __restrict__ B, long long ldb, double* __restrict__ Cm, long long ldc) {
  typedef Gemm3Cfg<WM, WN, NI> C;
  constexpr int BM = C::BM, BN = C::BN, BK = C::BK, STAGES = C::STAGES, T = C::T, SA_LD = C::SA_LD, SB_LD = C::SB_LD;
  constexpr int SA_STAGE = BK * SA_LD, SB_STAGE = BN * SB_LD;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + STAGES * SA_STAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int gid = lane >> 2, tig = lane & 3;
  const int KT = (K + BK - 1) / BK;
  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * BK;
    double* a = sA + stage * SA_STAGE;
    double* b = sB + stage * SB_STAGE;
#pragma unroll
    for (int i = 0; i < (BK * (BM / 2) + T - 1) / T; i++) {
      int idx = tid + T * i;
      if ((BK * (BM / 2)) % T != 0 && idx >= BK * (BM / 2)) break;
      int k = idx / (BM / 2), c2 = idx % (BM / 2);
      int m = m0 + 2 * c2, kk = k0 + k;
      const double* src = A + (long long)kk * lda + m;
      int bytes = (kk < K) ? max(0, min(16, (M - m) * 8)) : 0;
      if (bytes == 0) src = A;
      cp_async16(a + k * SA_LD + 2 * c2, src, bytes);
    }
#pragma unroll
    for (int i = 0; i < (BN * (BK / 2) + T - 1) / T; i++) {
      int idx = tid + T * i;
      if ((BN * (BK / 2)) % T != 0 && idx >= BN * (BK / 2)) break;
      int nn = idx / (BK / 2), c2 = idx % (BK / 2);
      int n = n0 + nn, kk = k0 + 2 * c2;
      const double* src = B + (long long)n * ldb + kk;
      int bytes = (n < N) ? max(0, min(16, (K - kk) * 8)) : 0;
      if (bytes == 0) src = B;
      cp_async16(b + nn * SB_LD + 2 * c2, src, bytes);
    }
  };
  for (int s = 0; s < STAGES - 1; s++) { if (s < KT) load_stage(s, s); cp_async_commit(); }
  double x[4][NI][2];
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < NI; ni++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int m = m0 + wm * 32 + mi * 8 + gid, n = n0 + wn * 8 * NI + ni * 8 + 2 * tig + h;
        x[mi][ni][h] = ((m < M) && (n < N)) ? -Cm[(long long)n * ldc + m] : 0.0;
      }
  for (int kt = 0; kt < KT; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    { int nk = kt + STAGES - 1; if (nk < KT) load_stage(nk % STAGES, nk); cp_async_commit(); }
    const double* a = sA + (kt % STAGES) * SA_STAGE;
    const double* b = sB + (kt % STAGES) * SB_STAGE;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; k4++) {
      double ar[4];
#pragma unroll
      for (int mi = 0; mi < 4; mi++) ar[mi] = a[(k4 * 4 + tig) * SA_LD + wm * 32 + mi * 8 + gid];
#pragma unroll
      for (int ni = 0; ni < NI; ni++) {
        const double br = b[(wn * 8 * NI + ni * 8 + gid) * SB_LD + k4 * 4 + tig];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) dmma(x[mi][ni][0], x[mi][ni][1], ar[mi], br);
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < NI; ni++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int m = m0 + wm * 32 + mi * 8 + gid, n = n0 + wn * 8 * NI + ni * 8 + 2 * tig + h;
        if (m < M && n < N) Cm[(long long)n * ldc + m] = -x[mi][ni][h];
      }
}
void dgemm_minus(int m, int n, int k, const double* A, long long lda, const double* B, long long ldb, double* Cm, long long ldc, cudaStream_t st) {
  if (m <= 0 || n <= 0 || k <= 0) return;
  typedef Gemm3Cfg<2, 2, 4> C;   // 64 x 64 CTA tile, warp 32 x 32, 2 CTAs per SM
  const int smem = C::STAGES * (C::BK * C::SA_LD + C::BN * C::SB_LD) * 8;
  dim3 grid((m + C::BM - 1) / C::BM, (n + C::BN - 1) / C::BN);
  k_dgemm_minus<2, 2, 4, 2><<<grid, C::T, smem, st>>>(m, n, k, A, lda, B, ldb, Cm, ldc);
}

template <int WM, int WN, int NI, int MINB>
static void launch_gemm3m(int m, int n, int k, const double* Are, const double* Aim, long long lda, const double* Bre, const double* Bim,
                          long long ldb, double* Cre, double* Cim, long long ldc, cudaStream_t st) {
  typedef Gemm3Cfg<WM, WN, NI> C;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_zgemm3m_minus<WM, WN, NI, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM); attr = true; }
  dim3 grid((m + C::BM - 1) / C::BM, (n + C::BN - 1) / C::BN);
  k_zgemm3m_minus<WM, WN, NI, MINB, false><<<grid, C::T, C::SMEM, st>>>(m, n, k, Are, Aim, nullptr, lda, Bre, Bim, ldb, Cre, Cim, ldc);
}
// trailing update with the pre-summed A plane (Asum = Are + Aim, same lda): the 64 x 32 / 3 CTAs-per-SM shape
void zgemm_minus_planar_psa(int m, int n, int k, const double* Are, const double* Aim, const double* Asum, long long lda, const double* Bre, const double* Bim,
                            long long ldb, double* Cre, double* Cim, long long ldc, cudaStream_t st) {
  if (m <= 0 || n <= 0 || k <= 0) return;
  typedef Gemm3Cfg<2, 2, 2> C;
  constexpr int SMEM = C::STAGES * (3 * C::BK * C::SA_LD + C::SB_STAGE) * 8;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_zgemm3m_minus<2, 2, 2, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM); attr = true; }
  dim3 grid((m + C::BM - 1) / C::BM, (n + C::BN - 1) / C::BN);
  k_zgemm3m_minus<2, 2, 2, 3, true><<<grid, C::T, SMEM, st>>>(m, n, k, Are, Aim, Asum, lda, Bre, Bim, ldb, Cre, Cim, ldc);
}
// S = Re + Im of an (rows x cols) block (column-major, ld each)
__global__ void k_sum_planes(const double* __restrict__ re, const double* __restrict__ im, long long ld, double* __restrict__ out, long long ldo, int rows, int cols) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / rows, r = i - c * rows;
    out[c * ldo + r] = re[c * ld + r] + im[c * ld + r];
  }
}
void launch_sum_planes(const double* re, const double* im, long long ld, double* out, long long ldo, int rows, int cols, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return;
  long long total = (long long)rows * cols; int g = (int)std::min<long long>((total + 255) / 256, 1184);
  k_sum_planes<<<g, 256, 0, st>>>(re, im, ld, out, ldo, rows, cols);
}

template <int WM, int WN, int BK, int STAGES, int MINB>
static void launch_gemm_cfg(int m, int n, int k, const double* Are, const double* Aim, long long lda, const double* Bre, const double* Bim,
                            long long ldb, double* Cre, double* Cim, long long ldc, cudaStream_t st) {
  typedef GemmCfg<WM, WN, BK, STAGES> C;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_zgemm_minus<WM, WN, BK, STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM); attr = true; }
  dim3 grid((m + C::BM - 1) / C::BM, (n + C::BN - 1) / C::BN);
  k_zgemm_minus<WM, WN, BK, STAGES, MINB><<<grid, C::T, C::SMEM, st>>>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc);
}

static int gemm_cfg() {
  static int cfg = -1;
  if (cfg < 0) { const char* e = getenv("MFB_GEMM_CFG"); cfg = e ? atoi(e) : 15; }
  return cfg;
}

void zgemm_minus_planar(int m, int n, int k, const double* Are, const double* Aim, long long lda, const double* Bre, const double* Bim,
                        long long ldb, double* Cre, double* Cim, long long ldc, cudaStream_t st) {
  if (m <= 0 || n <= 0 || k <= 0) return;
  if (!Aim) { dgemm_minus(m, n, k, Are, lda, Bre, ldb, Cre, ldc, st); return; }   // real system (imaginary planes absent)
  switch (gemm_cfg()) {
    case 10: launch_gemm3m<2, 2, 4, 2>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;   // 3M, 64 x 64, warp 32 x 32, 2 CTA/SM
    case 11: launch_gemm3m<2, 4, 2, 1>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;   // 3M, 64 x 64, 8 warps of 32 x 16, 1 CTA/SM
    case 12: launch_gemm3m<2, 2, 2, 2>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;   // 3M, 64 x 32, warp 32 x 16, 2 CTA/SM
    case 13: launch_gemm3m<4, 2, 2, 1>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;   // 3M, 128 x 32, 8 warps of 32 x 16, 1 CTA/SM
    case 14: launch_gemm3m<2, 2, 3, 2>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;   // 3M, 64 x 48, warp 32 x 24, 2 CTA/SM
    case 17: launch_gemm3m<4, 4, 2, 1>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;   // 3M, 128 x 64, 16 warps of 32 x 16
    case 18: launch_gemm3m<4, 2, 2, 2>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;   // 3M, 128 x 32, 8 warps of 32 x 16, 2 CTA/SM
    case 19: launch_gemm3m<2, 4, 2, 3>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;   // 3M, 64 x 64, 8 warps of 32 x 16, 3 CTA/SM
    case 15: launch_gemm3m<2, 2, 2, 3>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;   // 3M, 64 x 32, warp 32 x 16, 3 CTA/SM
    case 16: launch_gemm3m<2, 4, 2, 2>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;   // 3M, 64 x 64, 8 warps of 32 x 16, 2 CTA/SM
    case 0: launch_gemm_cfg<4, 2, 16, 3, 1>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;   // 128 x 64, 8 warps, 1 CTA/SM
    case 2: launch_gemm_cfg<2, 2, 16, 2, 2>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;   // 64 x 64, 2-stage
    case 3: launch_gemm_cfg<4, 1, 16, 3, 2>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;   // 128 x 32, 4 warps
    case 4: launch_gemm_cfg<2, 2, 8, 4, 2>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;    // 64 x 64, BK 8, 4-stage
    default: launch_gemm_cfg<2, 2, 16, 3, 2>(m, n, k, Are, Aim, lda, Bre, Bim, ldb, Cre, Cim, ldc, st); break;  // 64 x 64, 4 warps, 2 CTA/SM
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Sub-panel factorisation (cooperative): LU with partial pivoting of the (n - c0) x ib block column starting at (c0, c0).
// Every CTA keeps its slab of rows in shared memory for the whole kernel; per column there is ONE grid barrier: pivot
// candidates travel together with a copy of their row, and the rank-1 update of column j is fused with the pivot search
// of column j+1.  Pivot = max |re|+|im|, first occurrence (izamax).  Row interchanges are applied inside the ib columns
// only; the caller applies them to the other columns with k_laswp.
// ------------------------------------------------------------------------------------------------------------------
struct SubPanelArgs {
  double *Are, *Aim; long long lda; int n, c0, ib, rpc, rpcp;
  int* ipiv; double* cand_val; int* cand_row; double* cand_data; double* diag_data; int* info; int ldc;   // ldc = stride of a candidate row record (>= 2*ib)
};

__device__ __forceinline__ void block_argmax(double v, int row, double* s_val, int* s_row, double& best, int& brow) {
  const int tid = threadIdx.x;
  // warp-level reduction first
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double v2 = __shfl_xor_sync(0xffffffffu, v, o); int r2 = __shfl_xor_sync(0xffffffffu, row, o);
    if (v2 > v || (v2 == v && r2 < row)) { v = v2; row = r2; }
  }
  if ((tid & 31) == 0) { s_val[tid >> 5] = v; s_row[tid >> 5] = row; }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  v = s_val[0]; row = s_row[0];
  for (int w = 1; w < nw; w++) { double v2 = s_val[w]; int r2 = s_row[w]; if (v2 > v || (v2 == v && r2 < row)) { v = v2; row = r2; } }
  best = v; brow = row;
  __syncthreads();
}

// Pivot-search key of an entry: |re| + |im| (izamax / idamax), with NaN ranked as +infinity so that the arg-max below is a total order: a NaN
// entry is taken as the pivot (the lowest such row) and propagates through the factors as it does in LAPACK, instead of no row being selected.
__device__ __forceinline__ double pivot_key(double t) { return (t != t) ? __longlong_as_double(0x7ff0000000000000LL) : t; }

const int SP_MAXIB = 32;
const int TS = 64;         // diagonal block of the triangular solves (zgetrs)
void lu_invert_diagonal_blocks(const double* Are, const double* Aim, long long lda, int n, double* inv, cudaStream_t st);
const int SP_CLUSTER_SMEM = 200 * 1024;   // dynamic shared memory a CTA of the cluster panel kernel may use for its slab

// CX = false: real matrix (a.Aim == NULL): the imaginary plane is neither read nor written, the pivot is max |re| (idamax).
template <bool CX>
__global__ void __launch_bounds__(256) k_subpanel(SubPanelArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) double slab[];      // [2][ib][rpcp]
  __shared__ double s_val[8]; __shared__ int s_row[8];
  __shared__ double s_ure[SP_MAXIB], s_uim[SP_MAXIB];
  const int tid = threadIdx.x, c = blockIdx.x, G = gridDim.x;
  const int ib = a.ib, c0 = a.c0, rpcp = a.rpcp;
  const int rs = c0 + c * a.rpc, re = min(rs + a.rpc, a.n), nloc = max(re - rs, 0);
  double* sre = slab; double* sim = slab + (size_t)ib * rpcp;
  const long long lda = a.lda;
  const int BIG = 0x7fffffff;

  for (int idx = tid; idx < ib * nloc; idx += blockDim.x) {
    int jj = idx / nloc, i = idx - jj * nloc;
    sre[jj * rpcp + i] = a.Are[(long long)(c0 + jj) * lda + rs + i];
    sim[jj * rpcp + i] = CX ? a.Aim[(long long)(c0 + jj) * lda + rs + i] : 0.0;
  }
  __syncthreads();
  // candidate for column 0
  {
    double v = -1.0; int r = BIG;
    for (int i = tid; i < nloc; i += blockDim.x) { double t = pivot_key(fabs(sre[i]) + fabs(sim[i])); if (t > v) { v = t; r = rs + i; } }
    double best; int brow; block_argmax(v, r, s_val, s_row, best, brow);
    if (tid == 0) { a.cand_val[c] = best; a.cand_row[c] = brow; }
    if (brow != BIG && tid < ib) {
      a.cand_data[(size_t)c * a.ldc + tid] = sre[tid * rpcp + brow - rs];
      a.cand_data[(size_t)c * a.ldc + ib + tid] = sim[tid * rpcp + brow - rs];
    }
    if (c == 0 && tid < ib) { a.diag_data[tid] = sre[tid * rpcp]; a.diag_data[ib + tid] = sim[tid * rpcp]; }
  }
  for (int j = 0; j < ib; j++) {
    const int buf = j & 1, nbuf = buf ^ 1;
    const int dj = c0 + j;
    __threadfence();
    grid.sync();
    // ---- reduce the G candidates (every CTA does it redundantly) ----
    double v = -1.0; int r = BIG;
    for (int i = tid; i < G; i += blockDim.x) {
      double t = a.cand_val[buf * G + i]; int rr = a.cand_row[buf * G + i];
      if (t > v || (t == v && rr < r)) { v = t; r = rr; }
    }
    double best; int p; block_argmax(v, r, s_val, s_row, best, p);
    const int cstar = (p - c0) / a.rpc;
    if (tid < ib) {
      const double* cd = a.cand_data + ((size_t)buf * G + cstar) * a.ldc;
      s_ure[tid] = cd[tid]; s_uim[tid] = cd[ib + tid];
    }
    __syncthreads();
    const double pr = s_ure[j], pi = s_uim[j];
    const bool zero_pivot = (pr == 0.0 && pi == 0.0);
    if (c == 0 && tid == 0) { a.ipiv[dj] = p + 1; if (zero_pivot) atomicCAS(a.info, 0, dj + 1); }
    // ---- row interchange inside the slab ----
    if (p != dj && tid < ib) {
      if (p >= rs && p < re) {   // row p receives the old diagonal row
        const double* dd = a.diag_data + (size_t)buf * a.ldc;
        sre[tid * rpcp + p - rs] = dd[tid]; sim[tid * rpcp + p - rs] = dd[ib + tid];
      }
    }
    __syncthreads();
    if (p != dj && tid < ib && c == 0) { sre[tid * rpcp + j] = s_ure[tid]; sim[tid * rpcp + j] = s_uim[tid]; }   // diagonal row receives the pivot row
    __syncthreads();
    // ---- scale column j, rank-1 update of the columns to its right, fused pivot search for column j+1 ----
    double ir = 0.0, ii = 0.0;
    if (!zero_pivot) {  // reciprocal of the pivot (zgetf2 scales by the reciprocal), Smith's algorithm
      if (fabs(pr) >= fabs(pi)) { double t = pi / pr, d = pr + pi * t; ir = 1.0 / d; ii = -t / d; }
      else { double t = pr / pi, d = pr * t + pi; ir = t / d; ii = -1.0 / d; }
    }
    double nv = -1.0; int nr = BIG;
    for (int i = tid; i < nloc; i += blockDim.x) {
      if (rs + i <= dj) continue;
      double lr = sre[j * rpcp + i], li = sim[j * rpcp + i];
      if (!zero_pivot) { double t = lr * ir - li * ii; li = lr * ii + li * ir; lr = t; sre[j * rpcp + i] = lr; sim[j * rpcp + i] = li; }
      for (int jj = j + 1; jj < ib; jj++) {
        double xr = sre[jj * rpcp + i], xi = sim[jj * rpcp + i];
        xr -= lr * s_ure[jj] - li * s_uim[jj];
        xi -= lr * s_uim[jj] + li * s_ure[jj];
        sre[jj * rpcp + i] = xr; sim[jj * rpcp + i] = xi;
        if (jj == j + 1) { double t = pivot_key(fabs(xr) + fabs(xi)); if (t > nv) { nv = t; nr = rs + i; } }
      }
    }
    if (j + 1 < ib) {
      double nbest; int nbrow; block_argmax(nv, nr, s_val, s_row, nbest, nbrow);   // its barrier also orders the slab updates
      if (tid == 0) { a.cand_val[nbuf * G + c] = nbest; a.cand_row[nbuf * G + c] = nbrow; }
      if (nbrow != BIG && tid < ib) {
        a.cand_data[((size_t)nbuf * G + c) * a.ldc + tid] = sre[tid * rpcp + nbrow - rs];
        a.cand_data[((size_t)nbuf * G + c) * a.ldc + ib + tid] = sim[tid * rpcp + nbrow - rs];
      }
      if (c == 0 && tid < ib) {
        a.diag_data[(size_t)nbuf * a.ldc + tid] = sre[tid * rpcp + j + 1];
        a.diag_data[(size_t)nbuf * a.ldc + ib + tid] = sim[tid * rpcp + j + 1];
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < ib * nloc; idx += blockDim.x) {
    int jj = idx / nloc, i = idx - jj * nloc;
    a.Are[(long long)(c0 + jj) * lda + rs + i] = sre[jj * rpcp + i];
    if (CX) a.Aim[(long long)(c0 + jj) * lda + rs + i] = sim[jj * rpcp + i];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Sub-panel factorisation inside ONE thread-block cluster (sm_90+ clusters; 16 CTAs = non-portable size on sm_100a).
// Same algorithm and pivot rule as k_subpanel, but the per-column exchange never leaves the cluster: every CTA publishes
// its pivot candidate (value, row, a copy of the row) in its OWN shared memory, a hardware cluster barrier replaces the
// grid barrier, and the winner's row and the diagonal row are read through distributed shared memory.  The grid-wide
// version pays a grid barrier plus two dependent L2 round trips per column (~5 us); this one ~1.5 us, and it occupies
// 16 SMs instead of all of them, so the trailing update of the look-ahead keeps the rest of the GPU.  The slab of a CTA
// (rows x ib columns, one plane for a real matrix) must fit its shared memory: the caller picks ib or falls back.
// ------------------------------------------------------------------------------------------------------------------
template <bool CX>
__global__ void __launch_bounds__(256) k_subpanel_cluster(SubPanelArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) double slab[];      // [planes][ib][rpcp]
  __shared__ double s_val[8]; __shared__ int s_row[8];
  __shared__ double s_ure[SP_MAXIB], s_uim[SP_MAXIB];
  __shared__ double c_val[2]; __shared__ int c_row[2];
  __shared__ double c_data[2][2 * SP_MAXIB];          // candidate row of this CTA (re | im), double buffered over the columns
  __shared__ double d_data[2][2 * SP_MAXIB];          // current diagonal row (CTA 0)
  __shared__ int s_p;
  const int tid = threadIdx.x, c = (int)cluster.block_rank(), G = (int)cluster.num_blocks();
  const int ib = a.ib, c0 = a.c0, rpcp = a.rpcp;
  const int rs = c0 + c * a.rpc, re = min(rs + a.rpc, a.n), nloc = max(re - rs, 0);
  double* sre = slab; double* sim = slab + (size_t)ib * rpcp;   // sim is used only when CX
  const long long lda = a.lda;
  const int BIG = 0x7fffffff;
  // slab load: eight independent loads in flight per thread (the one-element-per-iteration loop of round 1 waited a full memory latency per element:
  // its store to shared memory held 18 % of the kernel's stall samples, profiles/r02_ncu_subpanel.txt)
  {
    const int total = ib * nloc;
    for (int base = tid; base < total; base += 8 * blockDim.x) {
      double vr[8], vi[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int idx = base + u * blockDim.x;
        if (idx < total) { const int jj = idx / nloc, i = idx - jj * nloc; vr[u] = a.Are[(long long)(c0 + jj) * lda + rs + i]; vi[u] = CX ? a.Aim[(long long)(c0 + jj) * lda + rs + i] : 0.0; }
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int idx = base + u * blockDim.x;
        if (idx < total) { const int jj = idx / nloc, i = idx - jj * nloc; sre[jj * rpcp + i] = vr[u]; if (CX) sim[jj * rpcp + i] = vi[u]; }
      }
    }
  }
  __syncthreads();
  {
    double v = -1.0; int r = BIG;
    for (int i = tid; i < nloc; i += blockDim.x) { double t = pivot_key(fabs(sre[i]) + (CX ? fabs(sim[i]) : 0.0)); if (t > v) { v = t; r = rs + i; } }
    double best; int brow; block_argmax(v, r, s_val, s_row, best, brow);
    if (tid == 0) { c_val[0] = best; c_row[0] = brow; }
    if (brow != BIG && tid < ib) { c_data[0][tid] = sre[tid * rpcp + brow - rs]; c_data[0][ib + tid] = CX ? sim[tid * rpcp + brow - rs] : 0.0; }
    if (c == 0 && tid < ib) { d_data[0][tid] = sre[tid * rpcp]; d_data[0][ib + tid] = CX ? sim[tid * rpcp] : 0.0; }
  }
  for (int j = 0; j < ib; j++) {
    const int buf = j & 1, nbuf = buf ^ 1;
    const int dj = c0 + j;
    cluster.sync();                                   // candidates of column j are visible cluster-wide
    if (tid < 32) {                                   // G <= 16 candidates: one warp
      double v = -1.0; int r = BIG;
      if (tid < G) { v = *cluster.map_shared_rank(&c_val[buf], tid); r = *cluster.map_shared_rank(&c_row[buf], tid); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double v2 = __shfl_xor_sync(0xffffffffu, v, o); int r2 = __shfl_xor_sync(0xffffffffu, r, o);
        if (v2 > v || (v2 == v && r2 < r)) { v = v2; r = r2; }
      }
      if (tid == 0) s_p = r;
    }
    __syncthreads();
    const int p = s_p;
    const int cstar = (p - c0) / a.rpc;
    if (tid < ib) {
      const double* cd = cluster.map_shared_rank(&c_data[buf][0], cstar);
      s_ure[tid] = cd[tid]; s_uim[tid] = cd[ib + tid];
    }
    if (p != dj && tid < ib && p >= rs && p < re) {   // row p receives the old diagonal row
      const double* dd = cluster.map_shared_rank(&d_data[buf][0], 0);
      sre[tid * rpcp + p - rs] = dd[tid]; if (CX) sim[tid * rpcp + p - rs] = dd[ib + tid];
    }
    __syncthreads();
    const double pr = s_ure[j], pi = s_uim[j];
    const bool zero_pivot = (pr == 0.0 && pi == 0.0);
    if (c == 0 && tid == 0) { a.ipiv[dj] = p + 1; if (zero_pivot) atomicCAS(a.info, 0, dj + 1); }
    if (p != dj && tid < ib && c == 0) { sre[tid * rpcp + j] = s_ure[tid]; if (CX) sim[tid * rpcp + j] = s_uim[tid]; }   // diagonal row receives the pivot row
    __syncthreads();
    double ir = 0.0, ii = 0.0;
    if (!zero_pivot) {
      if (!CX) ir = 1.0 / pr;
      else if (fabs(pr) >= fabs(pi)) { double t = pi / pr, d = pr + pi * t; ir = 1.0 / d; ii = -t / d; }
      else { double t = pr / pi, d = pr * t + pi; ir = t / d; ii = -1.0 / d; }
    }
    double nv = -1.0; int nr = BIG;
    for (int i = tid; i < nloc; i += blockDim.x) {
      if (rs + i <= dj) continue;
      double lr = sre[j * rpcp + i], li = CX ? sim[j * rpcp + i] : 0.0;
      if (!zero_pivot) {
        if (CX) { double t = lr * ir - li * ii; li = lr * ii + li * ir; lr = t; sre[j * rpcp + i] = lr; sim[j * rpcp + i] = li; }
        else { lr = lr * ir; sre[j * rpcp + i] = lr; }
      }
      for (int jj = j + 1; jj < ib; jj++) {
        double xr = sre[jj * rpcp + i];
        if (CX) {
          double xi = sim[jj * rpcp + i];
          xr -= lr * s_ure[jj] - li * s_uim[jj];
          xi -= lr * s_uim[jj] + li * s_ure[jj];
          sre[jj * rpcp + i] = xr; sim[jj * rpcp + i] = xi;
          if (jj == j + 1) { double t = pivot_key(fabs(xr) + fabs(xi)); if (t > nv) { nv = t; nr = rs + i; } }
        } else {
          xr -= lr * s_ure[jj];
          sre[jj * rpcp + i] = xr;
          if (jj == j + 1) { double t = pivot_key(fabs(xr)); if (t > nv) { nv = t; nr = rs + i; } }
        }
      }
    }
    if (j + 1 < ib) {
      double nbest; int nbrow; block_argmax(nv, nr, s_val, s_row, nbest, nbrow);   // its barrier also orders the slab updates
      if (tid == 0) { c_val[nbuf] = nbest; c_row[nbuf] = nbrow; }
      if (nbrow != BIG && tid < ib) { c_data[nbuf][tid] = sre[tid * rpcp + nbrow - rs]; c_data[nbuf][ib + tid] = CX ? sim[tid * rpcp + nbrow - rs] : 0.0; }
      if (c == 0 && tid < ib) { d_data[nbuf][tid] = sre[tid * rpcp + j + 1]; d_data[nbuf][ib + tid] = CX ? sim[tid * rpcp + j + 1] : 0.0; }
    }
  }
  cluster.sync();    // nobody leaves (and frees its shared memory) while a neighbour may still read it
  for (int idx = tid; idx < ib * nloc; idx += blockDim.x) {
    int jj = idx / nloc, i = idx - jj * nloc;
    a.Are[(long long)(c0 + jj) * lda + rs + i] = sre[jj * rpcp + i];
    if (CX) a.Aim[(long long)(c0 + jj) * lda + rs + i] = sim[jj * rpcp + i];
  }
}
// launch on a cluster of `cl` CTAs; returns cudaError
template <bool CX>
static cudaError_t launch_subpanel_cluster(const SubPanelArgs& pa, int cl, size_t smem, cudaStream_t st) {
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cl); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_subpanel_cluster<CX>, pa);
}

// row interchanges ipiv[k0 .. k0+nbw) applied to columns [c0,c1)
__global__ void k_laswp(double* Are, double* Aim, long long lda, int c0, int c1, int k0, int nbw, const int* __restrict__ ipiv) {
  int col = c0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= c1) return;
  double* ar = Are + (long long)col * lda; double* ai = Aim + (long long)col * lda;
  for (int j = 0; j < nbw; j++) {
    int p = ipiv[k0 + j] - 1, d = k0 + j;
    if (p != d) { double t = ar[d]; ar[d] = ar[p]; ar[p] = t; if (Aim) { t = ai[d]; ai[d] = ai[p]; ai[p] = t; } }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// In-panel update after a sub-panel of ib columns at (c0, c0): for the panel columns [cA, cB) to its right,
// U12 = inv(L11) A12 and A22 -= L21 U12 in ONE launch (these used to be a TRSM kernel plus a GEMM kernel per sub-panel;
// inside a panel both are latency, not throughput).  CTA (x, y): column block y of PU_TC columns; every CTA redoes the
// tiny forward substitution of its column block in shared memory; x = 0 stores U12, x >= 1 updates PU_RB rows: one row per
// thread, its L21 row (ib values) in registers, U12 broadcast from shared memory, C read and written once, coalesced.
// U12 overwrites A12, which every CTA of the column block reads first: the CTA that arrives LAST at the block's counter
// (after its own read) stores it and re-arms the counter.
// k_laswp2: the row interchanges of the sub-panel on the panel columns left and right of it, one launch.
// ------------------------------------------------------------------------------------------------------------------
const int PU_TC = 32, PU_RB = 128;
template <bool CX, int IBM>
__global__ void __launch_bounds__(256) k_panel_update(double* Are, double* Aim, long long lda, int n, int c0, int ib, int cA, int cB, int* __restrict__ arrive) {
  __shared__ double ur[IBM][PU_TC + 1], ui[CX ? IBM : 1][PU_TC + 1];
  __shared__ double lr[IBM][IBM + 1], li[CX ? IBM : 1][IBM + 1];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cb = cA + blockIdx.y * PU_TC, ncol = min(PU_TC, cB - cb);
  for (int idx = tid; idx < IBM * PU_TC; idx += 256) {
    const int cc = idx / IBM, i = idx - cc * IBM;
    const bool ok = cc < ncol && i < ib;
    ur[i][cc] = ok ? Are[(long long)(cb + cc) * lda + c0 + i] : 0.0;
    if (CX) ui[i][cc] = ok ? Aim[(long long)(cb + cc) * lda + c0 + i] : 0.0;
  }
  for (int idx = tid; idx < IBM * IBM; idx += 256) {
    const int j = idx / IBM, i = idx - j * IBM;
    const bool ok = i < ib && j < ib && i > j;
    lr[i][j] = ok ? Are[(long long)(c0 + j) * lda + c0 + i] : 0.0;
    if (CX) li[i][j] = ok ? Aim[(long long)(c0 + j) * lda + c0 + i] : 0.0;
  }
  __syncthreads();
  if (tid == 0) {                                          // A12 has been read: arrive
    __threadfence();
    const int prev = atomicAdd(arrive + blockIdx.y, 1);
    s_last = (prev == (int)gridDim.x - 1);
    if (s_last) arrive[blockIdx.y] = 0;
  }
  for (int j = 0; j < ib - 1; j++) {                       // forward substitution, unit lower L11; lanes = columns, warps = rows
    const double xr = ur[j][lane], xi = CX ? ui[j][lane] : 0.0;
    for (int i = j + 1 + warp; i < ib; i += 8) {
      if (CX) { ur[i][lane] -= lr[i][j] * xr - li[i][j] * xi; ui[i][lane] -= lr[i][j] * xi + li[i][j] * xr; }
      else ur[i][lane] -= lr[i][j] * xr;
    }
    __syncthreads();
  }
  __syncthreads();
  if (s_last) {
    for (int idx = tid; idx < ib * PU_TC; idx += 256) {
      const int cc = idx / ib, i = idx - cc * ib;
      if (cc < ncol) { Are[(long long)(cb + cc) * lda + c0 + i] = ur[i][cc]; if (CX) Aim[(long long)(cb + cc) * lda + c0 + i] = ui[i][cc]; }
    }
  }
  const int r = c0 + ib + blockIdx.x * PU_RB + (tid & (PU_RB - 1)), half = tid / PU_RB;
  if (r >= n) return;
  double Lr[IBM], Li[CX ? IBM : 1];
#pragma unroll
  for (int k = 0; k < IBM; k++) {
    Lr[k] = (k < ib) ? Are[(long long)(c0 + k) * lda + r] : 0.0;
    if (CX) Li[k] = (k < ib) ? Aim[(long long)(c0 + k) * lda + r] : 0.0;
  }
  const int cc0 = half * (PU_TC / 2), cc1 = min(cc0 + PU_TC / 2, ncol);
#pragma unroll 4
  for (int cc = cc0; cc < cc1; cc++) {
    const long long o = (long long)(cb + cc) * lda + r;
    double sr = Are[o], si = CX ? Aim[o] : 0.0;
#pragma unroll
    for (int k = 0; k < IBM; k++) {
      if (CX) { sr = fma(-Lr[k], ur[k][cc], fma(Li[k], ui[k][cc], sr)); si = fma(-Lr[k], ui[k][cc], fma(-Li[k], ur[k][cc], si)); }
      else sr = fma(-Lr[k], ur[k][cc], sr);
    }
    Are[o] = sr; if (CX) Aim[o] = si;
  }
}
template <bool CX>
static void launch_panel_update(double* Are, double* Aim, long long lda, int n, int c0, int ib, int cA, int cB, int* arrive, cudaStream_t st) {
  const int mrest = n - c0 - ib;
  dim3 grid(mrest > 0 ? (mrest + PU_RB - 1) / PU_RB : 1, (cB - cA + PU_TC - 1) / PU_TC);
  if (ib <= 8) k_panel_update<CX, 8><<<grid, 256, 0, st>>>(Are, Aim, lda, n, c0, ib, cA, cB, arrive);
  else if (ib <= 16) k_panel_update<CX, 16><<<grid, 256, 0, st>>>(Are, Aim, lda, n, c0, ib, cA, cB, arrive);
  else k_panel_update<CX, 32><<<grid, 256, 0, st>>>(Are, Aim, lda, n, c0, ib, cA, cB, arrive);
}
// interchanges ipiv[k0 .. k0+nbw) on the columns [a0,a1) and [b0,b1)
__global__ void k_laswp2(double* Are, double* Aim, long long lda, int a0, int a1, int b0, int b1, int k0, int nbw, const int* __restrict__ ipiv) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int col = (t < a1 - a0) ? a0 + t : b0 + (t - (a1 - a0));
  if (t >= (a1 - a0) + (b1 - b0)) return;
  double* ar = Are + (long long)col * lda; double* ai = Aim + (long long)col * lda;
  for (int j = 0; j < nbw; j++) {
    int p = ipiv[k0 + j] - 1, d = k0 + j;
    if (p != d) { double t2 = ar[d]; ar[d] = ar[p]; ar[p] = t2; if (Aim) { t2 = ai[d]; ai[d] = ai[p]; ai[p] = t2; } }
  }
}

// X = inv(L) * B in place: L = unit lower nbw x nbw block at (r0,r0), B = rows r0..r0+nbw of columns [c0,c1).
// One CTA per TRSM_TC columns; B tile in shared memory; warps own rows (warp-uniform L loads), lanes own columns.
const int TRSM_TC = 32;
const int TRSM_TB = 32;    // rows solved per k_trsm_lu launch (the L block is staged in shared memory)
// L (Lre/Lim, ldl) points at the top-left of the unit lower block, B (Bre/Bim, ldb) at the first of its nbw rows in column 0 of
// the ncols columns to solve (the two may live in different arrays: the distributed LU keeps L in the broadcast panel).
__global__ void __launch_bounds__(256) k_trsm_lu(const double* __restrict__ Lre, const double* __restrict__ Lim, long long ldl, double* Bre, double* Bim,
                                                 long long ldb, int nbw, int ncols) {
  extern __shared__ __align__(16) double sb[];     // [2][nbw][TRSM_TC+1]
  __shared__ double slr[TRSM_TB][TRSM_TB + 1], sli[TRSM_TB][TRSM_TB + 1];   // the L block (nbw <= TRSM_TB): a global load per substitution step was the latency of this kernel
  const int LD = TRSM_TC + 1;
  double* br = sb; double* bi = sb + (size_t)nbw * LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int cb = blockIdx.x * TRSM_TC, ncol = min(TRSM_TC, ncols - cb);
  for (int idx = tid; idx < nbw * nbw; idx += blockDim.x) {
    const int j = idx / nbw, i = idx - j * nbw;
    slr[i][j] = Lre[(long long)j * ldl + i]; sli[i][j] = Lim ? Lim[(long long)j * ldl + i] : 0.0;
  }
  // load: thread (i = tid % nbw-chunk, col) coalesced along rows
  for (int idx = tid; idx < nbw * TRSM_TC; idx += blockDim.x) {
    int cc = idx / nbw, i = idx - cc * nbw;
    bool ok = cc < ncol;
    br[i * LD + cc] = ok ? Bre[(long long)(cb + cc) * ldb + i] : 0.0;
    bi[i * LD + cc] = (ok && Bim) ? Bim[(long long)(cb + cc) * ldb + i] : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < nbw - 1; j++) {
    const double xr = br[j * LD + lane], xi = bi[j * LD + lane];
#pragma unroll 4
    for (int i = j + 1 + warp; i < nbw; i += nw) {
      const double lr = slr[i][j], li = sli[i][j];
      br[i * LD + lane] -= lr * xr - li * xi;
      bi[i * LD + lane] -= lr * xi + li * xr;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < nbw * TRSM_TC; idx += blockDim.x) {
    int cc = idx / nbw, i = idx - cc * nbw;
    if (cc < ncol) { Bre[(long long)(cb + cc) * ldb + i] = br[i * LD + cc]; if (Bim) Bim[(long long)(cb + cc) * ldb + i] = bi[i * LD + cc]; }
  }
}
// U12 = inv(L11) A12 for the nbw x nbw unit lower block at (r0,r0) and columns [c0,c1): blocked forward substitution,
// TRSM_TB rows at a time by substitution in shared memory, the rows below updated on the tensor pipe (returns launches).
static int launch_trsm_ext(const double* Lre, const double* Lim, long long ldl, double* Bre, double* Bim, long long ldb, int nbw, int ncols, cudaStream_t st) {
  if (ncols <= 0 || nbw <= 1) return 0;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_trsm_lu, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * (TRSM_TC + 1) * 8); attr = true; }
  int launches = 0;
  for (int jb = 0; jb < nbw; jb += TRSM_TB) {
    const int tb = (nbw - jb < TRSM_TB) ? (nbw - jb) : TRSM_TB;
    if (tb > 1) {
      size_t smem = (size_t)2 * tb * (TRSM_TC + 1) * 8;
      k_trsm_lu<<<(ncols + TRSM_TC - 1) / TRSM_TC, 256, smem, st>>>(Lre + (long long)jb * ldl + jb, poff(Lim, (long long)jb * ldl + jb), ldl, Bre + jb, poff(Bim, jb), ldb, tb, ncols);
      launches++;
    }
    const int mrest = nbw - jb - tb;
    if (mrest > 0) {
      zgemm_minus_planar(mrest, ncols, tb, Lre + (long long)jb * ldl + jb + tb, poff(Lim, (long long)jb * ldl + jb + tb), ldl, Bre + jb, poff(Bim, jb), ldb,
                         Bre + jb + tb, poff(Bim, jb + tb), ldb, st);
      launches++;
    }
  }
  return launches;
}
static int launch_trsm(double* Are, double* Aim, long long lda, int r0, int nbw, int c0, int c1, cudaStream_t st) {
  if (c1 <= c0) return 0;
  return launch_trsm_ext(Are + (long long)r0 * lda + r0, poff(Aim, (long long)r0 * lda + r0), lda, Are + (long long)c0 * lda + r0, poff(Aim, (long long)c0 * lda + r0), lda, nbw,
                         c1 - c0, st);
}

int lu_work_alloc(LuWork& w, int n, int nb) {
  w.nb = nb;
  const char* e_ib = getenv("MFB_LU_IB");
  w.ib = e_ib ? atoi(e_ib) : 16;
  if (w.ib < 4 || w.ib > SP_MAXIB || (w.ib & 3) || nb % w.ib) w.ib = 16;
  cudaDeviceProp prop; int dev; cudaGetDevice(&dev); cudaGetDeviceProperties(&prop, dev);
  w.n_sm = prop.multiProcessorCount;
  const char* e_la = getenv("MFB_LU_LOOKAHEAD");
  w.lookahead = e_la ? atoi(e_la) : 1;
  const char* e_pc = getenv("MFB_LU_PANEL_CTAS");
  w.panel_ctas = e_pc ? atoi(e_pc) : w.n_sm;
  if (w.panel_ctas < 1 || w.panel_ctas > w.n_sm) w.panel_ctas = w.n_sm;
  // cluster-resident panel: 16 CTAs (non-portable cluster size) if the device takes it, else 8; MFB_LU_CLUSTER=0 disables it
  {
    const char* e_cl = getenv("MFB_LU_CLUSTER");
    int want = e_cl ? atoi(e_cl) : 16;
    if (want != 0 && want != 8 && want != 16) want = 16;
    w.cluster = 0;
    const char* e_mr = getenv("MFB_LU_CLUSTER_MAX_ROWS");
    w.cluster_max_rows = e_mr ? atoi(e_mr) : (1 << 30);
    { const char* e_r = getenv("MFB_LU_CLUSTER_MIN_ROWS"); w.cluster_min_rows = e_r ? atoi(e_r) : 64; if (w.cluster_min_rows < 32) w.cluster_min_rows = 32; }
    const char* e_ci = getenv("MFB_LU_CLUSTER_IB");
    w.cluster_ib = e_ci ? atoi(e_ci) : 32;
    if (want > 0) {
      bool ok = cudaFuncSetAttribute(k_subpanel_cluster<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_CLUSTER_SMEM) == cudaSuccess &&
                cudaFuncSetAttribute(k_subpanel_cluster<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_CLUSTER_SMEM) == cudaSuccess;
      if (ok && want == 16)
        ok = cudaFuncSetAttribute(k_subpanel_cluster<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
             cudaFuncSetAttribute(k_subpanel_cluster<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
      if (ok) w.cluster = want;
      cudaGetLastError();
    }
  }
  { const char* e_fp = getenv("MFB_LU_FUSED_PANEL_UPDATE"); w.fused_panel_update = e_fp ? atoi(e_fp) : 2; }
  w.tma.ok = 0; w.tma_key = nullptr;
  size_t G = (size_t)w.n_sm;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
  A((void**)&w.cand_val, 2 * G * sizeof(double)); A((void**)&w.cand_row, 2 * G * sizeof(int));
  A((void**)&w.cand_data, 2 * G * 2 * SP_MAXIB * sizeof(double)); A((void**)&w.diag_data, 2 * 2 * SP_MAXIB * sizeof(double));
  A((void**)&w.info, sizeof(int));
  A((void**)&w.pu_arrive, 64 * sizeof(int));
  { const char* e_ps = getenv("MFB_GEMM_PRESUM"); w.asum[0] = w.asum[1] = nullptr;
    if (e_ps && atoi(e_ps) != 0) { const size_t ldn = ((size_t)n + 31) / 32 * 32; A((void**)&w.asum[0], ldn * nb * sizeof(double)); A((void**)&w.asum[1], ldn * nb * sizeof(double)); } }
  w.solve_ws = nullptr; A((void**)&w.solve_ws, (size_t)2 * n * sizeof(double));
  { const char* e_si = getenv("MFB_LU_SOLVE_INV"); w.inv = nullptr; if (!e_si || atoi(e_si) != 0) A((void**)&w.inv, (size_t)((n + TS - 1) / TS) * 4 * TS * TS * sizeof(double)); }
  // (no cudaMemset here: it runs on the legacy default stream, which synchronises implicitly with every blocking stream of the process and is an error
  // while another host thread captures a graph; pu_arrive is cleared on the panel stream below)
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_subpanel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_subpanel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  w.n_evs = 5 * ((n + nb - 1) / nb);
  w.evs = new cudaEvent_t[w.n_evs];
  for (int i = 0; i < w.n_evs; i++) cudaEventCreate(&w.evs[i]);
  const int n_steps = (n + nb - 1) / nb;
  w.pevs = new cudaEvent_t[2 * n_steps];
  for (int i = 0; i < 2 * n_steps; i++) cudaEventCreate(&w.pevs[i]);
  int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&w.panel_stream, cudaStreamNonBlocking, hi);
  if (e == cudaSuccess) e = cudaMemsetAsync(w.pu_arrive, 0, 64 * sizeof(int), w.panel_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(w.panel_stream);
  cudaEventCreateWithFlags(&w.ev_next_cols, cudaEventDisableTiming); cudaEventCreateWithFlags(&w.ev_panel_done, cudaEventDisableTiming);
  w.ms_panel = w.ms_swap = w.ms_trsm = w.ms_gemm = 0.f; w.n_steps_timed = 0; w.gemm_launches = 0; w.gemm_flops = 0.0;
  return (int)e;
}
void lu_work_free(LuWork& w) {
  cudaFree(w.cand_val); cudaFree(w.cand_row); cudaFree(w.cand_data); cudaFree(w.diag_data); cudaFree(w.info); cudaFree(w.pu_arrive); cudaFree(w.inv); cudaFree(w.solve_ws); cudaFree(w.asum[0]); cudaFree(w.asum[1]);
  for (int i = 0; i < w.n_evs; i++) cudaEventDestroy(w.evs[i]);
  for (int i = 0; i < 2 * (w.n_evs / 5); i++) cudaEventDestroy(w.pevs[i]);
  cudaEventDestroy(w.ev_next_cols); cudaEventDestroy(w.ev_panel_done); cudaStreamDestroy(w.panel_stream);
  delete[] w.evs; delete[] w.pevs;
}
void lu_collect_times(LuWork& w) {
  w.ms_panel = w.ms_swap = w.ms_trsm = w.ms_gemm = 0.f;
  for (int s = 0; s < w.n_steps_timed; s++) {
    cudaEvent_t* ev = w.evs + 5 * s; float t;
    cudaEventElapsedTime(&t, w.pevs[2 * s], w.pevs[2 * s + 1]); w.ms_panel += t;   // on its own stream: overlaps the trailing update under look-ahead
    cudaEventElapsedTime(&t, ev[1], ev[2]); w.ms_swap += t;
    cudaEventElapsedTime(&t, ev[2], ev[3]); w.ms_trsm += t;
    cudaEventElapsedTime(&t, ev[3], ev[4]); w.ms_gemm += t;
  }
}

// Panel = block column [k0, k0+nbw): sub-panels of ib columns (cooperative kernel above); after each sub-panel its row
// interchanges are applied to the rest of the panel, then U12' = inv(L11') A12' and A22' -= L21' U12' inside the panel.
static int factor_panel(double* Are, double* Aim, long long lda, int n, int k0, int nbw, int* ipiv, LuWork& w, cudaStream_t st) {
  // Sub-panel width of this panel: the cluster kernel needs the slab of a CTA (rows/cluster x ib, one or two planes) in
  // shared memory; 32 columns if they fit, else the configured width, else the grid-wide kernel.
  const int planes = Aim ? 2 : 1;
  int ib_panel = w.ib; bool use_cluster = false;
  if (w.cluster > 1 && (n - k0) <= w.cluster_max_rows) {
    // widest first (every sub-panel launch has a fixed cost); narrower than the configured width never pays: measured, the
    // cluster kernel beats the grid-wide one only while a CTA's slab holds >= 16 columns (profiles/r01_panel_vs_m.log)
    const int cands[2] = {w.cluster_ib, w.ib};
    for (int t = 0; t < 2 && !use_cluster; t++) {
      const int rpc = (n - k0 + w.cluster - 1) / w.cluster;
      if (cands[t] >= 8 && cands[t] <= SP_MAXIB && nbw % cands[t] == 0 &&
          (size_t)planes * cands[t] * (rpc | 1) * sizeof(double) <= (size_t)SP_CLUSTER_SMEM) { ib_panel = cands[t]; use_cluster = true; }
    }
  }
  for (int j0 = 0; j0 < nbw; j0 += ib_panel) {
    const int ib = (nbw - j0 < ib_panel) ? (nbw - j0) : ib_panel;
    const int c0 = k0 + j0, m = n - c0;
    SubPanelArgs pa; pa.Are = Are; pa.Aim = Aim; pa.lda = lda; pa.n = n; pa.c0 = c0; pa.ib = ib;
    pa.ipiv = ipiv; pa.cand_val = w.cand_val; pa.cand_row = w.cand_row; pa.cand_data = w.cand_data; pa.diag_data = w.diag_data; pa.info = w.info;
    pa.ldc = 2 * SP_MAXIB;
    cudaError_t e = cudaSuccess;
    bool done = false;
    if (use_cluster) {
      // a smaller (power of two) cluster for a short sub-panel: at least cluster_min_rows rows per CTA as long as its slab fits shared memory.  Small
      // clusters also matter when several systems are factorised side by side (capi.ProblemLanes): a 16-CTA cluster needs 16 free SMs inside ONE GPC
      // at the same moment, an 8- or 4-CTA cluster (portable sizes) is placed far more easily between the kernels of the other lanes.
      int G = w.cluster;
      while (G > 1 && (m + G - 1) / G < w.cluster_min_rows && (size_t)planes * ib * (((m + G / 2 - 1) / (G / 2)) | 1) * sizeof(double) <= (size_t)SP_CLUSTER_SMEM) G >>= 1;
      const int rpc = (m + G - 1) / G;
      pa.rpc = rpc; pa.rpcp = rpc | 1;
      const size_t smem = (size_t)planes * ib * pa.rpcp * sizeof(double);
      e = Aim ? launch_subpanel_cluster<true>(pa, G, smem, st) : launch_subpanel_cluster<false>(pa, G, smem, st);
      if (e == cudaSuccess) done = true;
      else { cudaGetLastError(); w.cluster = 0; use_cluster = false; }   // cluster launch not available: grid-wide kernel from now on
    }
    if (!done) {
      int G = w.panel_ctas;
      int rpc = (m + G - 1) / G; if (rpc < 64) rpc = 64;
      G = (m + rpc - 1) / rpc;
      pa.rpc = rpc; pa.rpcp = rpc | 1;
      void* args[] = {&pa};
      size_t smem = (size_t)2 * ib * pa.rpcp * sizeof(double);
      e = Aim ? cudaLaunchCooperativeKernel((void*)k_subpanel<true>, dim3(G), dim3(256), args, smem, st)
              : cudaLaunchCooperativeKernel((void*)k_subpanel<false>, dim3(G), dim3(256), args, smem, st);
      if (e != cudaSuccess) return (int)e;
    }
    w.launches += 1;
    // interchanges of this sub-panel on the other columns of the panel, then U12' = inv(L11') A12' and A22' -= L21' U12'
    const int nright = nbw - j0 - ib;
    if (j0 > 0 || nright > 0) {
      const int cnt = j0 + nright;
      k_laswp2<<<(cnt + 127) / 128, 128, 0, st>>>(Are, Aim, lda, k0, c0, c0 + ib, k0 + nbw, c0, ib, ipiv); w.launches++;
    }
    if (nright > 0) {
      if (w.fused_panel_update == 1 || (w.fused_panel_update == 2 && n - c0 <= 16384)) {   // 2 = auto: short panels are latency, tall ones belong on the tensor pipe
        if (Aim) launch_panel_update<true>(Are, Aim, lda, n, c0, ib, c0 + ib, k0 + nbw, w.pu_arrive, st);
        else launch_panel_update<false>(Are, Aim, lda, n, c0, ib, c0 + ib, k0 + nbw, w.pu_arrive, st);
        w.launches++;
      } else {
        w.launches += launch_trsm(Are, Aim, lda, c0, ib, c0 + ib, k0 + nbw, st);
        const int mrest = n - c0 - ib;
        if (mrest > 0) {
          zgemm_minus_planar(mrest, nright, ib, Are + (long long)c0 * lda + c0 + ib, poff(Aim, (long long)c0 * lda + c0 + ib), lda,
                             Are + (long long)(c0 + ib) * lda + c0, poff(Aim, (long long)(c0 + ib) * lda + c0), lda,
                             Are + (long long)(c0 + ib) * lda + c0 + ib, poff(Aim, (long long)(c0 + ib) * lda + c0 + ib), lda, st);
          w.launches++;
        }
      }
    }
  }
  return 0;
}

// Right-looking blocked LU with one panel of look-ahead: as soon as the trailing update of step k has finished the columns of
// panel k+1, that panel is factorised on a high-priority stream while the main stream updates the remaining columns.
int zgetrf_planar(double* Are, double* Aim, long long lda, int n, int* ipiv, LuWork& w, cudaStream_t st, bool timing) {
  const int nb = w.nb;
  cudaMemsetAsync(w.info, 0, sizeof(int), st);
  w.launches = 0; w.gemm_launches = 0; w.gemm_flops = 0.0; w.gemm_exec_flops = 0.0; w.n_steps_timed = 0;
  cudaStream_t ps = w.lookahead ? w.panel_stream : st;
  // tensor maps of the two planes for the TMA trailing-update kernel (gemm_tma.cu); rebuilt only when the matrix moves
  if (Aim && (w.tma_key != Are || !w.tma.ok)) { gemm_tma_make_maps(w.tma, Are, Aim, lda, n, n, Are, Aim, lda, n, n); w.tma_key = Are; }
  auto gemm = [&](int r0, int k0, int kw, int c0, int c1) {   // A[r0:n, c0:c1] -= A[r0:n, k0:k0+kw] * A[k0:k0+kw, c0:c1]
    if (c1 <= c0 || r0 >= n) return;
    if (Aim && gemm_tma_usable(w.tma, n - r0, c1 - c0, kw) && gemm_cfg() == 15)
      zgemm_minus_planar_tma(w.tma, n - r0, c1 - c0, kw, r0, k0, k0, c0, Are + (long long)c0 * lda + r0, Aim + (long long)c0 * lda + r0, lda, st);
    else if (Aim && w.asum[0] && gemm_cfg() == 15)
      zgemm_minus_planar_psa(n - r0, c1 - c0, kw, Are + (long long)k0 * lda + r0, Aim + (long long)k0 * lda + r0, w.asum[(k0 / nb) & 1] + r0, lda,
                             Are + (long long)c0 * lda + k0, Aim + (long long)c0 * lda + k0, lda, Are + (long long)c0 * lda + r0, Aim + (long long)c0 * lda + r0, lda, st);
    else
    zgemm_minus_planar(n - r0, c1 - c0, kw, Are + (long long)k0 * lda + r0, poff(Aim, (long long)k0 * lda + r0), lda,
                       Are + (long long)c0 * lda + k0, poff(Aim, (long long)c0 * lda + k0), lda, Are + (long long)c0 * lda + r0, poff(Aim, (long long)c0 * lda + r0), lda, st);
    const double mnk = (double)(n - r0) * (double)(c1 - c0) * (double)kw;
    w.launches++; w.gemm_launches++; w.gemm_flops += (Aim ? 8.0 : 2.0) * mnk;
    w.gemm_exec_flops += (Aim ? (gemm_cfg() >= 10 ? 6.0 : 8.0) : 2.0) * mnk;
  };
  // first panel
  {
    if (w.lookahead) { cudaEventRecord(w.ev_next_cols, st); cudaStreamWaitEvent(ps, w.ev_next_cols, 0); }
    if (timing) cudaEventRecord(w.pevs[0], ps);
    int e = factor_panel(Are, Aim, lda, n, 0, (n < nb) ? n : nb, ipiv, w, ps);
    if (e) return e;
    if (Aim && w.asum[0] && n > nb) launch_sum_planes(Are + nb, Aim + nb, lda, w.asum[0] + nb, lda, n - nb, nb, ps);   // L21 of panel 0
    if (timing) cudaEventRecord(w.pevs[1], ps);
    if (w.lookahead) { cudaEventRecord(w.ev_panel_done, ps); cudaStreamWaitEvent(st, w.ev_panel_done, 0); }
  }
  for (int k0 = 0, step = 0; k0 < n; k0 += nb, step++) {
    cudaEvent_t* ev = w.evs + 5 * step;
    const int nbw = (n - k0 < nb) ? (n - k0) : nb;
    const int c_next = k0 + nbw;                                  // first column of the next panel
    const int nbw_next = (n - c_next < nb) ? (n - c_next) : nb;   // its width (0 at the end)
    if (timing) cudaEventRecord(ev[1], st);
    // ---- interchanges of panel k outside the panel ----
    if (k0 > 0) { k_laswp<<<(k0 + 127) / 128, 128, 0, st>>>(Are, Aim, lda, 0, k0, k0, nbw, ipiv); w.launches++; }
    if (c_next < n) { k_laswp<<<(n - c_next + 127) / 128, 128, 0, st>>>(Are, Aim, lda, c_next, n, k0, nbw, ipiv); w.launches++; }
    if (timing) cudaEventRecord(ev[2], st);
    if (c_next < n) {
      // ---- U12 = inv(L11) * A12 ----
      w.launches += launch_trsm(Are, Aim, lda, k0, nbw, c_next, n, st);
      if (timing) cudaEventRecord(ev[3], st);
      // ---- trailing update, the columns of the next panel first ----
      gemm(c_next, k0, nbw, c_next, c_next + nbw_next);
      if (w.lookahead) { cudaEventRecord(w.ev_next_cols, st); cudaStreamWaitEvent(ps, w.ev_next_cols, 0); }
      if (timing) cudaEventRecord(w.pevs[2 * (step + 1)], ps);
      int e = factor_panel(Are, Aim, lda, n, c_next, nbw_next, ipiv, w, ps);
      if (e) return e;
      if (Aim && w.asum[0] && c_next + nbw_next < n)   // L21 of panel step+1: rows below it, for the trailing update of the next step
        launch_sum_planes(Are + (long long)c_next * lda + c_next + nbw_next, Aim + (long long)c_next * lda + c_next + nbw_next, lda,
                          w.asum[(step + 1) & 1] + c_next + nbw_next, lda, n - c_next - nbw_next, nbw_next, ps);
      if (timing) cudaEventRecord(w.pevs[2 * (step + 1) + 1], ps);
      if (w.lookahead) cudaEventRecord(w.ev_panel_done, ps);
      gemm(c_next, k0, nbw, c_next + nbw_next, n);
      if (w.lookahead) cudaStreamWaitEvent(st, w.ev_panel_done, 0);
    } else if (timing) cudaEventRecord(ev[3], st);
    if (timing) { cudaEventRecord(ev[4], st); w.n_steps_timed++; }
  }
  if (w.inv) lu_invert_diagonal_blocks(Are, Aim, lda, n, w.inv, st);   // for the solves (zgetrs_planar)
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// Triangular solves (zgetrs, 'N'): b := P b ; L y = b (unit lower) ; U x = y.  Blocked by TS rows: the diagonal block is
// staged in shared memory, the off-diagonal update is a bandwidth-bound GEMV (64 rows x 4 column groups per CTA).
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_permute(const double* __restrict__ sre, const double* __restrict__ sim, double* dre, double* dim_, const int* __restrict__ perm, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { dre[i] = sre[perm[i]]; if (sim) dim_[i] = sim[perm[i]]; }
}
__global__ void __launch_bounds__(256) k_trsv_diag(const double* __restrict__ Are, const double* __restrict__ Aim, long long lda, int kb, int nbw, double* bre, double* bim, int lower) {
  extern __shared__ __align__(16) double sdiag[];   // [2][TS][TS+1]
  double* sr = sdiag; double* si = sdiag + TS * (TS + 1);
  __shared__ double yr[TS], yi[TS];
  const int i = threadIdx.x;
  for (int idx = threadIdx.x; idx < nbw * nbw; idx += blockDim.x) {      // 256 threads fetch the block; the first TS solve
    const int j = idx / nbw, ii = idx - j * nbw;
    sr[j * (TS + 1) + ii] = Are[(long long)(kb + j) * lda + kb + ii]; si[j * (TS + 1) + ii] = Aim ? Aim[(long long)(kb + j) * lda + kb + ii] : 0.0;
  }
  double vr = 0.0, vi = 0.0;
  if (i < nbw) { vr = bre[kb + i]; vi = bim ? bim[kb + i] : 0.0; }
  __syncthreads();
  if (lower) {
    for (int j = 0; j < nbw; j++) {
      if (i == j) { yr[j] = vr; yi[j] = vi; }
      __syncthreads();
      if (i > j && i < nbw) {
        double lr = sr[j * (TS + 1) + i], li = si[j * (TS + 1) + i];
        vr -= lr * yr[j] - li * yi[j]; vi -= lr * yi[j] + li * yr[j];
      }
    }
  } else {
    for (int j = nbw - 1; j >= 0; j--) {
      if (i == j) {
        double ur = sr[j * (TS + 1) + j], ui = si[j * (TS + 1) + j];
        double qr, qi;   // (vr + i vi)/(ur + i ui), Smith
        if (fabs(ur) >= fabs(ui)) { double t = ui / ur, d = ur + ui * t; qr = (vr + vi * t) / d; qi = (vi - vr * t) / d; }
        else { double t = ur / ui, d = ur * t + ui; qr = (vr * t + vi) / d; qi = (vi * t - vr) / d; }
        vr = qr; vi = qi; yr[j] = vr; yi[j] = vi;
      }
      __syncthreads();
      if (i < j) {
        double ur = sr[j * (TS + 1) + i], ui = si[j * (TS + 1) + i];
        vr -= ur * yr[j] - ui * yi[j]; vi -= ur * yi[j] + ui * yr[j];
      }
    }
  }
  if (i < nbw) { bre[kb + i] = vr; if (bim) bim[kb + i] = vi; }
}
// b[r0:r1) -= A[r0:r1, kb:kb+nbw) * x[kb:kb+nbw); CTA = 64 rows x 4 column groups
__global__ void __launch_bounds__(256) k_gemv_update(const double* __restrict__ Are, const double* __restrict__ Aim, long long lda, int r0, int r1, int kb, int nbw, double* bre, double* bim) {
  __shared__ double xr[TS], xi[TS];
  __shared__ double pr[4][64], pi[4][64];
  const int tid = threadIdx.x, li = tid & 63, cg_ = tid >> 6;
  if (tid < nbw) { xr[tid] = bre[kb + tid]; xi[tid] = bim ? bim[kb + tid] : 0.0; }
  __syncthreads();
  const int i = r0 + blockIdx.x * 64 + li;
  double sr = 0.0, si = 0.0;
  if (i < r1) {
    const int per = (nbw + 3) / 4, j0 = cg_ * per, j1 = min(j0 + per, nbw);
#pragma unroll 8
    for (int j = j0; j < j1; j++) {
      double ar = Are[(long long)(kb + j) * lda + i], ai = Aim ? Aim[(long long)(kb + j) * lda + i] : 0.0;
      sr += ar * xr[j] - ai * xi[j]; si += ar * xi[j] + ai * xr[j];
    }
  }
  pr[cg_][li] = sr; pi[cg_][li] = si;
  __syncthreads();
  if (cg_ == 0 && i < r1) {
    bre[i] -= pr[0][li] + pr[1][li] + pr[2][li] + pr[3][li];
    if (bim) bim[i] -= pi[0][li] + pi[1][li] + pi[2][li] + pi[3][li];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Solves with precomputed inverses of the TS x TS diagonal blocks (computed once per factorisation): a substitution step is
// then one small matrix-vector product instead of a TS-step serial chain, fused into the kernel that updates the rows
// below / above.  inv layout: [block][0 = L, 1 = U][plane][TS*TS] column-major.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TS) k_trtri_blocks(const double* __restrict__ Are, const double* __restrict__ Aim, long long lda, int n, double* __restrict__ inv) {
  extern __shared__ __align__(16) double sd[];      // D[2][TS][TS+1], X[2][TS][TS+1]
  double* dr = sd; double* di = sd + TS * (TS + 1); double* xr = sd + 2 * TS * (TS + 1); double* xi = sd + 3 * TS * (TS + 1);
  const int blk = blockIdx.x, upper = blockIdx.y, kb = blk * TS, nbw = min(TS, n - kb), j = threadIdx.x;
  for (int c = 0; c < nbw; c++) if (j < nbw) { dr[j * (TS + 1) + c] = Are[(long long)(kb + c) * lda + kb + j]; di[j * (TS + 1) + c] = Aim ? Aim[(long long)(kb + c) * lda + kb + j] : 0.0; }
  for (int i = 0; i < TS; i++) { xr[i * (TS + 1) + j] = 0.0; xi[i * (TS + 1) + j] = 0.0; }     // X[i][j], column j is private to thread j
  __syncthreads();
  if (j < nbw) {
    if (!upper) {                                   // unit lower: x_j = 1, x_i = -sum_{k=j}^{i-1} L[i][k] x_k
      xr[j * (TS + 1) + j] = 1.0;
      for (int i = j + 1; i < nbw; i++) {
        double sr = 0.0, si = 0.0;
        for (int k = j; k < i; k++) {
          const double lr = dr[i * (TS + 1) + k], li = di[i * (TS + 1) + k], vr = xr[k * (TS + 1) + j], vi = xi[k * (TS + 1) + j];
          sr += lr * vr - li * vi; si += lr * vi + li * vr;
        }
        xr[i * (TS + 1) + j] = -sr; xi[i * (TS + 1) + j] = -si;
      }
    } else {                                        // upper: x_j = 1/U[j][j], x_i = -(sum_{k=i+1}^{j} U[i][k] x_k) / U[i][i]
      auto cdiv = [](double ar, double ai, double br, double bi, double& qr, double& qi) {
        if (fabs(br) >= fabs(bi)) { double t = bi / br, d = br + bi * t; qr = (ar + ai * t) / d; qi = (ai - ar * t) / d; }
        else { double t = br / bi, d = br * t + bi; qr = (ar * t + ai) / d; qi = (ai * t - ar) / d; }
      };
      double qr, qi; cdiv(1.0, 0.0, dr[j * (TS + 1) + j], di[j * (TS + 1) + j], qr, qi);
      xr[j * (TS + 1) + j] = qr; xi[j * (TS + 1) + j] = qi;
      for (int i = j - 1; i >= 0; i--) {
        double sr = 0.0, si = 0.0;
        for (int k = i + 1; k <= j; k++) {
          const double ur = dr[i * (TS + 1) + k], ui = di[i * (TS + 1) + k], vr = xr[k * (TS + 1) + j], vi = xi[k * (TS + 1) + j];
          sr += ur * vr - ui * vi; si += ur * vi + ui * vr;
        }
        cdiv(-sr, -si, dr[i * (TS + 1) + i], di[i * (TS + 1) + i], qr, qi);
        xr[i * (TS + 1) + j] = qr; xi[i * (TS + 1) + j] = qi;
      }
    }
  }
  __syncthreads();
  double* o = inv + ((size_t)blk * 2 + upper) * 2 * TS * TS;
  for (int c = 0; c < TS; c++) { o[c * TS + j] = xr[j * (TS + 1) + c]; o[TS * TS + c * TS + j] = xi[j * (TS + 1) + c]; }   // o[col c][row j]
}
// one block step: x_k = inv(D_k) w_k (redone by every CTA; CTA 0 stores it to xout), then w[rows of this CTA] -= A[rows, kb:kb+nbw) x_k
__global__ void __launch_bounds__(256) k_solve_step(const double* __restrict__ Are, const double* __restrict__ Aim, long long lda, const double* __restrict__ invb, int kb, int nbw,
                                                    int r0, int r1, double* wre, double* wim, double* xore, double* xoim) {
  __shared__ double xr[TS], xi[TS], vr[TS], vi[TS];
  __shared__ double pr[4][TS], pi[4][TS];
  const int tid = threadIdx.x, li = tid & 63, part = tid >> 6;
  if (tid < TS) { vr[tid] = (tid < nbw) ? wre[kb + tid] : 0.0; vi[tid] = (tid < nbw && wim) ? wim[kb + tid] : 0.0; }
  __syncthreads();
  {
    double sr = 0.0, si = 0.0;
    const double* ir = invb; const double* ii = invb + TS * TS;
#pragma unroll 4
    for (int c = part * 16; c < part * 16 + 16; c++) {
      const double ar = ir[c * TS + li], ai = ii[c * TS + li];
      sr += ar * vr[c] - ai * vi[c]; si += ar * vi[c] + ai * vr[c];
    }
    pr[part][li] = sr; pi[part][li] = si;
  }
  __syncthreads();
  if (tid < TS) { xr[tid] = pr[0][tid] + pr[1][tid] + pr[2][tid] + pr[3][tid]; xi[tid] = pi[0][tid] + pi[1][tid] + pi[2][tid] + pi[3][tid]; }
  __syncthreads();
  if (blockIdx.x == 0 && tid < nbw) { xore[kb + tid] = xr[tid]; if (xoim) xoim[kb + tid] = xi[tid]; }
  const int i = r0 + blockIdx.x * 64 + li;
  double sr = 0.0, si = 0.0;
  if (i < r1) {
    const int per = (nbw + 3) / 4, j0 = part * per, j1 = min(j0 + per, nbw);
#pragma unroll 8
    for (int j = j0; j < j1; j++) {
      const double ar = Are[(long long)(kb + j) * lda + i], ai = Aim ? Aim[(long long)(kb + j) * lda + i] : 0.0;
      sr += ar * xr[j] - ai * xi[j]; si += ar * xi[j] + ai * xr[j];
    }
  }
  __syncthreads();
  pr[part][li] = sr; pi[part][li] = si;
  __syncthreads();
  if (part == 0 && i < r1) {
    wre[i] -= pr[0][li] + pr[1][li] + pr[2][li] + pr[3][li];
    if (wim) wim[i] -= pi[0][li] + pi[1][li] + pi[2][li] + pi[3][li];
  }
}
void lu_invert_diagonal_blocks(const double* Are, const double* Aim, long long lda, int n, double* inv, cudaStream_t st) {
  const size_t sm = (size_t)4 * TS * (TS + 1) * sizeof(double);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_trtri_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); attr = true; }
  k_trtri_blocks<<<dim3((n + TS - 1) / TS, 2), TS, sm, st>>>(Are, Aim, lda, n, inv);
}

// perm[i] = source row of row i after the interchanges ipiv[0 .. n) (what the host builds in factor_device), formed on the device by one thread in
// shared memory: lets a whole factorise + solve sequence run without a host round trip (CUDA graph of small systems).  n <= 12000.
__global__ void k_perm_from_ipiv(const int* __restrict__ ipiv, int* __restrict__ perm, int n, int* __restrict__ bad) {
  extern __shared__ int sp[];
  for (int i = threadIdx.x; i < n; i += blockDim.x) sp[i] = i;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < n; i++) {
      const int q = ipiv[i] - 1;
      if (q < i || q >= n) { atomicExch(bad, 1); continue; }      // never index with a pivot the factorisation did not produce
      if (q != i) { const int t = sp[i]; sp[i] = sp[q]; sp[q] = t; }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = sp[i];
}
int launch_perm_from_ipiv(const int* ipiv, int* perm, int n, int* bad, cudaStream_t st) {
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_perm_from_ipiv, cudaFuncAttributeMaxDynamicSharedMemorySize, 48000); attr = true; }
  if ((size_t)n * sizeof(int) > 48000) return -1;
  k_perm_from_ipiv<<<1, 256, (size_t)n * sizeof(int), st>>>(ipiv, perm, n, bad);
  return (int)cudaGetLastError();
}

int zgetrs_planar(const double* Are, const double* Aim, long long lda, int n, const int* ipiv_host_perm_dev, double* bre, double* bim, long long ldb,
                  int nrhs, cudaStream_t st, const double* inv, double* ws) {
  // ipiv_host_perm_dev: device array perm[i] = source row of row i after all interchanges (built on the host from ipiv)
  // ws: 2 n doubles of caller-owned scratch (LuWork::solve_ws): no allocation, no cudaFree (a device-wide synchronisation) and no stream
  // synchronisation inside the call, so that solves of different systems on different streams overlap
  double* tmp = ws;
  if (!tmp && cudaMalloc((void**)&tmp, (size_t)2 * n * sizeof(double)) != cudaSuccess) return (int)cudaGetLastError();
  if (inv) {   // diagonal-block inverses available: one fused launch per block step
    const int nblk = (n + TS - 1) / TS;
    for (int c = 0; c < nrhs; c++) {
      double* br = bre + (long long)c * ldb; double* bi = poff(bim, (long long)c * ldb);
      double* wr = tmp; double* wi = bim ? tmp + n : nullptr;
      k_permute<<<(n + 255) / 256, 256, 0, st>>>(br, bi, wr, wi, ipiv_host_perm_dev, n);
      for (int b = 0; b < nblk; b++) {          // L y = P b: w = tmp is the running right-hand side, y goes to b
        const int kb = b * TS, nbw = (n - kb < TS) ? n - kb : TS, r0 = kb + nbw;
        k_solve_step<<<r0 < n ? (n - r0 + 63) / 64 : 1, 256, 0, st>>>(Are, Aim, lda, inv + ((size_t)b * 2 + 0) * 2 * TS * TS, kb, nbw, r0, n, wr, wi, br, bi);
      }
      for (int b = nblk - 1; b >= 0; b--) {     // U x = y: b (holding y) is the running right-hand side, x goes to tmp
        const int kb = b * TS, nbw = (n - kb < TS) ? n - kb : TS;
        k_solve_step<<<kb > 0 ? (kb + 63) / 64 : 1, 256, 0, st>>>(Are, Aim, lda, inv + ((size_t)b * 2 + 1) * 2 * TS * TS, kb, nbw, 0, kb, br, bi, wr, wi);
      }
      cudaMemcpyAsync(br, wr, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
      if (bi) cudaMemcpyAsync(bi, wi, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
    }
    if (!ws) { cudaStreamSynchronize(st); cudaFree(tmp); }
    return (int)cudaGetLastError();
  }
  const size_t dsm = (size_t)2 * TS * (TS + 1) * sizeof(double);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_trsv_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm); attr = true; }
  for (int c = 0; c < nrhs; c++) {
    double* br = bre + (long long)c * ldb; double* bi = poff(bim, (long long)c * ldb);
    k_permute<<<(n + 255) / 256, 256, 0, st>>>(br, bi, tmp, tmp + n, ipiv_host_perm_dev, n);
    cudaMemcpyAsync(br, tmp, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
    if (bi) cudaMemcpyAsync(bi, tmp + n, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
    for (int kb = 0; kb < n; kb += TS) {
      int nbw = (n - kb < TS) ? n - kb : TS;
      k_trsv_diag<<<1, 256, dsm, st>>>(Are, Aim, lda, kb, nbw, br, bi, 1);
      int r0 = kb + nbw;
      if (r0 < n) k_gemv_update<<<(n - r0 + 63) / 64, 256, 0, st>>>(Are, Aim, lda, r0, n, kb, nbw, br, bi);
    }
    int nblk = (n + TS - 1) / TS;
    for (int b = nblk - 1; b >= 0; b--) {
      int kb = b * TS, nbw = (n - kb < TS) ? n - kb : TS;
      k_trsv_diag<<<1, 256, dsm, st>>>(Are, Aim, lda, kb, nbw, br, bi, 0);
      if (kb > 0) k_gemv_update<<<(kb + 63) / 64, 256, 0, st>>>(Are, Aim, lda, 0, kb, kb, nbw, br, bi);
    }
  }
  if (!ws) { cudaStreamSynchronize(st); cudaFree(tmp); }
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// Distributed LU over P ranks, 1-D block-cyclic columns (see lu.cuh).  The kernels are the single-GPU ones: a panel is
// factorised in place in the owner's local columns (a column-shifted base pointer lets the global-index code address the
// local storage), packed (rows k0..n only) and broadcast together with its pivots; the other ranks read L11 / L21 from
// the packed copy.
// ------------------------------------------------------------------------------------------------------------------
int dist_rank_alloc(DistRank& R, int rank, int n, long long lda, int nb, int P, cudaStream_t main_stream, bool separate_comm) {
  R.rank = rank; R.ncl = dist_ncols_local(n, nb, P, rank); R.gemm_flops = 0.0;
  R.Lre = R.Lim = nullptr; R.pbuf[0] = R.pbuf[1] = R.pbuf[2] = nullptr; R.xfin = nullptr; R.ipiv = nullptr;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
  double* L = nullptr;
  A((void**)&L, (size_t)2 * lda * (R.ncl + 1) * sizeof(double));
  R.Lre = L; R.Lim = L + (size_t)lda * (R.ncl + 1);
  for (int i = 0; i < 3; i++) A((void**)&R.pbuf[i], (size_t)2 * nb * lda * sizeof(double) + (size_t)nb * sizeof(int) + 16);
  A((void**)&R.xfin, (size_t)2 * lda * sizeof(double));
  A((void**)&R.ipiv, (size_t)n * sizeof(int));
  if (e != cudaSuccess) return (int)e;
  int le = lu_work_alloc(R.w, n, nb);
  if (le) return le;
  cudaEventCreateWithFlags(&R.ev_panel, cudaEventDisableTiming); cudaEventCreateWithFlags(&R.ev_cols, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&R.ev_free[0], cudaEventDisableTiming); cudaEventCreateWithFlags(&R.ev_free[1], cudaEventDisableTiming);
  R.main = main_stream; R.comm = separate_comm ? R.w.panel_stream : main_stream;
  return 0;
}
void dist_rank_free(DistRank& R) {
  cudaFree(R.Lre); cudaFree(R.pbuf[0]); cudaFree(R.pbuf[1]); cudaFree(R.pbuf[2]); cudaFree(R.xfin); cudaFree(R.ipiv);
  lu_work_free(R.w);
  cudaEventDestroy(R.ev_panel); cudaEventDestroy(R.ev_cols); cudaEventDestroy(R.ev_free[0]); cudaEventDestroy(R.ev_free[1]);
}

int zgetrf_dist(DistLU& D) {
  const int n = D.n, nb = D.nb, P = D.P, nblk = D.nblk, NL = (int)D.r.size();
  const long long lda = D.lda;
  std::vector<int> ranks(NL); std::vector<void*> bufs(NL); std::vector<cudaStream_t> sts(NL);
  for (int i = 0; i < NL; i++) { ranks[i] = D.r[i].rank; sts[i] = D.r[i].comm; D.r[i].gemm_flops = 0.0; D.r[i].w.launches = 0; cudaMemsetAsync(D.r[i].w.info, 0, sizeof(int), D.r[i].main); }
  // MFB_DIST_TRACE=<file prefix>: per-step CUDA events of this rank (main stream: panel received, next-panel columns updated,
  // step finished; panel stream: panel factorised + packed, broadcast finished), written as CSV after the factorisation
  const char* trace = (NL == 1) ? getenv("MFB_DIST_TRACE") : nullptr;
  std::vector<cudaEvent_t> tev;
  if (trace) { tev.resize((size_t)6 * (nblk + 1)); for (auto& e : tev) cudaEventCreate(&e); cudaEventRecord(tev[0], D.r[0].main); }
  auto mark = [&](int k, int which, cudaStream_t st) { if (trace) cudaEventRecord(tev[(size_t)6 * (k + 1) + which], st); };
  auto width = [&](int k) { return (n - k * nb < nb) ? (n - k * nb) : nb; };
  // first local column that belongs to a global block > k
  auto nright = [&](const DistRank& R, int k) { int c = (k >= R.rank) ? ((k - R.rank) / P + 1) * nb : 0; return c < R.ncl ? c : R.ncl; };
  auto panel_bytes = [&](int k) { const long long m = n - (long long)k * nb, mp = (m + 1) & ~1ll; return (size_t)2 * width(k) * mp * sizeof(double) + (size_t)width(k) * sizeof(int); };
  // owner: factorise panel k in place (stream R.comm), pack rows k0..n + pivots into pbuf[k % 3]
  auto factor_and_pack = [&](DistRank& R, int k) -> int {
    const int k0 = k * nb, nbw = width(k), lc = dist_local_col(k, P, nb);
    const long long m = n - k0, mp = (m + 1) & ~1ll;
    double* Are = R.Lre + ((long long)lc - k0) * lda; double* Aim = R.Lim + ((long long)lc - k0) * lda;   // global column k0 -> local column lc
    int e = factor_panel(Are, Aim, lda, n, k0, nbw, R.ipiv, R.w, R.comm);
    if (e) return e;
    mark(k, 3, R.comm);
    double* pb = R.pbuf[k % 3];
    cudaMemcpy2DAsync(pb, (size_t)mp * 8, R.Lre + (long long)lc * lda + k0, (size_t)lda * 8, (size_t)m * 8, nbw, cudaMemcpyDeviceToDevice, R.comm);
    cudaMemcpy2DAsync(pb + (size_t)nbw * mp, (size_t)mp * 8, R.Lim + (long long)lc * lda + k0, (size_t)lda * 8, (size_t)m * 8, nbw, cudaMemcpyDeviceToDevice, R.comm);
    cudaMemcpyAsync(pb + (size_t)2 * nbw * mp, R.ipiv + k0, (size_t)nbw * sizeof(int), cudaMemcpyDeviceToDevice, R.comm);
    return 0;
  };
  auto bcast_panel = [&](int k) -> int {
    for (int i = 0; i < NL; i++) bufs[i] = D.r[i].pbuf[k % 3];
    int e = D.comm->bcast_bytes(dist_owner(k, P), ranks.data(), bufs.data(), panel_bytes(k), sts.data(), NL);
    if (e) return e;
    const int k0 = k * nb, nbw = width(k);
    const long long m = n - k0, mp = (m + 1) & ~1ll;
    for (int i = 0; i < NL; i++) {
      DistRank& R = D.r[i];
      if (R.rank != dist_owner(k, P)) cudaMemcpyAsync(R.ipiv + k0, R.pbuf[k % 3] + (size_t)2 * nbw * mp, (size_t)nbw * sizeof(int), cudaMemcpyDeviceToDevice, R.comm);
      cudaEventRecord(R.ev_panel, R.comm);
      mark(k, 4, R.comm);
    }
    return 0;
  };
  // A[k0:n, c0:c1) of the local columns: U12 = inv(L11) A12, A22 -= L21 U12, with L from the packed panel k
  auto update_cols = [&](DistRank& R, int k, int c0, int c1) {
    if (c1 <= c0) return;
    const int k0 = k * nb, nbw = width(k);
    const long long m = n - k0, mp = (m + 1) & ~1ll;
    const double* Pre = R.pbuf[k % 3]; const double* Pim = Pre + (size_t)nbw * mp;
    double* Bre = R.Lre + (long long)c0 * lda + k0; double* Bim = R.Lim + (long long)c0 * lda + k0;
    R.w.launches += launch_trsm_ext(Pre, Pim, mp, Bre, Bim, lda, nbw, c1 - c0, R.main);
    const int mrest = (int)m - nbw;
    if (mrest > 0) {
      // TMA kernel (gemm_tma.cu): A from the packed panel (its leading dimension changes with the step, so its two maps are encoded per call: host work of a
      // microsecond), B and C from the local columns.  Besides the faster kernel this removes the re-reads of the A panel: with the old tile order every rank
      // streamed the whole 123 MB panel once per 32-column tile of its local columns.
      GemmTmaMaps tm; tm.ok = 0;
      if (gemm_cfg() == 15 && (nbw % 16) == 0)
        gemm_tma_make_maps(tm, Pre, Pim, mp, (int)m, nbw, R.Lre, R.Lim, lda, n, R.ncl + 1);
      if (gemm_tma_usable(tm, mrest, c1 - c0, nbw))
        zgemm_minus_planar_tma(tm, mrest, c1 - c0, nbw, nbw, 0, k0, c0, Bre + nbw, Bim + nbw, lda, R.main);
      else
        zgemm_minus_planar(mrest, c1 - c0, nbw, Pre + nbw, Pim + nbw, mp, Bre, Bim, lda, Bre + nbw, Bim + nbw, lda, R.main);
      R.w.launches++; R.gemm_flops += 8.0 * (double)mrest * (double)(c1 - c0) * (double)nbw;
    }
  };
  // ---- panel 0 ----
  for (int i = 0; i < NL; i++) {
    DistRank& R = D.r[i];
    if (R.comm != R.main) { cudaEventRecord(R.ev_free[0], R.main); cudaStreamWaitEvent(R.comm, R.ev_free[0], 0); }   // the local columns are complete
    if (R.rank == dist_owner(0, P)) { int e = factor_and_pack(R, 0); if (e) return e; }
  }
  { int e = bcast_panel(0); if (e) return e; }
  for (int k = 0; k < nblk; k++) {
    const int k0 = k * nb, nbw = width(k);
    const bool has_next = k + 1 < nblk;
    for (int i = 0; i < NL; i++) {
      DistRank& R = D.r[i];
      // three panel buffers: the broadcast of panel k+1 overwrites the buffer last read by step k-2, i.e. it may start as soon as
      // this rank's main stream has begun step k-1 (event recorded there) -- one more step of slack than with two buffers
      if (R.comm != R.main) cudaEventRecord(R.ev_free[k & 1], R.main);
      cudaStreamWaitEvent(R.main, R.ev_panel, 0);
      mark(k, 0, R.main);
      // ---- interchanges of panel k on every local column outside the panel (the right-hand side column included) ----
      const bool mine = R.rank == dist_owner(k, P);
      const int lc = dist_local_col(k, P, nb), ctot = R.ncl + 1;
      const int a1 = mine ? lc : ctot;                // [0, a1) and [a0, ctot)
      const int a0 = mine ? lc + nbw : ctot;
      if (a1 > 0) { k_laswp<<<(a1 + 127) / 128, 128, 0, R.main>>>(R.Lre, R.Lim, lda, 0, a1, k0, nbw, R.ipiv); R.w.launches++; }
      if (a0 < ctot) { k_laswp<<<(ctot - a0 + 127) / 128, 128, 0, R.main>>>(R.Lre, R.Lim, lda, a0, ctot, k0, nbw, R.ipiv); R.w.launches++; }
      // ---- look-ahead: the owner of panel k+1 updates that panel's columns first and factorises it on the other stream ----
      if (has_next && R.rank == dist_owner(k + 1, P)) {
        const int cr = nright(R, k), nbw_next = width(k + 1);
        update_cols(R, k, cr, cr + nbw_next);
        mark(k, 1, R.main);
        if (R.comm != R.main) { cudaEventRecord(R.ev_cols, R.main); cudaStreamWaitEvent(R.comm, R.ev_cols, 0); }
        int e = factor_and_pack(R, k + 1); if (e) return e;
        // The panel is the critical path of the distributed factorisation (every rank waits for its broadcast), and run beside
        // this rank's own trailing update it takes three times as long (traced: 3.5 instead of 1.2 ms at m = 20k): the owner
        // holds its trailing update back until the panel is packed.  It pays for it one step in eight.
        if (R.comm != R.main && D.owner_waits_for_panel) { cudaEventRecord(R.ev_cols, R.comm); cudaStreamWaitEvent(R.main, R.ev_cols, 0); }
      } else if (has_next && R.comm != R.main && k >= 1) cudaStreamWaitEvent(R.comm, R.ev_free[(k - 1) & 1], 0);
    }
    if (has_next) { int e = bcast_panel(k + 1); if (e) return e; }
    for (int i = 0; i < NL; i++) {
      DistRank& R = D.r[i];
      int cr = nright(R, k);
      if (has_next && R.rank == dist_owner(k + 1, P)) cr += width(k + 1);
      update_cols(R, k, cr, R.ncl + 1);
      mark(k, 2, R.main);
    }
  }
  if (trace) {
    cudaStreamSynchronize(D.r[0].main); cudaStreamSynchronize(D.r[0].comm);
    char fn[512]; snprintf(fn, sizeof(fn), "%s_rank%d.csv", trace, D.r[0].rank);
    if (FILE* f = fopen(fn, "w")) {
      fprintf(f, "step,owner,t_panel_received,t_next_cols_updated,t_step_done,t_panel_factorised,t_bcast_done\n");
      for (int k = 0; k < nblk; k++) {
        fprintf(f, "%d,%d", k, dist_owner(k, P));
        for (int w = 0; w < 5; w++) {
          float t = -1.f;
          if (cudaEventQuery(tev[(size_t)6 * (k + 1) + w]) == cudaSuccess && cudaEventElapsedTime(&t, tev[0], tev[(size_t)6 * (k + 1) + w]) != cudaSuccess) t = -1.f;
          fprintf(f, ",%.3f", t);
        }
        fprintf(f, "\n");
      }
      fclose(f);
    }
    cudaGetLastError();
    for (auto& e : tev) cudaEventDestroy(e);
  }
  return (int)cudaGetLastError();
}

__global__ void k_vec_copy2(double* dre, double* dim_, const double* __restrict__ sre, const double* __restrict__ sim, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { dre[i] = sre[i]; dim_[i] = sim[i]; }
}

// U x = y by column blocks from the last to the first: the partial sums  -U(:, j) x_j  of the blocks already solved live
// with the rank that owns column block j; the block's right-hand side is their sum over the ranks (one small reduce per
// block), the owner solves its diagonal block and updates its own partial sums above it.
int zgetrs_dist(DistLU& D) {
  const int n = D.n, nb = D.nb, P = D.P, nblk = D.nblk, NL = (int)D.r.size();
  const long long lda = D.lda;
  std::vector<int> ranks(NL); std::vector<double*> a(NL), b(NL); std::vector<cudaStream_t> sts(NL);
  const size_t dsm = (size_t)2 * TS * (TS + 1) * sizeof(double);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_trsv_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm); attr = true; }
  for (int i = 0; i < NL; i++) {
    DistRank& R = D.r[i];
    ranks[i] = R.rank; sts[i] = R.main;
    double* vre = R.Lre + (long long)R.ncl * lda; double* vim = R.Lim + (long long)R.ncl * lda;
    if (R.rank != 0) { cudaMemsetAsync(vre, 0, (size_t)lda * 8, R.main); cudaMemsetAsync(vim, 0, (size_t)lda * 8, R.main); }   // y is counted once
    cudaMemsetAsync(R.xfin, 0, (size_t)2 * lda * 8, R.main);
  }
  for (int k = nblk - 1; k >= 0; k--) {
    const int k0 = k * nb, nbw = (n - k0 < nb) ? (n - k0) : nb, own = dist_owner(k, P);
    for (int i = 0; i < NL; i++) { a[i] = D.r[i].Lre + (long long)D.r[i].ncl * lda + k0; b[i] = D.r[i].Lim + (long long)D.r[i].ncl * lda + k0; }
    if (P > 1) { int e = D.comm->reduce_sum2(own, ranks.data(), a.data(), b.data(), (size_t)nbw, sts.data(), NL); if (e) return e; }
    for (int i = 0; i < NL; i++) {
      DistRank& R = D.r[i];
      if (R.rank != own) continue;
      const int lc = dist_local_col(k, P, nb);
      const double* Are = R.Lre + ((long long)lc - k0) * lda; const double* Aim = R.Lim + ((long long)lc - k0) * lda;
      double* vre = R.Lre + (long long)R.ncl * lda; double* vim = R.Lim + (long long)R.ncl * lda;
      const int nsub = (nbw + TS - 1) / TS;
      for (int sb = nsub - 1; sb >= 0; sb--) {
        const int kb = k0 + sb * TS, w = (k0 + nbw - kb < TS) ? (k0 + nbw - kb) : TS;
        k_trsv_diag<<<1, 256, dsm, R.main>>>(Are, Aim, lda, kb, w, vre, vim, 0);
        if (kb > 0) k_gemv_update<<<(kb + 63) / 64, 256, 0, R.main>>>(Are, Aim, lda, 0, kb, kb, w, vre, vim);
      }
      k_vec_copy2<<<(nbw + 255) / 256, 256, 0, R.main>>>(R.xfin + k0, R.xfin + lda + k0, vre + k0, vim + k0, nbw);
    }
  }
  if (P > 1) {
    for (int i = 0; i < NL; i++) a[i] = D.r[i].xfin;
    int e = D.comm->allreduce_sum(ranks.data(), a.data(), (size_t)2 * lda, sts.data(), NL); if (e) return e;
  }
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
