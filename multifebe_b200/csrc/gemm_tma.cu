// gemm_tma.cu -- trailing update of the LU (C -= A*B, planar complex, 3M form) with TMA operand staging, sm_100a.
//
// Same arithmetic as k_zgemm3m_minus of lu.cu (three real products X = Ar*Br, Y = Ai*Bi, Z = (Ar+Ai)*(Br+Bi) on the FP64 tensor pipe,
// mma.sync m8n8k4 = SASS DMMA.8x8x4; tcgen05 has no FP64 kind), but the operand tiles reach shared memory through the TMA
// (cp.async.bulk.tensor, SASS UTMALDG) instead of per-thread cp.async:
//   * the round-1 kernel spent ~190 integer / predicate instructions per k-tile and thread on the addresses and bounds of its 12 LDGSTS,
//     between a __syncthreads and the first DMMA of the tile (ncu source page, profiles/r02_ncu_zgemm_20k_sass.txt); here ONE thread issues
//     12 bulk tensor copies per k-tile, bounds are the tensor map's (out-of-range rows are zero-filled), and the consumers only run
//     LDS + DMMA;
//   * stages are handed over with mbarriers (full: TMA transaction bytes; empty: one arrival per warp), no block-wide barrier in the loop;
//   * the tiles are stored with the 128-byte TMA swizzle and the rows / columns of the 8x8x4 fragments are PERMUTED inside the warp tile so
//     that every fragment load is bank-conflict free without padding (derivation below);
//   * CTAs are rasterised in strips of GT_STRIP n-tiles (all m-tiles of a strip before the next strip): the B strip stays in L2 while the
//     A panel streams through once per strip.  With the plain (m fastest, then n) order the whole A panel (123 MB at 30k DOF, about the
//     size of the L2) cycled through the cache once per n-tile: ncu measured 50 GB of DRAM reads for a 20480 x 20480 x 256 update whose
//     operands and C tile are 7 GB.
//
// Shared-memory layout of a stage (24 KB): A_re | A_im (each 4 blocks of [16 k][16 m] doubles = 2 KB, one TMA box each) | B_re | B_im
// (each 2 blocks of [16 n][16 k]).  A block row is 128 bytes = 32 banks; SWIZZLE_128B XORs the 16-byte chunk index (address bits 4-6) with
// the row index mod 8 (bits 7-9).  An A fragment load reads element (k = 4*k4 + tig, m) for the 16 lanes (4 values of gid, 4 of tig) of a
// half warp: its 8-byte slot inside the row is p = m_in ^ ((k & 7) << 1) with m_in = m mod 16, so tig moves bits 1-2 of p and the four
// rows of the half warp must differ in bits 0 and 3: gid -> m_in = g0 | g2 << 1 | ml << 2 | g1 << 3 (g0..g2 the bits of gid, ml the parity
// of the 8-row mma tile inside the 16-row block).  A B fragment reads (n, k = 4*k4 + tig): p = k ^ ((n & 7) << 1), tig owns bits 0-1, so
// the four columns of a half warp must differ in bits 2-3 of p: gid -> n mod 8 = (gid & 3) << 1 | gid >> 2.  The C fragment of a lane
// follows the same two permutations (row gid, columns 2*tig and 2*tig + 1 of the mma tile).
#include "lu.cuh"
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mfbd {

constexpr int GT_BM = 64, GT_BN = 32, GT_BK = 16, GT_STRIP = 32;
constexpr int GT_BLOCK = 16 * 16 * 8;                               // one TMA box: 16 x 16 doubles
constexpr int GT_A_PLANE = (GT_BM / 16) * GT_BLOCK;                 // 8 KB
constexpr int GT_B_PLANE = (GT_BN / 16) * GT_BLOCK;                 // 4 KB
constexpr int GT_STAGE_BYTES = 2 * GT_A_PLANE + 2 * GT_B_PLANE;     // 24 KB
constexpr int gt_smem(int stages) { return stages * GT_STAGE_BYTES + 1024 + 2 * stages * 8; }

__device__ __forceinline__ unsigned gt_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gt_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void gt_mbar_init(unsigned bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void gt_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gt_mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void gt_mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void gt_tma_box(unsigned dst, const CUtensorMap* map, int row, int col, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(row), "r"(col), "r"(bar)
               : "memory");
}

// A operand = rows [a_row0 + ..), columns [a_col0, a_col0 + K) of the planes behind tAre / tAim; B operand = rows [b_row0, b_row0 + K),
// columns [b_col0 + ..) of the planes behind tBre / tBim (tensor maps of whole planes: rows x columns, box 16 x 16, SWIZZLE_128B).
// K is a multiple of 16; a_row0 any row whose tiles the map covers (out-of-range rows are zero-filled by the TMA and never stored).
// GT_STAGES / MINB: depth of the operand ring and CTAs per SM.  (3, 3) = 12 warps per SM, (4, 2) = 8 warps per SM: the DMMA micro-benchmark with this
// kernel's instruction mix (tools/microbench/dmma_sweep.cu, profiles/r02_dmma_sweep.log) runs 3.6 % faster with 2 or 4 warps per sub-partition than with 3.
template <int GT_STAGES, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_zgemm3m_tma(const __grid_constant__ CUtensorMap tAre, const __grid_constant__ CUtensorMap tAim, const __grid_constant__ CUtensorMap tBre,
              const __grid_constant__ CUtensorMap tBim, int M, int N, int K, int a_row0, int a_col0, int b_row0, int b_col0, double* __restrict__ Cre,
              double* __restrict__ Cim, long long ldc, int mt, int nt, int pf_dist) {
  extern __shared__ unsigned char gt_raw[];
  const unsigned raw = gt_smem_u32(gt_raw), base = (raw + 1023u) & ~1023u;      // SWIZZLE_128B atoms are 1024-byte aligned
  const unsigned bars = base + GT_STAGES * GT_STAGE_BYTES;                       // full[s] at bars + 8 s, empty[s] at bars + 8 (STAGES + s)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  const int gid = lane >> 2, tig = lane & 3;
  // strip rasterisation: tiles of GT_STRIP consecutive n-tiles, m-tiles outer inside the strip
  int mtile, ntile;
  auto tile_of = [&](int t, int& mtl, int& ntl) {
    const int per = GT_STRIP * mt, strip = t / per, w = t - strip * per;
    const int sw = min(GT_STRIP, nt - strip * GT_STRIP);
    mtl = w / sw; ntl = strip * GT_STRIP + (w - mtl * sw);
  };
  tile_of(blockIdx.x, mtile, ntile);
  const int m0 = mtile * GT_BM, n0 = ntile * GT_BN;
  const int KT = K / GT_BK;
  // The C tile is read once, at the start of a CTA, straight from DRAM (~1.5 us during which this CTA issues no DMMA: ~4 % of warp time in the ncu source
  // page).  Each CTA therefore pulls into L2 the C tile of the CTA that will take its place: pf_dist tiles ahead (about the number of co-resident CTAs).
  if (pf_dist > 0 && (int)blockIdx.x + pf_dist < mt * nt) {
    int pm, pn; tile_of((int)blockIdx.x + pf_dist, pm, pn);
    // 64 rows x 32 columns x 2 planes: a column of the tile is 512 bytes = 4 lines; thread -> (plane, column, line)
    for (int q = tid; q < 2 * GT_BN * 4; q += 128) {
      const int plane = q / (GT_BN * 4), rem = q - plane * (GT_BN * 4), col = rem >> 2, line = rem & 3;
      const int rr = pm * GT_BM + line * 16, cc = pn * GT_BN + col;
      if (rr < M && cc < N) asm volatile("prefetch.global.L2 [%0];" ::"l"((plane ? Cim : Cre) + (long long)cc * ldc + rr));
    }
  }

  auto issue = [&](int s, int kt) {
    const unsigned full = bars + 8u * s, st = base + (unsigned)s * GT_STAGE_BYTES;
    gt_mbar_expect_tx(full, GT_STAGE_BYTES);
    const int ac = a_col0 + kt * GT_BK, br = b_row0 + kt * GT_BK;
#pragma unroll
    for (int q = 0; q < GT_BM / 16; q++) {
      gt_tma_box(st + q * GT_BLOCK, &tAre, a_row0 + m0 + 16 * q, ac, full);
      gt_tma_box(st + GT_A_PLANE + q * GT_BLOCK, &tAim, a_row0 + m0 + 16 * q, ac, full);
    }
#pragma unroll
    for (int q = 0; q < GT_BN / 16; q++) {
      gt_tma_box(st + 2 * GT_A_PLANE + q * GT_BLOCK, &tBre, br, b_col0 + n0 + 16 * q, full);
      gt_tma_box(st + 2 * GT_A_PLANE + GT_B_PLANE + q * GT_BLOCK, &tBim, br, b_col0 + n0 + 16 * q, full);
    }
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GT_STAGES; s++) { gt_mbar_init(bars + 8u * s, 1); gt_mbar_init(bars + 8u * (GT_STAGES + s), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0)
    for (int s = 0; s < GT_STAGES && s < KT; s++) issue(s, s);

  // rows / columns of this lane's fragments (see the header)
  const int min0 = (gid & 1) | ((gid >> 2) << 1) | (((gid >> 1) & 1) << 3);       // m mod 16 for the even mma tile of a 16-row block; the odd one adds 4
  int mrow[4], ncol[2][2];
#pragma unroll
  for (int mi = 0; mi < 4; mi++) mrow[mi] = m0 + wm * 32 + (mi >> 1) * 16 + (min0 | ((mi & 1) << 2));
#pragma unroll
  for (int ni = 0; ni < 2; ni++)
#pragma unroll
    for (int h = 0; h < 2; h++) { const int j = 2 * tig + h; ncol[ni][h] = n0 + wn * 16 + ni * 8 + (((j & 3) << 1) | (j >> 2)); }

  // C enters through the accumulators (as in k_zgemm3m_minus): x = -Cr + X, y = Y, z = -(Ci + Cr) + Z  =>  Cr' = y - x, Ci' = x + y - z
  double x[4][2][2], y[4][2][2], z[4][2][2];
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < 2; ni++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const bool ok = (mrow[mi] < M) && (ncol[ni][h] < N);
        const long long o = (long long)ncol[ni][h] * ldc + mrow[mi];
        const double cr = ok ? Cre[o] : 0.0, ci = ok ? Cim[o] : 0.0;
        x[mi][ni][h] = -cr; y[mi][ni][h] = 0.0; z[mi][ni][h] = -(ci + cr);
      }

  // byte offsets of this lane inside a stage.  The k4- and tile-dependent parts of the swizzle are XORs of bits 5-6, folded into a few per-lane
  // variants so that every fragment load is [register + immediate]
  const int cA = (((gid >> 1) & 1) << 2) | (gid >> 2);                             // (m_in >> 1) of the even mma tile
  const unsigned aoff0 = (unsigned)(wm * 2 * GT_BLOCK + tig * 128 + ((cA ^ tig) << 4) + ((gid & 1) << 3));
  const int nl = ((gid & 3) << 1) | (gid >> 2);
  const unsigned boff0 = (unsigned)(2 * GT_A_PLANE + wn * GT_BLOCK + nl * 128 + ((((tig >> 1) ^ nl) & 7) << 4) + ((tig & 1) << 3));
  unsigned aoffv[2][2], boffv[4];
#pragma unroll
  for (int ml = 0; ml < 2; ml++)
#pragma unroll
    for (int ko = 0; ko < 2; ko++) aoffv[ml][ko] = aoff0 ^ (unsigned)((ml << 5) | (ko << 6));   // odd mma tile: m_in + 4 flips bit 1 of the chunk; odd k4: bit 2
#pragma unroll
  for (int k4 = 0; k4 < 4; k4++) boffv[k4] = boff0 ^ (unsigned)(k4 << 5);                        // chunk = ((tig >> 1) ^ nl) ^ (2 k4)
  const unsigned char* stage0 = gt_raw + (base - raw);

  for (int kt = 0; kt < KT; kt++) {
    const int s = kt % GT_STAGES;
    if (kt >= 1) {
      // Release of the stage consumed in the PREVIOUS iteration.  It must not be signalled right after that iteration's last LDS: the arrival
      // does not wait for loads in flight, ptxas hoists it above the DMMAs that consume them, and the refill then overwrote fragments that had
      // been requested but not yet read (first version of this kernel: one tile in ~600 wrong, always the last-read block of the straggling
      // warp; profiles/r02_gemm_tma_race.md).  Here every load of that iteration has completed: its value fed a DMMA issued before the loop branch.
      const int sp = (kt - 1) % GT_STAGES;
      if (lane == 0) gt_mbar_arrive(bars + 8u * (GT_STAGES + sp));
      if (tid == 0 && kt - 1 + GT_STAGES < KT) {            // refill it with the k-tile GT_STAGES ahead
        gt_mbar_wait(bars + 8u * (GT_STAGES + sp), (unsigned)(((kt - 1) / GT_STAGES) & 1));
        issue(sp, kt - 1 + GT_STAGES);
      }
    }
    __syncwarp();
    gt_mbar_wait(bars + 8u * s, (unsigned)((kt / GT_STAGES) & 1));
    const unsigned char* st = stage0 + (size_t)s * GT_STAGE_BYTES;
#pragma unroll
    for (int k4 = 0; k4 < GT_BK / 4; k4++) {
      double ar[4], ai[4], sa[4];
#pragma unroll
      for (int mi = 0; mi < 4; mi++) {
        const unsigned char* a = st + aoffv[mi & 1][k4 & 1] + (mi >> 1) * GT_BLOCK + k4 * 512;
        ar[mi] = *reinterpret_cast<const double*>(a); ai[mi] = *reinterpret_cast<const double*>(a + GT_A_PLANE);
        sa[mi] = ar[mi] + ai[mi];
      }
#pragma unroll
      for (int ni = 0; ni < 2; ni++) {
        const unsigned char* b = st + boffv[k4] + ni * 1024;
        const double br = *reinterpret_cast<const double*>(b), bi = *reinterpret_cast<const double*>(b + GT_B_PLANE), sb = br + bi;
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
          gt_dmma(x[mi][ni][0], x[mi][ni][1], ar[mi], br);
          gt_dmma(y[mi][ni][0], y[mi][ni][1], ai[mi], bi);
          gt_dmma(z[mi][ni][0], z[mi][ni][1], sa[mi], sb);
        }
      }
    }
  }
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < 2; ni++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        if (mrow[mi] < M && ncol[ni][h] < N) {
          const long long o = (long long)ncol[ni][h] * ldc + mrow[mi];
          Cre[o] = y[mi][ni][h] - x[mi][ni][h];
          Cim[o] = (x[mi][ni][h] + y[mi][ni][h]) - z[mi][ni][h];
        }
      }
}

// ---- host side -------------------------------------------------------------------------------------------------------
typedef CUresult (*GtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static GtEncodeFn gt_encode() {
  static GtEncodeFn fn = nullptr; static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = (GtEncodeFn)p;
    cudaGetLastError();
  }
  return fn;
}
static_assert(sizeof(CUtensorMap) == 128 && alignof(GemmTmaMaps) >= 64, "GemmTmaMaps layout");

// one plane (rows x cols doubles, column-major, leading dimension ld) -> 2-D tensor map, box 16 x 16, SWIZZLE_128B
static int gt_plane_map(void* out, const double* plane, long long ld, int rows, int cols) {
  GtEncodeFn enc = gt_encode();
  if (!enc || !plane || (ld & 1) || (reinterpret_cast<uintptr_t>(plane) & 15)) return 1;
  cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {16, 16}, estr[2] = {1, 1};
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(plane), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}
int gemm_tma_make_maps(GemmTmaMaps& t, const double* Are, const double* Aim, long long lda, int a_rows, int a_cols, const double* Bre, const double* Bim, long long ldb,
                       int b_rows, int b_cols) {
  t.ok = 0;
  if (const char* e = getenv("MFB_GEMM_TMA")) if (e[0] == '0') return 1;
  if (gt_plane_map(t.m[0], Are, lda, a_rows, a_cols) || gt_plane_map(t.m[1], Aim, lda, a_rows, a_cols) || gt_plane_map(t.m[2], Bre, ldb, b_rows, b_cols) ||
      gt_plane_map(t.m[3], Bim, ldb, b_rows, b_cols))
    return 2;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(k_zgemm3m_tma<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, gt_smem(3)) != cudaSuccess ||
        cudaFuncSetAttribute(k_zgemm3m_tma<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, gt_smem(4)) != cudaSuccess) { cudaGetLastError(); return 3; }
    attr = true;
  }
  t.ok = 1;
  return 0;
}
bool gemm_tma_usable(const GemmTmaMaps& t, int m, int n, int k) { return t.ok && m > 0 && n > 0 && k >= GT_BK && (k % GT_BK) == 0; }
void zgemm_minus_planar_tma(const GemmTmaMaps& t, int m, int n, int k, int a_row0, int a_col0, int b_row0, int b_col0, double* Cre, double* Cim, long long ldc, cudaStream_t st) {
  const int mt = (m + GT_BM - 1) / GT_BM, nt = (n + GT_BN - 1) / GT_BN;
  const CUtensorMap* mp = reinterpret_cast<const CUtensorMap*>(t.m);
  static int pf = -1, cfg = -1;
  if (pf < 0) { const char* e = getenv("MFB_GEMM_C_PREFETCH"); pf = e ? atoi(e) : 3 * 148; }
  if (cfg < 0) { const char* e = getenv("MFB_GEMM_TMA_CFG"); cfg = e ? atoi(e) : 0; }
  if (cfg == 1) k_zgemm3m_tma<4, 2><<<mt * nt, 128, gt_smem(4), st>>>(mp[0], mp[1], mp[2], mp[3], m, n, k, a_row0, a_col0, b_row0, b_col0, Cre, Cim, ldc, mt, nt, pf > 0 ? 2 * 148 : 0);
  else k_zgemm3m_tma<3, 3><<<mt * nt, 128, gt_smem(3), st>>>(mp[0], mp[1], mp[2], mp[3], m, n, k, a_row0, a_col0, b_row0, b_col0, Cre, Cim, ldc, mt, nt, pf);
}

}  // namespace mfbd
