// combine.cu -- column combination of device-resident systems: the coupled-region system built from single-region assemblies
// (multifebe_b200/host/coupled.py) without leaving the device.
//   k_combine      dst(row_map[r], dst_col[i]) += coef[i] * src(r, src_col[i])   for every term i and source row r (dst_col == -1: right-hand side)
//   k_add_entries  dst(rows[i], cols[i]) += v[i]                                  (free terms; cols == -1: right-hand side)
// All indices arrive already translated to the internal (permuted) order of each system.  Several terms may hit the same destination entry
// (three displacement columns feeding one pressure column ...), hence RED.ADD.
// Parity: tests/test_gpu_coupled.py (resident=True) against the multi-region oracle, green on a B200 (profiles/r02_first_contact.log).
#include "combine.cuh"

namespace mfbd {

__global__ void k_combine(DevSystem src, DevSystem dst, int n_rows, const int* __restrict__ src_row, const int* __restrict__ dst_row, int n_terms,
                          const int* __restrict__ src_col, const int* __restrict__ dst_col, const double* __restrict__ coef) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int sr = src_row[r], dr = dst_row[r];
  for (int i = blockIdx.y; i < n_terms; i += gridDim.y) {
    const size_t so = (size_t)src_col[i] * src.lda + sr;
    const double ar = src.Are[so], ai = src.Aim[so], cr = coef[2 * i], ci = coef[2 * i + 1];
    const double vr = cr * ar - ci * ai, vi = cr * ai + ci * ar;
    if (vr == 0.0 && vi == 0.0) continue;
    const int dc = dst_col[i];
    if (dc >= 0) { atomicAdd(dst.Are + (size_t)dc * dst.lda + dr, vr); atomicAdd(dst.Aim + (size_t)dc * dst.lda + dr, vi); }
    else { atomicAdd(dst.bre + dr, vr); atomicAdd(dst.bim + dr, vi); }
  }
}
void launch_combine(const DevSystem& src, const DevSystem& dst, int n_rows, const int* src_row, const int* dst_row, int n_terms, const int* src_col,
                    const int* dst_col, const double* coef, cudaStream_t st) {
  if (n_rows <= 0 || n_terms <= 0) return;
  dim3 grid((n_rows + 127) / 128, n_terms < 4096 ? n_terms : 4096);
  k_combine<<<grid, 128, 0, st>>>(src, dst, n_rows, src_row, dst_row, n_terms, src_col, dst_col, coef);
}

__global__ void k_add_entries(DevSystem dst, int n, const int* __restrict__ rows, const int* __restrict__ cols, const double* __restrict__ v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (cols[i] >= 0) { atomicAdd(dst.Are + (size_t)cols[i] * dst.lda + rows[i], v[2 * i]); atomicAdd(dst.Aim + (size_t)cols[i] * dst.lda + rows[i], v[2 * i + 1]); }
  else { atomicAdd(dst.bre + rows[i], v[2 * i]); atomicAdd(dst.bim + rows[i], v[2 * i + 1]); }
}
void launch_add_entries(const DevSystem& dst, int n, const int* rows, const int* cols, const double* v, cudaStream_t st) {
  if (n > 0) k_add_entries<<<(n + 255) / 256, 256, 0, st>>>(dst, n, rows, cols, v);
}

}  // namespace mfbd
