// tmared.cu -- does cp.reduce.async.bulk.tensor (.add, FLOAT64 tensor map) work on sm_100a, and how fast is it?
// box = 96 rows x 3 columns x 2 planes of a planar column-major complex matrix; one op per (tile, node).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(128) k_tma(const __grid_constant__ CUtensorMap tmap, int nrb, int nelem, int echunk, int nnode) {
  extern __shared__ __align__(128) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rb = blockIdx.x * 4 + warp;
  if (rb >= nrb) return;
  double* buf = sm + warp * (18 * 96);
  const int row0 = rb * 96;
  const int e0 = blockIdx.y * echunk, e1 = min(e0 + echunk, nelem);
  double v = 1.0;
  for (int e = e0; e < e1; e++) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    // layout [node j][plane][k][row]
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int l = 0; l < 3; l++) { buf[((j * 2 + 0) * 3 + k) * 96 + 3 * lane + l] = v; buf[((j * 2 + 1) * 3 + k) * 96 + 3 * lane + l] = 2.0 * v; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 3; j++) {
        int c = e >> 1; int nd = ((j == 0) ? c : (j == 1 ? c + 1 : c + 41)) % nnode;
        uint32_t s = (uint32_t)__cvta_generic_to_shared(buf + j * 576);
        asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                     :: "l"(&tmap), "r"(row0), "r"(3 * nd), "r"(0), "r"(s) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const int n = 16128; const long long lda = n; const int nnode = 5376, ncol = 3 * nnode;
  double* A; CK(cudaMalloc(&A, (size_t)2 * lda * ncol * 8)); CK(cudaMemset(A, 0, (size_t)2 * lda * ncol * 8));
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap tmap;
  cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)ncol, 2};
  cuuint64_t strides[2] = {(cuuint64_t)lda * 8, (cuuint64_t)lda * ncol * 8};
  cuuint32_t box[3] = {96, 3, 2}, estr[3] = {1, 1, 1};
  CUresult r = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, A, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode result %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  const int nrb = n / 96, nelem = 2 * nnode, echunk = 32;
  dim3 grid((nrb + 3) / 4, (nelem + echunk - 1) / echunk), block(128);
  const size_t smem = 4 * 18 * 96 * 8;
  CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    k_tma<<<grid, block, smem>>>(tmap, nrb, nelem, echunk, nnode);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  const double lane_updates = (double)nrb * 32 * nelem * 54;
  printf("tensor reduce: %8.3f ms  %.3e lane-updates/s  %.1f GB/s\n", best, lane_updates / best * 1e3, lane_updates * 8 / best / 1e6);
  // numerics: every entry of plane 0 must be (#contributions) * 3 reps, plane 1 twice that
  std::vector<double> h(4 * 96);
  CK(cudaMemcpy(h.data(), A + (size_t)300 * lda + 96 * 5, 96 * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h.data() + 96, A + (size_t)lda * ncol + (size_t)300 * lda + 96 * 5, 96 * 8, cudaMemcpyDeviceToHost));
  printf("sample re %g %g %g  im %g %g %g (im must be 2 x re, all rows equal)\n", h[0], h[1], h[95], h[96], h[97], h[191]);
  return 0;
}
