// redbench.cu -- micro-benchmarks that decide the K1 matrix-update design (development aid, not product code).
//   mode 0: RED.ADD.F64 from registers (lane = row triple, like the first K1)      -> lane-updates/s
//   mode 1: stage 9 columns x 2 planes x 96 rows in shared memory, UBLKRED.ADD.F64 (cp.reduce.async.bulk) 768 B chunks
//   mode 2: plain ST.64 of the same pattern (upper bound of the LSU path)
//   mode 3: mode 1 with plain bulk store (cp.async.bulk shared->global) instead of reduce
//   mode 4: DFMA only ; mode 5: DMMA only ; mode 6: DFMA + DMMA interleaved in the same warp ; 7: alternate warps
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o redbench redbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ int node_of(int e, int j, int nnode) {  // structured-strip like sharing: nodes e/2, e/2+1, e/2+41
  int c = e >> 1; int n = (j == 0) ? c : (j == 1 ? c + 1 : c + 41);
  return n % nnode;
}

template <int MODE>
__global__ void __launch_bounds__(128) k_upd(double* Are, double* Aim, long long lda, int nrb, int nelem, int echunk, int nnode) {
  extern __shared__ __align__(128) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rb = blockIdx.x * 4 + warp;
  if (rb >= nrb) return;
  double* buf = sm + warp * (18 * 96);
  const int row0 = rb * 96;
  const int e0 = blockIdx.y * echunk, e1 = min(e0 + echunk, nelem);
  double v = 1.0 + lane * 1e-3;
  for (int e = e0; e < e1; e++) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int j = 0; j < 3; j++) {
        int nd = node_of(e, j, nnode);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          size_t col = (size_t)(3 * nd + k) * lda;
#pragma unroll
          for (int l = 0; l < 3; l++) {
            if (MODE == 0) { atomicAdd(Are + col + row0 + 3 * lane + l, v); atomicAdd(Aim + col + row0 + 3 * lane + l, v); }
            else { Are[col + row0 + 3 * lane + l] = v; Aim[col + row0 + 3 * lane + l] = v; }
          }
        }
      }
    } else {
      if (lane < 18) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 9; c++)
#pragma unroll
        for (int l = 0; l < 3; l++) { buf[(2 * c) * 96 + 3 * lane + l] = v; buf[(2 * c + 1) * 96 + 3 * lane + l] = v; }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane < 18) {
        int c = lane >> 1, j = c / 3, k = c - 3 * j;
        int nd = node_of(e, j, nnode);
        double* dst = ((lane & 1) ? Aim : Are) + (size_t)(3 * nd + k) * lda + row0;
        uint32_t s = (uint32_t)__cvta_generic_to_shared(buf + lane * 96);
        if (MODE == 1) asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" :: "l"(dst), "r"(s), "r"(768) : "memory");
        else asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(s), "r"(768) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (MODE == 13) {}
    }
    v += 1e-9;
  }
  if (MODE == 1 || MODE == 3) { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
}

// mode 11: one elected lane issues all 18 (serial issue from one thread)
__global__ void __launch_bounds__(128) k_upd_lane0(double* Are, double* Aim, long long lda, int nrb, int nelem, int echunk, int nnode) {
  extern __shared__ __align__(128) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rb = blockIdx.x * 4 + warp;
  if (rb >= nrb) return;
  double* buf = sm + warp * (18 * 96);
  const int row0 = rb * 96;
  const int e0 = blockIdx.y * echunk, e1 = min(e0 + echunk, nelem);
  double v = 1.0 + lane * 1e-3;
  for (int e = e0; e < e1; e++) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 9; c++)
#pragma unroll
      for (int l = 0; l < 3; l++) { buf[(2 * c) * 96 + 3 * lane + l] = v; buf[(2 * c + 1) * 96 + 3 * lane + l] = v; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 18; q++) {
        int c = q >> 1, j = c / 3, k = c - 3 * j;
        int nd = node_of(e, j, nnode);
        double* dst = ((q & 1) ? Aim : Are) + (size_t)(3 * nd + k) * lda + row0;
        uint32_t s = (uint32_t)__cvta_generic_to_shared(buf + q * 96);
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" :: "l"(dst), "r"(s), "r"(768) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    v += 1e-9;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- pipe concurrency -------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(256) k_pipes(double* out, int iters) {
  double f[8], m[16];
  const double a = 1.0000001, b = 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) f[i] = i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; i++) m[i] = i;
  const bool do_f = (MODE == 4) || (MODE == 6) || (MODE == 7 && ((threadIdx.x >> 5) & 1) == 0);
  const bool do_m = (MODE == 5) || (MODE == 6) || (MODE == 7 && ((threadIdx.x >> 5) & 1) == 1);
  for (int it = 0; it < iters; it++) {
    if (do_f) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] = fma(f[i], a, b);
    }
    if (do_m) {
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) dmma(m[2 * i], m[2 * i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += f[i];
#pragma unroll
  for (int i = 0; i < 16; i++) s += m[i];
  if (s == 1.2345) out[0] = s;
}

int main(int argc, char** argv) {
  const int n = 16128;  // rows; 168 row blocks of 96
  const long long lda = n;
  const int nnode = 5376, ncol = 3 * nnode;
  double *Are, *Aim;
  CK(cudaMalloc(&Are, (size_t)lda * ncol * 8)); CK(cudaMalloc(&Aim, (size_t)lda * ncol * 8));
  CK(cudaMemset(Are, 0, (size_t)lda * ncol * 8)); CK(cudaMemset(Aim, 0, (size_t)lda * ncol * 8));
  const int nrb = n / 96, nelem = 2 * nnode, echunk = 32;
  dim3 grid((nrb + 3) / 4, (nelem + echunk - 1) / echunk), block(128);
  const size_t smem = 4 * 18 * 96 * 8;
  CK(cudaFuncSetAttribute(k_upd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_upd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_upd_lane0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double lane_updates = (double)nrb * 32 * nelem * 54;
  printf("matrix 2 x %.2f GB, %d row blocks x %d elements, %.3e lane updates (%.1f GB of update traffic)\n", lda * ncol * 8 / 1e9, nrb, nelem, lane_updates, lane_updates * 8 / 1e9);
  for (int mode : {0, 1, 11, 2, 3}) {
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaEventRecord(e0));
      switch (mode) {
        case 0: k_upd<0><<<grid, block, 0>>>(Are, Aim, lda, nrb, nelem, echunk, nnode); break;
        case 1: k_upd<1><<<grid, block, smem>>>(Are, Aim, lda, nrb, nelem, echunk, nnode); break;
        case 11: k_upd_lane0<<<grid, block, smem>>>(Are, Aim, lda, nrb, nelem, echunk, nnode); break;
        case 2: k_upd<2><<<grid, block, 0>>>(Are, Aim, lda, nrb, nelem, echunk, nnode); break;
        case 3: k_upd<3><<<grid, block, smem>>>(Are, Aim, lda, nrb, nelem, echunk, nnode); break;
      }
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf("mode %2d: %8.3f ms  %.3e lane-updates/s  %.1f GB/s of f64 updates  (%.2f SM-cycles per lane update at 1.965 GHz x 148)\n", mode, best, lane_updates / best * 1e3,
           lane_updates * 8 / best / 1e6, best * 1e-3 * 1.965e9 * 148 / lane_updates);
  }
  // check mode 1 numerics quickly: zero, run once, every touched entry must be a multiple of ~1
  double* out; CK(cudaMalloc(&out, 8));
  for (int mode : {4, 5, 6, 7}) {
    const int iters = 20000; float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaEventRecord(e0));
      switch (mode) {
        case 4: k_pipes<4><<<148 * 4, 256>>>(out, iters); break;
        case 5: k_pipes<5><<<148 * 4, 256>>>(out, iters); break;
        case 6: k_pipes<6><<<148 * 4, 256>>>(out, iters); break;
        case 7: k_pipes<7><<<148 * 4, 256>>>(out, iters); break;
      }
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double thr = 148.0 * 4 * 256, wf = (mode == 7 ? 0.5 : 1.0);
    double ffl = (mode == 5) ? 0 : thr * wf * iters * 32.0 * 2;                 // 32 DFMA per iter per thread
    double mfl = (mode == 4) ? 0 : thr * wf / 32 * iters * 16.0 * (2.0 * 8 * 8 * 4);  // 16 DMMA per iter per warp
    printf("mode %d: %8.3f ms  DFMA %.2f TFLOP/s  DMMA %.2f TFLOP/s  sum %.2f\n", mode, best, ffl / best / 1e9, mfl / best / 1e9, (ffl + mfl) / best / 1e9);
  }
  return 0;
}
