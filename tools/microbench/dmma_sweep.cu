// DMMA pipe micro-benchmark: how many warps per SM sub-partition does mma.sync.m8n8k4.f64 need to saturate the FP64 tensor pipe, with the instruction mix of the
// trailing-update kernel (24 independent accumulators per k-step; optionally the 6 DADDs and 12 shared-memory loads that go with them)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench/dmma_sweep.cu -o tools/microbench/dmma_sweep && tools/microbench/dmma_sweep
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MIX>   // 0: DMMA only; 1: + 6 DADD per 24 DMMA; 2: + 6 DADD + 12 LDS.64; 3: + 6 DADD + 6 LDS.128 (the same bytes in half the instructions)
__global__ void k(double* out, int iters) {
  __shared__ double sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
  __syncthreads();
  double x[4][2][2], y[4][2][2], z[4][2][2];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 2; b++) { x[a][b][0] = x[a][b][1] = y[a][b][0] = y[a][b][1] = z[a][b][0] = z[a][b][1] = 0.0; }
  double ar[4], ai[4], br[2], bi[2];
#pragma unroll
  for (int a = 0; a < 4; a++) { ar[a] = 1.0 + threadIdx.x * 1e-6 + a; ai[a] = 0.5 - a * 1e-3; }
  br[0] = 0.3; br[1] = 0.7; bi[0] = -0.2; bi[1] = 0.9;
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; it++) {
    if (MIX == 2) {
      const double* p = sm + ((it * 64 + lane) & 1023);
#pragma unroll
      for (int a = 0; a < 4; a++) { ar[a] = p[a * 32]; ai[a] = p[a * 32 + 128]; }
      br[0] = p[512]; br[1] = p[544]; bi[0] = p[640]; bi[1] = p[672];
    }
    if (MIX == 3) {
      const double2* p = reinterpret_cast<const double2*>(sm) + ((it * 32 + lane) & 511);
      double2 v;
      v = p[0]; ar[0] = v.x; ar[1] = v.y; v = p[32]; ar[2] = v.x; ar[3] = v.y;
      v = p[64]; ai[0] = v.x; ai[1] = v.y; v = p[96]; ai[2] = v.x; ai[3] = v.y;
      v = p[128]; br[0] = v.x; br[1] = v.y; v = p[160]; bi[0] = v.x; bi[1] = v.y;
    }
    double sa[4], sb[2];
#pragma unroll
    for (int a = 0; a < 4; a++) sa[a] = (MIX >= 1) ? ar[a] + ai[a] : ar[a];
#pragma unroll
    for (int b = 0; b < 2; b++) sb[b] = (MIX >= 1) ? br[b] + bi[b] : br[b];
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
      for (int a = 0; a < 4; a++) { dmma(x[a][b][0], x[a][b][1], ar[a], br[b]); dmma(y[a][b][0], y[a][b][1], ai[a], bi[b]); dmma(z[a][b][0], z[a][b][1], sa[a], sb[b]); }
  }
  double s = 0.0;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 2; b++) s += x[a][b][0] + x[a][b][1] + y[a][b][0] + y[a][b][1] + z[a][b][0] + z[a][b][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MIX> void run(int warps_per_cta, int ctas_per_sm, double* out) {
  const int iters = 20000, nsm = 148;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MIX><<<nsm * ctas_per_sm, 32 * warps_per_cta>>>(out, 100);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k<MIX><<<nsm * ctas_per_sm, 32 * warps_per_cta>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float t; cudaEventElapsedTime(&t, e0, e1); if (t < best) best = t; }
  const double flops = (double)nsm * ctas_per_sm * warps_per_cta * iters * 24.0 * 512.0;
  printf("mix %d  warps/CTA %2d  CTAs/SM %d  warps/SMSP %4.1f : %6.2f TFLOP/s\n", MIX, warps_per_cta, ctas_per_sm, warps_per_cta * ctas_per_sm / 4.0, flops / (best * 1e-3) / 1e12);
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * 8);
  const int cfg[][2] = {{4, 1}, {4, 2}, {4, 3}, {4, 4}, {8, 2}, {8, 3}, {8, 4}, {16, 2}};
  for (auto& c : cfg) run<0>(c[0], c[1], out);
  for (auto& c : cfg) run<1>(c[0], c[1], out);
  for (auto& c : cfg) run<2>(c[0], c[1], out);
  for (auto& c : cfg) run<3>(c[0], c[1], out);
  return 0;
}
