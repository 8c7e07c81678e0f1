// DMMA (mma.sync.m8n8k4.f64) issue-rate microbenchmark for sm_100a.
// Measures DMMA instructions per cycle per SM as a function of resident warps per
// SM sub-partition and of the number of independent accumulators per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int U>
__global__ void k(double *out, int iters, long long *cycles) {
  double c0[U], c1[U];
#pragma unroll
  for (int u = 0; u < U; ++u) { c0[u] = threadIdx.x * 1e-3 + u; c1[u] = u * 0.5; }
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[u]), "+d"(c1[u]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) s += c0[u] + c1[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int U>
void run(int threads, int blocks_per_sm) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int grid = sms * blocks_per_sm;
  double *out; long long *cyc;
  cudaMalloc(&out, sizeof(double) * grid * threads);
  cudaMalloc(&cyc, sizeof(long long) * grid);
  int iters = 20000;
  k<U><<<grid, threads>>>(out, 100, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<U><<<grid, threads>>>(out, iters, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[1]; cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
  double dmma = (double)iters * U * (threads / 32) * grid;
  double tflops = dmma * 512.0 / (ms * 1e-3) / 1e12;
  double per_smsp_interval = (double)h[0] / ((double)iters * U * (threads / 32) * blocks_per_sm / 4.0);
  printf("U=%2d threads=%4d blocks/SM=%d warps/SMSP=%.1f : %.2f TFLOP/s, %.1f cycles per DMMA per SMSP (clock64)\n",
         U, threads, blocks_per_sm, threads / 32.0 * blocks_per_sm / 4.0, tflops, per_smsp_interval);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int t : {128, 256, 384, 512, 768, 1024}) { run<4>(t, 1); run<8>(t, 1); run<16>(t, 1); run<34>(t, 1); }
  run<34>(256, 2); run<16>(128, 3); run<32>(128, 2);
  return 0;
}
