// Micro-probe (debug aid): per-SM throughput of DFMA and of the fp64 tensor op (mma.m8n8k4.f64) on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters, long long* cyc) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.000001, c = 0.5;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void dmma_kernel(double* out, int iters, long long* cyc) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c1[0] + c0[1] + c1[1] + c0[2] + c1[2] + c0[3] + c1[3];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 << 20); cudaMalloc(&cyc, 1024);
  long long h;
  for (int threads : {128, 384, 512, 1024}) {
    const int iters = 2000;
    dfma_kernel<<<1, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    dfma_kernel<<<1, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA  threads=%4d: %lld cycles for %d x 8 DFMA per thread -> %.2f DFMA lanes/clk/SM\n", threads, h, iters, (double)threads * iters * 8 / h);
    dmma_kernel<<<1, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    dmma_kernel<<<1, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA  threads=%4d: %lld cycles for %d x 4 mma.m8n8k4 per warp -> %.2f FMA/clk/SM (256 FMA per mma)\n", threads, h, iters, (double)(threads / 32) * iters * 4 * 256 / h);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
