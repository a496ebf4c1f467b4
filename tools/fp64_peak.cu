// Microbenchmark: FP64 FMA (vector pipe) vs DMMA m8n8k4 (tensor pipe) throughput on the current GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void dmma_kernel(double* out, int iters) {
  double c[8][2];
  for (int q = 0; q < 8; ++q) c[q][0] = c[q][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int q = 0; q < 8; ++q) s += c[q][0] + c[q][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dsqrt_kernel(double* out, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  for (int i = 0; i < iters; ++i) {
    a0 = sqrt(a0 + 1.5); a1 = sqrt(a1 + 1.5); a2 = sqrt(a2 + 1.5); a3 = sqrt(a3 + 1.5);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int blocks = p.multiProcessorCount * 4, threads = 512, iters = 20000;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  dfma_kernel<<<blocks, threads>>>(out, 100); cudaDeviceSynchronize();
  cudaEventRecord(e0); dfma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%s SMs=%d  DFMA: %.2f TFLOP/s (%.3f ms)\n", p.name, p.multiProcessorCount, 2.0 * 8 * iters * blocks * threads / ms / 1e9, ms);
  dmma_kernel<<<blocks, threads>>>(out, 100); cudaDeviceSynchronize();
  cudaEventRecord(e0); dmma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  printf("DMMA m8n8k4: %.2f TFLOP/s (%.3f ms)\n", 2.0 * 256 * 8 * iters * (double)blocks * threads / 32 / ms / 1e9, ms);
  dsqrt_kernel<<<blocks, threads>>>(out, 100); cudaDeviceSynchronize();
  cudaEventRecord(e0); dsqrt_kernel<<<blocks, threads>>>(out, iters / 10); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  printf("DSQRT(+add): %.2f Gop/s (%.3f ms)\n", 4.0 * (iters / 10) * blocks * threads / ms / 1e6, ms);
  return 0;
}
