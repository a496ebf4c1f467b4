// Dependent-chain latencies on the current GPU (one warp, one block): DFMA, DMUL, rcp_fast, LDS, and
// __syncthreads() cost with 8 warps.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rcp_fast(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  r = fma(fma(-a, r, 1.0), r, r);
  r = fma(fma(-a, r, 1.0), r, r);
  r = fma(fma(-a, r, 1.0), r, r);
  return r;
}
__global__ void lat_kernel(double* out, long long* clk, int iters) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double a = 1.0 + threadIdx.x * 1e-9;
  const double b = 1.0000001, c = 1e-12;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) a = fma(a, b, c);
  long long t1 = clock64();
  for (int i = 0; i < iters; ++i) a = a * b;
  long long t2 = clock64();
  for (int i = 0; i < iters; ++i) a = rcp_fast(a) + 0.5;
  long long t3 = clock64();
  int idx = threadIdx.x;
  for (int i = 0; i < iters; ++i) { double v = sm[idx & 1023]; idx = (int)(v * 3.0) + threadIdx.x; a += v; }
  long long t4 = clock64();
  for (int i = 0; i < iters; ++i) { __syncthreads(); }
  long long t5 = clock64();
  for (int i = 0; i < iters; ++i) a = 1.0 / a + 0.5;
  long long t6 = clock64();
  for (int i = 0; i < iters; ++i) a = sqrt(a) + 0.5;
  long long t7 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
  if (threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; clk[4] = t5 - t4; clk[5] = t6 - t5; clk[6] = t7 - t6; }
}
int main() {
  double* out; long long* clk; cudaMalloc(&out, 8 * 1024); cudaMalloc(&clk, 64);
  long long h[8];
  const int iters = 2000;
  for (int threads : {32, 256}) {
    lat_kernel<<<1, threads>>>(out, clk, iters); cudaDeviceSynchronize();
    lat_kernel<<<1, threads>>>(out, clk, iters); cudaDeviceSynchronize();
    cudaMemcpy(h, clk, 56, cudaMemcpyDeviceToHost);
    printf("threads=%d cycles/op: DFMA %.1f DMUL %.1f rcp_fast(+add) %.1f LDS(dep chain) %.1f syncthreads %.1f div(+add) %.1f sqrt(+add) %.1f\n", threads,
           (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters, (double)h[3] / iters, (double)h[4] / iters, (double)h[5] / iters, (double)h[6] / iters);
  }
  return 0;
}
