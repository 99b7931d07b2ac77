// Dependent-chain latencies on this GPU (cycles per op): DFMA, rsqrt(double), 1/x, sqrt, 64-bit shuffle, LDS round trip, __syncthreads with 8 warps.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(double* out, long long* cyc, int n) {
  __shared__ double sm[256];
  double x = out[0] + 1.5, y = 1.0000001, z = 0.5;
  long long t0, t1;
  const int tid = threadIdx.x;
  t0 = clock64(); for (int i = 0; i < n; i++) x = fma(x, y, z); t1 = clock64(); if (tid == 0) cyc[0] = t1 - t0;
  t0 = clock64(); for (int i = 0; i < n; i++) x = rsqrt(x + 2.0); t1 = clock64(); if (tid == 0) cyc[1] = t1 - t0;
  t0 = clock64(); for (int i = 0; i < n; i++) x = 1.0 / (x + 2.0); t1 = clock64(); if (tid == 0) cyc[2] = t1 - t0;
  t0 = clock64(); for (int i = 0; i < n; i++) x = sqrt(x + 2.0); t1 = clock64(); if (tid == 0) cyc[3] = t1 - t0;
  t0 = clock64(); for (int i = 0; i < n; i++) x = __shfl_sync(0xffffffffu, x, (tid + 1) & 31); t1 = clock64(); if (tid == 0) cyc[4] = t1 - t0;
  t0 = clock64(); for (int i = 0; i < n; i++) { sm[tid] = x; __syncwarp(); x = sm[(tid + 1) & 255] + 1.0; __syncwarp(); } t1 = clock64(); if (tid == 0) cyc[5] = t1 - t0;
  t0 = clock64(); for (int i = 0; i < n; i++) { sm[tid] = x; __syncthreads(); x = sm[(tid + 33) & 255] + 1.0; __syncthreads(); } t1 = clock64(); if (tid == 0) cyc[6] = t1 - t0;
  float xf = (float)x;
  t0 = clock64(); for (int i = 0; i < n; i++) xf = fmaf(xf, 1.0000001f, 0.5f); t1 = clock64(); if (tid == 0) cyc[7] = t1 - t0;
  t0 = clock64(); for (int i = 0; i < n; i++) x = x * y; t1 = clock64(); if (tid == 0) cyc[8] = t1 - t0;
  out[tid] = x + xf;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 16 * 8); cudaMemset(out, 0, 256 * 8);
  const int n = 1000;
  for (int th : {32, 256}) {
    k<<<1, th>>>(out, cyc, n); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[] = {"dfma", "rsqrt(double)", "1/x", "sqrt", "shfl64", "sts+syncwarp+lds", "sts+bar+lds+bar", "ffma", "dmul"};
    printf("threads %d:", th);
    for (int i = 0; i < 9; i++) printf(" %s=%.1f", names[i], (double)h[i] / n);
    printf(" cycles/op\n");
  }
  return 0;
}
