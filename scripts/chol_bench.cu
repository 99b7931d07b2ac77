// Developer micro-benchmark (GPU box): the reduced-system solvers of csrc/lm_kernel.cuh on random SPD systems, one CTA, timed with clock64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multical_b200/csrc -o /tmp/chol_bench scripts/chol_bench.cu && /tmp/chol_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
#include <vector>
#include "../include/mcba.h"
#include "lm_kernel.cuh"
using namespace mcba;

template <int R>
__global__ void __launch_bounds__(LM_THREADS, 1) k_small(int n, const double* S, const double* rhs, const double* gh, double* out, long long* cyc) {
  extern __shared__ double sm[];
  __shared__ int fail;
  if (threadIdx.x == 0) fail = 0;
  __syncthreads();
  const long long t0 = clock64();
  chol_small_body<R>(n, S, rhs, gh, 0.0, &fail, out, sm);
  const long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = fail; }
}
__global__ void __launch_bounds__(LM_THREADS, 1) k_diag(int n, double* S, double* Linv, long long* cyc) {
  extern __shared__ double sm[];
  __shared__ int fail;
  if (threadIdx.x == 0) fail = 0;
  __syncthreads();
  const long long t0 = clock64();
  chol_diag_body(n, 0, S, Linv, &fail, sm);
  const long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = fail; }
}
template <int R>
__global__ void __launch_bounds__(LM_THREADS, 1) k_rot(int n, const double* S, const double* rhs, const double* gh, double* out, long long* cyc) {
  extern __shared__ double sm[];
  __shared__ int fail;
  if (threadIdx.x == 0) fail = 0;
  __syncthreads();
  const long long t0 = clock64();
  chol_rot_body<R>(n, S, rhs, gh, 0.0, &fail, out, sm);
  const long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = fail; }
}
template <int CQ>
__global__ void __launch_bounds__(LM_THREADS, 1) k_v3(int n, const double* S, const double* rhs, const double* gh, double* out, long long* cyc) {
  extern __shared__ double sm[];
  __shared__ int fail;
  if (threadIdx.x == 0) fail = 0;
  __syncthreads();
  const long long t0 = clock64();
  chol_v3_body<CQ>(n, S, rhs, gh, 0.0, &fail, out, sm);
  const long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = fail; }
}
#ifdef HAVE_CTA
__global__ void __launch_bounds__(LM_THREADS, 1) k_cta(int n, const double* S, const double* rhs, const double* gh, double* out, long long* cyc) {
  extern __shared__ double sm[];
  __shared__ int fail;
  if (threadIdx.x == 0) fail = 0;
  __syncthreads();
  const long long t0 = clock64();
  chol_solve_cta(n, S, rhs, gh, 0.0, &fail, out, sm);
  const long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = fail; }
}
#endif

int main() {
  for (int n : {32, 70, 126}) {
    std::vector<double> A((size_t)n * n), b(n), x(n), z(n, 0.0);
    srand(1);
    std::vector<double> G((size_t)n * n);
    for (auto& v : G) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double s = i == j ? 0.5 : 0.0; for (int k = 0; k < n; k++) s += G[i * n + k] * G[j * n + k] / n; A[i * n + j] = s; }
    for (int i = 0; i < n; i++) b[i] = rand() / (double)RAND_MAX;
    double *dS, *db, *dz, *dx, *dLi; long long* dc;
    cudaMalloc(&dS, A.size() * 8); cudaMalloc(&db, n * 8); cudaMalloc(&dz, n * 8); cudaMalloc(&dx, n * 8); cudaMalloc(&dLi, 32 * 32 * 8 * 8); cudaMalloc(&dc, 64);
    cudaMemcpy(dS, A.data(), A.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dz, z.data(), n * 8, cudaMemcpyHostToDevice);
    auto check = [&](const char* name) {
      long long c[2]; cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost); cudaMemcpy(x.data(), dx, n * 8, cudaMemcpyDeviceToHost);
      double err = 0; for (int i = 0; i < n; i++) { double s = -b[i]; for (int j = 0; j < n; j++) s += A[i * n + j] * x[j]; err = fmax(err, fabs(s)); }
      printf("n=%3d %-12s %8lld cycles (%.2f us @1.965GHz)  fail=%lld  max|Ax-b|=%.2e  [%s]\n", n, name, c[0], c[0] / 1965.0, c[1], err, cudaGetErrorString(cudaGetLastError()));
    };
    const size_t smsz = 190 * 1024;
    const int R = (n + 15) / 16;
    for (int rep = 0; rep < 2; rep++) {
#define CS(RR) case RR: cudaFuncSetAttribute(k_small<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smsz); k_small<RR><<<1, LM_THREADS, smsz>>>(n, dS, db, dz, dx, dc); break;
      switch (R) { CS(1) CS(2) CS(3) CS(4) CS(5) CS(6) CS(7) CS(8) }
      cudaDeviceSynchronize();
      if (rep) check("chol_small");
#define CR(RR) case RR: cudaFuncSetAttribute(k_rot<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smsz); k_rot<RR><<<1, LM_THREADS, smsz>>>(n, dS, db, dz, dx, dc); break;
      switch (R) { CR(1) CR(2) CR(3) CR(4) CR(5) CR(6) CR(7) CR(8) }
      cudaDeviceSynchronize();
      if (rep) check("chol_rot");
      {
        const int CQ = ((n + 7) / 8 + 1) / 2 * 2;
#define CV(CC) case CC: cudaFuncSetAttribute(k_v3<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smsz); k_v3<CC><<<1, LM_THREADS, smsz>>>(n, dS, db, dz, dx, dc); break;
        switch (CQ) { CV(2) CV(4) CV(6) CV(8) CV(10) CV(12) CV(14) CV(16) }
        cudaDeviceSynchronize();
        if (rep) check("chol_v3");
      }
#ifdef HAVE_CTA
      cudaFuncSetAttribute(k_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smsz);
      k_cta<<<1, LM_THREADS, smsz>>>(n, dS, db, dz, dx, dc); cudaDeviceSynchronize();
      if (rep) check("chol_cta");
#endif
    }
    if (n == 32) {
      cudaFuncSetAttribute(k_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smsz);
      for (int rep = 0; rep < 2; rep++) { cudaMemcpy(dS, A.data(), A.size() * 8, cudaMemcpyHostToDevice); k_diag<<<1, LM_THREADS, smsz>>>(n, dS, dLi, dc); cudaDeviceSynchronize(); }
      long long c[2]; cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
      printf("n= 32 chol_diag_body %8lld cycles (%.2f us)\n", c[0], c[0] / 1965.0);
    }
  }
  return 0;
}
