// Developer micro-benchmark (GPU box): the reduced-system solvers of csrc/lm_kernel.cuh on random SPD systems, one CTA, timed with clock64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multical_b200/csrc -I scripts -o /tmp/chol_bench scripts/chol_bench.cu && /tmp/chol_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
#include <vector>
#include "../include/mcba.h"
#include "lm_kernel.cuh"
#include "chol_variants.cuh"
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

// candidate: the 32x32 diagonal block by ONE warp (block in registers, lane i = row i), then its inverse
__device__ __forceinline__ void warp_potrf32_inverse(double (*Lm)[CHOL_NB + 1], double (*Liv)[CHOL_NB + 1], int lane, int nb, int* fail) {
  double a[CHOL_NB], rsd[CHOL_NB];
#pragma unroll
  for (int j = 0; j < CHOL_NB; j++) a[j] = j <= lane ? Lm[lane][j] : 0.0;
#pragma unroll
  for (int k = 0; k < CHOL_NB; k++) {
    const double akk = __shfl_sync(0xffffffffu, a[k], k);
    if (lane == 0 && k < nb && !(akk > 0.0)) *fail += 1;
    const double rs = fast_rsqrt(fmin(fmax(akk, 1e-30), 1e30));
    rsd[k] = rs;
    const double l = lane >= k ? a[k] * rs : 0.0;
    a[k] = l;
#pragma unroll
    for (int j = k + 1; j < CHOL_NB; j++) { const double ljk = __shfl_sync(0xffffffffu, l, j); a[j] -= l * ljk; }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < CHOL_NB; j++) Lm[lane][j] = j <= lane ? a[j] : 0.0;
  __syncwarp();
  double z[CHOL_NB];
#pragma unroll
  for (int i = 0; i < CHOL_NB; i++) {
    double t0 = (i == lane) ? 1.0 : 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
    for (int m = 0; m < CHOL_NB; m += 4) {
      if (m < i) t0 -= Lm[i][m] * z[m];
      if (m + 1 < i) t1 -= Lm[i][m + 1] * z[m + 1];
      if (m + 2 < i) t2 -= Lm[i][m + 2] * z[m + 2];
      if (m + 3 < i) t3 -= Lm[i][m + 3] * z[m + 3];
    }
    z[i] = (i >= lane) ? ((t0 + t1) + (t2 + t3)) * rsd[i] : 0.0;
  }
#pragma unroll
  for (int i = 0; i < CHOL_NB; i++) Liv[i][lane] = z[i];
  __syncwarp();
}
__global__ void __launch_bounds__(LM_THREADS, 1) k_diag_warp(int n, double* S, double* Linv, long long* cyc) {
  extern __shared__ double sm[];
  __shared__ int fail;
  double (*Lm)[CHOL_NB + 1] = reinterpret_cast<double (*)[CHOL_NB + 1]>(sm);
  double (*Liv)[CHOL_NB + 1] = reinterpret_cast<double (*)[CHOL_NB + 1]>(sm + CHOL_NB * (CHOL_NB + 1));
  const int tid = threadIdx.x;
  if (tid == 0) fail = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int o = tid; o < CHOL_NB * CHOL_NB; o += LM_THREADS) { const int i = o / CHOL_NB, j = o % CHOL_NB; Lm[i][j] = j <= i ? __ldcg(&S[(size_t)i * n + j]) : 0.0; }
  __syncthreads();
  const long long t1 = clock64();
  if (tid < 32) warp_potrf32_inverse(Lm, Liv, tid, 32, &fail);
  __syncthreads();
  const long long t2 = clock64();
  for (int o = tid; o < CHOL_NB * CHOL_NB; o += LM_THREADS) { const int i = o / CHOL_NB, j = o % CHOL_NB; if (j <= i) S[(size_t)i * n + j] = Lm[i][j]; Linv[o] = Liv[i][j]; }
  __syncthreads();
  const long long t3 = clock64();
  if (tid == 0) { cyc[0] = t3 - t0; cyc[1] = fail; cyc[2] = t2 - t1; }
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
      cudaFuncSetAttribute(k_diag_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smsz);
      for (int rep = 0; rep < 2; rep++) { cudaMemcpy(dS, A.data(), A.size() * 8, cudaMemcpyHostToDevice); k_diag_warp<<<1, LM_THREADS, smsz>>>(n, dS, dLi, dc); cudaDeviceSynchronize(); }
      long long c3[3]; cudaMemcpy(c3, dc, 24, cudaMemcpyDeviceToHost);
      printf("n= 32 diag_warp      %8lld cycles (%.2f us), factor+inverse alone %lld cycles (%.2f us) fail=%lld [%s]\n", c3[0], c3[0] / 1965.0, c3[2], c3[2] / 1965.0, c3[1], cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
