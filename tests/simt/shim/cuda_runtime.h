// cuda_runtime.h -- TEST INFRASTRUCTURE ONLY: host stand-in for the CUDA runtime calls multical_b200/csrc/solver.cu makes, used
// when tests/simt builds the kernels for the SIMT interpreter (simt.h).  "Device" memory is host memory, streams are
// synchronous, the one reported device calls itself sm_100 so that mcba_create accepts it.  Peer / IPC calls fail.
#pragma once
#include <chrono>
#include "simt.h"

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
struct CUstream_st;
typedef CUstream_st* cudaStream_t;
struct SimtEvent { std::chrono::steady_clock::time_point t; };
typedef SimtEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp {
  char name[256];
  int major, minor, multiProcessorCount;
  size_t sharedMemPerBlockOptin, totalGlobalMem;
};

static inline const char* cudaGetErrorString(cudaError_t e) {
  return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : e == cudaErrorNotSupported ? "operation not supported (SIMT host build)" : "error";
}
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof(*p));
  snprintf(p->name, sizeof(p->name), "SIMT interpreter (host)");
  p->major = 10; p->minor = 0;
  const char* e = getenv("SIMT_SMS");
  p->multiProcessorCount = e ? std::max(1, atoi(e)) : 148;
  p->sharedMemPerBlockOptin = 227 * 1024;
  p->totalGlobalMem = (size_t)8 << 30;
  return cudaSuccess;
}
// device allocations are poisoned (0xFF bytes: NaN doubles, -1 integers) like uninitialised HBM may be
template <class T>
static inline cudaError_t cudaMalloc(T** p, size_t bytes) {
  void* q = malloc(bytes ? bytes : 1);
  if (!q) return cudaErrorMemoryAllocation;
  memset(q, 0xFF, bytes);
  *p = (T*)q;
  return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new SimtEvent(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F*, int, int) { return cudaSuccess; }
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorNotSupported; }
