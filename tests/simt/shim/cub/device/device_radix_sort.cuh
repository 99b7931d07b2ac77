// TEST INFRASTRUCTURE ONLY: host stand-in for cub::DeviceRadixSort::SortKeys (two-phase temp-storage protocol kept).
#pragma once
#include <algorithm>
#include "../../cuda_runtime.h"
namespace cub {
struct DeviceRadixSort {
  template <class K, class N>
  static cudaError_t SortKeys(void* tmp, size_t& tmp_bytes, const K* in, K* out, N n, int = 0, int = sizeof(K) * 8, cudaStream_t = nullptr) {
    if (!tmp) { tmp_bytes = 16; return cudaSuccess; }
    if (out != in) memmove(out, in, sizeof(K) * (size_t)n);
    std::sort(out, out + n);
    return cudaSuccess;
  }
};
}  // namespace cub
