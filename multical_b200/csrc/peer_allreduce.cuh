// peer_allreduce.cuh — one-shot all-reduce over NVLink peer memory (push model), replacing NCCL for the latency-bound
// exchanges of an LM iteration (a few scalars; the reduced normal equations S (n_s^2 doubles) while they fit the slot).
//
// Every rank owns one cudaMalloc'd buffer, IPC-mapped into all peers:   [flags: 2 parities x world x 64 B]
//                                                                       [data : 2 parities x world (source rank) x cap doubles]
// all-reduce #seq (parity = seq & 1):
//   1. every CTA stores its share of the local values into slot [parity][my rank] of EVERY rank's buffer (remote stores)
//   2. the last CTA to finish fences (system scope) and stores seq into flag [parity][my rank] of every rank
//   3. every CTA spins on its OWN flags until all sources show seq, then reduces the world slots in rank order -- the
//      same order on every rank, so all ranks obtain bit-identical results (they run the same scalar logic on them).
// Two parities suffice: a rank cannot finish #seq+1 before every peer has posted #seq+1, i.e. finished reading #seq.
#pragma once
#include <stdint.h>
#include "solver_kernels.cuh"

namespace mcba {

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// The same exchange executed by ONE CTA as the tail of the kernel that produced the values (MCBA_FUSE=1 on several GPUs): the last
// CTA of k_quad, the single CTA of k_cost_from_moments / k_scale_exchange.  Reduction -> all-reduce over NVLink -> the scalar
// trust-region step that consumes it, in one launch.  All threads of the CTA must call it; the values must be visible to the CTA
// (a __syncthreads after they were written).
__device__ __noinline__ void peer_allreduce_block(const PeerArgs& a) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int parity = a.seq & 1u;
  int total = 0;
  for (int s = 0; s < a.nseg; s++) total += a.seg[s].count;
  for (int idx = tid; idx < total; idx += nt) {
    int s = 0, off = idx;
    while (off >= a.seg[s].count) { off -= a.seg[s].count; s++; }
    const double v = a.seg[s].buf[off];
    const size_t o = peer_data_off(a.world, a.cap, parity, a.rank) + idx;
    for (int p = 0; p < a.world; p++) a.base[p][o] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < a.world)
    st_volatile_u64(reinterpret_cast<unsigned long long*>(a.base[tid] + peer_flag_off(a.world, parity, a.rank)), (unsigned long long)a.seq);
  if (tid < a.world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(a.base[a.rank] + peer_flag_off(a.world, parity, tid));
    while (ld_volatile_u64(f) != (unsigned long long)a.seq) { }
  }
  __syncthreads();
  __threadfence_system();
  const double* mine = a.base[a.rank];
  for (int idx = tid; idx < total; idx += nt) {
    int s = 0, off = idx;
    while (off >= a.seg[s].count) { off -= a.seg[s].count; s++; }
    double acc = __ldcv(mine + peer_data_off(a.world, a.cap, parity, 0) + idx);
    for (int src = 1; src < a.world; src++) {
      const double v = __ldcv(mine + peer_data_off(a.world, a.cap, parity, src) + idx);
      acc = a.seg[s].op == 0 ? acc + v : fmax(acc, v);
    }
    a.seg[s].buf[off] = acc;
  }
  __syncthreads();
  if (a.epilogue && tid == 0) run_epilogue(a.epilogue, a.st, a.red);
}

__global__ void __launch_bounds__(256)
k_peer_allreduce(PeerArgs a) {
  __shared__ int is_last;
  const int tid = threadIdx.x;
  const int parity = a.seq & 1u;
  int total = 0;
  for (int s = 0; s < a.nseg; s++) total += a.seg[s].count;
  // 1. push my values into my slot on every rank
  for (int idx = blockIdx.x * blockDim.x + tid; idx < total; idx += gridDim.x * blockDim.x) {
    int s = 0, off = idx;
    while (off >= a.seg[s].count) { off -= a.seg[s].count; s++; }
    const double v = a.seg[s].buf[off];
    const size_t o = peer_data_off(a.world, a.cap, parity, a.rank) + idx;
    for (int p = 0; p < a.world; p++) a.base[p][o] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (tid == 0) { const unsigned t = atomicAdd(a.counter, 1u); is_last = (t == gridDim.x - 1); }
  __syncthreads();
  if (is_last) {
    // 2. all CTAs of this rank have pushed and fenced: publish
    __threadfence_system();
    if (tid < a.world)
      st_volatile_u64(reinterpret_cast<unsigned long long*>(a.base[tid] + peer_flag_off(a.world, parity, a.rank)), (unsigned long long)a.seq);
    if (tid == 0) *a.counter = 0;
  }
  // 3. wait for every source, then reduce in rank order
  if (tid < a.world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(a.base[a.rank] + peer_flag_off(a.world, parity, tid));
    while (ld_volatile_u64(f) != (unsigned long long)a.seq) { }
  }
  __syncthreads();
  __threadfence_system();
  const double* mine = a.base[a.rank];
  for (int idx = blockIdx.x * blockDim.x + tid; idx < total; idx += gridDim.x * blockDim.x) {
    int s = 0, off = idx;
    while (off >= a.seg[s].count) { off -= a.seg[s].count; s++; }
    double acc = __ldcv(mine + peer_data_off(a.world, a.cap, parity, 0) + idx);
    for (int src = 1; src < a.world; src++) {
      const double v = __ldcv(mine + peer_data_off(a.world, a.cap, parity, src) + idx);
      acc = a.seg[s].op == 0 ? acc + v : fmax(acc, v);
    }
    a.seg[s].buf[off] = acc;
  }
  if (a.epilogue) {                      // launched with one CTA: all reduced values are in place after this barrier
    __syncthreads();
    if (tid == 0) run_epilogue(a.epilogue, a.st, a.red);
  }
}

}  // namespace mcba
