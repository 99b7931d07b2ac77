// simt.h -- TEST INFRASTRUCTURE ONLY (never part of the product, never loaded by multical_b200).
//
// A small SIMT interpreter that lets the CPU test-suite execute the *unmodified* CUDA kernels of multical_b200/csrc on the
// host: tests/simt/translate.py rewrites `kernel<<<grid, block, smem, stream>>>(args)` into simt::launch(...) and
// `__shared__` into per-block storage, and this header supplies threadIdx/blockIdx, barriers, warp shuffles, the fp64 MMA
// fragment semantics and atomics.  Every CUDA thread of a block is a fiber (ucontext) of ONE host thread; blocks run one
// after the other, so execution is deterministic and data races cannot be observed -- what this checks is indexing, math and
// host-side plumbing of the real kernels (the GPU suite, -m gpu, checks the hardware execution).  A block whose fibers all
// wait on a barrier that cannot complete (divergent __syncthreads / __syncwarp, a lane missing from a full-mask shuffle)
// aborts with a message instead of hanging.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <signal.h>
#include <execinfo.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 { double x, y; };
struct alignas(8) float2 { float x, y; };      // 16-byte vector accesses fault on hardware when misaligned: let the host build use aligned moves too
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }

namespace simt {

constexpr int MAX_THREADS = 1024;
constexpr size_t STACK_BYTES = 512 * 1024;
enum { WAIT_NONE = 0, WAIT_BLOCK = 1, WAIT_WARP = 2 };

struct Warp {
  int alive = 0, arrived = 0;
  unsigned gen = 0;
  alignas(16) unsigned char xbuf[2][32][16];
};
struct Fiber {
  ucontext_t ctx;
  uint3 tid;
  int lane = 0;
  Warp* warp = nullptr;
  bool done = false;
  int wait = WAIT_NONE;
  unsigned wait_gen = 0;
  unsigned xcount = 0;
  int mark = 0;                 // debugging aid: set by set_mark(), reported with a deadlock
  char* stack = nullptr;
};
struct Block {
  int nthreads = 0, alive = 0, arrived = 0;
  unsigned gen = 0;
  uint3 idx{0, 0, 0};
  dim3 bdim, gdim;
  unsigned char* dyn = nullptr;
  size_t dyn_bytes = 0;
};

// thread_local: several "ranks" (one host thread each, tests/test_simt_multirank.py) may run kernels at the same time
static thread_local Fiber* cur = nullptr;
static thread_local Block blk;
static thread_local ucontext_t sched_ctx;
static thread_local std::vector<Fiber> fibers;
static thread_local std::vector<Warp> warps;
static thread_local const void* cur_body = nullptr;
static thread_local void (*cur_invoke)(const void*) = nullptr;
static thread_local const char* cur_name = "";
static thread_local long long total_launches = 0;
inline void set_mark(int m);
// per-kernel dynamic counts (SIMT_STATS=<file>: appended at process exit): launches, block barriers completed, warp collectives
// (shuffles / ballots / mma fragments exchanged) -- the serialisation points of a kernel, independent of any clock
struct KernelStats { long long launches = 0, blocks = 0, block_barriers = 0, warp_collectives = 0; };
inline std::map<std::string, KernelStats>& stats() { static std::map<std::string, KernelStats> m; return m; }
static thread_local KernelStats* cur_stats = nullptr;
inline bool stats_on() { static const bool on = getenv("SIMT_STATS") != nullptr; return on; }
inline void dump_stats();
inline void register_stats_dump() {       // once per process (launch<> is a template: a static in there exists once per launch site)
  static bool registered = false;
  if (!registered) { registered = true; stats(); atexit(dump_stats); }      // the map first: it must outlive the exit handler
}
inline void dump_stats() {
  const char* path = getenv("SIMT_STATS");
  FILE* f = path ? fopen(path, "a") : nullptr;
  if (!f) return;
  for (auto& kv : stats())
    fprintf(f, "%s\t%lld\t%lld\t%lld\t%lld\n", kv.first.c_str(), kv.second.launches, kv.second.blocks, kv.second.block_barriers, kv.second.warp_collectives);
  fclose(f);
}

[[noreturn]] inline void die(const char* what) {
  fprintf(stderr, "[simt] %s in kernel %s, block (%u,%u,%u)\n", what, cur_name, blk.idx.x, blk.idx.y, blk.idx.z);
  abort();
}
inline void yield() { swapcontext(&cur->ctx, &sched_ctx); }
inline void set_mark(int m) { if (cur) cur->mark = m; }
inline void spin_yield() { if (cur) yield(); }          // stays runnable: the scheduler comes back to it after the other fibers

inline void block_barrier() {
  Fiber* f = cur;
  blk.arrived++;
  if (blk.arrived >= blk.alive) { blk.arrived = 0; blk.gen++; if (cur_stats) cur_stats->block_barriers++; return; }
  f->wait = WAIT_BLOCK; f->wait_gen = blk.gen;
  yield();
  f->wait = WAIT_NONE;
}
inline void warp_barrier() {
  Fiber* f = cur;
  Warp& w = *f->warp;
  w.arrived++;
  if (w.arrived >= w.alive) { w.arrived = 0; w.gen++; return; }
  f->wait = WAIT_WARP; f->wait_gen = w.gen;
  yield();
  f->wait = WAIT_NONE;
}
// every live lane of the warp deposits n <= 16 bytes; returns the table of all lanes' deposits (valid until the lane's next
// but one collective: two buffers alternate, and no lane can run two collectives ahead of another)
inline unsigned char (*exchange(const void* src, size_t n))[16] {
  Fiber* f = cur;
  Warp& w = *f->warp;
  const int par = f->xcount & 1;
  f->xcount++;
  memcpy(w.xbuf[par][f->lane], src, n);
  if (cur_stats && f->lane == 0) cur_stats->warp_collectives++;
  warp_barrier();
  return w.xbuf[par];
}
inline void check_mask(unsigned mask) { if (mask != 0xffffffffu) die("partial-mask warp collective (not supported by the interpreter)"); }

inline void trampoline() {
  cur_invoke(cur_body);
  Fiber* f = cur;
  f->done = true;
  blk.alive--;
  f->warp->alive--;
  if (blk.arrived > 0 && blk.arrived >= blk.alive) { blk.arrived = 0; blk.gen++; }
  Warp& w = *f->warp;
  if (w.arrived > 0 && w.arrived >= w.alive) { w.arrived = 0; w.gen++; }
  swapcontext(&f->ctx, &sched_ctx);
}

inline char* stack_of(int i) {
  static thread_local std::vector<char*> pool;
  if ((int)pool.size() <= i) pool.resize(i + 1, nullptr);
  if (!pool[i]) {
    void* p = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { fprintf(stderr, "[simt] mmap of a fiber stack failed\n"); abort(); }
    pool[i] = (char*)p;
  }
  return pool[i];
}

inline void run_block() {
  const int n = blk.nthreads;
  const int nw = (n + 31) / 32;
  if ((int)fibers.size() < n) fibers.resize(n);
  if ((int)warps.size() < nw) warps.resize(nw);
  for (int w = 0; w < nw; w++) { warps[w].alive = std::min(32, n - 32 * w); warps[w].arrived = 0; warps[w].gen = 0; memset(warps[w].xbuf, 0, sizeof(warps[w].xbuf)); }
  blk.alive = n; blk.arrived = 0; blk.gen = 0;
  for (int i = 0; i < n; i++) {
    Fiber& f = fibers[i];
    f.tid.x = i % blk.bdim.x; f.tid.y = (i / blk.bdim.x) % blk.bdim.y; f.tid.z = i / (blk.bdim.x * blk.bdim.y);
    f.lane = i & 31; f.warp = &warps[i >> 5];
    f.done = false; f.wait = WAIT_NONE; f.xcount = 0; f.mark = 0;
    f.stack = stack_of(i);
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK_BYTES; f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
  }
  // SIMT_ORDER=reverse | shuffle:<seed>: the order in which runnable threads are resumed.  A kernel that is correct on hardware
  // gives the same result for every order; a missing __syncwarp / __syncthreads between a write and a read by another thread is
  // hidden by one order and exposed by another.
  static thread_local int order_mode = -1; static thread_local unsigned order_seed = 1;
  if (order_mode < 0) {
    const char* e = getenv("SIMT_ORDER");
    order_mode = !e ? 0 : !strcmp(e, "reverse") ? 1 : !strncmp(e, "shuffle", 7) ? 2 : 0;
    if (order_mode == 2 && e[7] == ':') order_seed = (unsigned)atoi(e + 8) * 2654435761u + 1u;
  }
  static thread_local std::vector<int> order;
  order.resize(n);
  for (int i = 0; i < n; i++) order[i] = order_mode == 1 ? n - 1 - i : i;
  int remaining = n;
  while (remaining > 0) {
    bool progress = false;
    if (order_mode == 2)
      for (int i = n - 1; i > 0; i--) { order_seed = order_seed * 1664525u + 1013904223u; std::swap(order[i], order[(order_seed >> 8) % (unsigned)(i + 1)]); }
    for (int oi = 0; oi < n; oi++) {
      const int i = order[oi];
      Fiber& f = fibers[i];
      if (f.done) continue;
      if (f.wait == WAIT_BLOCK && blk.gen == f.wait_gen) continue;
      if (f.wait == WAIT_WARP && f.warp->gen == f.wait_gen) continue;
      cur = &f;
      swapcontext(&sched_ctx, &f.ctx);
      progress = true;
      if (f.done) remaining--;
    }
    if (!progress) {
      // who waits where: block barrier (arrived / alive) and every warp's barrier
      fprintf(stderr, "[simt] block barrier: %d arrived of %d alive\n", blk.arrived, blk.alive);
      for (int w = 0; w < nw; w++) {
        int nb = 0, nwp = 0, nd = 0;
        for (int i = 32 * w; i < std::min(n, 32 * w + 32); i++) { if (fibers[i].done) nd++; else if (fibers[i].wait == WAIT_BLOCK) nb++; else if (fibers[i].wait == WAIT_WARP) nwp++; }
        fprintf(stderr, "[simt]   warp %d: %d at the block barrier, %d at a warp barrier / collective, %d finished; marks:", w, nb, nwp, nd);
        for (int i = 32 * w; i < std::min(n, 32 * w + 32); i++) fprintf(stderr, " %d%s", fibers[i].mark, fibers[i].wait == WAIT_WARP ? "w" : "");
        fprintf(stderr, "\n");
      }
      die("deadlock: every live thread waits on a barrier that cannot complete");
    }
  }
  cur = nullptr;
}

template <class F>
inline void invoke_body(const void* p) { (*(const F*)p)(); }

// SIGSEGV inside a kernel: say which kernel / block / thread, and whether the fault address lies in the fiber's stack
inline void on_segv(int, siginfo_t* si, void*) {
  char msg[512];
  const char* a = (const char*)si->si_addr;
  int n;
  if (cur) {
    const bool in_stack = a >= cur->stack - 4096 && a < cur->stack + STACK_BYTES;
    n = snprintf(msg, sizeof(msg), "[simt] SIGSEGV at %p in kernel %s, block (%u,%u,%u), thread (%u,%u,%u)%s\n", (const void*)a, cur_name,
                 blk.idx.x, blk.idx.y, blk.idx.z, cur->tid.x, cur->tid.y, cur->tid.z, in_stack ? " -- fiber stack overflow" : "");
  } else n = snprintf(msg, sizeof(msg), "[simt] SIGSEGV at %p outside any kernel (last kernel %s)\n", (const void*)a, cur_name);
  if (write(2, msg, (size_t)n) < 0) {}
  void* bt[32];
  backtrace_symbols_fd(bt, backtrace(bt, 32), 2);
  _exit(139);
}
inline void install_segv_handler() {
  static bool done = false;
  if (done) return;
  done = true;
  static char alt[64 * 1024];
  stack_t ss; ss.ss_sp = alt; ss.ss_size = sizeof(alt); ss.ss_flags = 0;
  sigaltstack(&ss, nullptr);
  struct sigaction sa; memset(&sa, 0, sizeof(sa));
  sa.sa_sigaction = on_segv; sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
  sigaction(SIGSEGV, &sa, nullptr);
}

// grid x block fibers; dynamic shared memory is poisoned (0xFF bytes = NaN doubles) at every block start
template <class F>
inline void launch(const char* name, dim3 grid, dim3 block, size_t smem, const F& body) {
  const size_t n = (size_t)block.x * block.y * block.z;
  cur_name = name;
  install_segv_handler();
  if (n == 0 || n > MAX_THREADS) die("bad block size");
  if ((size_t)grid.x * grid.y * grid.z == 0) die("empty grid");
  if (smem > 227 * 1024) die("more than 227 KB of dynamic shared memory");
  total_launches++;
  // Blocks run one after the other, so a block that spins on a flag another block of the same launch publishes would never return.
  // k_peer_allreduce (csrc/peer_allreduce.cuh) is such a kernel and is written grid-size agnostic (grid-stride loops, "last block"
  // by counter): the interpreter runs it with ONE block; its spin on the flags of the OTHER ranks is served by their host threads.
  // k_lm (csrc/lm_kernel.cuh) walks virtual blocks between grid barriers and is grid-size agnostic the same way.
  if (strstr(name, "k_lm")) grid = dim3(1);
  blk.nthreads = (int)n; blk.bdim = block; blk.gdim = grid;
  std::vector<unsigned char> dyn(smem + 64);
  blk.dyn = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
  blk.dyn_bytes = smem;
  cur_body = &body; cur_invoke = &invoke_body<F>;
  if (stats_on()) {          // single-rank runs only (the map is not locked)
    register_stats_dump();
    cur_stats = &stats()[name];
    cur_stats->launches++; cur_stats->blocks += (long long)grid.x * grid.y * grid.z;
  } else cur_stats = nullptr;
  for (unsigned z = 0; z < grid.z; z++)
    for (unsigned y = 0; y < grid.y; y++)
      for (unsigned x = 0; x < grid.x; x++) {
        blk.idx.x = x; blk.idx.y = y; blk.idx.z = z;
        memset(blk.dyn, 0xFF, smem);
        run_block();
      }
  blk.dyn = nullptr;
}
template <class F> inline void launch(const char* name, dim3 g, dim3 b, const F& body) { launch(name, g, b, 0, body); }
template <class F, class S> inline void launch(const char* name, dim3 g, dim3 b, size_t smem, S /*stream*/, const F& body) { launch(name, g, b, smem, body); }

inline void* dyn_smem() { return blk.dyn; }

// fp64 tensor-core fragment semantics of mma.sync.aligned.m8n8k4.row.col.f64 (PTX ISA "Matrix Fragments for mma.m8n8k4 with
// .f64"): lane = 4*grp + tig holds A[grp][tig], B[tig][grp] and C[grp][2*tig], C[grp][2*tig+1]
inline void dmma884(double& c0, double& c1, double a, double b) {
  const double ab[2] = {a, b};
  unsigned char (*t)[16] = exchange(ab, 16);
  const int grp = cur->lane >> 2, tig = cur->lane & 3;
  for (int k = 0; k < 4; k++) {
    double A, B0, B1;
    memcpy(&A, t[4 * grp + k], 8);                         // A[grp][k]
    memcpy(&B0, t[4 * (2 * tig) + k] + 8, 8);              // B[k][2*tig]
    memcpy(&B1, t[4 * (2 * tig + 1) + k] + 8, 8);          // B[k][2*tig+1]
    c0 += A * B0;
    c1 += A * B1;
  }
}

}  // namespace simt

// ---------------------------------------------------------------- CUDA language surface used by multical_b200/csrc
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __noinline__
#define threadIdx (simt::cur->tid)
#define blockIdx (simt::blk.idx)
#define blockDim (simt::blk.bdim)
#define gridDim (simt::blk.gdim)
constexpr int warpSize = 32;

static inline void __syncthreads() { simt::block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::check_mask(mask); simt::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}

template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  static_assert(sizeof(T) <= 16, "shuffle payload");
  simt::check_mask(mask);
  const int lane = simt::cur->lane;
  unsigned char (*t)[16] = simt::exchange(&v, sizeof(T));
  const int s = (lane & ~(width - 1)) | (src & (width - 1));
  T r; memcpy(&r, t[s], sizeof(T)); return r;
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  simt::check_mask(mask);
  const int lane = simt::cur->lane;
  unsigned char (*t)[16] = simt::exchange(&v, sizeof(T));
  const int s = lane ^ lanemask;
  if ((s & ~(width - 1)) != (lane & ~(width - 1))) return v;
  T r; memcpy(&r, t[s], sizeof(T)); return r;
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  simt::check_mask(mask);
  const int lane = simt::cur->lane;
  unsigned char (*t)[16] = simt::exchange(&v, sizeof(T));
  const int s = lane - (int)delta;
  if (s < (lane & ~(width - 1))) return v;
  T r; memcpy(&r, t[s], sizeof(T)); return r;
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  simt::check_mask(mask);
  const int lane = simt::cur->lane;
  unsigned char (*t)[16] = simt::exchange(&v, sizeof(T));
  const int s = lane + (int)delta;
  if (s >= (lane & ~(width - 1)) + width) return v;
  T r; memcpy(&r, t[s], sizeof(T)); return r;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  simt::check_mask(mask);
  // a lane that voted may have left the kernel before a slower lane reads the table: votes carry the collective's sequence
  // number, stale or never-written slots do not match it
  struct { int p; unsigned tag; } mine = {pred ? 1 : 0, simt::cur->xcount + 1u}, v;
  unsigned char (*t)[16] = simt::exchange(&mine, sizeof(mine));
  unsigned r = 0;
  for (int l = 0; l < 32; l++) {
    memcpy(&v, t[l], sizeof(v));
    if (v.tag == mine.tag && v.p) r |= 1u << l;
  }
  return r;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcv(const T* p) { return *(const volatile T*)p; }

// blocks run one after the other and fibers only switch at barriers / collectives, so plain read-modify-write is atomic
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline long long clock64() { return 0; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
#define __align__(n) __attribute__((aligned(n)))
static inline unsigned atomicAdd(unsigned* p, int v) { unsigned o = *p; *p = o + (unsigned)v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline double min(double a, double b) { return fmin(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
