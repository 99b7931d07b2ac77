// solver_kernels.cuh — on-device trust-region machinery.
//
// Follows scipy.optimize._lsq.trf.trf_no_bounds (the solver behind calibration.py:209-210) step by step:
// jac scaling (common.py compute_jac_scale), regularised Gauss-Newton direction, 2-D subspace trust-region
// step, ratio test / radius update (update_tr_radius), termination tests (check_termination).
// The one deliberate change: scipy's LSMR inner iteration is replaced by an EXACT solve of
//     (A + reg I) gn = g_h ,   A = D H D  (H = J^T J block-arrow: shared | per-frame 6x6)
// through the Schur complement of the per-frame blocks (S = A_ss + reg I - sum_f Y_f Y_f^T).
#pragma once
#include "kernels.cuh"

namespace mcba {

struct SolverState {
  double ftol, xtol, gtol, reg_floor;
  int max_nfev;
  int nfev, njev, iteration, status, accepted, first_scale, chol_fail, done;
  double cost, cost_new, Delta, reg;
  double g_norm, gh_norm;
  double alpha, beta;            // step_h = alpha*gh + beta*gn
  double B11, B12, B22, gS1, gS2, n1, n2, mu;
  double step_h_norm, predicted, step_norm, x_norm, actual_reduction, ratio;
  double last_step_norm, last_reduction;
};

// slots of the cross-rank reduction scratch `red` (doubles). Entries marked F hold only the contribution of
// this rank's frames and are summed (or maxed) over ranks; S entries are computed from replicated data.
enum {
  RED_GH2_F = 0, RED_XS2_F,                                     // k_scale, frame parts        (1 all-reduce, sum)
  RED_GMAX_F,                                                    // k_scale                     (all-reduce, max)
  RED_AGG, RED_AGN, RED_ANN, RED_DOTGN_F, RED_GN2_F,           // quadratic forms + dots      (1 all-reduce, sum)
  RED_COSTNEW, RED_STEP2_F, RED_XN2_F,                          // trial step                  (1 all-reduce, sum)
  RED_COST,                                                      // cost at the linearisation   (grouped with g_s, diag_s)
  RED_GH2_S, RED_XS2_S, RED_GMAX_S, RED_DOTGN_S, RED_GN2_S, RED_STEP2_S, RED_XN2_S,   // replicated (shared) parts: never reduced
  RED_COUNT
};

// ---- NVLink peer-memory exchange (peer_allreduce.cuh): argument block, also taken by the kernels that run an exchange as their tail
constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_MAX_SEG = 6;
constexpr int PEER_FLAG_STRIDE = 8;      // doubles (64 B) between flags

struct PeerSeg { double* buf; int count; int op; };     // op 0 = sum, 1 = max
struct SolverState;
struct PeerArgs {
  PeerSeg seg[PEER_MAX_SEG];
  int nseg, rank, world, cap;
  unsigned seq;
  double* base[PEER_MAX_WORLD];          // peer-mapped base pointer of every rank's buffer (base[rank] = own)
  unsigned* counter;
  int epilogue;                          // EPI_* scalar step run by thread 0 after the reduction (single-CTA exchanges only)
  SolverState* st;
  double* red;
};

__host__ __device__ inline size_t peer_flag_off(int world, int parity, int src) { return (size_t)(parity * world + src) * PEER_FLAG_STRIDE; }
__host__ __device__ inline size_t peer_data_off(int world, int cap, int parity, int src) {
  return (size_t)2 * world * PEER_FLAG_STRIDE + ((size_t)parity * world + src) * cap;
}
__host__ __device__ inline size_t peer_buffer_doubles(int world, int cap) { return (size_t)2 * world * PEER_FLAG_STRIDE + (size_t)2 * world * cap; }

__device__ void peer_allreduce_block(const PeerArgs& a);

__device__ __forceinline__ double block_sum(double v, double* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = lane < nw ? sm[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;   // valid in thread 0
}
__device__ __forceinline__ double block_max(double v, double* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = lane < nw ? sm[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
  }
  return r;
}

// deterministic sum of per-CTA partials: out[j] = sum_i part[i*stride + j]
__global__ void k_sum_partials(const double* part, int count, int stride, int nout, double* out) {
  __shared__ double sm[32];
  for (int j = 0; j < nout; j++) {
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += part[(size_t)i * stride + j];
    s = block_sum(s, sm);
    if (threadIdx.x == 0) out[j] = s;
    __syncthreads();
  }
}

// extract diag(H_ss) so that it can be all-reduced as a vector
__global__ void k_diag(const double* Hss, int n_s, double* diag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_s) diag[i] = Hss[(size_t)i * n_s + i];
}

// trf.py top of the outer loop: ||g||_inf, gtol / max_nfev exits, first-iteration cost and Delta.
__device__ inline void begin_iteration(SolverState* st, const double* red) {
  const double gh2 = red[RED_GH2_S] + red[RED_GH2_F];
  st->gh_norm = sqrt(gh2);
  st->g_norm = fmax(red[RED_GMAX_S], red[RED_GMAX_F]);
  if (st->first_scale) {
    st->cost = red[RED_COST];
    double D0 = sqrt(red[RED_XS2_S] + red[RED_XS2_F]);
    st->Delta = (D0 == 0.0) ? 1.0 : D0;
    st->first_scale = 0;
  }
  if (st->g_norm < st->gtol) st->status = 1;
  st->done = (st->status != -99) || (st->nfev >= st->max_nfev);
}
__global__ void k_begin_iteration(SolverState* st, double* red) { begin_iteration(st, red); }

// common.py compute_jac_scale: scale_inv = ||J[:,i]|| = sqrt(H_ii); zero -> 1 on the first call, running max after;
// also g_h = d*g, ||g||_inf, ||g_h||^2 and ||x*scale_inv||^2 (initial Delta, trf.py).  Single CTA.
// Single-GPU fast path (fused != 0): reads diag(H_ss) in place, sums the per-CTA cost partials of k_expand_shared and
// runs the begin-of-iteration logic itself; with several ranks those three need all-reduces in between.
__device__ __forceinline__ void scale_body(int n, int n_s, const double* diag_s, const double* Hss, const double* Hff, const double* g, const double* x,
                                           double* sinv, double* d, double* gh, int first, double* red,
                                           int fused, const double* cost_part, int n_cost_part, SolverState* st, int fb) {
  __shared__ double sm[32];
  if (fused && st->done) return;
  double gh2s = 0, gh2f = 0, gms = 0, gmf = 0, xs2s = 0, xs2f = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double hd;
    if (i < n_s) hd = fused ? __ldcg(&Hss[(size_t)i * n_s + i]) : diag_s[i];       // (written by other CTAs' atomics when this runs as an epilogue)
    else { const int f = (i - n_s) / fb, j = (i - n_s) % fb; hd = Hff[(size_t)f * fb * fb + j * (fb + 1)]; }
    double nrm = sqrt(fmax(hd, 0.0));
    double si;
    if (first) si = (nrm == 0.0) ? 1.0 : nrm; else si = fmax(nrm, sinv[i]);
    sinv[i] = si;
    const double di = 1.0 / si;
    d[i] = di;
    const double gi = __ldcg(&g[i]), ghi = di * gi, xs = x[i] * si;
    gh[i] = ghi;
    if (i < n_s) { gh2s += ghi * ghi; gms = fmax(gms, fabs(gi)); xs2s += xs * xs; }
    else { gh2f += ghi * ghi; gmf = fmax(gmf, fabs(gi)); xs2f += xs * xs; }
  }
  double r;
  r = block_sum(gh2s, sm); if (threadIdx.x == 0) red[RED_GH2_S] = r;
  r = block_sum(gh2f, sm); if (threadIdx.x == 0) red[RED_GH2_F] = r;
  r = block_max(gms, sm);  if (threadIdx.x == 0) red[RED_GMAX_S] = r;
  r = block_max(gmf, sm);  if (threadIdx.x == 0) red[RED_GMAX_F] = r;
  r = block_sum(xs2s, sm); if (threadIdx.x == 0) red[RED_XS2_S] = r;
  r = block_sum(xs2f, sm); if (threadIdx.x == 0) red[RED_XS2_F] = r;
  if (fused) {
    double c = 0.0;
    for (int i = threadIdx.x; i < n_cost_part; i += blockDim.x) c += __ldcg(&cost_part[i]);
    c = block_sum(c, sm);
    if (threadIdx.x == 0) { if (n_cost_part > 0) red[RED_COST] = c; begin_iteration(st, red); }      // no partials: the cost is already in place
  }
}
__global__ void k_scale(int n, int n_s, const double* diag_s, const double* Hss, const double* Hff, const double* g, const double* x,
                        double* sinv, double* d, double* gh, int first, double* red,
                        int fused, const double* cost_part, int n_cost_part, SolverState* st, int fb) {
  scale_body(n, n_s, diag_s, Hss, Hff, g, x, sinv, d, gh, first, red, fused, cost_part, n_cost_part, st, fb);
}
// Multi-GPU with MCBA_FUSE=1: the scaling in two parts around ONE exchange instead of two.  part 1 = the frame entries, which only
// need local data (H_ff, g_f, x_f): their sinv / d / gh and RED_GH2_F, RED_XS2_F, RED_GMAX_F, which then travel with g_s, diag(H_ss)
// and the cost; part 2 = the shared entries from the reduced diagonal / gradient, then begin_iteration -- every sum it reads is
// reduced by then, so the second exchange of the iteration disappears.
__global__ void k_scale_part(int part, int n, int n_s, const double* diag_s, const double* Hff, const double* g, const double* x,
                             double* sinv, double* d, double* gh, int first, double* red, SolverState* st, int fb) {
  __shared__ double sm[32];
  const int lo = part == 1 ? n_s : 0, hi = part == 1 ? n : n_s;
  double gh2 = 0, gm = 0, xs2 = 0;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    double hd;
    if (i < n_s) hd = diag_s[i];
    else { const int f = (i - n_s) / fb, j = (i - n_s) % fb; hd = Hff[(size_t)f * fb * fb + j * (fb + 1)]; }
    const double nrm = sqrt(fmax(hd, 0.0));
    double si;
    if (first) si = (nrm == 0.0) ? 1.0 : nrm; else si = fmax(nrm, sinv[i]);
    sinv[i] = si;
    const double di = 1.0 / si;
    d[i] = di;
    const double gi = g[i], ghi = di * gi, xs = x[i] * si;
    gh[i] = ghi;
    gh2 += ghi * ghi; gm = fmax(gm, fabs(gi)); xs2 += xs * xs;
  }
  double r;
  r = block_sum(gh2, sm); if (threadIdx.x == 0) red[part == 1 ? RED_GH2_F : RED_GH2_S] = r;
  r = block_max(gm, sm);  if (threadIdx.x == 0) red[part == 1 ? RED_GMAX_F : RED_GMAX_S] = r;
  r = block_sum(xs2, sm); if (threadIdx.x == 0) red[part == 1 ? RED_XS2_F : RED_XS2_S] = r;
  if (part == 2 && threadIdx.x == 0) begin_iteration(st, red);
}

// Both parts, the cost sum, diag(H_ss) and the exchange in between as ONE single-CTA launch (MCBA_FUSE=1 with peer buffers).
__global__ void __launch_bounds__(1024)
k_scale_exchange(int n, int n_s, const double* Hss, double* diag_s, const double* Hff, const double* g, const double* x, double* sinv, double* d, double* gh,
                 int first, double* red, const double* cost_part, int n_cost_part, SolverState* st, int fb, PeerArgs pa) {
  __shared__ double sm[32];
  double c = 0.0;
  for (int i = threadIdx.x; i < n_cost_part; i += blockDim.x) c += cost_part[i];
  c = block_sum(c, sm);
  if (threadIdx.x == 0) red[RED_COST] = c;
  for (int i = threadIdx.x; i < n_s; i += blockDim.x) diag_s[i] = Hss[(size_t)i * n_s + i];
  double gh2 = 0, gm = 0, xs2 = 0;
  for (int i = n_s + threadIdx.x; i < n; i += blockDim.x) {              // frame entries: local data only
    const int f = (i - n_s) / fb, j = (i - n_s) % fb;
    const double nrm = sqrt(fmax(Hff[(size_t)f * fb * fb + j * (fb + 1)], 0.0));
    double si;
    if (first) si = (nrm == 0.0) ? 1.0 : nrm; else si = fmax(nrm, sinv[i]);
    sinv[i] = si;
    const double di = 1.0 / si;
    d[i] = di;
    const double gi = g[i], ghi = di * gi, xs = x[i] * si;
    gh[i] = ghi;
    gh2 += ghi * ghi; gm = fmax(gm, fabs(gi)); xs2 += xs * xs;
  }
  double r;
  r = block_sum(gh2, sm); if (threadIdx.x == 0) red[RED_GH2_F] = r;
  r = block_max(gm, sm);  if (threadIdx.x == 0) red[RED_GMAX_F] = r;
  r = block_sum(xs2, sm); if (threadIdx.x == 0) red[RED_XS2_F] = r;
  __syncthreads();
  peer_allreduce_block(pa);                                              // g_s, diag_s, cost, frame sums (sum) and the frame maximum (max)
  __syncthreads();
  gh2 = 0; gm = 0; xs2 = 0;
  for (int i = threadIdx.x; i < n_s; i += blockDim.x) {                  // shared entries from the reduced diagonal / gradient
    const double nrm = sqrt(fmax(diag_s[i], 0.0));
    double si;
    if (first) si = (nrm == 0.0) ? 1.0 : nrm; else si = fmax(nrm, sinv[i]);
    sinv[i] = si;
    const double di = 1.0 / si;
    d[i] = di;
    const double gi = g[i], ghi = di * gi, xs = x[i] * si;
    gh[i] = ghi;
    gh2 += ghi * ghi; gm = fmax(gm, fabs(gi)); xs2 += xs * xs;
  }
  r = block_sum(gh2, sm); if (threadIdx.x == 0) red[RED_GH2_S] = r;
  r = block_max(gm, sm);  if (threadIdx.x == 0) red[RED_GMAX_S] = r;
  r = block_sum(xs2, sm); if (threadIdx.x == 0) red[RED_XS2_S] = r;
  if (threadIdx.x == 0) begin_iteration(st, red);
}

// tail of k_expand_shared (MCBA_FUSE=1, single GPU, no later kernel adds to H_ss): the last CTA to finish does what k_scale does
__device__ __noinline__ void scale_epilogue(const ScaleEpilogue& e, int n_s, const double* Hss, const double* g) {
  __shared__ int scale_is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) { const unsigned t = atomicAdd(e.counter, 1u); scale_is_last = (t == gridDim.x - 1); }
  __syncthreads();
  if (!scale_is_last) return;
  if (threadIdx.x == 0) *e.counter = 0;
  scale_body(e.n, n_s, nullptr, Hss, e.Hff, g, e.x, e.sinv, e.d, e.gh, e.first, e.red, 1, e.cost_part, e.n_cost_part, e.st, e.fb);
}

// quadratic forms u^T A v, A = D H D, for (u,u) [, (u,v), (v,v)].  Grid = F frame CTAs + shared CTAs.
// partial[cta][3].  u, v are scaled-space vectors (gh, gn).
__device__ inline void reg_compute(SolverState* st, const double* red);
__device__ inline void subspace_compute(SolverState* st, const double* red);

constexpr int QUAD_THREADS = 128;
constexpr int QUAD_WARPS = QUAD_THREADS / 32;
// blocks [0, frame_blocks): one WARP per frame (W_f^T u_s by lane-strided sums + shuffles, then the 6x6 part);
// remaining blocks: one thread per shared row (column walk of the symmetric H_ss is coalesced).
// partial[frame or F + shared block][QS].  FB = parameters per frame block (6; 12 for RollingFrames' start+end pose).
// DOTS (opt-in MCBA_FUSE=1, with two != 0): the records also carry u.v and v.v of the same rows (QS = 5), which is what k_dots
// computes for (gh, gn) in a launch of its own; the last block then fills RED_DOTGN_* / RED_GN2_* as well.
// XCHG (finalize 4, MCBA_FUSE=1 on several GPUs): the last block also all-reduces the sums over the ranks and runs pa.epilogue.
template <int FB, bool DOTS = false, bool XCHG = false>
__global__ void __launch_bounds__(QUAD_THREADS)
k_quad(int n_s, int F, int motion_on, const double* Hss, const double* Hff, const double* W, const double* d,
       const double* u, const double* v, int two, double* partial,
       int finalize /*0 none, 1 sum, 2 sum+reg, 3 sum+subspace, 4 sum + all-reduce over the ranks + pa.epilogue*/, unsigned* counter, double* red,
       SolverState* st, PeerArgs pa) {
  __shared__ double sm[32];
  __shared__ int is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int QS = DOTS ? 5 : 3;
  const int nframe = motion_on ? F : 0;
  const int frame_blocks = (nframe + QUAD_WARPS - 1) / QUAD_WARPS;
  if ((int)blockIdx.x < frame_blocks) {
    const int f = blockIdx.x * QUAD_WARPS + warp;
    if (f < nframe) {
      const double* Wf = W + (size_t)f * n_s * FB;
      double tu[FB], tv[FB];
#pragma unroll
      for (int j = 0; j < FB; j++) { tu[j] = 0.0; tv[j] = 0.0; }
      for (int s = lane; s < n_s; s += 32) {
        const double us = d[s] * u[s], vs = two ? d[s] * v[s] : 0.0;
#pragma unroll
        for (int j = 0; j < FB; j++) { const double w = Wf[s * FB + j]; tu[j] += w * us; tv[j] += w * vs; }
      }
#pragma unroll
      for (int j = 0; j < FB; j++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { tu[j] += __shfl_xor_sync(0xffffffffu, tu[j], o); tv[j] += __shfl_xor_sync(0xffffffffu, tv[j], o); }
      }
      if (lane == 0) {
        double uf[FB], vf[FB], uu = 0, uv = 0, vv = 0;
#pragma unroll
        for (int j = 0; j < FB; j++) { const int i = n_s + FB * f + j; uf[j] = d[i] * u[i]; vf[j] = two ? d[i] * v[i] : 0.0; }
        const double* H = Hff + (size_t)f * FB * FB;
#pragma unroll
        for (int i = 0; i < FB; i++) {
          double hu = 0, hv = 0;
#pragma unroll
          for (int j = 0; j < FB; j++) { hu += H[i * FB + j] * uf[j]; hv += H[i * FB + j] * vf[j]; }
          uu += uf[i] * (hu + 2.0 * tu[i]);
          uv += uf[i] * hv + uf[i] * tv[i] + vf[i] * tu[i];
          vv += vf[i] * (hv + 2.0 * tv[i]);
        }
        partial[(size_t)f * QS + 0] = uu; partial[(size_t)f * QS + 1] = uv; partial[(size_t)f * QS + 2] = vv;
        if constexpr (DOTS) {
          double dt = 0, g2 = 0;
#pragma unroll
          for (int j = 0; j < FB; j++) { const int i = n_s + FB * f + j; dt += u[i] * v[i]; g2 += v[i] * v[i]; }
          partial[(size_t)f * QS + 3] = dt; partial[(size_t)f * QS + 4] = g2;
        }
      }
    }
  } else {
    const int sb = blockIdx.x - frame_blocks;
    const int i = sb * QUAD_THREADS + tid;
    double uu = 0, uv = 0, vv = 0;
    if (i < n_s) {
      double hu = 0, hv = 0;
      for (int j = 0; j < n_s; j++) {
        const double h = Hss[(size_t)j * n_s + i];
        hu += h * d[j] * u[j];
        if (two) hv += h * d[j] * v[j];
      }
      const double ui = d[i] * u[i], vi = two ? d[i] * v[i] : 0.0;
      uu = ui * hu; uv = ui * hv; vv = vi * hv;
    }
    double r;
    r = block_sum(uu, sm); if (tid == 0) partial[(size_t)(nframe + sb) * QS + 0] = r;
    r = block_sum(uv, sm); if (tid == 0) partial[(size_t)(nframe + sb) * QS + 1] = r;
    r = block_sum(vv, sm); if (tid == 0) partial[(size_t)(nframe + sb) * QS + 2] = r;
    if constexpr (DOTS) {
      const double dt = i < n_s ? u[i] * v[i] : 0.0, g2 = i < n_s ? v[i] * v[i] : 0.0;
      r = block_sum(dt, sm); if (tid == 0) partial[(size_t)(nframe + sb) * QS + 3] = r;
      r = block_sum(g2, sm); if (tid == 0) partial[(size_t)(nframe + sb) * QS + 4] = r;
    }
  }
  if (!finalize) return;
  // last-block reduction: deterministic (index-ordered) sum of all partial records, then the scalar step that needs it
  __threadfence();
  __syncthreads();
  if (tid == 0) { const unsigned t = atomicAdd(counter, 1u); is_last = (t == gridDim.x - 1); }
  __syncthreads();
  if (!is_last) return;
  const int nparts = nframe + (int)gridDim.x - frame_blocks;
  const int nout = two ? 3 : 1;
  for (int j = 0; j < nout; j++) {
    double acc = 0.0;
    for (int i = tid; i < nparts; i += QUAD_THREADS) acc += __ldcg(&partial[(size_t)i * QS + j]);
    acc = block_sum(acc, sm);
    if (tid == 0) red[RED_AGG + j] = acc;
    __syncthreads();
  }
  if constexpr (DOTS) {          // frame records -> the _F slots (summed over ranks), shared records -> the replicated _S slots
    for (int j = 0; j < 2; j++) {
      double af = 0.0, as = 0.0;
      for (int i = tid; i < nparts; i += QUAD_THREADS) { const double val = __ldcg(&partial[(size_t)i * QS + 3 + j]); if (i < nframe) af += val; else as += val; }
      af = block_sum(af, sm);
      if (tid == 0) red[j == 0 ? RED_DOTGN_F : RED_GN2_F] = af;
      __syncthreads();
      as = block_sum(as, sm);
      if (tid == 0) red[j == 0 ? RED_DOTGN_S : RED_GN2_S] = as;
      __syncthreads();
    }
  }
  if (tid == 0) {
    *counter = 0;
    if (finalize == 2) reg_compute(st, red);
    else if (finalize == 3) subspace_compute(st, red);
  }
  if constexpr (XCHG) { if (finalize == 4) { __syncthreads(); peer_allreduce_block(pa); } }      // the sums of all ranks, then the scalar step, in this launch
}

// trf.py: reg_term = -ag_value / Delta^2 with ag_value = min over [0, Delta/||g_h||] of a t^2 + b t,
// a = g_h^T A g_h, b = -||g_h||^2 (build_quadratic_1d / minimize_quadratic_1d).
__device__ inline void reg_compute(SolverState* st, const double* red) {
  const double a = red[RED_AGG];
  const double gh2 = st->gh_norm * st->gh_norm;
  const double b = -gh2;
  const double ub = st->Delta / st->gh_norm;
  // minimize a t^2 + b t on [0, ub]
  double best_t = 0.0, best = 0.0;
  { const double yv = a * ub * ub + b * ub; if (yv < best) { best = yv; best_t = ub; } }
  if (a != 0.0) { const double ext = -0.5 * b / a; if (ext > 0.0 && ext < ub) { const double yv = a * ext * ext + b * ext; if (yv < best) { best = yv; best_t = ext; } } }
  (void)best_t;
  double reg = -best / (st->Delta * st->Delta);
  if (!(reg > st->reg_floor)) reg = st->reg_floor;
  st->reg = reg;
}

__global__ void k_reg(SolverState* st, const double* red) { reg_compute(st, red); }

// per frame: L L^T = D_f H_ff D_f + reg I ; Y_f = (D_s W_f D_f) L^-T (n_s x 6) ; z_f = L^-1 (D_f g_f)
constexpr int SCHUR_THREADS = 128;
template <int FB>
__global__ void __launch_bounds__(SCHUR_THREADS)
k_schur_frames(int n_s, const double* Hff, const double* W, const double* d, const double* gh,
               const SolverState* st, double* Y, double* Lf, double* zf, const double* Hss, double* S, double* rhs) {
  __shared__ double L[FB * FB];
  __shared__ double df[FB];
  const int f = blockIdx.x, tid = threadIdx.x;
  // S_local = D_s H_ss D_s, rhs_local = 0 (grid-stride; the SYRK kernel that follows subtracts sum_f Y_f Y_f^T)
  for (size_t idx = (size_t)blockIdx.x * SCHUR_THREADS + tid; idx < (size_t)n_s * n_s; idx += (size_t)gridDim.x * SCHUR_THREADS) {
    const int i = idx / n_s, j = idx % n_s;
    S[idx] = d[i] * d[j] * Hss[idx];
    if (idx < (size_t)n_s) rhs[idx] = 0.0;
  }
  if (tid == 0) {
    const double reg = st->reg;
    const double* H = Hff + (size_t)f * FB * FB;
    double A[FB * FB];
    for (int j = 0; j < FB; j++) df[j] = d[n_s + FB * f + j];
    for (int i = 0; i < FB; i++) for (int j = 0; j < FB; j++) A[i * FB + j] = df[i] * df[j] * H[i * FB + j] + (i == j ? reg : 0.0);
    for (int j = 0; j < FB; j++) {
      double s = A[j * FB + j];
      for (int k = 0; k < j; k++) s -= L[j * FB + k] * L[j * FB + k];
      const double piv = sqrt(fmax(s, 1e-300));
      L[j * FB + j] = piv;
      for (int i = j + 1; i < FB; i++) {
        double t = A[i * FB + j];
        for (int k = 0; k < j; k++) t -= L[i * FB + k] * L[j * FB + k];
        L[i * FB + j] = t / piv;
      }
      for (int i = 0; i < j; i++) L[i * FB + j] = 0.0;
    }
    double z[FB];
    for (int i = 0; i < FB; i++) {
      double t = gh[n_s + FB * f + i];
      for (int k = 0; k < i; k++) t -= L[i * FB + k] * z[k];
      z[i] = t / L[i * FB + i];
      zf[(size_t)f * FB + i] = z[i];
    }
    for (int i = 0; i < FB * FB; i++) Lf[(size_t)f * FB * FB + i] = L[i];
  }
  __syncthreads();
  const double* Wf = W + (size_t)f * n_s * FB;
  double* Yf = Y + (size_t)f * n_s * FB;
  for (int s = tid; s < n_s; s += SCHUR_THREADS) {
    const double ds = d[s];
    double y[FB];
#pragma unroll
    for (int i = 0; i < FB; i++) {
      double t = ds * Wf[s * FB + i] * df[i];
#pragma unroll
      for (int k = 0; k < i; k++) t -= L[i * FB + k] * y[k];
      y[i] = t / L[i * FB + i];
    }
#pragma unroll
    for (int i = 0; i < FB; i++) Yf[s * FB + i] = y[i];
  }
}

// S_local = D_s H_ss D_s ; rhs_local = 0      (the reg*I and D_s g_s terms are added after the all-reduce)
__global__ void k_schur_init(int n_s, const double* Hss, const double* d, double* S, double* rhs) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (size_t)n_s * n_s) { const int i = idx / n_s, j = idx % n_s; S[idx] = d[i] * d[j] * Hss[idx]; }
  if (idx < (size_t)n_s) rhs[idx] = 0.0;
}

// S -= sum_f Y_f Y_f^T over this CTA's frame chunk; 32x32 output tile per CTA, 2x2 micro-tile per thread.
constexpr int SYRK_TILE = 32;
__host__ __device__ constexpr int syrk_fr(int fb) { return 48 / fb; }       // frames staged per step: 2 x 12 KB of shared memory
template <int FB>
__global__ void __launch_bounds__(256)
k_schur_syrk(int n_s, int F, int chunk_frames, const double* Y, double* S, const double* zf, double* rhs) {
  constexpr int SYRK_FR = syrk_fr(FB);
  __shared__ double Yi[SYRK_FR][SYRK_TILE][FB];
  __shared__ double Yj[SYRK_FR][SYRK_TILE][FB];
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj < ti) return;
  const int f0 = blockIdx.z * chunk_frames, f1 = min(F, f0 + chunk_frames);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  double acc[2][2] = {{0, 0}, {0, 0}};
  double racc = 0.0;                     // diagonal tiles also accumulate rhs -= Y_f z_f for their 32 rows
  for (int fb = f0; fb < f1; fb += SYRK_FR) {
    const int nf = min(SYRK_FR, f1 - fb);
    for (int o = threadIdx.x; o < SYRK_FR * SYRK_TILE * FB; o += 256) {
      const int ff = o / (SYRK_TILE * FB), rem = o % (SYRK_TILE * FB), r = rem / FB, k = rem % FB;
      const int gi = ti * SYRK_TILE + r, gj = tj * SYRK_TILE + r;
      (&Yi[0][0][0])[o] = (ff < nf && gi < n_s) ? Y[((size_t)(fb + ff) * n_s + gi) * FB + k] : 0.0;
      (&Yj[0][0][0])[o] = (ff < nf && gj < n_s) ? Y[((size_t)(fb + ff) * n_s + gj) * FB + k] : 0.0;
    }
    __syncthreads();
    for (int ff = 0; ff < nf; ff++) {
#pragma unroll
      for (int k = 0; k < FB; k++) {
        const double a0 = Yi[ff][ty][k], a1 = Yi[ff][ty + 16][k], b0 = Yj[ff][tx][k], b1 = Yj[ff][tx + 16][k];
        acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
      }
    }
    if (ti == tj && threadIdx.x < SYRK_TILE) {
      for (int ff = 0; ff < nf; ff++) {
        const double* z = zf + (size_t)(fb + ff) * FB;
#pragma unroll
        for (int k = 0; k < FB; k++) racc += Yi[ff][threadIdx.x][k] * z[k];
      }
    }
    __syncthreads();
  }
  if (ti == tj && threadIdx.x < SYRK_TILE) {
    const int i = ti * SYRK_TILE + threadIdx.x;
    if (i < n_s) atomicAdd(&rhs[i], -racc);
  }
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const int i = ti * SYRK_TILE + ty + 16 * a, j = tj * SYRK_TILE + tx + 16 * b;
      if (i < n_s && j < n_s && (ti != tj || j >= i)) {
        atomicAdd(&S[(size_t)i * n_s + j], -acc[a][b]);
        if (i != j) atomicAdd(&S[(size_t)j * n_s + i], -acc[a][b]);
      }
    }
}

// Dense SPD solve of the reduced (shared-parameter) system   (S + reg I) p_s = rhs + D_s g_s  -> gn[0..n_s)
//  * n_s <= CHOL_SMALL_MAX : one CTA, matrix resident in shared memory (k_chol_small)
//  * larger                : right-looking blocked Cholesky, NB=32 panels: k_chol_diag (1 CTA) -> k_chol_trsm
//                            (row chunks) -> k_chol_syrk (tiles), then k_chol_substitute (1 CTA)
constexpr int CHOL_SMALL_MAX = 128;
constexpr int CHOL_SMALL_THREADS = 256;
// One CTA of 16x16 threads; the matrix lives in REGISTERS, cyclically distributed: thread (ty,tx) owns A[ty+16p][tx+16q],
// p,q < R (R = ceil(n/16) <= 8).  Per column: the pivot and the scaled column go through shared memory (2 barriers), the
// rank-1 update is R*R predicated FMAs on registers.  The factor is then written to shared memory and warp 0 does both
// substitutions with the right-hand side in registers and one shuffle broadcast per column.
template <int R>
__global__ void __launch_bounds__(CHOL_SMALL_THREADS)
k_chol_small(int n, const double* Sg, const double* rhs, const double* gh, SolverState* st, double* out) {
  extern __shared__ double shm[];
  const int ld = n | 1;
  double* Lm = shm;                       // n x ld   (written after the factorisation)
  double* colbuf = Lm + (size_t)n * ld;   // n
  double* invd = colbuf + n;              // n        1 / L_kk
  double* piv = invd + n;                 // 1 (+1 pad)
  // column index is the SLOW thread index: all owners of a matrix column sit in one half-warp, so only that warp
  // executes the pivot / column-scaling code
  const int tid = threadIdx.x, ty = tid & 15, tx = tid >> 4;
  const double reg = st->reg;
  double a[R][R];
#pragma unroll
  for (int p = 0; p < R; p++)
#pragma unroll
    for (int q = 0; q < R; q++) {
      const int i = ty + 16 * p, j = tx + 16 * q;
      a[p][q] = (i < n && j < n) ? Sg[(size_t)j * n + i] + (i == j ? reg : 0.0) : 0.0;     // S is symmetric: coalesced read
    }
  if (tid == 0) piv[0] = a[0][0];
  __syncthreads();
  for (int k = 0; k < n; k++) {
    const int kq = k >> 4, kt = k & 15;
    if (tx == kt) {
      const double akk = piv[0];
      if (ty == kt && !(akk > 0.0)) st->chol_fail += 1;
      const double rs = rsqrt(fmax(akk, 1e-300));
      if (ty == kt) invd[k] = rs;
#define MCBA_SCALE_Q(Q) case Q: if constexpr (Q < R) { _Pragma("unroll") for (int p = 0; p < R; p++) { const int i = ty + 16 * p; \
        if (i >= k && i < n) { const double l = a[p][Q < R ? Q : 0] * rs; a[p][Q < R ? Q : 0] = l; colbuf[i] = l; } } } break;
      switch (kq) { MCBA_SCALE_Q(0) MCBA_SCALE_Q(1) MCBA_SCALE_Q(2) MCBA_SCALE_Q(3) MCBA_SCALE_Q(4) MCBA_SCALE_Q(5) MCBA_SCALE_Q(6) MCBA_SCALE_Q(7) }
#undef MCBA_SCALE_Q
    }
    __syncthreads();
    double ci[R], cj[R];
#pragma unroll
    for (int p = 0; p < R; p++) { const int i = ty + 16 * p; ci[p] = (i > k && i < n) ? colbuf[i] : 0.0; }
#pragma unroll
    for (int q = 0; q < R; q++) { const int j = tx + 16 * q; cj[q] = (j > k && j < n) ? colbuf[j] : 0.0; }
#pragma unroll
    for (int p = 0; p < R; p++)
#pragma unroll
      for (int q = 0; q < R; q++) a[p][q] -= ci[p] * cj[q];       // (also touches the unused upper triangle: harmless)
    // publish the next pivot
    {
      const int k1 = k + 1, q1 = k1 >> 4, t1 = k1 & 15;
      if (k1 < n && ty == t1 && tx == t1) {
#pragma unroll
        for (int p = 0; p < R; p++) if (p == q1) piv[0] = a[p][p];
      }
    }
    __syncthreads();
  }
  // factor -> shared memory (lower triangle incl. diagonal)
#pragma unroll
  for (int p = 0; p < R; p++)
#pragma unroll
    for (int q = 0; q < R; q++) {
      const int i = ty + 16 * p, j = tx + 16 * q;
      if (i < n && j <= i) Lm[i * ld + j] = a[p][q];
    }
  __syncthreads();
  if (tid < 32) {
    constexpr int RS = (R * 16 + 31) / 32;       // rows per lane
    const int lane = tid;
    double bs[RS];
#pragma unroll
    for (int s2 = 0; s2 < RS; s2++) { const int i = lane + 32 * s2; bs[s2] = i < n ? rhs[i] + gh[i] : 0.0; }
    // forward: L y = b
#pragma unroll
    for (int s1 = 0; s1 < RS; s1++) {
      for (int kk = 0; kk < 32; kk++) {
        const int k = 32 * s1 + kk;
        if (k >= n) break;
        const double yk = __shfl_sync(0xffffffffu, bs[s1] * invd[k], kk);
        if (lane == kk) bs[s1] = yk;
#pragma unroll
        for (int s2 = s1; s2 < RS; s2++) { const int i = lane + 32 * s2; if (i > k && i < n) bs[s2] -= Lm[i * ld + k] * yk; }
      }
    }
    // backward: L^T x = y
#pragma unroll
    for (int s1 = RS - 1; s1 >= 0; s1--) {
      for (int kk = 31; kk >= 0; kk--) {
        const int k = 32 * s1 + kk;
        if (k >= n) continue;
        const double xk = __shfl_sync(0xffffffffu, bs[s1] * invd[k], kk);
        if (lane == kk) bs[s1] = xk;
#pragma unroll
        for (int s2 = 0; s2 <= s1; s2++) { const int i = lane + 32 * s2; if (i < k) bs[s2] -= Lm[k * ld + i] * xk; }
      }
    }
#pragma unroll
    for (int s2 = 0; s2 < RS; s2++) { const int i = lane + 32 * s2; if (i < n) out[i] = bs[s2]; }
  }
}

// k_chol_blocked (opt-in: MCBA_CHOL=blocked, n <= 128): the same reduced solve as k_chol_small with the per-column block
// barriers taken out of the critical path.  k_chol_small pays two __syncthreads and a shared-memory round trip per COLUMN
// (profiles/r01_ncu_small_kernels_cfg2.csv: 70 columns x 0.73 us); here a column step of the 16x16 diagonal block is warp-level
// (lane j holds column j in registers: pivot broadcast, rsqrt, scaled column broadcast by 16 shuffles, rank-1 update on
// registers -- no barrier), the diagonal block is inverted by the same warp, and the panel (thread per row, in place) and the
// trailing update (16x16 thread tiling) are plain products between 3 block barriers per 16 columns.  Both substitutions use the
// inverted diagonal blocks: per block one 16-long product and one row-parallel update.  Matrix (lower triangle) in shared memory.
constexpr int CB = 16;
__host__ __device__ inline size_t chol_blocked_smem_doubles(int n) {
  const int nblk = (n + CB - 1) / CB;
  return (size_t)n * (n | 1) + (size_t)nblk * CB * CB + (size_t)nblk * CB;
}
__global__ void __launch_bounds__(256)
k_chol_blocked(int n, const double* Sg, const double* rhs, const double* gh, SolverState* st, double* out) {
  extern __shared__ double cbs[];
  const int ld = n | 1;                          // odd leading dimension: rows walked by consecutive threads hit different banks
  const int nblk = (n + CB - 1) / CB;
  double* A = cbs;                                // [n][ld], lower triangle
  double* Li = A + (size_t)n * ld;                // [nblk][CB][CB] inverses of the diagonal blocks (row-major, lower triangular)
  double* b = Li + (size_t)nblk * CB * CB;        // [nblk*CB] right-hand side -> y -> x
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double reg = st->reg;
  for (int idx = tid; idx < n * n; idx += 256) {
    const int i = idx / n, j = idx % n;
    if (j <= i) A[i * ld + j] = Sg[(size_t)i * n + j] + (i == j ? reg : 0.0);
  }
  for (int i = tid; i < nblk * CB; i += 256) b[i] = i < n ? rhs[i] + gh[i] : 0.0;
  __syncthreads();

  for (int kb = 0, blk = 0; kb < n; kb += CB, blk++) {
    const int nb = min(CB, n - kb);
    double* Lb = Li + (size_t)blk * CB * CB;
    if (warp == 0) {
      // ---- diagonal block: lanes j and j+16 both hold column j (full-mask shuffles stay uniform); identity padding beyond nb
      const int j = lane & 15;
      double col[CB], rsd[CB];
#pragma unroll
      for (int i = 0; i < CB; i++)
        col[i] = (i < nb && j < nb) ? (i >= j ? A[(kb + i) * ld + kb + j] : A[(kb + j) * ld + kb + i]) : (i == j ? 1.0 : 0.0);
#pragma unroll
      for (int k = 0; k < CB; k++) {
        const double dkk = __shfl_sync(0xffffffffu, col[k], k);
        if (lane == 0 && k < nb && !(dkk > 0.0)) st->chol_fail += 1;
        const double rs = rsqrt(fmax(dkk, 1e-300));
        rsd[k] = rs;
        double lik[CB];
#pragma unroll
        for (int i = k; i < CB; i++) {
          double v = col[i];
          if (j == k) { v *= rs; col[i] = v; }
          lik[i] = __shfl_sync(0xffffffffu, v, k);
        }
        double ljk = 0.0;
#pragma unroll
        for (int i = k + 1; i < CB; i++) if (i == j) ljk = lik[i];
        if (j > k) {
#pragma unroll
          for (int i = k + 1; i < CB; i++) col[i] -= lik[i] * ljk;
        }
      }
      if (lane < CB && j < nb) {
#pragma unroll
        for (int i = 0; i < CB; i++) if (i >= j && i < nb) A[(kb + i) * ld + kb + j] = col[i];
      }
      __syncwarp();
      // ---- inverse of the diagonal block: lane j solves L z = e_j (1 / L_ii = rsd[i] from the factorisation)
      if (lane < CB) {
        double z[CB];
#pragma unroll
        for (int i = 0; i < CB; i++) {
          double sacc = (i == j) ? 1.0 : 0.0;
#pragma unroll
          for (int m = 0; m < i; m++) {
            const double lim = (i < nb) ? A[(kb + i) * ld + kb + m] : 0.0;     // m < i < nb; padding rows are identity
            sacc -= lim * z[m];
          }
          z[i] = (i >= j) ? sacc * rsd[i] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < CB; i++) Lb[i * CB + j] = z[i];
      }
    }
    __syncthreads();
    // ---- panel: rows below the block, thread per row, in place:  P = A_panel L_kk^-T   (P[i][jj] = sum_{m<=jj} A[i][m] Linv[jj][m])
    for (int i = kb + nb + tid; i < n; i += 256) {
      double a[CB];
#pragma unroll
      for (int m = 0; m < CB; m++) a[m] = m < nb ? A[i * ld + kb + m] : 0.0;
#pragma unroll
      for (int jj = 0; jj < CB; jj++) {
        if (jj < nb) {
          double pacc = 0.0;
#pragma unroll
          for (int m = 0; m <= jj; m++) pacc += a[m] * Lb[jj * CB + m];
          A[i * ld + kb + jj] = pacc;
        }
      }
    }
    __syncthreads();
    // ---- trailing update (lower triangle): A[i][jj] -= P[i][:] . P[jj][:]
    {
      const int base = kb + nb, ty = tid & 15, tx = tid >> 4;
      for (int i = base + ty; i < n; i += 16) {
        double pi[CB];
#pragma unroll
        for (int m = 0; m < CB; m++) pi[m] = A[i * ld + kb + m];
        for (int jj = base + tx; jj <= i; jj += 16) {
          double acc = 0.0;
#pragma unroll
          for (int m = 0; m < CB; m++) acc += pi[m] * A[jj * ld + kb + m];
          A[i * ld + jj] -= acc;
        }
      }
    }
    __syncthreads();
  }
  // ---- forward substitution  L y = b, block by block
  for (int kb = 0, blk = 0; kb < n; kb += CB, blk++) {
    const int nb = min(CB, n - kb);
    const double* Lb = Li + (size_t)blk * CB * CB;
    double yv = 0.0;
    if (tid < CB) {
#pragma unroll
      for (int m = 0; m < CB; m++) if (m <= tid) yv += Lb[tid * CB + m] * b[kb + m];
    }
    __syncthreads();
    if (tid < CB) b[kb + tid] = yv;
    __syncthreads();
    for (int i = kb + nb + tid; i < n; i += 256) {
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < CB; m++) acc += A[i * ld + kb + m] * b[kb + m];       // columns beyond nb only exist in the last block
      b[i] -= acc;
    }
    __syncthreads();
  }
  // ---- backward substitution  L^T x = y
  for (int blk = nblk - 1; blk >= 0; blk--) {
    const int kb = blk * CB, nb = min(CB, n - kb);
    const double* Lb = Li + (size_t)blk * CB * CB;
    double xv = 0.0;
    if (tid < CB) {
#pragma unroll
      for (int m = 0; m < CB; m++) if (m >= tid) xv += Lb[m * CB + tid] * b[kb + m];
    }
    __syncthreads();
    if (tid < CB) b[kb + tid] = xv;
    __syncthreads();
    for (int i = tid; i < kb; i += 256) {
      double acc = 0.0;
      for (int m = 0; m < nb; m++) acc += A[(kb + m) * ld + i] * b[kb + m];
      b[i] -= acc;
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += 256) out[i] = b[i];
}

constexpr int CHOL_NB = 32;
// add reg to the diagonal (once, before the blocked factorisation)
__global__ void k_chol_addreg(int n, double* S, const SolverState* st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) S[(size_t)i * n + i] += st->reg;
}
// Factor the 32x32 diagonal block at kb in registers (16x16 threads, 2x2 entries each, same scheme as k_chol_small),
// write L_kk back (lower) and its inverse to Linv[kb/32] (row-major 32x32, identity-padded for a short last block):
// the panel solve and both substitutions then become plain matrix products.
__global__ void __launch_bounds__(256)
k_chol_diag(int n, int kb, double* S, double* Linv_all, SolverState* st) {
  __shared__ double Lm[CHOL_NB][CHOL_NB + 1];
  __shared__ double colbuf[CHOL_NB], invd[CHOL_NB], piv[2];
  const int nb = min(CHOL_NB, n - kb);
  const int tid = threadIdx.x, ty = tid & 15, tx = tid >> 4;
  double a[2][2];
#pragma unroll
  for (int p = 0; p < 2; p++)
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int i = ty + 16 * p, j = tx + 16 * q;
      a[p][q] = (i < nb && j < nb) ? (j <= i ? S[(size_t)(kb + i) * n + kb + j] : S[(size_t)(kb + j) * n + kb + i]) : (i == j ? 1.0 : 0.0);
    }
  if (tid == 0) piv[0] = a[0][0];
  __syncthreads();
  for (int k = 0; k < CHOL_NB; k++) {
    const int kq = k >> 4, kt = k & 15;
    if (tx == kt) {
      const double akk = piv[0];
      if (ty == kt && k < nb && !(akk > 0.0)) st->chol_fail += 1;
      const double rs = rsqrt(fmax(akk, 1e-300));
      if (ty == kt) invd[k] = rs;
#pragma unroll
      for (int p = 0; p < 2; p++) {
        const int i = ty + 16 * p;
        if (i >= k) {
          if (kq == 0) { const double l = a[p][0] * rs; a[p][0] = l; colbuf[i] = l; }
          else { const double l = a[p][1] * rs; a[p][1] = l; colbuf[i] = l; }
        }
      }
    }
    __syncthreads();
    double ci[2], cj[2];
#pragma unroll
    for (int p = 0; p < 2; p++) { const int i = ty + 16 * p; ci[p] = i > k ? colbuf[i] : 0.0; }
#pragma unroll
    for (int q = 0; q < 2; q++) { const int j = tx + 16 * q; cj[q] = j > k ? colbuf[j] : 0.0; }
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
      for (int q = 0; q < 2; q++) a[p][q] -= ci[p] * cj[q];
    {
      const int k1 = k + 1, q1 = k1 >> 4, t1 = k1 & 15;
      if (k1 < CHOL_NB && ty == t1 && tx == t1) piv[0] = q1 == 0 ? a[0][0] : a[1][1];
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < 2; p++)
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int i = ty + 16 * p, j = tx + 16 * q;
      Lm[i][j] = j <= i ? a[p][q] : 0.0;
      if (i < nb && j <= i) S[(size_t)(kb + i) * n + kb + j] = a[p][q];
    }
  __syncthreads();
  // inverse: thread j < 32 solves L z = e_j
  if (tid < CHOL_NB) {
    const int j = tid;
    double z[CHOL_NB];
#pragma unroll
    for (int i = 0; i < CHOL_NB; i++) {
      double t = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int k2 = 0; k2 < CHOL_NB; k2++) if (k2 < i) t -= Lm[i][k2] * z[k2];
      z[i] = (i >= j) ? t * invd[i] : 0.0;
    }
    double* Li = Linv_all + (size_t)(kb / CHOL_NB) * CHOL_NB * CHOL_NB;
#pragma unroll
    for (int i = 0; i < CHOL_NB; i++) Li[i * CHOL_NB + j] = z[i];
  }
}
// panel below the diagonal block: X = A[:, kb:kb+32] L_kk^-T as a product with the block inverse; 32 rows per CTA.
// Also mirrors the panel into the upper triangle (S[kb+j][i] = X[i][j]) so that L^T is readable row-wise.
__global__ void __launch_bounds__(256)
k_chol_trsm(int n, int kb, double* S, const double* Linv_all) {
  __shared__ double Li[CHOL_NB][CHOL_NB + 1];
  __shared__ double At[CHOL_NB][CHOL_NB + 1];
  const int nb = min(CHOL_NB, n - kb);
  const double* Lg = Linv_all + (size_t)(kb / CHOL_NB) * CHOL_NB * CHOL_NB;
  const int i0 = kb + nb + blockIdx.x * CHOL_NB;
  for (int o = threadIdx.x; o < CHOL_NB * CHOL_NB; o += 256) {
    const int r = o / CHOL_NB, c = o % CHOL_NB;
    Li[r][c] = Lg[o];
    At[r][c] = (i0 + r < n && c < nb) ? S[(size_t)(i0 + r) * n + kb + c] : 0.0;
  }
  __syncthreads();
  const int r = threadIdx.x >> 3, cg = threadIdx.x & 7;
  double x[4] = {0, 0, 0, 0};
#pragma unroll 8
  for (int k = 0; k < CHOL_NB; k++) {
    const double av = At[r][k];
#pragma unroll
    for (int q = 0; q < 4; q++) x[q] += av * Li[cg * 4 + q][k];       // X[i][j] = sum_k A[i][k] Linv[j][k]
  }
  if (i0 + r < n) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int j = cg * 4 + q;
      if (j < nb) { S[(size_t)(i0 + r) * n + kb + j] = x[q]; S[(size_t)(kb + j) * n + i0 + r] = x[q]; }
    }
  }
}
// trailing update (lower triangle): A[i][j] -= sum_k L[i][kb+k] L[j][kb+k], 32x32 tiles, 2x2 per thread
__global__ void __launch_bounds__(256)
k_chol_syrk(int n, int kb, double* S) {
  __shared__ double Li[CHOL_NB][CHOL_NB + 1];
  __shared__ double Lj[CHOL_NB][CHOL_NB + 1];
  const int nb = min(CHOL_NB, n - kb);
  const int base = kb + nb;
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti) return;
  const int i0 = base + ti * 32, j0 = base + tj * 32;
  if (i0 >= n) return;
  for (int o = threadIdx.x; o < 32 * nb; o += 256) {
    const int r = o / nb, k = o % nb;
    Li[r][k] = (i0 + r < n) ? S[(size_t)(i0 + r) * n + kb + k] : 0.0;
    Lj[r][k] = (j0 + r < n) ? S[(size_t)(j0 + r) * n + kb + k] : 0.0;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[2][2] = {{0, 0}, {0, 0}};
  for (int k = 0; k < nb; k++) {
    const double a0 = Li[ty][k], a1 = Li[ty + 16][k], b0 = Lj[tx][k], b1 = Lj[tx + 16][k];
    acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
  }
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
      if (i < n && j < n && j <= i) S[(size_t)i * n + j] -= acc[a][b];
    }
}
// Both substitutions with the factor in global memory (lower = L, strict upper = L^T mirror) and the inverted diagonal
// blocks: per 32-block a 32x32 product by warp 0, then every thread updates one remaining row with 32 FMAs.
__global__ void __launch_bounds__(512)
k_chol_substitute(int n, const double* L, const double* Linv_all, const double* rhs, const double* gh, double* out) {
  extern __shared__ double bsh[];                    // n (+32 pad)
  __shared__ double yb[CHOL_NB];
  const int tid = threadIdx.x;
  const int nblk = (n + CHOL_NB - 1) / CHOL_NB;
  for (int i = tid; i < nblk * CHOL_NB; i += 512) bsh[i] = i < n ? rhs[i] + gh[i] : 0.0;
  __syncthreads();
  for (int blk = 0; blk < nblk; blk++) {              // forward: L y = b
    const int kb = blk * CHOL_NB;
    const double* Li = Linv_all + (size_t)blk * CHOL_NB * CHOL_NB;
    if (tid < CHOL_NB) {
      double acc = 0.0;
#pragma unroll 8
      for (int k = 0; k < CHOL_NB; k++) acc += Li[tid * CHOL_NB + k] * bsh[kb + k];
      yb[tid] = acc;
    }
    __syncthreads();
    if (tid < CHOL_NB) bsh[kb + tid] = yb[tid];
    for (int i = kb + CHOL_NB + tid; i < n; i += 512) {
      const double* row = L + (size_t)i * n + kb;
      double acc = 0.0;
#pragma unroll 8
      for (int k = 0; k < CHOL_NB; k++) acc += row[k] * yb[k];
      bsh[i] -= acc;
    }
    __syncthreads();
  }
  for (int blk = nblk - 1; blk >= 0; blk--) {         // backward: L^T x = y
    const int kb = blk * CHOL_NB;
    const double* Li = Linv_all + (size_t)blk * CHOL_NB * CHOL_NB;
    if (tid < CHOL_NB) {
      double acc = 0.0;
#pragma unroll 8
      for (int k = 0; k < CHOL_NB; k++) acc += Li[k * CHOL_NB + tid] * bsh[kb + k];      // (L_kk^-1)^T
      yb[tid] = acc;
    }
    __syncthreads();
    if (tid < CHOL_NB) bsh[kb + tid] = yb[tid];
    const int nbv = min(CHOL_NB, n - kb);
    for (int i = tid; i < kb; i += 512) {
      const double* row = L + (size_t)i * n + kb;      // mirrored panel: row[k] = L[kb+k][i]
      double acc = 0.0;
      for (int k = 0; k < nbv; k++) acc += row[k] * yb[k];
      bsh[i] -= acc;
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += 512) out[i] = bsh[i];
}

// back-substitution of the eliminated frame blocks: gn_f = L^-T (z_f - Y_f^T gn_s)
template <int FB>
__global__ void __launch_bounds__(SCHUR_THREADS)
k_backsub(int n_s, const double* Y, const double* Lf, const double* zf, double* gn) {
  __shared__ double sm[32];
  __shared__ double t[FB];
  const int f = blockIdx.x, tid = threadIdx.x;
  const double* Yf = Y + (size_t)f * n_s * FB;
  double acc[FB];
#pragma unroll
  for (int k = 0; k < FB; k++) acc[k] = 0.0;
  for (int s = tid; s < n_s; s += SCHUR_THREADS) {
    const double ps = gn[s];
#pragma unroll
    for (int k = 0; k < FB; k++) acc[k] += Yf[s * FB + k] * ps;
  }
#pragma unroll
  for (int k = 0; k < FB; k++) { const double r = block_sum(acc[k], sm); if (tid == 0) t[k] = r; }
  if (tid == 0) {
    const double* L = Lf + (size_t)f * FB * FB;
    double y[FB];
    for (int i = 0; i < FB; i++) y[i] = zf[(size_t)f * FB + i] - t[i];
    for (int i = FB - 1; i >= 0; i--) {
      double v = y[i];
      for (int k = i + 1; k < FB; k++) v -= L[k * FB + i] * y[k];
      y[i] = v / L[i * FB + i];
    }
    for (int i = 0; i < FB; i++) gn[n_s + FB * f + i] = y[i];
  }
}

// gh.gn and ||gn||^2, split shared / frame.  Single CTA.
__global__ void k_dots(int n, int n_s, const double* gh, const double* gn, double* red) {
  __shared__ double sm[32];
  double ds = 0, df = 0, ns = 0, nf = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double a = gh[i], b = gn[i];
    if (i < n_s) { ds += a * b; ns += b * b; } else { df += a * b; nf += b * b; }
  }
  double r;
  r = block_sum(ds, sm); if (threadIdx.x == 0) red[RED_DOTGN_S] = r;
  r = block_sum(df, sm); if (threadIdx.x == 0) red[RED_DOTGN_F] = r;
  r = block_sum(ns, sm); if (threadIdx.x == 0) red[RED_GN2_S] = r;
  r = block_sum(nf, sm); if (threadIdx.x == 0) red[RED_GN2_F] = r;
}

// trf.py: S = qr([g_h, gn_h]); B_S = (J_h S)^T (J_h S); g_S = S^T g_h   -- expressed through Gram-Schmidt
// coefficients so that no basis vectors are materialised: q1 = gh/n1, q2 = (gn - mu q1)/n2.
__device__ inline void subspace_compute(SolverState* st, const double* red) {
  const double n1 = st->gh_norm;
  const double dot = red[RED_DOTGN_S] + red[RED_DOTGN_F];
  const double gn2 = red[RED_GN2_S] + red[RED_GN2_F];
  const double mu = dot / n1;
  double n2sq = gn2 - mu * mu;
  const double agg = red[RED_AGG], agn = red[RED_AGN], ann = red[RED_ANN];
  st->n1 = n1; st->mu = mu;
  st->B11 = agg / (n1 * n1);
  st->gS1 = n1; st->gS2 = 0.0;
  if (!(n2sq > 1e-28 * gn2) || !(n2sq > 0.0)) {     // gn parallel to gh: 1-D subspace
    st->n2 = 0.0; st->B12 = 0.0; st->B22 = 1.0;
  } else {
    const double n2 = sqrt(n2sq), c = mu / n1;
    st->n2 = n2;
    st->B12 = (agn - c * agg) / (n1 * n2);
    st->B22 = (ann - 2.0 * c * agn + c * c * agg) / (n2 * n2);
  }
}

__global__ void k_subspace(SolverState* st, const double* red) { subspace_compute(st, red); }

// common.py solve_trust_region_2d: minimise 0.5 p^T B p + g^T p, ||p|| <= Delta  (B 2x2 symmetric).
// Interior Newton point if B is positive definite and inside; otherwise the global boundary minimiser via
// the secular equation in the eigenbasis of B (equivalent to scipy's argmin over the quartic's real roots).
__host__ __device__ inline void solve_tr_2d(double b11, double b12, double b22, double g1, double g2, double Delta, double& p1, double& p2) {
  const double det = b11 * b22 - b12 * b12;
  if (b11 > 0.0 && det > 0.0) {
    const double q1 = -(b22 * g1 - b12 * g2) / det, q2 = -(b11 * g2 - b12 * g1) / det;
    if (q1 * q1 + q2 * q2 <= Delta * Delta) { p1 = q1; p2 = q2; return; }
  }
  // eigen-decomposition
  const double tr = b11 + b22, df = b11 - b22;
  const double rad = sqrt(df * df + 4.0 * b12 * b12);
  const double l1 = 0.5 * (tr - rad), l2 = 0.5 * (tr + rad);       // l1 <= l2
  double v1x, v1y;
  if (fabs(b12) > 1e-300 * fmax(fabs(tr), 1.0)) { v1x = l1 - b22; v1y = b12; const double nn = hypot(v1x, v1y); if (nn > 0) { v1x /= nn; v1y /= nn; } else { v1x = 1; v1y = 0; } }
  else if (b11 <= b22) { v1x = 1; v1y = 0; } else { v1x = 0; v1y = 1; }
  const double v2x = -v1y, v2y = v1x;
  const double h1 = v1x * g1 + v1y * g2, h2 = v2x * g1 + v2y * g2;
  // find sigma >= max(0,-l1) with h1^2/(l1+s)^2 + h2^2/(l2+s)^2 = Delta^2
  const double gnorm = hypot(h1, h2);
  double lo = fmax(0.0, -l1);
  double hi = fmax(lo, gnorm / Delta - l1) + 1e-300;
  auto pn2 = [&](double s) { const double a = h1 / (l1 + s), b = h2 / (l2 + s); return a * a + b * b; };
  double c1, c2;
  // hard case: h1 ~ 0 and the l2-component alone stays inside at s = -l1
  const bool hard = (fabs(h1) <= 1e-14 * gnorm) && (l2 + lo > 0.0) && (h2 * h2 / ((l2 + lo) * (l2 + lo)) <= Delta * Delta);
  if (hard || gnorm == 0.0) {
    c2 = (l2 + lo > 0.0) ? -h2 / (l2 + lo) : 0.0;
    const double rem = Delta * Delta - c2 * c2;
    c1 = sqrt(fmax(rem, 0.0));
  } else {
    while (pn2(hi) > Delta * Delta) hi = 2.0 * hi + 1e-12;
    double s = hi;
    for (int it = 0; it < 200; it++) {
      s = 0.5 * (lo + hi);
      if (pn2(s) > Delta * Delta) lo = s; else hi = s;
      if (hi - lo <= 1e-16 * fmax(hi, 1e-300)) break;
    }
    s = 0.5 * (lo + hi);
    c1 = -h1 / (l1 + s); c2 = -h2 / (l2 + s);
    const double nn = hypot(c1, c2);
    if (nn > 0.0) { c1 *= Delta / nn; c2 *= Delta / nn; }
  }
  p1 = c1 * v1x + c2 * v2x;
  p2 = c1 * v1y + c2 * v2y;
}

__device__ inline void tr_step_compute(SolverState* st) {
  double p1, p2;
  solve_tr_2d(st->B11, st->B12, st->B22, st->gS1, st->gS2, st->Delta, p1, p2);
  if (st->n2 == 0.0) p2 = 0.0;
  st->step_h_norm = sqrt(p1 * p1 + p2 * p2);
  st->predicted = -(0.5 * (st->B11 * p1 * p1 + 2.0 * st->B12 * p1 * p2 + st->B22 * p2 * p2) + st->gS1 * p1 + st->gS2 * p2);
  // step_h = p1 q1 + p2 q2 = alpha gh + beta gn
  if (st->n2 == 0.0) { st->alpha = p1 / st->n1; st->beta = 0.0; }
  else { st->beta = p2 / st->n2; st->alpha = p1 / st->n1 - st->beta * st->mu / st->n1; }
}

// x_new = x + d*(alpha gh + beta gn); norms of step and x (split shared / frame). Single CTA.
__global__ void k_step(int n, int n_s, SolverState* st, const double* x, const double* d, const double* gh,
                       const double* gn, double* x_new, double* red) {
  __shared__ double sm[32];
  if (st->done) return;
  if (threadIdx.x == 0) tr_step_compute(st);      // 2-D trust-region subproblem for the current Delta
  __syncthreads();
  const double al = st->alpha, be = st->beta;
  double s2s = 0, s2f = 0, x2s = 0, x2f = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double stp = d[i] * (al * gh[i] + be * gn[i]);
    const double xi = x[i];
    x_new[i] = xi + stp;
    if (i < n_s) { s2s += stp * stp; x2s += xi * xi; } else { s2f += stp * stp; x2f += xi * xi; }
  }
  double r;
  r = block_sum(s2s, sm); if (threadIdx.x == 0) red[RED_STEP2_S] = r;
  r = block_sum(s2f, sm); if (threadIdx.x == 0) red[RED_STEP2_F] = r;
  r = block_sum(x2s, sm); if (threadIdx.x == 0) red[RED_XN2_S] = r;
  r = block_sum(x2f, sm); if (threadIdx.x == 0) red[RED_XN2_F] = r;
}

// k_step and k_make_trial as ONE single-CTA launch (opt-in MCBA_FUSE=1): the trial state only needs x_new, which this CTA has
// just written; a few hundred poses are nothing for 1024 threads, and the launch in between disappears.
__global__ void __launch_bounds__(1024)
k_step_trial(int n, int n_s, SolverState* st, const double* x, const double* d, const double* gh, const double* gn, double* x_new, double* red,
             DeviceProblem p, double* cam_o, double* board_o, double* frame_o, double* intr_o, double* bpts_o, double* he_o, int n_items) {
  __shared__ double sm[32];
  if (st->done) return;
  if (threadIdx.x == 0) tr_step_compute(st);
  __syncthreads();
  const double al = st->alpha, be = st->beta;
  double s2s = 0, s2f = 0, x2s = 0, x2f = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double stp = d[i] * (al * gh[i] + be * gn[i]);
    const double xi = x[i];
    x_new[i] = xi + stp;
    if (i < n_s) { s2s += stp * stp; x2s += xi * xi; } else { s2f += stp * stp; x2f += xi * xi; }
  }
  double r;
  r = block_sum(s2s, sm); if (threadIdx.x == 0) red[RED_STEP2_S] = r;
  r = block_sum(s2f, sm); if (threadIdx.x == 0) red[RED_STEP2_F] = r;
  r = block_sum(x2s, sm); if (threadIdx.x == 0) red[RED_XN2_S] = r;
  r = block_sum(x2f, sm); if (threadIdx.x == 0) red[RED_XN2_F] = r;
  __syncthreads();                      // x_new complete (block-wide visibility of this CTA's global writes)
  for (int i = threadIdx.x; i < n_items; i += blockDim.x) make_trial_item(p, x_new, cam_o, board_o, frame_o, intr_o, bpts_o, he_o, i);
}

// trf.py inner loop after fun(x_new): actual reduction, update_tr_radius, check_termination.
// sum of the per-view cost entries of the moment records (moments[v][T-1]) -> red[RED_COSTNEW]
__global__ void k_cost_from_moments(const double* moments, int V, int T, double* red, PeerArgs pa) {
  __shared__ double sm[32];
  double c = 0.0;
  for (int v = threadIdx.x; v < V; v += blockDim.x) c += moments[(size_t)v * T];       // T == 1: compact per-view costs
  c = block_sum(c, sm);
  if (threadIdx.x == 0) red[RED_COSTNEW] = c;
  if (pa.world > 1) { __syncthreads(); peer_allreduce_block(pa); }       // MCBA_FUSE=1: trial cost and step norms of all ranks + the acceptance test
}

// trf.py inner loop after fun(x_new): actual reduction, update_tr_radius, check_termination (one thread)
__device__ inline void accept_compute(SolverState* st, const double* red) {
  st->nfev += 1;
  const double cost_new = red[RED_COSTNEW];
  st->cost_new = cost_new;
  const double shn = st->step_h_norm;
  if (!isfinite(cost_new)) {            // trf.py: non-finite f_new -> shrink and retry
    st->Delta = 0.25 * shn;
    st->actual_reduction = -1.0;
    st->accepted = 0;
    return;
  }
  const double actual = st->cost - cost_new;
  const double pred = st->predicted;
  double ratio;
  if (pred > 0.0) ratio = actual / pred; else if (pred == 0.0 && actual == 0.0) ratio = 1.0; else ratio = 0.0;
  double Dn = st->Delta;
  if (ratio < 0.25) Dn = 0.25 * shn;
  else if (ratio > 0.75 && shn > 0.95 * st->Delta) Dn = 2.0 * st->Delta;
  const double step_norm = sqrt(red[RED_STEP2_S] + red[RED_STEP2_F]);
  const double x_norm = sqrt(red[RED_XN2_S] + red[RED_XN2_F]);
  st->step_norm = step_norm; st->x_norm = x_norm; st->actual_reduction = actual; st->ratio = ratio;
  const bool ft = (actual < st->ftol * st->cost) && (ratio > 0.25);
  const bool xt = step_norm < st->xtol * (st->xtol + x_norm);
  int status = -99;
  if (ft && xt) status = 4; else if (ft) status = 2; else if (xt) status = 3;
  st->status = status;
  if (status == -99) st->Delta = Dn;
  st->accepted = actual > 0.0;
}

// tail of the moment kernels (MCBA_FUSE=1, single GPU): the last CTA to finish sums the per-view costs and runs the acceptance test
__device__ __noinline__ void view_accept_epilogue(SolverState* st, double* red, unsigned* counter, const double* view_cost, int V) {
  __shared__ double acc_sm[32];
  __shared__ int acc_is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) { const unsigned t = atomicAdd(counter, 1u); acc_is_last = (t == gridDim.x - 1); }
  __syncthreads();
  if (!acc_is_last) return;
  if (threadIdx.x == 0) *counter = 0;
  if (st->done) return;
  double c = 0.0;
  for (int v = threadIdx.x; v < V; v += blockDim.x) c += __ldcg(&view_cost[v]);
  c = block_sum(c, acc_sm);
  if (threadIdx.x == 0) { red[RED_COSTNEW] = c; accept_compute(st, red); }
}

// single-GPU: the cost sum is done by the same CTA (moments != nullptr); multi-GPU: the sum, the exchange and the test are
// separate (k_cost_from_moments, then the exchange kernel runs accept_compute as its epilogue)
__global__ void k_accept(SolverState* st, double* red, const double* moments, int V, int T) {
  __shared__ double sm[32];
  if (st->done) return;
  if (moments) {
    double c = 0.0;
    for (int v = threadIdx.x; v < V; v += blockDim.x) c += moments[(size_t)v * T];     // T == 1: compact per-view costs
    c = block_sum(c, sm);
    if (threadIdx.x == 0) red[RED_COSTNEW] = c;
  }
  if (threadIdx.x != 0) return;
  accept_compute(st, red);
}

// scalar step that follows an exchange (run inside the exchange kernel when it goes over peer memory, as a 1-thread kernel after NCCL)
enum { EPI_NONE = 0, EPI_BEGIN = 1, EPI_REG = 2, EPI_SUBSPACE = 3, EPI_ACCEPT = 4 };
__device__ inline void run_epilogue(int epi, SolverState* st, double* red) {
  switch (epi) {
    case EPI_BEGIN: begin_iteration(st, red); break;
    case EPI_REG: reg_compute(st, red); break;
    case EPI_SUBSPACE: subspace_compute(st, red); break;
    case EPI_ACCEPT: if (!st->done) accept_compute(st, red); break;
    default: break;
  }
}
__global__ void k_epilogue(int epi, SolverState* st, double* red) { run_epilogue(epi, st, red); }

__global__ void k_axpby_copy(int n, const double* src, double* dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

}  // namespace mcba
