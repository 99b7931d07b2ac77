// linearize.cuh — the fused linearisation pass: residuals, analytic Jacobian rows, per-view moments on the fp64 tensor path AND their
// expansion into the block-arrow normal equations, in ONE kernel that reads every corner once and writes nothing per corner or per view.
//
//   k_linearize<MODEL, ROLL>   one CTA per frame (persistent, static frame -> CTA map), warps own the frame's views by camera:
//                              per view   M = sum [G r]^T [G r]  (mma.sync.m8n8k4.f64, view_chunk below)
//                              epilogue   the view's 6x6 twist maps (camera / frame / board pose -> camera-frame twist) applied to M:
//                                         H_ff, g_f, W_f (this frame's block row of the arrow) are COMPLETE when the CTA leaves the frame and
//                                         are stored once; the shared-parameter blocks are accumulated per (CTA, camera) in a record that
//                                         only this CTA's owning warp ever touches (plain read-modify-write in L2, no atomics).
//   k_reduce_shared<NP>        sums the per-CTA records in CTA order (fixed -> bit-reproducible run to run) and stores H_ss, g_s, cost.
//
// Replaces k_views_mma + k_expand_frames + k_expand_shared of round 1 (1 KB/view moment records written and read back twice:
// 1.45x the algorithmic bytes at 5.5 M corners, and fp64 atomics into H_ss / g that made two identical solves differ in the 8th digit).
// Reference semantics: the Jacobian scipy builds by finite differences over calibration.py:173-196's sparsity pattern, here analytic.
#pragma once
#include "solver_kernels.cuh"

namespace mcba {

constexpr int LIN_WARPS = 8;
constexpr int LIN_THREADS = LIN_WARPS * 32;

// compile-time shape of a residual row's local Jacobian [twist block(s) | fx fy cx cy dist | r]
template <int MODEL, bool ROLL>
struct LinShape {
  static constexpr int ND = model_nd(MODEL);
  static constexpr int NP = ROLL ? 2 : 1;
  static constexpr int KO = 6 * NP;                 // twist columns
  static constexpr int NIN = 4 + ND;                // intrinsic columns
  static constexpr int D = KO + NIN;
  static constexpr int E = D * (D + 1) / 2;
  static constexpr int T = E + D + 1;               // moment record: upper triangle | G^T r | cost
  static constexpr int NC = mma_nc(MODEL, ROLL);    // staged columns (D + 1 padded to 8)
  static constexpr int NT = NC / 8;
  static constexpr int NPAIR = NT * (NT + 1) / 2;
  static constexpr int KINT = 5 + ND;
  static constexpr int FB = KO;                     // parameters of one eliminated frame block
  static constexpr int UB = D * 6 + 42;             // per (camera, board): M[:,xi] Ab (D x 6) | Ab^T M Ab (36) | Ab^T g (6)
};

// One 32-corner chunk of a view: every lane computes its corner's residual pair and 2 x D Jacobian rows (fp64), the warp stages
// [G | r] (64 rows x NC columns) in shared memory and accumulates M += [G r]^T [G r] into the DMMA C fragments `acc`.
template <int MODEL, bool ROLL>
__device__ __forceinline__ void view_chunk(const DeviceProblem& p, int loss, double f_scale, const ViewPose& vp, const ViewPose& vpe, const double* k,
                                           const double* bp, double inv_h, int base, int end, int lane, double* stage,
                                           double (&acc)[LinShape<MODEL, ROLL>::NPAIR][2], double& cost_acc) {
  using S = LinShape<MODEL, ROLL>;
  constexpr int ND = S::ND, KO = S::KO, D = S::D, NC = S::NC, NT = S::NT;
  const int grp = lane >> 2, tig = lane & 3;
  const int idx = base + lane;
  double gu[NC], gv[NC];
#pragma unroll
  for (int i = 0; i < NC; i++) { gu[i] = 0.0; gv[i] = 0.0; }
  if (idx < end) {
    const double2 ob = p.obs[idx];
    const int pi = p.pid[idx];
    const double X[3] = {bp[3 * pi], bp[3 * pi + 1], bp[3 * pi + 2]};
    const double tau = ob.y * inv_h;
    double Xc[3], Xs[3], Xe[3];
    corner_point<ROLL>(vp, vpe, X, tau, Xc, Xs, Xe);
    double u, w_;
    double Ju[3], Jv[3], ku[4 + ND], kv[4 + ND];
    project<MODEL, true>(Xc, k, u, w_, Ju, Jv, ku, kv);
    double ru = u - ob.x, rv = w_ - ob.y;              // projected - observed (calibration.py:206)
    double wu = 1.0, wv = 1.0;
    if (loss == 0) {
      cost_acc += 0.5 * (ru * ru + rv * rv);
    } else {
      // robust loss per scalar residual (least_squares.py construct_loss_function, common.py:720-731)
      const double is = 1.0 / f_scale, fs2 = f_scale * f_scale;
      double zu = ru * is, zv = rv * is;
      zu *= zu; zv *= zv;
      double r0u, r1u, r2u, r0v, r1v, r2v;
      loss_rho(loss, zu, r0u, r1u, r2u);
      loss_rho(loss, zv, r0v, r1v, r2v);
      cost_acc += 0.5 * fs2 * (r0u + r0v);
      double ju = r1u + 2.0 * r2u * zu, jv = r1v + 2.0 * r2v * zv;
      ju = fmax(fmax(ju, TRIGGS_FLOOR * r1u), SCIPY_EPS);
      jv = fmax(fmax(jv, TRIGGS_FLOOR * r1v), SCIPY_EPS);
      wu = sqrt(ju); wv = sqrt(jv);
      ru *= r1u / wu; rv *= r1v / wv;
    }
    if constexpr (!ROLL) {
      gu[0] = (Xc[1] * Ju[2] - Xc[2] * Ju[1]) * wu; gu[1] = (Xc[2] * Ju[0] - Xc[0] * Ju[2]) * wu; gu[2] = (Xc[0] * Ju[1] - Xc[1] * Ju[0]) * wu;
      gv[0] = (Xc[1] * Jv[2] - Xc[2] * Jv[1]) * wv; gv[1] = (Xc[2] * Jv[0] - Xc[0] * Jv[2]) * wv; gv[2] = (Xc[0] * Jv[1] - Xc[1] * Jv[0]) * wv;
#pragma unroll
      for (int i = 0; i < 3; i++) { gu[3 + i] = Ju[i] * wu; gv[3 + i] = Jv[i] * wv; }
    } else {
      const double su = (1.0 - tau) * wu, sv = (1.0 - tau) * wv, eu = tau * wu, ev = tau * wv;
      gu[0] = (Xs[1] * Ju[2] - Xs[2] * Ju[1]) * su; gu[1] = (Xs[2] * Ju[0] - Xs[0] * Ju[2]) * su; gu[2] = (Xs[0] * Ju[1] - Xs[1] * Ju[0]) * su;
      gv[0] = (Xs[1] * Jv[2] - Xs[2] * Jv[1]) * sv; gv[1] = (Xs[2] * Jv[0] - Xs[0] * Jv[2]) * sv; gv[2] = (Xs[0] * Jv[1] - Xs[1] * Jv[0]) * sv;
      gu[6] = (Xe[1] * Ju[2] - Xe[2] * Ju[1]) * eu; gu[7] = (Xe[2] * Ju[0] - Xe[0] * Ju[2]) * eu; gu[8] = (Xe[0] * Ju[1] - Xe[1] * Ju[0]) * eu;
      gv[6] = (Xe[1] * Jv[2] - Xe[2] * Jv[1]) * ev; gv[7] = (Xe[2] * Jv[0] - Xe[0] * Jv[2]) * ev; gv[8] = (Xe[0] * Jv[1] - Xe[1] * Jv[0]) * ev;
#pragma unroll
      for (int i = 0; i < 3; i++) { gu[3 + i] = Ju[i] * su; gv[3 + i] = Jv[i] * sv; gu[9 + i] = Ju[i] * eu; gv[9 + i] = Jv[i] * ev; }
    }
    gu[KO] = ku[0] * wu; gu[KO + 2] = wu;
    gv[KO + 1] = kv[1] * wv; gv[KO + 3] = wv;
#pragma unroll
    for (int i = 0; i < ND; i++) { gu[KO + 4 + i] = ku[4 + i] * wu; gv[KO + 4 + i] = kv[4 + i] * wv; }
    gu[D] = ru; gv[D] = rv;                      // residual column: Gt^T Gt then carries G^T r as well
  }
  // stage: rows 2*lane (u) and 2*lane+1 (v); one 16-byte store per column, consecutive lanes -> consecutive addresses
#pragma unroll
  for (int j = 0; j < NC; j++)
    *reinterpret_cast<double2*>(stage + j * MMA_KPAD + 2 * lane) = make_double2(gu[j], gv[j]);
  __syncwarp();
  // k-steps of 4 residual rows; fragment of column tile I = Gt[k0 + tig][8 I + grp] serves as A (row tile) and B (col tile)
  const int ksteps = (2 * min(32, end - base) + 3) >> 2;       // ragged last chunk: skip all-zero row groups
#pragma unroll 4
  for (int ks = 0; ks < ksteps; ks++) {
    double fr[NT];
#pragma unroll
    for (int I = 0; I < NT; I++) fr[I] = stage[(8 * I + grp) * MMA_KPAD + 4 * ks + tig];
    int t = 0;
#pragma unroll
    for (int I = 0; I < NT; I++)
#pragma unroll
      for (int J = I; J < NT; J++) { dmma884(acc[t][0], acc[t][1], fr[I], fr[J]); t++; }
  }
  __syncwarp();
}

struct LinArgs {
  int loss; double f_scale;
  int split;             // warps per view: LIN_WARPS / split camera slots per CTA (split > 1 when the rig has fewer cameras than a CTA has warps)
  double* Hff;           // [F][FB*FB]
  double* g;             // [n]: frame parts written here (shared part by k_reduce_shared)
  double* W;             // [F][n_s][FB]
  double* spart;         // [grid][C][rec]   rec = T + B*UB   per-(CTA, camera) partial sums of the shared blocks
  double* frame_cost;    // [F]
};

__host__ __device__ inline int lin_record_doubles(int T, int D, int B) { return T + B * (D * 6 + 42); }
// dynamic shared memory of k_linearize in doubles
__host__ __device__ inline size_t lin_warp_doubles(int NC, int T, int D, int FB, int nin, int B, int NP) {
  (void)NP;
  return ((size_t)NC * MMA_KPAD           // stage (chunk loop) = Ms | Tf | Tb | Ac | Af | Ab | map scratch (epilogue)
       + (size_t)B * 6 * FB               // Wb: this warp's partial board rows of W_f
       + FB * FB + FB                     // hacc: H_ff | g_f partial
       + (6 + nin) * FB                   // wacc: camera-pose and intrinsics rows of W_f (current camera)
       + T                                // macc: raw moment sum of the current camera
       + (D * 6 + 42)                     // ub: board coupling of the current (camera, board)
       + 2                                // cost of this frame's views | pad
       + 1) & ~(size_t)1;                 // even: every warp's stage buffer takes 16-byte stores
}
__host__ __device__ inline size_t lin_smem_doubles(int NC, int T, int D, int FB, int nin, int B, int NP, int npair, int split) {
  (void)npair; (void)split;      // split > 1: a warp's fragments meet in its own stage buffer, (2 npair + 1) x 32 <= NC x MMA_KPAD doubles
  return (size_t)LIN_WARPS * lin_warp_doubles(NC, T, D, FB, nin, B, NP);
}

template <int MODEL, bool ROLL>
__global__ void __launch_bounds__(LIN_THREADS, 2)
k_linearize(DeviceProblem p, LinArgs a) {
  using S = LinShape<MODEL, ROLL>;
  constexpr int NP = S::NP, KO = S::KO, NIN = S::NIN, D = S::D, E = S::E, T = S::T, NC = S::NC, NT = S::NT, NPAIR = S::NPAIR, KINT = S::KINT, FB = S::FB, UB = S::UB;
  constexpr int NHF = FB * FB + FB;
  constexpr int NWC = (6 + NIN) * FB;
  extern __shared__ double lsm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane >> 2, tig = lane & 3;
  const int B = p.B, n_s = p.n_s;
  const int split = a.split, slots = LIN_WARPS / split, slot = warp / split, sub = warp % split;
  const bool leader = sub == 0;
  const bool frames_on = p.motion_on != 0;
  const size_t wd = lin_warp_doubles(NC, T, D, FB, NIN, B, NP);
  double* stage = lsm + (size_t)warp * wd;
  double* Ms = stage;                         // [NC][NC] full symmetric, column D = G^T r (epilogue view of the stage buffer)
  double* Tf = Ms + NC * NC;                  // [D][FB]   M[:, xi_a] Af_a
  double* Tb = Tf + D * FB;                   // [D][6]    sum_a M[:, xi_a] Ab_a
  double* Ac = Tb + D * 6;                    // twist maps of the view, also inside the stage buffer
  double* Af = Ac + 36;
  double* Ab = Af + 36 * NP;
  double* scr = Ab + 36 * NP;
  static_assert(NC * NC + D * FB + D * 6 + 36 + 72 * NP + 33 * NP <= NC * MMA_KPAD, "the epilogue's scratch must fit the stage buffer");
  double* Wb = stage + (size_t)NC * MMA_KPAD; // [B][6][FB]
  double* hacc = Wb + (size_t)B * 6 * FB;     // [NHF]
  double* wacc = hacc + NHF;                  // [NWC]
  double* macc = wacc + NWC;                  // [T]
  double* ub = macc + T;                      // [UB]
  double* wcost = ub + UB;                    // [1]
  static_assert((2 * NPAIR + 1) * 32 <= NC * MMA_KPAD, "the fragments of a view part must fit the warp's stage buffer");
  const int rec = lin_record_doubles(T, D, B);
  double* myrec = a.spart + (size_t)blockIdx.x * p.C * rec;

  // this CTA's records start at zero; the leader warps' running sums too
  for (int i = tid; i < p.C * rec; i += LIN_THREADS) myrec[i] = 0.0;
  if (leader) {
    for (int i = lane; i < NWC + T + UB; i += 32) wacc[i] = 0.0;       // wacc | macc | ub are contiguous
  }
  int cur_cam = -1, cur_board = -1;           // what macc / wacc (camera) and ub (camera, board) currently hold   (leader warps)
  bool wdirty = false;                        // wacc holds rows of cur_cam for the frame in progress
  __syncthreads();

  auto flush_ub = [&]() {                     // ub -> record of (cur_cam, cur_board)
    if (cur_cam < 0 || cur_board < 0 || p.off_bp < 0) return;
    double* r = myrec + (size_t)cur_cam * rec + T + (size_t)cur_board * UB;
    for (int i = lane; i < UB; i += 32) { r[i] = r[i] + ub[i]; ub[i] = 0.0; }
  };
  auto flush_macc = [&]() {
    if (cur_cam < 0) return;
    double* r = myrec + (size_t)cur_cam * rec;
    for (int i = lane; i < T; i += 32) { r[i] = r[i] + macc[i]; macc[i] = 0.0; }
  };
  auto flush_wacc = [&](int f) {              // camera rows of W_f (complete for this frame: one warp owns a camera)
    if (cur_cam < 0 || !frames_on || !wdirty) return;
    wdirty = false;
    double* Wf = a.W + (size_t)f * n_s * FB;
    for (int o = lane; o < NWC; o += 32) {
      const double val = wacc[o];
      const int row = o / FB, col = o % FB;
      if (row < 6) { if (p.off_cp >= 0) Wf[(size_t)(p.off_cp + 6 * cur_cam + row) * FB + col] = val; }
      else if (p.off_in >= 0) {
        const int i = row - 6;
        if (p.fix_aspect && i == 1) continue;                            // fy follows fx (camera.py:159-160): its row is folded below
        const double v2 = (p.fix_aspect && i == 0) ? val + wacc[(6 + 1) * FB + col] : val;
        Wf[(size_t)(p.off_in + p.kint * cur_cam + intr_param_index(p, i)) * FB + col] = v2;
      }
    }
    __syncwarp();
    for (int o = lane; o < NWC; o += 32) wacc[o] = 0.0;
  };

  for (int f = blockIdx.x; f < p.F; f += gridDim.x) {
    const int v0 = p.frame_view_start[f], v1 = p.frame_view_start[f + 1];
    if (frames_on) {
      double* Wf = a.W + (size_t)f * n_s * FB;
      for (int i = tid; i < n_s * FB; i += LIN_THREADS) Wf[i] = 0.0;
    }
    if (leader) {
      for (int i = lane; i < B * 6 * FB + NHF; i += 32) Wb[i] = 0.0;     // Wb | hacc are contiguous
      if (lane == 0) wcost[0] = 0.0;
    }
    __syncthreads();

    int v = v0;
    while (true) {
      while (v < v1 && (p.view_cam[v] % slots) != slot) v++;
      const bool have = v < v1;
      if (split == 1) { if (!have) break; }
      else {
        // the warps of a slot walk the same views; the CTA meets twice per round, so every warp runs the same number of rounds
        __shared__ int any_view;
        if (tid == 0) any_view = 0;
        __syncthreads();
        if (have && lane == 0 && leader) any_view = 1;
        __syncthreads();
        if (!any_view) break;
      }
      double acc[NPAIR][2];
#pragma unroll
      for (int i = 0; i < NPAIR; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
      double cost_acc = 0.0;
      int c = 0, b = 0;
      if (have) {
        c = p.view_cam[v]; b = p.view_board[v];
        const int beg = p.view_start[v], end = p.view_start[v + 1];
        ViewPose vp, vpe;
        compose_views<ROLL>(p, c, f, b, vp, vpe);
        const double inv_h = ROLL ? 1.0 / p.img_h[c] : 0.0;
        double k[KINT];
#pragma unroll
        for (int i = 0; i < KINT; i++) k[i] = p.intr[c * KINT + i];
        const double* bp = p.board_pts + (size_t)b * p.P * 3;
        for (int base = beg + 32 * sub; base < end; base += 32 * split)
          view_chunk<MODEL, ROLL>(p, a.loss, a.f_scale, vp, vpe, k, bp, inv_h, base, end, lane, stage, acc, cost_acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cost_acc += __shfl_xor_sync(0xffffffffu, cost_acc, o);
      }
      if (split > 1) {        // meet: the slot's leader adds the other warps' fragments in warp order (same lane -> same matrix element)
        if (have && !leader) {
          double* xr = stage;                  // this warp's stage buffer is free between two views
#pragma unroll
          for (int t = 0; t < NPAIR; t++) { xr[(2 * t) * 32 + lane] = acc[t][0]; xr[(2 * t + 1) * 32 + lane] = acc[t][1]; }
          if (lane == 0) xr[2 * NPAIR * 32] = cost_acc;
        }
        __syncthreads();
        if (have && leader) {
          for (int w = 1; w < split; w++) {
            const double* xr = lsm + (size_t)(warp + w) * wd;
#pragma unroll
            for (int t = 0; t < NPAIR; t++) { acc[t][0] += xr[(2 * t) * 32 + lane]; acc[t][1] += xr[(2 * t + 1) * 32 + lane]; }
            cost_acc += xr[2 * NPAIR * 32];
          }
        }
      }
      if (have && leader) {
        // ---------------------------------------------------------------- epilogue of one view (camera c, frame f, board b)
        if (c != cur_cam) { flush_ub(); flush_wacc(f); flush_macc(); cur_cam = c; cur_board = -1; }
        if (b != cur_board) { flush_ub(); cur_board = b; }
        __syncwarp();
        // fragments -> Ms (both triangles) and the camera's raw moment sum
        {
          int t = 0;
#pragma unroll
          for (int I = 0; I < NT; I++)
#pragma unroll
            for (int J = I; J < NT; J++) {
#pragma unroll
              for (int h = 0; h < 2; h++) {
                const int i = 8 * I + grp, j = 8 * J + 2 * tig + h;
                const double val = acc[t][h];
                Ms[i * NC + j] = val;
                if (I != J) Ms[j * NC + i] = val;
                if (i < D && j < D && i <= j) macc[tri_index(D, i, j)] += val;
                else if (i < D && j == D) macc[E + i] += val;
              }
              t++;
            }
          if (lane == 0) { macc[T - 1] += cost_acc; wcost[0] += cost_acc; }
        }
        view_twist_maps_par<NP>(p, c, f, b, lane, Ac, frames_on ? Af : nullptr, p.off_bp >= 0 ? Ab : nullptr, scr);
        __syncwarp();
        if (frames_on) {
          for (int o = lane; o < D * FB; o += 32) {           // Tf[:, 6a+k] = M[:, xi_a] Af_a
            const int i = o / FB, col = o % FB, ablk = col / 6, kk0 = col % 6;
            double s = 0.0;
#pragma unroll
            for (int kk = 0; kk < 6; kk++) s += Ms[i * NC + 6 * ablk + kk] * Af[36 * ablk + kk * 6 + kk0];
            Tf[o] = s;
          }
        }
        if (p.off_bp >= 0) {
          for (int o = lane; o < D * 6; o += 32) {            // Tb = sum_a M[:, xi_a] Ab_a
            const int i = o / 6, j = o % 6;
            double s = 0.0;
#pragma unroll
            for (int kk = 0; kk < KO; kk++) s += Ms[i * NC + kk] * Ab[36 * (kk / 6) + (kk % 6) * 6 + j];
            Tb[o] = s;
          }
        }
        __syncwarp();
        if (frames_on) {
          // H_ff[6a+i, col] += Af_a^T Tf[xi_a rows, col] ; g_f[6a+i] += Af_a^T g_xi_a
          for (int o = lane; o < NHF; o += 32) {
            double s = 0.0;
            if (o < FB * FB) {
              const int r = o / FB, col = o % FB, ablk = r / 6, i = r % 6;
#pragma unroll
              for (int kk = 0; kk < 6; kk++) s += Af[36 * ablk + kk * 6 + i] * Tf[(6 * ablk + kk) * FB + col];
            } else {
              const int r = o - FB * FB, ablk = r / 6, i = r % 6;
#pragma unroll
              for (int kk = 0; kk < 6; kk++) s += Af[36 * ablk + kk * 6 + i] * Ms[(6 * ablk + kk) * NC + D];
            }
            hacc[o] += s;
          }
          // camera rows of W_f: sum_a Ac^T Tf[xi_a rows] (6 x FB) | Tf[kappa rows] (NIN x FB)
          for (int o = lane; o < NWC; o += 32) {
            const int row = o / FB, col = o % FB;
            double s = 0.0;
            if (row < 6) {
#pragma unroll
              for (int kk = 0; kk < KO; kk++) s += Ac[(kk % 6) * 6 + row] * Tf[kk * FB + col];
            } else s = Tf[(KO + row - 6) * FB + col];
            wacc[o] += s;
          }
          wdirty = true;
          // board rows of W_f (shared between cameras): this warp's partial
          if (p.off_bp >= 0) {
            for (int o = lane; o < 6 * FB; o += 32) {
              const int i = o / FB, col = o % FB;
              double s = 0.0;
#pragma unroll
              for (int kk = 0; kk < KO; kk++) s += Ab[36 * (kk / 6) + (kk % 6) * 6 + i] * Tf[kk * FB + col];
              Wb[(size_t)b * 6 * FB + o] += s;
            }
          }
        }
        if (p.off_bp >= 0) {
          // shared blocks that need this view's board map: (camera | intrinsics) x board pose, board x board, board gradient
          for (int o = lane; o < UB; o += 32) {
            double s;
            if (o < D * 6) s = Tb[o];
            else if (o < D * 6 + 36) {
              const int q = o - D * 6, i = q / 6, j = q % 6;
              s = 0.0;
#pragma unroll
              for (int kk = 0; kk < KO; kk++) s += Ab[36 * (kk / 6) + (kk % 6) * 6 + i] * Tb[kk * 6 + j];
            } else {
              const int i = o - D * 6 - 36;
              s = 0.0;
#pragma unroll
              for (int kk = 0; kk < KO; kk++) s += Ab[36 * (kk / 6) + (kk % 6) * 6 + i] * Ms[kk * NC + D];
            }
            ub[o] += s;
          }
        }
        __syncwarp();
      }
      if (have) v++;
    }
    // ---- end of the frame: camera rows out, then the CTA sums the slots' partials in slot order
    if (leader) { flush_wacc(f); }
    __syncthreads();
    if (frames_on) {
      for (int o = tid; o < NHF; o += LIN_THREADS) {
        double s = 0.0;
        for (int w = 0; w < slots; w++) s += lsm[(size_t)(w * split) * wd + (hacc - stage) + o];
        if (o < FB * FB) a.Hff[(size_t)f * FB * FB + o] = s; else a.g[n_s + FB * f + o - FB * FB] = s;
      }
      if (p.off_bp >= 0) {
        double* Wf = a.W + (size_t)f * n_s * FB;
        for (int o = tid; o < B * 6 * FB; o += LIN_THREADS) {
          double s = 0.0;
          for (int w = 0; w < slots; w++) s += lsm[(size_t)(w * split) * wd + (Wb - stage) + o];
          const int bb = o / (6 * FB), i = (o % (6 * FB)) / FB, j = o % FB;
          Wf[(size_t)(p.off_bp + 6 * bb + i) * FB + j] = s;
        }
      }
    }
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < slots; w++) s += lsm[(size_t)(w * split) * wd + (wcost - stage)];
      a.frame_cost[f] = s;
    }
    __syncthreads();
  }
  if (leader) { flush_ub(); flush_macc(); }
}

// ------------------------------------------------------------------------------------------------
// k_reduce_shared: the per-CTA records of k_linearize -> H_ss, g_s, cost.  No atomics on data: two runs give bit-identical results.
//   step 1  grid = cameras x slices of 32 record entries; 256 threads = 32 entries x 8 groups: group g adds records g, g+8, ... in
//           order, the 8 group sums are added in order -> the camera's summed record (global scratch `sred`)
//   step 2  the LAST slice CTA of a camera (arrival counter) applies the camera's own twist map (it does not depend on the view) and
//           STORES the camera's blocks of H_ss / g_s: (pose | intrinsics) x (pose | intrinsics), and x every board pose
//   step 3  board x board blocks and board gradients are sums over the cameras: the last camera to finish step 2 adds the per-camera
//           partials in camera order, and the per-frame costs in frame order
constexpr int RED_THREADS = 256;
constexpr int RED_ENT = 32, RED_GROUPS = RED_THREADS / RED_ENT;
struct ReduceArgs {
  const double* spart; int nparts;            // [nparts][C][rec]
  double* sred;                               // [C][rec]
  double* Hss; double* g;
  double* bpart;                              // [C][B][42]
  const double* frame_cost; int F; double* cost_out;
  unsigned* cam_counter;                      // [C + 1]: slices done per camera | cameras done
};
__host__ __device__ inline int reduce_slices(int rec) { return (rec + RED_ENT - 1) / RED_ENT; }
__host__ __device__ inline size_t reduce_smem_doubles(int T, int D, int B) { return (size_t)lin_record_doubles(T, D, B) + D * 6 + 36; }

template <int NP>
__global__ void __launch_bounds__(RED_THREADS)
k_reduce_shared(DeviceProblem p, ReduceArgs a) {
  constexpr int KO = 6 * NP;
  extern __shared__ double rsm[];
  __shared__ double sm[32];
  __shared__ double gsum[RED_GROUPS][RED_ENT];
  __shared__ int is_last;
  const int D = p.D, T = p.T, B = p.B, n_s = p.n_s;
  const int E = D * (D + 1) / 2;
  const int UB = D * 6 + 42;
  const int rec = lin_record_doubles(T, D, B);
  const int nsl = reduce_slices(rec);
  const int c = blockIdx.x / nsl, slice = blockIdx.x % nsl, tid = threadIdx.x;
  // ---- step 1
  {
    const int e = slice * RED_ENT + (tid % RED_ENT), grp = tid / RED_ENT;
    double s = 0.0;
    if (e < rec) {
      const double* src = a.spart + (size_t)c * rec + e;
      const size_t stride = (size_t)p.C * rec;
      int q = grp;
      for (; q + 3 * RED_GROUPS < a.nparts; q += 4 * RED_GROUPS) {       // four loads in flight, added in record order
        const double v0 = src[(size_t)q * stride], v1 = src[(size_t)(q + RED_GROUPS) * stride];
        const double v2 = src[(size_t)(q + 2 * RED_GROUPS) * stride], v3 = src[(size_t)(q + 3 * RED_GROUPS) * stride];
        s = (((s + v0) + v1) + v2) + v3;
      }
      for (; q < a.nparts; q += RED_GROUPS) s += src[(size_t)q * stride];
    }
    gsum[grp][tid % RED_ENT] = s;
    __syncthreads();
    if (tid < RED_ENT && slice * RED_ENT + tid < rec) {
      double t = 0.0;
#pragma unroll
      for (int gq = 0; gq < RED_GROUPS; gq++) t += gsum[gq][tid];
      a.sred[(size_t)c * rec + slice * RED_ENT + tid] = t;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) { const unsigned t = atomicAdd(a.cam_counter + c, 1u); is_last = (t == (unsigned)nsl - 1); }
  __syncthreads();
  if (!is_last) return;
  if (tid == 0) a.cam_counter[c] = 0;
  __threadfence();
  // ---- step 2: this camera's blocks
  double* Msum = rsm;                         // [T] then per board [UB]
  double* Um = rsm + rec;                     // [D][6]
  double* Ac = Um + D * 6;                    // 36
  for (int i = tid; i < rec; i += RED_THREADS) Msum[i] = __ldcg(&a.sred[(size_t)c * rec + i]);
  if (tid == 0) { const PoseT& pc = p.cam_T[c]; const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; twist_map(I3, pc.JL, pc.t, Ac); }
  __syncthreads();
  for (int o = tid; o < D * 6; o += RED_THREADS) {       // Um = sum_a Msum[:, xi_a] Ac  (the camera's map is the same for every chain)
    const int i = o / 6, j = o % 6; double acc = 0.0;
    for (int kk = 0; kk < KO; kk++) acc += msym(Msum, D, i, kk) * Ac[(kk % 6) * 6 + j];
    Um[o] = acc;
  }
  __syncthreads();
  const int nin = 4 + p.nd;
  const int cp = p.off_cp >= 0 ? p.off_cp + 6 * c : -1;
  const int in0 = p.off_in >= 0 ? p.off_in + p.kint * c : -1;
  auto put = [&](int i, int j, double val) { a.Hss[(size_t)i * n_s + j] = val; a.Hss[(size_t)j * n_s + i] = val; };
  if (cp >= 0) {
    for (int o = tid; o < 36; o += RED_THREADS) {
      const int i = o / 6, j = o % 6;
      if (j < i) continue;
      double acc = 0.0;
      for (int kk = 0; kk < KO; kk++) acc += Ac[(kk % 6) * 6 + i] * Um[kk * 6 + j];
      put(cp + i, cp + j, acc);
    }
    if (tid < 6) {
      double acc = 0.0;
      for (int kk = 0; kk < KO; kk++) acc += Ac[(kk % 6) * 6 + tid] * Msum[E + kk];
      a.g[cp + tid] = acc;
    }
  }
  if (in0 >= 0) {
    // one thread per TARGET parameter pair: rows folded onto one parameter (fix_aspect: fy onto fx, camera.py:159-160) are added in a fixed order
    const int kint = p.kint;
    if (cp >= 0)
      for (int o = tid; o < kint * 6; o += RED_THREADS) {
        const int pi = o / 6, j = o % 6;
        double acc = 0.0;
        for (int i = 0; i < nin; i++) if (intr_param_index(p, i) == pi) acc += Um[(KO + i) * 6 + j];
        put(in0 + pi, cp + j, acc);
      }
    for (int o = tid; o < kint * kint; o += RED_THREADS) {
      const int pi = o / kint, pj = o % kint;
      double acc = 0.0;
      for (int i = 0; i < nin; i++) if (intr_param_index(p, i) == pi)
        for (int j = 0; j < nin; j++) if (intr_param_index(p, j) == pj) acc += msym(Msum, D, KO + i, KO + j);
      a.Hss[(size_t)(in0 + pi) * n_s + in0 + pj] = acc;
    }
    for (int pi = tid; pi < kint; pi += RED_THREADS) {
      double acc = 0.0;
      for (int i = 0; i < nin; i++) if (intr_param_index(p, i) == pi) acc += Msum[E + KO + i];
      a.g[in0 + pi] = acc;
    }
  }
  if (p.off_bp >= 0) {
    for (int b = 0; b < B; b++) {
      const int bp = p.off_bp + 6 * b;
      const double* U = Msum + T + (size_t)b * UB;       // [D][6] | Hbb 36 | gb 6
      if (cp >= 0)
        for (int o = tid; o < 36; o += RED_THREADS) {     // camera pose x board pose = sum_a Ac^T U_xi_a
          const int i = o / 6, j = o % 6; double acc = 0.0;
          for (int kk = 0; kk < KO; kk++) acc += Ac[(kk % 6) * 6 + i] * U[kk * 6 + j];
          put(cp + i, bp + j, acc);
        }
      if (in0 >= 0)
        for (int o = tid; o < p.kint * 6; o += RED_THREADS) { // intrinsics x board pose = U_kappa
          const int pi = o / 6, j = o % 6;
          double acc = 0.0;
          for (int i = 0; i < nin; i++) if (intr_param_index(p, i) == pi) acc += U[(KO + i) * 6 + j];
          put(in0 + pi, bp + j, acc);
        }
      for (int o = tid; o < 42; o += RED_THREADS) a.bpart[((size_t)c * B + b) * 42 + o] = U[D * 6 + o];
    }
  }
  // ---- step 3 (last camera): board x board blocks, board gradients (camera order) and the cost (frame order)
  __threadfence();
  __syncthreads();
  if (tid == 0) { const unsigned t = atomicAdd(a.cam_counter + p.C, 1u); is_last = (t == (unsigned)p.C - 1); }
  __syncthreads();
  if (!is_last) return;
  if (tid == 0) a.cam_counter[p.C] = 0;
  __threadfence();
  if (p.off_bp >= 0)
    for (int o = tid; o < B * 42; o += RED_THREADS) {
      const int b = o / 42, q = o % 42;
      double acc = 0.0;
      for (int cc = 0; cc < p.C; cc++) acc += __ldcg(&a.bpart[((size_t)cc * B + b) * 42 + q]);
      const int bp = p.off_bp + 6 * b;
      if (q < 36) a.Hss[(size_t)(bp + q / 6) * n_s + bp + q % 6] = acc; else a.g[bp + q - 36] = acc;
    }
  double cs = 0.0;
  for (int f = tid; f < a.F; f += RED_THREADS) cs += a.frame_cost[f];
  cs = block_sum(cs, sm);
  if (tid == 0) *a.cost_out = cs;
}

}  // namespace mcba
