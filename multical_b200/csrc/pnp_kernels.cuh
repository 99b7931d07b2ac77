// pnp_kernels.cuh — batched board-pose initialisation: the C*F*B calls of
//   board.estimate_pose_points   board/common.py:36-47   (undistort the detected corners, cv2.solvePnPGeneric with the camera
//                                                          matrix and no distortion, RMS reprojection error)
// that tables.make_pose_table (tables.py:44-66) makes one view at a time, as one launch with one warp per detection list.
//
// Per view (camera c, board b, n detected corners with ids):
//   1. has_min_detections_grid (board/common.py:30-34; charuco.py:104-106, aprilgrid.py:197-199): at least min_points corners
//      and min_rows distinct rows AND columns of the id grid, else the view is invalid (tables.py:38 invalid_pose).
//   2. camera.undistort_points (camera.py:119-122 = cv2.undistortPoints(pts, K, dist, P=K): exactly 5 fixed-point iterations of
//      the inverse distortion, as OpenCV's default criteria; camera_fisheye.py:108-111 = cv2.fisheye.undistortPoints: up to 10
//      Newton steps on theta), result rounded to float32 like `.astype('float32')` (board/common.py:40).
//   3. cv2.solvePnPGeneric(objPoints, undistorted, K, no distortion) (board/common.py:42), default SOLVEPNP_ITERATIVE: for a planar
//      target a homography initialisation refined by Levenberg-Marquardt on the pixel reprojection error.  OpenCV's own run ends
//      within ~3e-10 of the minimiser (measured), so the device iterates its own LM to convergence from its own
//      homography initialisation: the same minimum, not the same trajectory.
//   4. error = RMS over the 2n scalar residuals (the reprojectionError solvePnPGeneric returns).
// Everything is fp64; the 8x8 (homography) and 6x6 (LM) systems are solved redundantly by every lane from warp-reduced sums, so
// the warp never diverges on data.
#pragma once
#include <stdint.h>
#include "geometry.cuh"

namespace mcba {

struct PnpArgs {
  int C, F, B, P, model, kint, nv;
  const int64_t* det_start;     // [nv+1] CSR over lists w = (c*F+f)*B+b
  const int32_t* det_ids;       // point ids on board b
  const double2* det_xy;        // detected pixel corners
  const double* board_pts;      // [B][P][3]
  const double* intr;           // [C][kint] = [fx fy cx cy skew dist...]
  const int32_t* grid;          // [B][5] = id-grid width, height, id divisor (4 for AprilGrid tags), min_points, min_rows
  double2* und;                 // scratch [total]: undistorted pixels (float32-rounded)
  double* poses;                // [nv][16]
  double* err;                  // [nv]
  int32_t* npts;                // [nv]
  uint8_t* valid;               // [nv]
  int max_iters;
};

template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cv2.undistortPoints(..., P = K) for one pixel; MODEL_TILTED un-tilts first (OpenCV computeTiltProjectionMatrix inverse)
template <int MODEL>
__host__ __device__ inline void undistort_pixel(const double* k, double u, double v, double& uo, double& vo) {
  constexpr int ND = model_nd(MODEL);
  const double fx = k[0], fy = k[1], cx = k[2], cy = k[3], skew = k[4];
  const double* d = k + 5;
  double x = (u - cx) / fx, y = (v - cy) / fy;
  if constexpr (MODEL == MODEL_FISHEYE) {
    // cv2.fisheye.undistortPoints: theta_d = |p| clipped to [-pi/2, pi/2], Newton on theta (<= 10 steps, eps 1e-8), scale = tan(theta)/theta_d
    const double PI_2 = 1.5707963267948966;
    double theta_d = sqrt(x * x + y * y);
    theta_d = fmin(fmax(-PI_2, theta_d), PI_2);
    bool converged = false;
    double theta = theta_d, scale = 0.0;
    if (fabs(theta_d) > 1e-8) {
      for (int j = 0; j < 10; j++) {
        const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
        const double k0 = d[0] * t2, k1 = d[1] * t4, k2 = d[2] * t6, k3 = d[3] * t8;
        const double fix = (theta * (1 + k0 + k1 + k2 + k3) - theta_d) / (1 + 3 * k0 + 5 * k1 + 7 * k2 + 9 * k3);
        theta -= fix;
        if (fabs(fix) < 1e-8) { converged = true; break; }
      }
      scale = tan(theta) / theta_d;
    } else converged = true;
    const bool flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
    if (converged && !flipped) { x *= scale; y *= scale; }
    else { uo = -1000000.0; vo = -1000000.0; return; }
  } else {
    if constexpr (ND >= 14) {
      double M[9], dMx[9], dMy[9];
      tilt_matrices(d[12], d[13], M, dMx, dMy);
      // inverse of the 3x3 tilt matrix by cofactors
      const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
      const double det = M[0] * c00 + M[1] * c01 + M[2] * c02, id = 1.0 / det;
      const double I0 = c00 * id, I1 = (M[2] * M[7] - M[1] * M[8]) * id, I2 = (M[1] * M[5] - M[2] * M[4]) * id;
      const double I3 = c01 * id, I4 = (M[0] * M[8] - M[2] * M[6]) * id, I5 = (M[2] * M[3] - M[0] * M[5]) * id;
      const double I6 = c02 * id, I7 = (M[1] * M[6] - M[0] * M[7]) * id, I8 = (M[0] * M[4] - M[1] * M[3]) * id;
      const double a = I0 * x + I1 * y + I2, b = I3 * x + I4 * y + I5, w = I6 * x + I7 * y + I8;
      const double iw = w != 0.0 ? 1.0 / w : 1.0;
      x = a * iw; y = b * iw;
    }
    const double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
    double k4 = 0, k5 = 0, k6 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    if constexpr (ND >= 8) { k4 = d[5]; k5 = d[6]; k6 = d[7]; }
    if constexpr (ND >= 12) { s1 = d[8]; s2 = d[9]; s3 = d[10]; s4 = d[11]; }
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
      const double r2 = x * x + y * y;
      const double icdist = (1 + ((k6 * r2 + k5) * r2 + k4) * r2) / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
      if (icdist < 0) { x = x0; y = y0; break; }       // OpenCV: "test: undistortPoints.regression_14583"
      const double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + s1 * r2 + s2 * r2 * r2;
      const double dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + s3 * r2 + s4 * r2 * r2;
      x = (x0 - dx) * icdist;
      y = (y0 - dy) * icdist;
    }
  }
  uo = fx * x + skew * y + cx;      // P = K as a 3x3 matrix
  vo = fy * y + cy;
}

// in-place Cholesky solve of the N x N SPD system A x = b (A full row-major, destroyed); false when a pivot is not positive
template <int N>
__host__ __device__ inline bool spd_solve(double* A, double* b) {
#pragma unroll
  for (int j = 0; j < N; j++) {
    double s = A[j * N + j];
#pragma unroll
    for (int k = 0; k < j; k++) s -= A[j * N + k] * A[j * N + k];
    if (!(s > 0.0)) return false;
    const double piv = sqrt(s);
    A[j * N + j] = piv;
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      double t = A[i * N + j];
#pragma unroll
      for (int k = 0; k < j; k++) t -= A[i * N + k] * A[j * N + k];
      A[i * N + j] = t / piv;
    }
  }
#pragma unroll
  for (int i = 0; i < N; i++) {
    double t = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) t -= A[i * N + k] * b[k];
    b[i] = t / A[i * N + i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    double t = b[i];
#pragma unroll
    for (int k = i + 1; k < N; k++) t -= A[k * N + i] * b[k];
    b[i] = t / A[i * N + i];
  }
  return true;
}

// pose from the plane-to-normalised-image homography H (columns h1 h2 h3 ~ r1 r2 t): Gram-Schmidt on (h1, h2), r3 = r1 x r2
__host__ __device__ inline void pose_from_homography(const double* H /*3x3 row-major*/, double* R, double* t) {
  double h1[3] = {H[0], H[3], H[6]}, h2[3] = {H[1], H[4], H[7]}, h3[3] = {H[2], H[5], H[8]};
  const double n1 = sqrt(h1[0] * h1[0] + h1[1] * h1[1] + h1[2] * h1[2]);
  const double n2 = sqrt(h2[0] * h2[0] + h2[1] * h2[1] + h2[2] * h2[2]);
  const double lam = 2.0 / (n1 + n2);
  double r1[3] = {h1[0] / n1, h1[1] / n1, h1[2] / n1};
  const double dp = r1[0] * h2[0] + r1[1] * h2[1] + r1[2] * h2[2];
  double r2[3] = {h2[0] - dp * r1[0], h2[1] - dp * r1[1], h2[2] - dp * r1[2]};
  const double m2 = sqrt(r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
  r2[0] /= m2; r2[1] /= m2; r2[2] /= m2;
  const double r3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
#pragma unroll
  for (int i = 0; i < 3; i++) { R[3 * i] = r1[i]; R[3 * i + 1] = r2[i]; R[3 * i + 2] = r3[i]; t[i] = lam * h3[i]; }
}

// T <- exp(omega) T + v  (first-order consistent with the left twist d x_cam = omega x x_cam + v)
__host__ __device__ inline void apply_twist(const double* delta, const double* R, const double* t, double* Rn, double* tn) {
  double E[9], JL[9];
  rodrigues(delta, E, JL);
  mat3_mul(E, R, Rn);
  mat3_vec(E, t, tn);
  tn[0] += delta[3]; tn[1] += delta[4]; tn[2] += delta[5];
}

constexpr int PNP_WARPS = 4;

template <int MODEL>
__global__ void __launch_bounds__(PNP_WARPS * 32)
k_pnp_views(PnpArgs a) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * PNP_WARPS + (threadIdx.x >> 5);
  if (w >= a.nv) return;
  const int b = w % a.B, c = w / (a.B * a.F);
  const int64_t beg = a.det_start[w], end = a.det_start[w + 1];
  const int n = (int)(end - beg);
  double* Tout = a.poses + (size_t)16 * w;
  auto invalid = [&]() {
    if (lane < 16) Tout[lane] = (lane % 5 == 0) ? 1.0 : 0.0;         // tables.py:38 invalid_pose: identity, no points, error 0
    if (lane == 0) { a.err[w] = 0.0; a.npts[w] = 0; a.valid[w] = 0; }
  };
  // ---- 1. has_min_detections_grid
  const int gw = a.grid[5 * b], gh = a.grid[5 * b + 1], gdiv = a.grid[5 * b + 2], min_points = a.grid[5 * b + 3], min_rows = a.grid[5 * b + 4];
  unsigned long long rows = 0ull, cols = 0ull;
  for (int64_t i = beg + lane; i < end; i += 32) {
    const int id = a.det_ids[i] / gdiv;
    rows |= 1ull << min(id / gw, 63);                                 // np.unravel_index(ids, (h, w)): row = id // w, col = id % w
    cols |= 1ull << min(id % gw, 63);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { rows |= __shfl_xor_sync(0xffffffffu, rows, o); cols |= __shfl_xor_sync(0xffffffffu, cols, o); }
  (void)gh;
  if (n < min_points || n < 4 || __popcll(rows) < min_rows || __popcll(cols) < min_rows) { invalid(); return; }

  // ---- 2. undistort (float32-rounded pixels) and the statistics of the object points (Hartley normalisation of the plane)
  const double* k = a.intr + (size_t)c * a.kint;
  const double fx = k[0], fy = k[1], cx = k[2], cy = k[3];
  const double* bp = a.board_pts + (size_t)b * a.P * 3;
  double sx = 0, sy = 0;
  for (int64_t i = beg + lane; i < end; i += 32) {
    const double2 p = a.det_xy[i];
    double uo, vo;
    undistort_pixel<MODEL>(k, p.x, p.y, uo, vo);
    a.und[i] = make_double2((double)(float)uo, (double)(float)vo);
    const int id = a.det_ids[i];
    sx += bp[3 * id]; sy += bp[3 * id + 1];
  }
  sx = warp_sum(sx) / n; sy = warp_sum(sy) / n;
  double sd = 0;
  for (int64_t i = beg + lane; i < end; i += 32) {
    const int id = a.det_ids[i];
    const double dx = bp[3 * id] - sx, dy = bp[3 * id + 1] - sy;
    sd += sqrt(dx * dx + dy * dy);
  }
  sd = warp_sum(sd) / n;
  const double sc = sd > 0 ? 1.4142135623730951 / sd : 1.0;
  __syncwarp();                                                       // a.und written by other lanes is read below

  // ---- 3. homography (h33 = 1) between the normalised board plane and the normalised image: 8x8 normal equations
  double R[9], t[3];
  {
    double N[44];                                                     // upper triangle of A^T A (36) | A^T b (8)
#pragma unroll
    for (int i = 0; i < 44; i++) N[i] = 0.0;
    for (int64_t i = beg + lane; i < end; i += 32) {
      const int id = a.det_ids[i];
      const double X = (bp[3 * id] - sx) * sc, Y = (bp[3 * id + 1] - sy) * sc;
      const double2 q = a.und[i];
      const double x = (q.x - cx) / fx, y = (q.y - cy) / fy;
      const double ru[8] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y}, rv[8] = {0, 0, 0, X, Y, 1, -y * X, -y * Y};
      int e = 0;
#pragma unroll
      for (int r = 0; r < 8; r++) {
#pragma unroll
        for (int s = r; s < 8; s++) { N[e] += ru[r] * ru[s] + rv[r] * rv[s]; e++; }
        N[36 + r] += ru[r] * x + rv[r] * y;
      }
    }
#pragma unroll
    for (int i = 0; i < 44; i++) N[i] = warp_sum(N[i]);
    double A[64], h[8];
    {
      int e = 0;
#pragma unroll
      for (int r = 0; r < 8; r++) {
#pragma unroll
        for (int s = r; s < 8; s++) { A[r * 8 + s] = N[e]; A[s * 8 + r] = N[e]; e++; }
        h[r] = N[36 + r];
      }
    }
    if (!spd_solve<8>(A, h)) { invalid(); return; }
    // H maps (X, Y, 1) of the NORMALISED plane; undo the normalisation: X_n = sc (X - sx)
    const double Hn[9] = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], 1.0};
    double H[9];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      H[3 * r] = Hn[3 * r] * sc;
      H[3 * r + 1] = Hn[3 * r + 1] * sc;
      H[3 * r + 2] = Hn[3 * r + 2] - sc * (Hn[3 * r] * sx + Hn[3 * r + 1] * sy);
    }
    if (H[8] < 0) {
#pragma unroll
      for (int i = 0; i < 9; i++) H[i] = -H[i];                       // the board is in front of the camera
    }
    pose_from_homography(H, R, t);
  }

  // ---- 4. Levenberg-Marquardt on the pixel reprojection error with the pinhole K (no distortion), left-twist updates
  auto cost_of = [&](const double* Rc, const double* tc) {
    double s = 0.0;
    for (int64_t i = beg + lane; i < end; i += 32) {
      const int id = a.det_ids[i];
      const double X[3] = {bp[3 * id], bp[3 * id + 1], bp[3 * id + 2]};
      double Xc[3];
      mat3_vec(Rc, X, Xc);
      Xc[0] += tc[0]; Xc[1] += tc[1]; Xc[2] += tc[2];
      const double iz = Xc[2] != 0.0 ? 1.0 / Xc[2] : 1.0;
      const double2 q = a.und[i];
      const double ru = fx * Xc[0] * iz + cx - q.x, rv = fy * Xc[1] * iz + cy - q.y;
      s += ru * ru + rv * rv;
    }
    return warp_sum(s);
  };
  double lambda = 1e-3;
  double cost = cost_of(R, t);
  for (int it = 0; it < a.max_iters; it++) {
    double N[27];                                                     // upper triangle of J^T J (21) | J^T r (6)
#pragma unroll
    for (int i = 0; i < 27; i++) N[i] = 0.0;
    for (int64_t i = beg + lane; i < end; i += 32) {
      const int id = a.det_ids[i];
      const double X[3] = {bp[3 * id], bp[3 * id + 1], bp[3 * id + 2]};
      double Xc[3];
      mat3_vec(R, X, Xc);
      Xc[0] += t[0]; Xc[1] += t[1]; Xc[2] += t[2];
      const double iz = Xc[2] != 0.0 ? 1.0 / Xc[2] : 1.0;
      const double xn = Xc[0] * iz, yn = Xc[1] * iz;
      const double2 q = a.und[i];
      const double ru = fx * xn + cx - q.x, rv = fy * yn + cy - q.y;
      const double Ju[3] = {fx * iz, 0.0, -fx * xn * iz}, Jv[3] = {0.0, fy * iz, -fy * yn * iz};
      const double gu[6] = {Xc[1] * Ju[2] - Xc[2] * Ju[1], Xc[2] * Ju[0] - Xc[0] * Ju[2], Xc[0] * Ju[1] - Xc[1] * Ju[0], Ju[0], Ju[1], Ju[2]};
      const double gv[6] = {Xc[1] * Jv[2] - Xc[2] * Jv[1], Xc[2] * Jv[0] - Xc[0] * Jv[2], Xc[0] * Jv[1] - Xc[1] * Jv[0], Jv[0], Jv[1], Jv[2]};
      int e = 0;
#pragma unroll
      for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int s = r; s < 6; s++) { N[e] += gu[r] * gu[s] + gv[r] * gv[s]; e++; }
        N[21 + r] += gu[r] * ru + gv[r] * rv;
      }
    }
#pragma unroll
    for (int i = 0; i < 27; i++) N[i] = warp_sum(N[i]);
    bool improved = false, stop = false;
    for (int tries = 0; tries < 12 && !improved && !stop; tries++) {
      double A[36], d[6];
      {
        int e = 0;
#pragma unroll
        for (int r = 0; r < 6; r++) {
#pragma unroll
          for (int s = r; s < 6; s++) { A[r * 6 + s] = N[e]; A[s * 6 + r] = N[e]; e++; }
          d[r] = -N[21 + r];
        }
      }
#pragma unroll
      for (int r = 0; r < 6; r++) A[r * 7] += lambda * fmax(A[r * 7], 1e-30);
      if (!spd_solve<6>(A, d)) { lambda *= 10.0; continue; }
      double Rn[9], tn[3];
      apply_twist(d, R, t, Rn, tn);
      const double cn = cost_of(Rn, tn);
      const double step2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3] + d[4] * d[4] + d[5] * d[5];
      if (cn <= cost) {
        const bool tiny = (cost - cn) <= 1e-15 * cost || step2 < 1e-26;
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = Rn[i];
        t[0] = tn[0]; t[1] = tn[1]; t[2] = tn[2];
        cost = cn;
        lambda = fmax(lambda * 0.1, 1e-12);
        improved = true;
        if (tiny) stop = true;
      } else {
        if (step2 < 1e-26) stop = true;
        lambda *= 10.0;
      }
    }
    if (stop || !improved) break;
  }
  if (!(cost == cost) || !(t[2] == t[2])) { invalid(); return; }       // NaN: solvePnP would have failed
  if (lane < 12) { const int r = lane / 4, cc = lane % 4; Tout[lane] = cc < 3 ? R[3 * r + cc] : t[r]; }
  else if (lane < 16) Tout[lane] = lane == 15 ? 1.0 : 0.0;
  if (lane == 0) { a.err[w] = sqrt(cost / (2.0 * n)); a.npts[w] = n; a.valid[w] = 1; }
}

}  // namespace mcba
