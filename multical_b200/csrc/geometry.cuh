// geometry.cuh — SE(3) pose chain and camera models, device side (fp64).
//
// Replaces, per corner, what the reference evaluates densely on the CPU:
//   pose chain   x_cam = T_cam[c] T_frame[f] T_board[b] X   motion/static_frames.py:16-25, tables.py:284-304,400-405
//   pinhole      cv2.projectPoints(rvec=0,tvec=0,K,dist)    camera.py:124-128   (5/8/12 coefficients, camera.py:43-48)
//   fisheye      cv2.fisheye.projectPoints(...)             camera_fisheye.py:113-117
// plus the analytic derivatives that scipy obtains by 2-point finite differences (calibration.py:209-210).
// The math here is __host__ __device__ so that tests/host_math can run the very same source on the CPU against cv2 / scipy
// (analytic Jacobians vs the Jacobian cv2.projectPoints returns); the product only ever calls it from kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace mcba {

enum { MODEL_STANDARD = 0, MODEL_RATIONAL = 1, MODEL_THIN_PRISM = 2, MODEL_FISHEYE = 3, MODEL_TILTED = 4 };

__host__ __device__ constexpr int model_nd(int model) {
  return model == MODEL_STANDARD ? 5 : model == MODEL_RATIONAL ? 8 : model == MODEL_THIN_PRISM ? 12 : model == MODEL_TILTED ? 14 : 4;
}
// local Jacobian layout of one residual row: [omega(3) v(3) fx fy cx cy dist(nd)]  (skew has no effect:
// cv2 ignores K[0,1]) -> D = 10 + nd
__host__ __device__ constexpr int model_D(int model) { return 10 + model_nd(model); }

// per-pose table entry built once per parameter vector: R, t and the SO(3) left Jacobian of the rotvec
struct PoseT {
  double R[9];
  double t[3];
  double JL[9];
  double pad[3];
};

__host__ __device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__host__ __device__ __forceinline__ void mat3_vec(const double* A, const double* x, double* y) {
#pragma unroll
  for (int i = 0; i < 3; i++) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}

// rotvec -> R (transform/rtvec.py:24-27, scipy Rotation.from_rotvec) and left Jacobian JL with
//   R(r + dr) ~= exp([JL dr]x) R(r)
__host__ __device__ inline void rodrigues(const double* r, double* R, double* JL) {
  const double x = r[0], y = r[1], z = r[2];
  const double th2 = x * x + y * y + z * z;
  double A, B, Cc;            // sin(th)/th, (1-cos th)/th^2, (th - sin th)/th^3
  if (th2 < 1e-6) {
    const double th4 = th2 * th2;
    A = 1.0 - th2 / 6.0 + th4 / 120.0;
    B = 0.5 - th2 / 24.0 + th4 / 720.0;
    Cc = 1.0 / 6.0 - th2 / 120.0 + th4 / 5040.0;
  } else {
    const double th = sqrt(th2);
    double s, c;
    sincos(th, &s, &c);
    A = s / th;
    B = (1.0 - c) / th2;
    Cc = (th - s) / (th2 * th);
  }
  // K = [r]x ; K^2 = r r^T - th2 I
  const double K[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  const double K2[9] = {x * x - th2, x * y, x * z, x * y, y * y - th2, y * z, x * z, y * z, z * z - th2};
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    R[i] = I + A * K[i] + B * K2[i];
    JL[i] = I + B * K[i] + Cc * K2[i];
  }
}

// 4x4 (row-major, last row ignored) -> rtvec [rx ry rz tx ty tz] with the rotation vector in canonical form (angle in [0, pi]):
// the reference's transform/rtvec.py:29-32 (scipy Rotation.from_matrix(...).as_rotvec()), restated with the same steps --
// largest-pivot quaternion extraction, w >= 0, angle = 2 atan2(|xyz|, w), series below 1e-3 rad.
__host__ __device__ inline void matrix_to_rtvec(const double* T, double* rt) {
  const double m00 = T[0], m01 = T[1], m02 = T[2], m10 = T[4], m11 = T[5], m12 = T[6], m20 = T[8], m21 = T[9], m22 = T[10];
  const double tr = m00 + m11 + m22;
  double q[4];
  int choice = 3; double best = tr;
  if (m00 > best) { best = m00; choice = 0; }
  if (m11 > best) { best = m11; choice = 1; }
  if (m22 > best) { best = m22; choice = 2; }
  if (choice == 0)      { q[0] = 1.0 - tr + 2.0 * m00; q[1] = m10 + m01; q[2] = m20 + m02; q[3] = m21 - m12; }
  else if (choice == 1) { q[1] = 1.0 - tr + 2.0 * m11; q[2] = m21 + m12; q[0] = m01 + m10; q[3] = m02 - m20; }
  else if (choice == 2) { q[2] = 1.0 - tr + 2.0 * m22; q[0] = m02 + m20; q[1] = m12 + m21; q[3] = m10 - m01; }
  else                  { q[0] = m21 - m12; q[1] = m02 - m20; q[2] = m10 - m01; q[3] = 1.0 + tr; }
  const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double s = (q[3] < 0.0 ? -1.0 : 1.0) / nq;
  const double x = q[0] * s, y = q[1] * s, z = q[2] * s, w = q[3] * s;
  const double nv = sqrt(x * x + y * y + z * z);
  const double angle = 2.0 * atan2(nv, w);
  double scale;
  if (angle <= 1e-3) { const double a2 = angle * angle; scale = 2.0 + a2 / 12.0 + 7.0 * a2 * a2 / 2880.0; }
  else scale = angle / sin(0.5 * angle);
  rt[0] = scale * x; rt[1] = scale * y; rt[2] = scale * z;
  rt[3] = T[3]; rt[4] = T[7]; rt[5] = T[11];
}

// 6x6 map from a pose-parameter increment (dr, dt) to the camera-frame twist (omega, v) it induces on
// x_cam.  With Rl, tl = rotation / translation of everything LEFT of the perturbed pose in the chain
// (identity for the camera pose) and tcur = translation of the chain up to and including this pose:
//   omega = Rl JL dr ,   v = [tcur]x Rl JL dr + Rl dt
__host__ __device__ __forceinline__ void twist_map(const double* Rl, const double* JL, const double* tcur, double* A /*6x6 row-major*/) {
  double RJ[9];
  mat3_mul(Rl, JL, RJ);
  const double tx = tcur[0], ty = tcur[1], tz = tcur[2];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double a0 = RJ[j], a1 = RJ[3 + j], a2 = RJ[6 + j];
    A[0 * 6 + j] = a0; A[1 * 6 + j] = a1; A[2 * 6 + j] = a2;
    A[3 * 6 + j] = ty * a2 - tz * a1;       // (t x a)
    A[4 * 6 + j] = tz * a0 - tx * a2;
    A[5 * 6 + j] = tx * a1 - ty * a0;
    A[0 * 6 + 3 + j] = 0; A[1 * 6 + 3 + j] = 0; A[2 * 6 + 3 + j] = 0;
    A[3 * 6 + 3 + j] = Rl[j]; A[4 * 6 + 3 + j] = Rl[3 + j]; A[5 * 6 + 3 + j] = Rl[6 + j];
  }
}

// Tilted-sensor matrix of the 14-coefficient model and its derivatives wrt (tauX, tauY): restatement of OpenCV's
// cv::detail::computeTiltProjectionMatrix (the reference reaches it through cv2.projectPoints, camera.py:43-48,124-128).
//   matTilt = P_z(R_y R_x) R_y R_x ,  P_z(R) = [[R22, 0, -R02], [0, R22, -R12], [0, 0, 1]]
__host__ __device__ inline void tilt_matrices(double tx, double ty, double* M, double* dMx, double* dMy) {
  double sx, cx, sy, cy;
  sincos(tx, &sx, &cx);
  sincos(ty, &sy, &cy);
  const double Rx[9] = {1, 0, 0, 0, cx, sx, 0, -sx, cx};
  const double Ry[9] = {cy, 0, -sy, 0, 1, 0, sy, 0, cy};
  const double dRx[9] = {0, 0, 0, 0, -sx, cx, 0, -cx, -sx};
  const double dRy[9] = {-sy, 0, -cy, 0, 0, 0, cy, 0, -sy};
  double R[9], dRa[9], dRb[9];
  mat3_mul(Ry, Rx, R);
  mat3_mul(Ry, dRx, dRa);        // d(RyRx)/dtauX
  mat3_mul(dRy, Rx, dRb);        // d(RyRx)/dtauY
  const double Pz[9] = {R[8], 0, -R[2], 0, R[8], -R[5], 0, 0, 1};
  const double dPa[9] = {dRa[8], 0, -dRa[2], 0, dRa[8], -dRa[5], 0, 0, 0};
  const double dPb[9] = {dRb[8], 0, -dRb[2], 0, dRb[8], -dRb[5], 0, 0, 0};
  double t1[9], t2[9];
  mat3_mul(Pz, R, M);
  mat3_mul(Pz, dRa, t1); mat3_mul(dPa, R, t2);
#pragma unroll
  for (int i = 0; i < 9; i++) dMx[i] = t1[i] + t2[i];
  mat3_mul(Pz, dRb, t1); mat3_mul(dPb, R, t2);
#pragma unroll
  for (int i = 0; i < 9; i++) dMy[i] = t1[i] + t2[i];
}

// ---------------------------------------------------------------------------------------------
// Camera models.  k points at [fx fy cx cy skew dist...] (camera.py:144-155).  Outputs u,v; when JAC:
//   Ju,Jv = d(u,v)/dX_cam (3 each);  ku,kv = d(u,v)/d[fx fy cx cy dist...] (4+nd each; fy,cy of ku and
//   fx,cx of kv are structurally zero and not written)
template <int MODEL, bool JAC>
__host__ __device__ __forceinline__ void project(const double* X, const double* __restrict__ k, double& u, double& v,
                                        double* Ju, double* Jv, double* ku, double* kv) {
  constexpr int ND = model_nd(MODEL);
  const double fx = k[0], fy = k[1], cx = k[2], cy = k[3];
  const double* d = k + 5;
  if constexpr (MODEL == MODEL_FISHEYE) {
    const double iz = 1.0 / X[2];
    const double a = X[0] * iz, b = X[1] * iz;
    const double r2 = a * a + b * b;
    const double r = sqrt(r2);
    const double th = atan(r);
    const double t2 = th * th, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
    const double thd = th * (1.0 + d[0] * t2 + d[1] * t4 + d[2] * t6 + d[3] * t8);
    const bool big = r > 1e-8;
    const double inv_r = big ? 1.0 / r : 1.0;
    const double cd = big ? thd * inv_r : 1.0;
    const double xd = a * cd, yd = b * cd;
    u = fx * xd + cx;
    v = fy * yd + cy;
    if constexpr (JAC) {
      const double dthd = 1.0 + 3.0 * d[0] * t2 + 5.0 * d[1] * t4 + 7.0 * d[2] * t6 + 9.0 * d[3] * t8;
      const double dcd_r = big ? (dthd / (1.0 + r2) - cd) * inv_r * inv_r : 0.0;   // (d cd/dr)/r
      const double xa = cd + a * a * dcd_r, xb = a * b * dcd_r, yb = cd + b * b * dcd_r;
      Ju[0] = fx * xa * iz; Ju[1] = fx * xb * iz; Ju[2] = -fx * (xa * a + xb * b) * iz;
      Jv[0] = fy * xb * iz; Jv[1] = fy * yb * iz; Jv[2] = -fy * (xb * a + yb * b) * iz;
      ku[0] = xd; ku[2] = 1.0;
      kv[1] = yd; kv[3] = 1.0;
      const double air = big ? a * inv_r : 0.0, bir = big ? b * inv_r : 0.0;
      const double t3 = th * t2;
      ku[4] = fx * air * t3; ku[5] = fx * air * t3 * t2; ku[6] = fx * air * t3 * t4; ku[7] = fx * air * t3 * t6;
      kv[4] = fy * bir * t3; kv[5] = fy * bir * t3 * t2; kv[6] = fy * bir * t3 * t4; kv[7] = fy * bir * t3 * t6;
    }
  } else {
    const double Z = X[2];
    const double iz = (Z != 0.0) ? 1.0 / Z : 1.0;      // cv2.projectPoints: z ? 1/z : 1
    const double x = X[0] * iz, y = X[1] * iz;
    const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    const double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
    const double cdist = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
    double icd2 = 1.0;
    if constexpr (ND >= 8) icd2 = 1.0 / (1.0 + d[5] * r2 + d[6] * r4 + d[7] * r6);
    const double a1 = 2.0 * x * y, a2 = r2 + 2.0 * x * x, a3 = r2 + 2.0 * y * y;
    const double rad = cdist * icd2;
    double xd = x * rad + p1 * a1 + p2 * a2;
    double yd = y * rad + p1 * a3 + p2 * a1;
    if constexpr (ND >= 12) {
      xd += d[8] * r2 + d[9] * r4;
      yd += d[10] * r2 + d[11] * r4;
    }
    if constexpr (ND < 14) {
      u = fx * xd + cx;
      v = fy * yd + cy;
    }
    // derivatives of the distorted normalised point (xd, yd) wrt (x, y) and wrt the distortion coefficients
    double xx = 0, xy = 0, yx = 0, yy = 0;
    double dxk[ND], dyk[ND];
    if constexpr (JAC) {
      const double dcd = k1 + 2.0 * k2 * r2 + 3.0 * k3 * r4;
      double drad = dcd;
      if constexpr (ND >= 8) drad = dcd * icd2 - cdist * icd2 * icd2 * (d[5] + 2.0 * d[6] * r2 + 3.0 * d[7] * r4);
      double tpx = 0.0, tpy = 0.0;
      if constexpr (ND >= 12) { tpx = d[8] + 2.0 * d[9] * r2; tpy = d[10] + 2.0 * d[11] * r2; }
      const double cross = a1 * drad + 2.0 * p1 * x + 2.0 * p2 * y;
      xx = rad + 2.0 * x * x * drad + 2.0 * p1 * y + 6.0 * p2 * x + 2.0 * x * tpx;
      xy = cross + 2.0 * y * tpx;
      yx = cross + 2.0 * x * tpy;
      yy = rad + 2.0 * y * y * drad + 6.0 * p1 * y + 2.0 * p2 * x + 2.0 * y * tpy;
      const double xi = x * icd2, yi = y * icd2;
      dxk[0] = xi * r2; dxk[1] = xi * r4; dxk[2] = a1; dxk[3] = a2; dxk[4] = xi * r6;
      dyk[0] = yi * r2; dyk[1] = yi * r4; dyk[2] = a3; dyk[3] = a1; dyk[4] = yi * r6;
      if constexpr (ND >= 8) {
        const double gx = -xi * rad, gy = -yi * rad;     // -x cdist icd2^2
        dxk[5] = gx * r2; dxk[6] = gx * r4; dxk[7] = gx * r6;
        dyk[5] = gy * r2; dyk[6] = gy * r4; dyk[7] = gy * r6;
      }
      if constexpr (ND >= 12) {
        dxk[8] = r2; dxk[9] = r4; dxk[10] = 0.0; dxk[11] = 0.0;
        dyk[8] = 0.0; dyk[9] = 0.0; dyk[10] = r2; dyk[11] = r4;
      }
    }
    if constexpr (ND < 14) {
      if constexpr (JAC) {
        Ju[0] = fx * xx * iz; Ju[1] = fx * xy * iz; Ju[2] = -fx * (xx * x + xy * y) * iz;
        Jv[0] = fy * yx * iz; Jv[1] = fy * yy * iz; Jv[2] = -fy * (yx * x + yy * y) * iz;
        ku[0] = xd; ku[2] = 1.0;
        kv[1] = yd; kv[3] = 1.0;
#pragma unroll
        for (int i = 0; i < ND; i++) { ku[4 + i] = fx * dxk[i]; kv[4 + i] = fy * dyk[i]; }
      }
    } else {
      // tilted sensor: (xt, yt) = perspective division of matTilt (xd, yd, 1)
      double M[9], dMx[9], dMy[9];
      tilt_matrices(d[12], d[13], M, dMx, dMy);
      const double v0 = M[0] * xd + M[1] * yd + M[2], v1 = M[3] * xd + M[4] * yd + M[5], v2 = M[6] * xd + M[7] * yd + M[8];
      const double ip = (v2 != 0.0) ? 1.0 / v2 : 1.0;
      const double xt = v0 * ip, yt = v1 * ip;
      u = fx * xt + cx;
      v = fy * yt + cy;
      if constexpr (JAC) {
        const double T00 = (M[0] - M[6] * xt) * ip, T01 = (M[1] - M[7] * xt) * ip;
        const double T10 = (M[3] - M[6] * yt) * ip, T11 = (M[4] - M[7] * yt) * ip;
        const double ux = T00 * xx + T01 * yx, uy = T00 * xy + T01 * yy;      // d xt / d(x, y)
        const double vx = T10 * xx + T11 * yx, vy = T10 * xy + T11 * yy;      // d yt / d(x, y)
        Ju[0] = fx * ux * iz; Ju[1] = fx * uy * iz; Ju[2] = -fx * (ux * x + uy * y) * iz;
        Jv[0] = fy * vx * iz; Jv[1] = fy * vy * iz; Jv[2] = -fy * (vx * x + vy * y) * iz;
        ku[0] = xt; ku[2] = 1.0;
        kv[1] = yt; kv[3] = 1.0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
          ku[4 + i] = fx * (T00 * dxk[i] + T01 * dyk[i]);
          kv[4 + i] = fy * (T10 * dxk[i] + T11 * dyk[i]);
        }
        const double ax0 = dMx[0] * xd + dMx[1] * yd + dMx[2], ax1 = dMx[3] * xd + dMx[4] * yd + dMx[5], ax2 = dMx[6] * xd + dMx[7] * yd + dMx[8];
        const double ay0 = dMy[0] * xd + dMy[1] * yd + dMy[2], ay1 = dMy[3] * xd + dMy[4] * yd + dMy[5], ay2 = dMy[6] * xd + dMy[7] * yd + dMy[8];
        ku[16] = fx * (ax0 - xt * ax2) * ip; kv[16] = fy * (ax1 - yt * ax2) * ip;
        ku[17] = fx * (ay0 - xt * ay2) * ip; kv[17] = fy * (ay1 - yt * ay2) * ip;
      }
    }
  }
}

// scipy robust losses on z = (f/f_scale)^2 (least_squares.py:183-219); returns rho0, rho1, rho2
__host__ __device__ __forceinline__ void loss_rho(int loss, double z, double& r0, double& r1, double& r2) {
  switch (loss) {
    case 1: { const double t = 1.0 + z; const double s = sqrt(t); r0 = 2.0 * (s - 1.0); r1 = 1.0 / s; r2 = -0.5 / (t * s); break; }
    case 2: if (z <= 1.0) { r0 = z; r1 = 1.0; r2 = 0.0; } else { const double s = sqrt(z); r0 = 2.0 * s - 1.0; r1 = 1.0 / s; r2 = -0.5 / (z * s); } break;
    case 3: r0 = log1p(z); r1 = 1.0 / (1.0 + z); r2 = -1.0 / ((1.0 + z) * (1.0 + z)); break;
    case 4: { const double t = 1.0 + z * z; r0 = atan(z); r1 = 1.0 / t; r2 = -2.0 * z / (t * t); break; }
    default: r0 = z; r1 = 1.0; r2 = 0.0;
  }
}

}  // namespace mcba
