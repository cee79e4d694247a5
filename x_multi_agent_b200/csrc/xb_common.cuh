// Shared host/device helpers for the xVIO hot path (fp64).  Quaternions are stored (x,y,z,w) like
// Eigen's coeffs() and the reference's q_array_ (reference: src/x/ekf/state.cpp:235-247).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#ifndef XB_HD
#ifdef __CUDACC__
#define XB_HD __host__ __device__ __forceinline__
#else
#define XB_HD inline
#endif
#endif

#define XB_CORE 15  // kSizeCoreErr, reference: include/x/common/types.h:39-47

// ---- xvec layout (one State's estimates as a flat double array) -------------------------------
// reference: include/x/ekf/state.h:240-337 (members), src/x/ekf/state.cpp:87-99,145-161
enum {
  XV_P = 0, XV_V = 3, XV_Q = 6, XV_BW = 10, XV_BA = 13, XV_QIC = 16, XV_PIC = 20,
  XV_WM = 23, XV_AM = 26, XV_TIME = 29, XV_SEQ = 30, XV_ARR = 32
};
XB_HD int xv_len(int M, int F) { return XV_ARR + 7 * M + 3 * F; }
XB_HD int xv_parr(int) { return XV_ARR; }
XB_HD int xv_qarr(int M) { return XV_ARR + 3 * M; }
XB_HD int xv_farr(int M) { return XV_ARR + 7 * M; }

// ---- small fixed-size math ---------------------------------------------------------------------
// Eigen::Quaterniond::toRotationMatrix() without normalisation (row-major 3x3 out).
XB_HD void xb_rot_raw(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}
// q.normalized().toRotationMatrix() -- the form used throughout src/x/vio/*.cpp.
XB_HD void xb_rot(const double* q, double* R) {
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double qn[4] = {q[0] / n, q[1] / n, q[2] / n, q[3] / n};
  xb_rot_raw(qn, R);
}
// Hamilton product a*b, (x,y,z,w) storage (Eigen operator*).
XB_HD void xb_qmul(const double* a, const double* b, double* o) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  o[0] = aw * bx + ax * bw + ay * bz - az * by;
  o[1] = aw * by + ay * bw + az * bx - ax * bz;
  o[2] = aw * bz + az * bw + ax * by - ay * bx;
  o[3] = aw * bw - ax * bx - ay * by - az * bz;
}
XB_HD void xb_qnormalize(double* q) {
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
// reference: src/x/ekf/state.cpp:273-283 (errorQuatFromSmallAngles: exact angle-axis).
XB_HD void xb_small_angle_quat(const double* d, double* q) {
  const double n = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (n == 0.0) { q[0] = q[1] = q[2] = 0.0; q[3] = 1.0; return; }
  const double s = sin(0.5 * n);
  q[0] = d[0] / n * s; q[1] = d[1] / n * s; q[2] = d[2] / n * s; q[3] = cos(0.5 * n);
}
// [v]x, row-major.  reference: include/x/common/eigen_matrix_base_plugin.h:32-41, include/x/vio/tools.h:57-66
XB_HD void xb_skew(const double* v, double* S) {
  S[0] = 0.0; S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2]; S[4] = 0.0; S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0]; S[8] = 0.0;
}
// C(3x3) = A(3x3) * B(3x3), row-major
XB_HD void xb_mm33(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[i * 3 + 0] * B[0 + j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
// y = A x
XB_HD void xb_mv33(const double* A, const double* x, double* y) {
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = A[i * 3] * x[0] + A[i * 3 + 1] * x[1] + A[i * 3 + 2] * x[2];
}
// y = A^T x
XB_HD void xb_mtv33(const double* A, const double* x, double* y) {
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2];
}
// C(2x3) = A(2x3) * B(3x3)
XB_HD void xb_mm23(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[i * 3 + 0] * B[0 + j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
// 3x3 inverse by cofactors (Eigen fixed-size inverse), row-major; returns determinant.
XB_HD double xb_inv33(const double* A, double* I) {
  const double c00 = A[4] * A[8] - A[5] * A[7];
  const double c01 = A[5] * A[6] - A[3] * A[8];
  const double c02 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  const double id = 1.0 / det;
  I[0] = c00 * id; I[1] = (A[2] * A[7] - A[1] * A[8]) * id; I[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  I[3] = c01 * id; I[4] = (A[0] * A[8] - A[2] * A[6]) * id; I[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  I[6] = c02 * id; I[7] = (A[1] * A[6] - A[0] * A[7]) * id; I[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  return det;
}
// The 3x3 `mat` of the inverse-depth Jacobians (reference: src/x/vio/slam_update.cpp:153-157).
XB_HD void xb_mat_ivd(double a, double b, double r, double* m) {
  m[0] = 1.0; m[1] = 0.0; m[2] = -a / r;
  m[3] = 0.0; m[4] = 1.0; m[5] = -b / r;
  m[6] = 0.0; m[7] = 0.0; m[8] = -1.0 / r;
}

#ifdef __CUDACC__
__device__ __forceinline__ double xb_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif
