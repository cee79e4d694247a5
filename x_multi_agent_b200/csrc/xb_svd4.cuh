// 4x4 one-sided Jacobi SVD shared by the single-agent (k_tracks.cu) and the joint multi-agent (k_multi_msckf.cu)
// two-view DLT.  reference call site: src/x/vision/triangulation.cpp:81-100 (cv::triangulatePoints).
#pragma once
#include "xb_common.cuh"

namespace xb {

// ---- 4x4 one-sided Jacobi: right singular vector of the smallest singular value -----------------
// cv::triangulatePoints (OpenCV calib3d/triangulate.cpp) solves the same 4x4 homogeneous system by SVD.
static __device__ void smallest_right_singular_vector4(double* A /*4x4 row-major, destroyed*/, double* v /*4*/) {
  double V[16];
  for (int i = 0; i < 16; ++i) V[i] = (i % 5 == 0) ? 1.0 : 0.0;
#pragma unroll 1
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {  // static indices keep A and V in registers
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int i = 0; i < 4; ++i) {
          al += A[i * 4 + p] * A[i * 4 + p];
          be += A[i * 4 + q] * A[i * 4 + q];
          ga += A[i * 4 + p] * A[i * 4 + q];
        }
        if (ga * ga <= 1e-28 * (al * be) || ga == 0.0) continue;  // |ga| <= 1e-14 sqrt(al be)
        off = 1.0;
        // tan of the rotation angle: t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (be - al) / (2 ga), written with one
        // square root and one division:  t = 2 ga / (d + sign(d) sqrt(d^2 + 4 ga^2)),  d = be - al  (sign(0) = +1)
        const double d = be - al, g2 = 2.0 * ga;
        const double tt = g2 / (d + copysign(sqrt(fma(d, d, g2 * g2)), d));
        const double c = rsqrt(fma(tt, tt, 1.0)), s = c * tt;
        for (int i = 0; i < 4; ++i) {
          const double ap = A[i * 4 + p], aq = A[i * 4 + q];
          A[i * 4 + p] = c * ap - s * aq;
          A[i * 4 + q] = s * ap + c * aq;
          const double vp = V[i * 4 + p], vq = V[i * 4 + q];
          V[i * 4 + p] = c * vp - s * vq;
          V[i * 4 + q] = s * vp + c * vq;
        }
      }
    if (off == 0.0) break;  // no rotation was needed in this sweep
  }
  int best = 0;
  double bn = 1e300;
  for (int p = 0; p < 4; ++p) {
    double n = 0.0;
    for (int i = 0; i < 4; ++i) n += A[i * 4 + p] * A[i * 4 + p];
    if (n < bn) { bn = n; best = p; }
  }
  for (int i = 0; i < 4; ++i) v[i] = V[i * 4 + best];
}

}  // namespace xb
