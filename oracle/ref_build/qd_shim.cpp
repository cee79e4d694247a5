// Build recipe helper (test infrastructure only, never shipped): wraps the reference's
// own symbolic process-noise function (reference: src/x/ekf/propagator.cpp:207-840,
// Propagator::discreteProcessNoiseCov) so that it can be compiled WHERE IT LIES, without
// Eigen, into oracle/_ref/libxref_qd.so.  The function body is pure scalar arithmetic; the
// only Eigen surface it touches is q.w()/x()/y()/z(), e_w(i), e_a(i) and a 15x15
// CoreCovMatrix with Zero() and operator()(r,c) -- reproduced by the three PODs below.
// build_ref.sh extracts the body into oracle/_ref/qd_body.inc (git-ignored) at build time;
// no reference source is copied into this repository.
#include <cstring>
namespace {
enum { kIdxP = 0, kIdxV = 3, kIdxQ = 6, kIdxBw = 9, kIdxBa = 12 };
struct Quaternion { double w_, x_, y_, z_;
  double w() const { return w_; } double x() const { return x_; }
  double y() const { return y_; } double z() const { return z_; } };
struct Vector3 { double v[3]; double operator()(int i) const { return v[i]; } };
struct CoreCovMatrix { double m[15][15];
  static CoreCovMatrix Zero() { CoreCovMatrix c; std::memset(c.m, 0, sizeof(c.m)); return c; }
  double& operator()(int r, int c) { return m[r][c]; } };
struct Propagator {
  CoreCovMatrix discreteProcessNoiseCov(const double dt, const Quaternion &q, const Vector3 &e_w,
      const Vector3 &e_a, const double n_w, const double n_bw, const double n_a,
      const double n_ba) const;
};
#include "qd_body.inc"
}  // namespace
// q is (w,x,y,z); out is row-major 15x15.
extern "C" void xref_qd(double dt, const double q[4], const double e_w[3], const double e_a[3],
                        double n_w, double n_bw, double n_a, double n_ba, double out[225]) {
  Quaternion qq{q[0], q[1], q[2], q[3]};
  Vector3 w{{e_w[0], e_w[1], e_w[2]}}, a{{e_a[0], e_a[1], e_a[2]}};
  Propagator p;
  CoreCovMatrix c = p.discreteProcessNoiseCov(dt, qq, w, a, n_w, n_bw, n_a, n_ba);
  std::memcpy(out, c.m, sizeof(c.m));
}
