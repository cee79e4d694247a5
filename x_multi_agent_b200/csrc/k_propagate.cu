// IMU propagation of state and covariance over a chain of ring-buffer slots.
// reference: src/x/ekf/propagator.cpp:30-205 (propagateState, quaternionIntegrator,
// discreteStateTransition, propagateCovarianceMatrices), :207-840 (Q_d, see qd_poly.cuh).
//
// Device layout: every slot owns its estimates (xvec) and the 15 x N strip [P_ii | P_iv]; the
// (N-15)^2 block P_vv is NOT copied per IMU step (the reference does, propagator.cpp:204): slots
// propagated from one another share a P_vv generation.  P_vi is the transpose of
// P_iv (the reference computes both with the same products, propagator.cpp:195-203).
#include "xb_kernels.h"
#include "qd_poly.cuh"

namespace xb {

// reference: propagator.cpp:74-98
__device__ void quat_integrator(const double* w0, const double* w1, double dt, double* D /*4x4*/) {
  auto omega = [](const double* v, double* O) {  // eigen_matrix_base_plugin.h:43-52
    const double x = v[0], y = v[1], z = v[2];
    O[0] = 0.0; O[1] = z; O[2] = -y; O[3] = x;
    O[4] = -z; O[5] = 0.0; O[6] = x; O[7] = y;
    O[8] = y; O[9] = -x; O[10] = 0.0; O[11] = z;
    O[12] = -x; O[13] = -y; O[14] = -z; O[15] = 0.0;
  };
  double O1[16], O0[16], Om[16], A[16], Ak[16], Tm[16];
  omega(w1, O1);
  omega(w0, O0);
  const double wm[3] = {(w1[0] + w0[0]) / 2.0, (w1[1] + w0[1]) / 2.0, (w1[2] + w0[2]) / 2.0};
  omega(wm, Om);
  for (int i = 0; i < 16; ++i) { A[i] = Om[i] * 0.5 * dt; Ak[i] = A[i]; D[i] = (i % 5 == 0) ? 1.0 : 0.0; }
  int fac = 1;
  for (int k = 1; k < 5; ++k) {
    fac *= k;
    for (int i = 0; i < 16; ++i) D[i] = D[i] + Ak[i] / fac;
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        double s = 0.0;
        for (int e = 0; e < 4; ++e) s += Ak[r * 4 + e] * A[e * 4 + c];
        Tm[r * 4 + c] = s;
      }
    for (int i = 0; i < 16; ++i) Ak[i] = Tm[i];
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      double s10 = 0.0, s01 = 0.0;
      for (int e = 0; e < 4; ++e) {
        s10 += O1[r * 4 + e] * O0[e * 4 + c];
        s01 += O0[r * 4 + e] * O1[e * 4 + c];
      }
      D[r * 4 + c] += 1.0 / 48.0 * (s10 - s01) * dt * dt;
    }
}

// reference: propagator.cpp:100-164.  F row-major 15x15.
__device__ void state_transition(double dt, const double* w, const double* a, const double* q, double* F) {
  double wx[9], ax[9], C[9], Ca[9], ww[9];
  xb_skew(w, wx);
  xb_skew(a, ax);
  xb_rot_raw(q, C);
  const double dt2 = dt * dt * 0.5, dt3 = dt2 * dt / 3.0, dt4 = dt3 * dt * 0.25, dt5 = dt4 * dt * 0.2;
  xb_mm33(C, ax, Ca);
  xb_mm33(wx, wx, ww);
  double m1[9], m2[9], A[9], B[9], E[9], Fm[9], Cm[9];
  for (int i = 0; i < 9; ++i) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    m1[i] = -dt2 * I + dt3 * wx[i] - dt4 * ww[i];
    m2[i] = dt3 * I - dt4 * wx[i] + dt5 * ww[i];
    E[i] = I - dt * wx[i] + dt2 * ww[i];
    Fm[i] = -dt * I + dt2 * wx[i] - dt3 * ww[i];
  }
  xb_mm33(Ca, m1, A);
  xb_mm33(Ca, m2, B);
  xb_mm33(Ca, Fm, Cm);
  for (int i = 0; i < 225; ++i) F[i] = (i % 16 == 0) ? 1.0 : 0.0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const int e = r * 3 + c;
      F[(0 + r) * 15 + 3 + c] = (r == c) ? dt : 0.0;
      F[(0 + r) * 15 + 6 + c] = A[e];
      F[(0 + r) * 15 + 9 + c] = B[e];
      F[(0 + r) * 15 + 12 + c] = -C[e] * dt2;
      F[(3 + r) * 15 + 6 + c] = Cm[e];
      F[(3 + r) * 15 + 9 + c] = -A[e];
      F[(3 + r) * 15 + 12 + c] = -C[e] * dt;
      F[(6 + r) * 15 + 6 + c] = E[e];
      F[(6 + r) * 15 + 9 + c] = Fm[e];
    }
}

// One launch propagates the estimates of `n_steps` consecutive slots starting after `start`
// and emits F_d / Q_d of every step into FQ (n_steps x 450).
__global__ void __launch_bounds__(128) k_prop_means(double* __restrict__ xv, int LX, int NS, int start, int n_steps,
                                                    ImuSample in, PropParams pp, double* __restrict__ FQ) {
  const int t = threadIdx.x;
  const double* x0 = xv + (size_t)start * LX;
  // the new IMU sample of a processImu call goes into the last slot of the chain (State::setImu, state.cpp:145-151)
  if (in.valid && t == 0) {
    double* xl = xv + (size_t)((start + n_steps) % NS) * LX;
    xl[XV_TIME] = in.t;
    xl[XV_SEQ] = in.seq;
    for (int e = 0; e < 3; ++e) { xl[XV_WM + e] = in.w[e]; xl[XV_AM + e] = in.a[e]; }
  }
  // State::setStaticStatesFrom (state.cpp:153-161): biases, extrinsics, window and feature arrays
  for (int k = 1; k <= n_steps; ++k) {
    double* x1 = xv + (size_t)((start + k) % NS) * LX;
    for (int e = XV_BW + t; e < XV_WM; e += blockDim.x) x1[e] = x0[e];
    for (int e = XV_ARR + t; e < LX; e += blockDim.x) x1[e] = x0[e];
  }
  __syncthreads();
  // the quaternion-integrator matrices depend on the IMU samples and the (static) gyro bias only: one thread per step
  __shared__ double Dm[128][16];
  if (t < n_steps) {
    const double* s0 = xv + (size_t)((start + t) % NS) * LX;
    const double* s1 = xv + (size_t)((start + t + 1) % NS) * LX;
    double w1[3], w0[3];
    for (int e = 0; e < 3; ++e) {  // State::computeUnbiasedImuMeasurements, state.cpp:177-182
      w1[e] = s1[XV_WM + e] - s1[XV_BW + e];
      w0[e] = s0[XV_WM + e] - s0[XV_BW + e];
    }
    quat_integrator(w0, w1, s1[XV_TIME] - s0[XV_TIME], Dm[t]);
  }
  __syncthreads();
  if (t == 0) {  // the short sequential chain: q, v, p
    for (int k = 1; k <= n_steps; ++k) {
      const double* s0 = xv + (size_t)((start + k - 1) % NS) * LX;
      double* s1 = xv + (size_t)((start + k) % NS) * LX;
      double a1[3], a0[3];
      for (int e = 0; e < 3; ++e) {
        a1[e] = s1[XV_AM + e] - s1[XV_BA + e];
        a0[e] = s0[XV_AM + e] - s0[XV_BA + e];
      }
      const double dt = s1[XV_TIME] - s0[XV_TIME];
      const double* D = Dm[k - 1];
      double q1[4];
      for (int r = 0; r < 4; ++r)
        q1[r] = D[r * 4] * s0[XV_Q] + D[r * 4 + 1] * s0[XV_Q + 1] + D[r * 4 + 2] * s0[XV_Q + 2] + D[r * 4 + 3] * s0[XV_Q + 3];
      xb_qnormalize(q1);
      double R1[9], R0[9], ra1[3], ra0[3];
      xb_rot_raw(q1, R1);
      xb_rot_raw(&s0[XV_Q], R0);
      xb_mv33(R1, a1, ra1);
      xb_mv33(R0, a0, ra0);
      for (int e = 0; e < 3; ++e) {
        const double dv = (ra1[e] + ra0[e]) / 2.0;
        const double v1 = s0[XV_V + e] + (dv + pp.g[e]) * dt;
        s1[XV_V + e] = v1;
        s1[XV_P + e] = s0[XV_P + e] + (v1 + s0[XV_V + e]) / 2.0 * dt;
      }
      for (int e = 0; e < 4; ++e) s1[XV_Q + e] = q1[e];
    }
  }
  __syncthreads();
  if (t < n_steps) {
    const double* s0 = xv + (size_t)((start + t) % NS) * LX;
    const double* s1 = xv + (size_t)((start + t + 1) % NS) * LX;
    double w1[3], a1[3];
    for (int e = 0; e < 3; ++e) {
      w1[e] = s1[XV_WM + e] - s1[XV_BW + e];
      a1[e] = s1[XV_AM + e] - s1[XV_BA + e];
    }
    const double dt = s1[XV_TIME] - s0[XV_TIME];
    double* F = FQ + (size_t)t * 450;
    double* Q = F + 225;
    state_transition(dt, w1, a1, &s1[XV_Q], F);
    double C[9];
    xb_rot_raw(&s1[XV_Q], C);
    for (int e = 0; e < 225; ++e) Q[e] = 0.0;
    xb_qd_poly(dt, C, w1, a1, pp.n_w, pp.n_bw, pp.n_a, pp.n_ba, Q);
  }
}

// Strip propagation: for every step k,  strip_k = [F P_ii F^T + Q | F P_iv]  (propagator.cpp:195-203).
// second != 0: the same recurrence on the column strips P_vi^T (P_vi' = P_vi F^T, propagator.cpp:203), whose core block
// is the transpose of P_ii and therefore takes Q_d^T.
__global__ void __launch_bounds__(128) k_prop_strips(double* __restrict__ strip, int N, int NS, int start, int n_steps,
                                                     const double* __restrict__ FQ, int second) {
  __shared__ double Fs[225], Qs[225], Pii[225], Tm[225];
  const int t = threadIdx.x;
  const int j = blockIdx.x * blockDim.x + t;  // column
  const bool core_block = blockIdx.x == 0;
  const size_t SS = (size_t)15 * N;
  const double* s0 = strip + (size_t)start * SS;
  double v[15];
  if (j >= XB_CORE && j < N)
    for (int r = 0; r < 15; ++r) v[r] = s0[(size_t)r * N + j];
  if (core_block)
    for (int e = t; e < 225; e += blockDim.x) Pii[e] = s0[(size_t)(e / 15) * N + (e % 15)];
  for (int k = 0; k < n_steps; ++k) {
    __syncthreads();
    for (int e = t; e < 225; e += blockDim.x) {
      Fs[e] = FQ[(size_t)k * 450 + e];
      Qs[e] = FQ[(size_t)k * 450 + 225 + (second ? (e % 15) * 15 + e / 15 : e)];
    }
    __syncthreads();
    double* s1 = strip + (size_t)((start + k + 1) % NS) * SS;
    if (j >= XB_CORE && j < N) {
      double u[15];
#pragma unroll
      for (int r = 0; r < 15; ++r) {
        double s = 0.0;
#pragma unroll
        for (int e = 0; e < 15; ++e) s = fma(Fs[r * 15 + e], v[e], s);
        u[r] = s;
      }
#pragma unroll
      for (int r = 0; r < 15; ++r) { v[r] = u[r]; s1[(size_t)r * N + j] = u[r]; }
    }
    if (core_block) {
      for (int e = t; e < 225; e += blockDim.x) {  // Tm = F * P_ii
        const int r = e / 15, c = e % 15;
        double s = 0.0;
        for (int a = 0; a < 15; ++a) s = fma(Fs[r * 15 + a], Pii[a * 15 + c], s);
        Tm[e] = s;
      }
      __syncthreads();
      for (int e = t; e < 225; e += blockDim.x) {  // P_ii' = Tm * F^T + Q
        const int r = e / 15, c = e % 15;
        double s = 0.0;
        for (int a = 0; a < 15; ++a) s = fma(Tm[r * 15 + a], Fs[c * 15 + a], s);
        s += Qs[e];
        s1[(size_t)r * N + c] = s;
        Pii[e] = s;  // each thread rewrites only the entries it owns; Tm already holds F*P_ii
      }
    }
  }
}

void launch_prop_means(cudaStream_t s, double* xv, int LX, int NS, int start, int n_steps, const ImuSample& in,
                       const PropParams& pp, double* FQ) {
  if (n_steps <= 0) return;
  k_prop_means<<<1, 128, 0, s>>>(xv, LX, NS, start, n_steps, in, pp, FQ);
  count_launch();
}
void launch_prop_strips(cudaStream_t s, double* strip, int N, int NS, int start, int n_steps, const double* FQ, int second) {
  if (n_steps <= 0) return;
  k_prop_strips<<<(N + 127) / 128, 128, 0, s>>>(strip, N, NS, start, n_steps, FQ, second);
  count_launch();
}
void launch_propagate(cudaStream_t s, double* xv, int LX, double* strip, int N, int NS, int start, int n_steps,
                      const ImuSample& in, const PropParams& pp, double* FQ) {
  launch_prop_means(s, xv, LX, NS, start, n_steps, in, pp, FQ);
  launch_prop_strips(s, strip, N, NS, start, n_steps, FQ);
}

}  // namespace xb
