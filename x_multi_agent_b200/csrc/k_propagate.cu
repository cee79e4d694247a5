// IMU propagation of state and covariance over a chain of ring-buffer slots.
// reference: src/x/ekf/propagator.cpp:30-205 (propagateState, quaternionIntegrator,
// discreteStateTransition, propagateCovarianceMatrices), :207-840 (Q_d, see qd_poly.cuh).
//
// Device layout: every slot owns its estimates (xvec) and the 15 x N strip [P_ii | P_iv]; the
// (N-15)^2 block P_vv is NOT copied per IMU step (the reference does, propagator.cpp:204): slots
// propagated from one another share a P_vv generation.  P_vi is the transpose of
// P_iv (the reference computes both with the same products, propagator.cpp:195-203).
#include <cstdlib>
#include "xb_kernels.h"
#include "qd_poly.cuh"
#include "qd_poly_levels.cuh"

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

__device__ __noinline__ void state_transition_call(double dt, const double* w, const double* a, const double* q, double* F);

// One launch propagates the estimates of `n_steps` consecutive slots starting after `start` and emits F_d / Q_d of every
// step into FQ (n_steps x 450): the re-propagation over the buffered IMU tail (Ekf::repropagateFromStateAtIdx,
// ekf.cpp:227-255) and processImu chains longer than one step.
// One CTA per step.  The only true dependency between steps is the q / v / p recurrence (a few hundred cycles per step
// once its operands sit in shared memory), so CTA k simply redoes that recurrence from the start slot up to its own
// step -- no grid-wide dependency -- and then spends its warps on what is expensive: F_d on one warp and the Q_d
// polynomial level by level over XB_QD_NWARP warps.  Round 1 evaluated the polynomial of every step on ONE thread each
// (47 us for 10 steps, the longest item next to the covariance downdate).
__global__ void __launch_bounds__(512) k_prop_means(double* __restrict__ xv, int LX, int NS, int start, int n_steps,
                                                    ImuSample in, PropParams pp, double* __restrict__ FQ) {
  XB_PDL_SHORT();
  __shared__ double imu[129][8];  // slots start .. start + k: w_m[3], a_m[3], time
  __shared__ double Dm[128][16];  // quaternion-integrator matrices of the steps 1 .. k
  __shared__ double wsc[16][80];  // per-warp scratch of the integrator: O1, O0, A, Ak, Tm4
  __shared__ double x0s[32], x1s[32], w1s[3], a1s[3], Cs[9], dts[1], Fs[225], Qs[225], Sq[XB_QD_NTEMP];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int k = blockIdx.x + 1;  // this CTA produces slot start + k
  const double* x0 = xv + (size_t)start * LX;
  double* xk = xv + (size_t)((start + k) % NS) * LX;
  for (int e = t; e < (k + 1) * 8; e += blockDim.x) {
    const int j = e >> 3, c = e & 7;
    const double* xj = xv + (size_t)((start + j) % NS) * LX;
    double v = 0.0;
    if (c < 3) v = xj[XV_WM + c];
    else if (c < 6) v = xj[XV_AM + c - 3];
    else if (c == 6) v = xj[XV_TIME];
    // the new IMU sample of a processImu call goes into the last slot of the chain (State::setImu, state.cpp:145-151)
    if (in.valid && j == n_steps) {
      if (c < 3) v = in.w[c];
      else if (c < 6) v = in.a[c - 3];
      else if (c == 6) v = in.t;
    }
    imu[j][c] = v;
  }
  if (t < 32) {
    x0s[t] = x0[t];
    x1s[t] = (t >= XV_WM && t < XV_ARR) ? xk[t] : x0[t];  // State::setStaticStatesFrom (state.cpp:153-161)
  }
  for (int e = t; e < 225; e += blockDim.x) Qs[e] = 0.0;
  __syncthreads();
  if (t == 0 && in.valid && k == n_steps) {
    x1s[XV_WM] = in.w[0]; x1s[XV_WM + 1] = in.w[1]; x1s[XV_WM + 2] = in.w[2];
    x1s[XV_AM] = in.a[0]; x1s[XV_AM + 1] = in.a[1]; x1s[XV_AM + 2] = in.a[2];
    x1s[XV_TIME] = in.t;
    x1s[XV_SEQ] = in.seq;
  }
  // quaternionIntegrator (propagator.cpp:74-98) of the steps 1..k: one warp per step, element (r, c) per lane, same
  // operation order as quat_integrator()
  for (int j = warp + 1; j <= k; j += 16) {
    double* O1 = wsc[warp];
    double* O0 = O1 + 16;
    double* A = O1 + 32;
    double* Ak = O1 + 48;
    double* Tm4 = O1 + 64;
    double* D = Dm[j - 1];
    const double dt = imu[j][6] - imu[j - 1][6];
    const int r = (lane & 15) >> 2, c = lane & 3;
    if (lane < 16) {
      const int kk[16] = {-1, 2, 1, 0, 2, -1, 0, 1, 1, 0, -1, 2, 0, 1, 2, -1};
      const double sg[16] = {0, 1, -1, 1, -1, 0, 1, 1, 1, -1, 0, 1, -1, -1, -1, 0};
      const int q = kk[lane];
      double o1 = 0.0, o0 = 0.0, om = 0.0;
      if (q >= 0) {
        const double w1 = imu[j][q] - x0s[XV_BW + q], w0 = imu[j - 1][q] - x0s[XV_BW + q];
        o1 = sg[lane] * w1;
        o0 = sg[lane] * w0;
        om = sg[lane] * ((w1 + w0) / 2.0);
      }
      O1[lane] = o1;
      O0[lane] = o0;
      const double a = om * 0.5 * dt;
      A[lane] = a;
      Ak[lane] = a;
      D[lane] = (r == c) ? 1.0 : 0.0;
    }
    __syncwarp();
    int fac = 1;
#pragma unroll 1
    for (int it = 1; it < 5; ++it) {
      fac *= it;
      if (lane < 16) {
        D[lane] = D[lane] + Ak[lane] / fac;
        double s = 0.0;
        for (int e = 0; e < 4; ++e) s += Ak[r * 4 + e] * A[e * 4 + c];
        Tm4[lane] = s;
      }
      __syncwarp();
      if (lane < 16) Ak[lane] = Tm4[lane];
      __syncwarp();
    }
    if (lane < 16) {
      double s10 = 0.0, s01 = 0.0;
      for (int e = 0; e < 4; ++e) {
        s10 += O1[r * 4 + e] * O0[e * 4 + c];
        s01 += O0[r * 4 + e] * O1[e * 4 + c];
      }
      D[lane] += 1.0 / 48.0 * (s10 - s01) * dt * dt;
    }
    __syncwarp();
  }
  __syncthreads();
  if (t == 0) {  // propagateState (propagator.cpp:30-51) for the steps 1..k, operands in shared memory
    double q0[4], v0[3], p0[3], R0[9];
    for (int e = 0; e < 4; ++e) q0[e] = x0s[XV_Q + e];
    for (int e = 0; e < 3; ++e) { v0[e] = x0s[XV_V + e]; p0[e] = x0s[XV_P + e]; }
    xb_rot_raw(q0, R0);
    const double gv[3] = {pp.g[0], pp.g[1], pp.g[2]};
    for (int j = 1; j <= k; ++j) {
      const double* D = Dm[j - 1];
      double a1[3], a0[3];
      for (int e = 0; e < 3; ++e) {
        a1[e] = imu[j][3 + e] - x0s[XV_BA + e];
        a0[e] = imu[j - 1][3 + e] - x0s[XV_BA + e];
      }
      const double dt = imu[j][6] - imu[j - 1][6];
      double q1[4];
      for (int r = 0; r < 4; ++r) q1[r] = D[r * 4] * q0[0] + D[r * 4 + 1] * q0[1] + D[r * 4 + 2] * q0[2] + D[r * 4 + 3] * q0[3];
      xb_qnormalize(q1);
      double R1[9], ra1[3], ra0[3];
      xb_rot_raw(q1, R1);
      xb_mv33(R1, a1, ra1);
      xb_mv33(R0, a0, ra0);
      for (int e = 0; e < 3; ++e) {
        const double dv = (ra1[e] + ra0[e]) / 2.0;
        const double v1 = v0[e] + (dv + gv[e]) * dt;
        p0[e] = p0[e] + (v1 + v0[e]) / 2.0 * dt;
        v0[e] = v1;
      }
      for (int e = 0; e < 4; ++e) q0[e] = q1[e];
      for (int e = 0; e < 9; ++e) R0[e] = R1[e];
      if (j == k) {
        for (int e = 0; e < 3; ++e) { w1s[e] = imu[j][e] - x0s[XV_BW + e]; a1s[e] = a1[e]; }
        dts[0] = dt;
      }
    }
    for (int e = 0; e < 4; ++e) x1s[XV_Q + e] = q0[e];
    for (int e = 0; e < 3; ++e) { x1s[XV_V + e] = v0[e]; x1s[XV_P + e] = p0[e]; }
    for (int e = 0; e < 9; ++e) Cs[e] = R0[e];
  }
  __syncthreads();
  // F_d on warp 0 (one lane); the Q_d polynomial on warps 1..XB_QD_NWARP, each executing its statements of every
  // dependency level with a named barrier among those warps between levels (qd_poly_levels.cuh)
  if (warp == 0) {
    if (lane == 0) state_transition_call(dts[0], w1s, a1s, &x1s[XV_Q], Fs);
  } else if (warp <= XB_QD_NWARP) {
    xb_qd_poly_warp(warp - 1, dts[0], Cs, w1s, a1s, pp.n_w, pp.n_bw, pp.n_a, pp.n_ba, Sq, Qs);
  }
  __syncthreads();
  // estimates of slot start + k (window and feature arrays: State::setStaticStatesFrom)
  if (t >= 32 && t < 64) xk[t - 32] = x1s[t - 32];
  for (int e = XV_ARR + t; e < LX; e += blockDim.x) xk[e] = x0[e];
  __syncthreads();
  double* Fo = FQ + (size_t)(k - 1) * 450;
  if (t < 225) { Fo[t] = Fs[t]; Fo[225 + t] = Qs[t]; }
}

// Strip propagation: for every step k,  strip_k = [F P_ii F^T + Q | F P_iv]  (propagator.cpp:195-203).
// The steps of a chain are NOT processed one after the other: slot k's off-diagonal strip is Phi_k P_iv(start) with the
// prefix product Phi_k = F_k ... F_1 (15x15), so blockIdx.y = k-1 picks the step, every CTA forms the prefix product it needs
// itself (k-1 products of 15x15 matrices, a few hundred cycles each) and all slots are written side by side; the core
// block P_ii needs the true recurrence (Q_d enters every step) and is chained inside the CTAs with blockIdx.x == 0.
// The serial form took 2.8 us per step (28 us for the 10-step re-propagation of an update).  (F_k (F_k-1 v)) and
// ((F_k F_k-1) v) differ in rounding only.
// second != 0: the same recurrence on the column strips P_vi^T (P_vi' = P_vi F^T, propagator.cpp:203), whose core block
// is the transpose of P_ii and therefore takes Q_d^T.
__global__ void __launch_bounds__(128) k_prop_strips(double* __restrict__ strip, int N, int NS, int start, int n_steps,
                                                     const double* __restrict__ FQ, int second) {
  XB_PDL_SHORT();
  // F_d (and Q_d for the core block) of several steps are brought into shared memory at once: one global-memory latency
  // per 8 / 16 steps of the chain instead of one per step
  __shared__ double buf[3600], Pa[225], Pb[225];
  const int t = threadIdx.x;
  const int k = blockIdx.y + 1;
  const size_t SS = (size_t)15 * N;
  const double* s0 = strip + (size_t)start * SS;
  double* s1 = strip + (size_t)((start + k) % NS) * SS;
  if (blockIdx.x == 0) {
    // core block: P_ii <- F P_ii F^T + Q over the steps 1..k
    for (int e = t; e < 225; e += blockDim.x) Pa[e] = s0[(size_t)(e / 15) * N + (e % 15)];
    for (int j0 = 0; j0 < k; j0 += 8) {
      const int nj = min(8, k - j0);
      __syncthreads();
      for (int e = t; e < nj * 450; e += blockDim.x) {
        const int jj = e / 450, r = e % 450;
        buf[e] = FQ[(size_t)(j0 + jj) * 450 + ((r >= 225 && second) ? 225 + ((r - 225) % 15) * 15 + (r - 225) / 15 : r)];
      }
      for (int jj = 0; jj < nj; ++jj) {
        const double* Fs = buf + jj * 450;
        const double* Qs = Fs + 225;
        __syncthreads();
        for (int e = t; e < 225; e += blockDim.x) {  // Pb = F * P_ii
          const int r = e / 15, c = e % 15;
          double s = 0.0;
          for (int a = 0; a < 15; ++a) s = fma(Fs[r * 15 + a], Pa[a * 15 + c], s);
          Pb[e] = s;
        }
        __syncthreads();
        for (int e = t; e < 225; e += blockDim.x) {  // P_ii' = Pb * F^T + Q
          const int r = e / 15, c = e % 15;
          double s = 0.0;
          for (int a = 0; a < 15; ++a) s = fma(Pb[r * 15 + a], Fs[c * 15 + a], s);
          Pa[e] = s + Qs[e];
        }
      }
    }
    __syncthreads();
    for (int e = t; e < 225; e += blockDim.x) s1[(size_t)(e / 15) * N + (e % 15)] = Pa[e];
    return;
  }
  // prefix product Phi_k = F_k ... F_1
  double* cur = Pa;
  double* nxt = Pb;
  for (int j0 = 0; j0 < k; j0 += 16) {
    const int nj = min(16, k - j0);
    __syncthreads();
    for (int e = t; e < nj * 225; e += blockDim.x) buf[e] = FQ[(size_t)(j0 + e / 225) * 450 + e % 225];
    __syncthreads();
    for (int jj = 0; jj < nj; ++jj) {
      const double* Fs = buf + jj * 225;
      if (j0 + jj == 0) {
        for (int e = t; e < 225; e += blockDim.x) cur[e] = Fs[e];
      } else {
        for (int e = t; e < 225; e += blockDim.x) {
          const int r = e / 15, c = e % 15;
          double s = 0.0;
          for (int a = 0; a < 15; ++a) s = fma(Fs[r * 15 + a], cur[a * 15 + c], s);
          nxt[e] = s;
        }
        double* tmp = cur; cur = nxt; nxt = tmp;
      }
      __syncthreads();
    }
  }
  const int j = XB_CORE + ((int)blockIdx.x - 1) * (int)blockDim.x + t;  // column
  if (j < N) {
    double v[15];
#pragma unroll
    for (int r = 0; r < 15; ++r) v[r] = s0[(size_t)r * N + j];
#pragma unroll 1
    for (int r0 = 0; r0 < 15; r0 += 5) {   // five independent accumulation chains per pass
      double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int e = 0; e < 15; ++e)
#pragma unroll
        for (int u = 0; u < 5; ++u) s[u] = fma(cur[(r0 + u) * 15 + e], v[e], s[u]);
#pragma unroll
      for (int u = 0; u < 5; ++u) s1[(size_t)(r0 + u) * N + j] = s[u];
    }
  }
}

__device__ __noinline__ void state_transition_call(double dt, const double* w, const double* a, const double* q, double* F) {
  state_transition(dt, w, a, q, F);
}

// ---- one IMU step in ONE launch (Ekf::processImu, ekf.cpp:66-140: hot loop #1) -------------------------------------
// Every CTA recomputes the step's small quantities itself -- quaternion integrator across 16 lanes, the q/v/p update,
// F_d on one warp and (CTA 0) the Q_d polynomial level by level over XB_QD_NWARP warps (qd_poly_levels.cuh) -- and then
// propagates its 512 columns of the covariance strip(s); CTA 0 also writes the estimates of the new slot.  No grid-wide dependency,
// so means and strips need no second launch, and the serial Q_d evaluation (the whole cost of the two-kernel form on a
// single sample) is cut to its longest partition.
__global__ void __launch_bounds__(512) k_prop_step(double* __restrict__ xv, int LX, double* __restrict__ strip, int N, int NS,
                                                   int start, ImuSample in, PropParams pp, double* __restrict__ FQ) {
  XB_PDL_LONG();
  __shared__ double x0s[32], x1s[32], O1[16], O0[16], A[16], Ak[16], Tm4[16], Dm[16], w1s[3], a1s[3], Cs[9];
  __shared__ double Fs[225], Qs[225], Pii[225], Tm[225], Sq[XB_QD_NTEMP];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const double* x0 = xv + (size_t)start * LX;
  double* x1 = xv + (size_t)((start + 1) % NS) * LX;
#ifdef XB_PROP_PROF
  long long ck[6];
  ck[0] = clock64();
#define PCK(i) if (blockIdx.x == 0 && t == 0) ck[i] = clock64();
#else
#define PCK(i)
#endif
  if (t < 32) {
    x0s[t] = x0[t];
    x1s[t] = (t >= XV_WM && t < XV_ARR) ? x1[t] : x0[t];   // biases, extrinsics: State::setStaticStatesFrom (state.cpp:153-161)
  }
  for (int e = t; e < 225; e += blockDim.x) Qs[e] = 0.0;
  __syncthreads();
  if (t == 0 && in.valid) {                                   // State::setImu (state.cpp:145-151)
    x1s[XV_WM] = in.w[0]; x1s[XV_WM + 1] = in.w[1]; x1s[XV_WM + 2] = in.w[2];
    x1s[XV_AM] = in.a[0]; x1s[XV_AM + 1] = in.a[1]; x1s[XV_AM + 2] = in.a[2];
    x1s[XV_TIME] = in.t;
    x1s[XV_SEQ] = in.seq;
  }
  __syncthreads();
  PCK(1)
  const double dt = x1s[XV_TIME] - x0s[XV_TIME];
  if (warp == 0) {
    // quaternionIntegrator (propagator.cpp:74-98), element (r, c) per lane, same operation order as quat_integrator()
    const int r = (lane & 15) >> 2, c = lane & 3;
    if (lane < 16) {
      // Omega(v)[r][c] (eigen_matrix_base_plugin.h:43-52): +-v[k] with k, sign from the (r, c) pattern
      //   [ 0  z -y  x ; -z  0  x  y ;  y -x  0  z ; -x -y -z  0 ]
      const int kk[16] = {-1, 2, 1, 0, 2, -1, 0, 1, 1, 0, -1, 2, 0, 1, 2, -1};
      const double sg[16] = {0, 1, -1, 1, -1, 0, 1, 1, 1, -1, 0, 1, -1, -1, -1, 0};
      const int k = kk[lane];
      double o1 = 0.0, o0 = 0.0, om = 0.0;
      if (k >= 0) {
        const double w1 = x1s[XV_WM + k] - x1s[XV_BW + k], w0 = x0s[XV_WM + k] - x0s[XV_BW + k];
        o1 = sg[lane] * w1;
        o0 = sg[lane] * w0;
        om = sg[lane] * ((w1 + w0) / 2.0);
      }
      O1[lane] = o1;
      O0[lane] = o0;
      const double a = om * 0.5 * dt;
      A[lane] = a;
      Ak[lane] = a;
      Dm[lane] = (r == c) ? 1.0 : 0.0;
    }
    __syncwarp();
    int fac = 1;
#pragma unroll 1
    for (int k = 1; k < 5; ++k) {
      fac *= k;
      if (lane < 16) {
        Dm[lane] = Dm[lane] + Ak[lane] / fac;
        double s = 0.0;
        for (int e = 0; e < 4; ++e) s += Ak[r * 4 + e] * A[e * 4 + c];
        Tm4[lane] = s;
      }
      __syncwarp();
      if (lane < 16) Ak[lane] = Tm4[lane];
      __syncwarp();
    }
    if (lane < 16) {
      double s10 = 0.0, s01 = 0.0;
      for (int e = 0; e < 4; ++e) {
        s10 += O1[r * 4 + e] * O0[e * 4 + c];
        s01 += O0[r * 4 + e] * O1[e * 4 + c];
      }
      Dm[lane] += 1.0 / 48.0 * (s10 - s01) * dt * dt;
    }
    __syncwarp();
    if (lane == 0) {   // propagateState (propagator.cpp:30-51)
      double q1[4], a1[3], a0[3];
      for (int e = 0; e < 3; ++e) {
        a1[e] = x1s[XV_AM + e] - x1s[XV_BA + e];
        a0[e] = x0s[XV_AM + e] - x0s[XV_BA + e];
      }
      for (int rr = 0; rr < 4; ++rr)
        q1[rr] = Dm[rr * 4] * x0s[XV_Q] + Dm[rr * 4 + 1] * x0s[XV_Q + 1] + Dm[rr * 4 + 2] * x0s[XV_Q + 2] + Dm[rr * 4 + 3] * x0s[XV_Q + 3];
      xb_qnormalize(q1);
      double R1[9], R0[9], ra1[3], ra0[3];
      xb_rot_raw(q1, R1);
      xb_rot_raw(&x0s[XV_Q], R0);
      xb_mv33(R1, a1, ra1);
      xb_mv33(R0, a0, ra0);
      const double gv[3] = {pp.g[0], pp.g[1], pp.g[2]};
      for (int e = 0; e < 3; ++e) {
        const double dv = (ra1[e] + ra0[e]) / 2.0;
        const double v1 = x0s[XV_V + e] + (dv + gv[e]) * dt;
        x1s[XV_V + e] = v1;
        x1s[XV_P + e] = x0s[XV_P + e] + (v1 + x0s[XV_V + e]) / 2.0 * dt;
      }
      for (int e = 0; e < 4; ++e) x1s[XV_Q + e] = q1[e];
      // operands of F_d / Q_d, staged in shared memory for the warps below
      for (int e = 0; e < 3; ++e) {
        w1s[e] = x1s[XV_WM + e] - x1s[XV_BW + e];
        a1s[e] = a1[e];
      }
      for (int e = 0; e < 9; ++e) Cs[e] = R1[e];
    }
  }
  __syncthreads();
  PCK(2)
  // F_d on warp 0 (one lane); CTA 0 (the only one that needs Q_d: it owns the core block) evaluates the polynomial on warps
  // 1..XB_QD_NWARP, each executing its statements of every dependency level with a named barrier among those warps between
  // levels (qd_poly_levels.cuh); the strip CTAs go straight to their columns
  if (warp == 0) {
    if (lane == 0) state_transition_call(dt, w1s, a1s, &x1s[XV_Q], Fs);
  } else if (blockIdx.x == 0 && warp <= XB_QD_NWARP) {
    xb_qd_poly_warp(warp - 1, dt, Cs, w1s, a1s, pp.n_w, pp.n_bw, pp.n_a, pp.n_ba, Sq, Qs);
  }
  __syncthreads();
  PCK(3)
  // strip columns of this thread: P_iv' = F P_iv (propagator.cpp:197); compact loops (the kernel runs once per launch on
  // cold instruction caches: code size is latency)
  // CTA 0: core block + estimates; CTAs 1..: 512 strip columns each (they run side by side on different SMs)
  const bool core_block = blockIdx.x == 0;
  const int j = XB_CORE + ((int)blockIdx.x - 1) * (int)blockDim.x + t;
  const size_t SS = (size_t)15 * N;
  const double* s0 = strip + (size_t)start * SS;
  double* s1 = strip + (size_t)((start + 1) % NS) * SS;
  if (!core_block && j < N) {
    double v[15];
#pragma unroll
    for (int r = 0; r < 15; ++r) v[r] = s0[(size_t)r * N + j];
#pragma unroll 1
    for (int r0 = 0; r0 < 15; r0 += 5) {   // five independent accumulation chains per pass
      double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int e = 0; e < 15; ++e)
#pragma unroll
        for (int u = 0; u < 5; ++u) s[u] = fma(Fs[(r0 + u) * 15 + e], v[e], s[u]);
#pragma unroll
      for (int u = 0; u < 5; ++u) s1[(size_t)(r0 + u) * N + j] = s[u];
    }
  }
  if (core_block) {
    if (t < 225) Pii[t] = s0[(size_t)(t / 15) * N + (t % 15)];
    __syncthreads();
    if (t < 225) {  // Tm = F * P_ii
      const int r = t / 15, c = t % 15;
      double s = 0.0;
      for (int a = 0; a < 15; ++a) s = fma(Fs[r * 15 + a], Pii[a * 15 + c], s);
      Tm[t] = s;
    }
    __syncthreads();
    if (t < 225) {  // P_ii' = Tm * F^T + Q
      const int r = t / 15, c = t % 15;
      double s = 0.0;
      for (int a = 0; a < 15; ++a) s = fma(Tm[r * 15 + a], Fs[c * 15 + a], s);
      s1[(size_t)r * N + c] = s + Qs[t];
      FQ[t] = Fs[t];            // F_d / Q_d of the step where the two-kernel form leaves them
      FQ[225 + t] = Qs[t];
    }
    // estimates of the new slot
    if (t < 32) x1[t] = x1s[t];
    for (int e = XV_ARR + t; e < LX; e += blockDim.x) x1[e] = x0[e];
  }
#ifdef XB_PROP_PROF
  PCK(4)
  if (blockIdx.x == 0 && t == 0)
    for (int i = 0; i < 5; ++i) FQ[460 + i] = (double)(ck[i] - ck[0]);
#endif
}

void launch_prop_step(cudaStream_t s, double* xv, int LX, double* strip, int N, int NS, int start, const ImuSample& in,
                      const PropParams& pp, double* FQ) {
  XB_LAUNCH(k_prop_step, 1 + (N - XB_CORE + 511) / 512, 512, 0, s, xv, LX, strip, N, NS, start, in, pp, FQ);
  count_launch();
}

void launch_prop_means(cudaStream_t s, double* xv, int LX, int NS, int start, int n_steps, const ImuSample& in,
                       const PropParams& pp, double* FQ) {
  if (n_steps <= 0) return;
  XB_LAUNCH(k_prop_means, n_steps, 512, 0, s, xv, LX, NS, start, n_steps, in, pp, FQ);
  count_launch();
}
void launch_prop_strips(cudaStream_t s, double* strip, int N, int NS, int start, int n_steps, const double* FQ, int second) {
  if (n_steps <= 0) return;
  dim3 grid(1 + (N - XB_CORE + 127) / 128, n_steps);
  XB_LAUNCH(k_prop_strips, grid, 128, 0, s, strip, N, NS, start, n_steps, FQ, second);
  count_launch();
}
// IMU fields (w_m, a_m, time, seq: xvec 23..30) of the n slots after `start`, written before a batched propagation
// (State::setImu, state.cpp:145-151): the samples travel as kernel arguments, no staging buffer
__global__ void k_imu_scatter(double* __restrict__ xv, int LX, int NS, int start, int n, ImuBatch b) {
  XB_PDL_SHORT();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * 8) return;
  const int j = e >> 3, c = e & 7;
  xv[(size_t)((start + 1 + j) % NS) * LX + XV_WM + c] = b.v[j][c];
}
void launch_imu_scatter(cudaStream_t s, double* xv, int LX, int NS, int start, int n, const ImuBatch& b) {
  XB_LAUNCH(k_imu_scatter, (n * 8 + 127) / 128, 128, 0, s, xv, LX, NS, start, n, b);
  count_launch();
}
void launch_propagate(cudaStream_t s, double* xv, int LX, double* strip, int N, int NS, int start, int n_steps,
                      const ImuSample& in, const PropParams& pp, double* FQ) {
  launch_prop_means(s, xv, LX, NS, start, n_steps, in, pp, FQ);
  launch_prop_strips(s, strip, N, NS, start, n_steps, FQ);
}

}  // namespace xb
