// Multi-agent covariance-intersection fusion on the device (SLAM-SLAM matches).
// reference: src/x/vio/multi_slam_update.cpp:61-246 (processOneMatch), src/x/ekf/ci.cpp:94-127 (pair fuseCI,
// fixed weight), src/x/ekf/updater.cpp:22-36,144-161 (collaborativeUpdate / applyCI).
//
// Compressed payload.  The reference ships a peer's whole SimpleState -- including its N x N
// covariance (simple_state.h:65-74) -- but a SLAM-SLAM match uses the peer only through
//     other_G_p_f  (3)        the matched feature in world coordinates
//     other_h P_other other_h^T  (3x3)   (other_h has 9 non-zero columns: anchor pos/att, feature)
// so every agent packs, per SLAM feature, 13 doubles [valid | G_p_f | h P h^T] computed ON ITS OWN GPU from
// its own state; the slots are exchanged with one all-gather (NCCL) and each agent runs the CI step locally
// with identical arithmetic.
#include "xb_kernels.h"

namespace xb {

// h (3x9 over [anchor pos | anchor att | feature]) and world point of SLAM feature f of the state in xv.
__device__ bool slam_world_jac(const double* xv, int M, int n_poses, int anchor, int f, double* h9, int* cols9, double* G) {
  const double* parr = xv + XV_ARR;
  const double* qarr = xv + XV_ARR + 3 * M;
  const double* farr = xv + XV_ARR + 7 * M;
  const double a = farr[3 * f], b = farr[3 * f + 1], r = farr[3 * f + 2];
  if (anchor < 0 || anchor >= n_poses || r == 0.0) return false;  // multi_slam_update.cpp:83-88
  double Ra[9], sk[9], m3[9], t3[3], A1[9], A2[9];
  xb_rot(qarr + 4 * anchor, Ra);
  const double ab1[3] = {a, b, 1.0};
  xb_mv33(Ra, ab1, t3);
  for (int e = 0; e < 3; ++e) G[e] = (1.0 / r) * t3[e] + parr[3 * anchor + e];
  xb_skew(ab1, sk);
  xb_mm33(Ra, sk, A1);
  xb_mat_ivd(a, b, r, m3);
  xb_mm33(Ra, m3, A2);
  for (int i = 0; i < 3; ++i)
    for (int c = 0; c < 3; ++c) {
      h9[i * 9 + c] = (i == c) ? 1.0 : 0.0;              // anchor position
      h9[i * 9 + 3 + c] = -(1.0 / r) * A1[i * 3 + c];    // anchor attitude
      h9[i * 9 + 6 + c] = (1.0 / r) * A2[i * 3 + c];     // feature (alpha, beta, rho)
    }
  for (int c = 0; c < 3; ++c) {
    cols9[c] = XB_CORE + 3 * anchor + c;
    cols9[3 + c] = XB_CORE + 3 * M + 3 * anchor + c;
    cols9[6 + c] = XB_CORE + (2 * M + f) * 3 + c;
  }
  return true;
}

// payload: [0]=n_features [1]=time [2..7] reserved | per feature f: 13 doubles
__global__ void k_ci_pack(const double* __restrict__ xv, const double* __restrict__ P, int N, int M, int F, int n_poses,
                          int n_features, const int* __restrict__ anchor, double* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f == 0) { out[0] = n_features; out[1] = xv[XV_TIME]; for (int e = 2; e < 8; ++e) out[e] = 0.0; }
  if (f >= F) return;
  double* o = out + 8 + 13 * f;
  for (int e = 0; e < 13; ++e) o[e] = 0.0;
  if (f >= n_features) return;
  double h9[27], G[3];
  int c9[9];
  if (!slam_world_jac(xv, M, n_poses, anchor[f], f, h9, c9, G)) return;
  double T[27];  // h P (3 x 9 over the same 9 columns)
  for (int i = 0; i < 3; ++i)
    for (int c = 0; c < 9; ++c) {
      double s = 0.0;
      for (int d = 0; d < 9; ++d) s = fma(h9[i * 9 + d], P[(size_t)c9[d] * N + c9[c]], s);
      T[i * 9 + c] = s;
    }
  o[0] = 1.0;
  for (int e = 0; e < 3; ++e) o[1 + e] = G[e];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) {
      double s = 0.0;
      for (int c = 0; c < 9; ++c) s = fma(T[i * 9 + c], h9[k * 9 + c], s);
      o[4 + i * 3 + k] = s;
    }
}
void launch_ci_pack(cudaStream_t s, const double* xv, const double* P, int N, int M, int F, int n_poses, int n_features,
                    const int* anchor, double* out) {
  k_ci_pack<<<(std::max(F, 1) + 63) / 64, 64, 0, s>>>(xv, P, N, M, F, n_poses, n_features, anchor, out);
  count_launch();
}

// One thread per match: residual, gate, CI-fused S^-1.  rec: per match 64 doubles
//   [0]=inlier [1]=gamma [2]=w_result [3..11]=S^-1 [12..14]=res [15..41]=h9 [42..50]=cols9 (as doubles)
__global__ void k_ci_slam_construct(const double* __restrict__ xv, const double* __restrict__ P, int N, int M, int n_poses,
                                    int n_features, const int* __restrict__ anchor, const double* __restrict__ gathered,
                                    int payload_len, const int* __restrict__ matches /*[n][3] peer,cur,recv*/, int n_matches,
                                    double var_lm, double w_other, double chi2_90_3, double* __restrict__ rec,
                                    int* __restrict__ last_inlier) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_matches) {
    double* o = rec + 64 * (size_t)j;
    for (int e = 0; e < 64; ++e) o[e] = 0.0;
    const int peer = matches[3 * j], cur = matches[3 * j + 1], rcv = matches[3 * j + 2];
    const double* po = gathered + (size_t)peer * payload_len + 8 + 13 * rcv;
    double h9[27], G[3];
    int c9[9];
    const bool ok = cur >= 0 && cur < n_features && po[0] == 1.0 && slam_world_jac(xv, M, n_poses, anchor[cur], cur, h9, c9, G);
    if (ok) {
      double T[27], Mo[9];
      for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 9; ++c) {
          double s = 0.0;
          for (int d = 0; d < 9; ++d) s = fma(h9[i * 9 + d], P[(size_t)c9[d] * N + c9[c]], s);
          T[i * 9 + c] = s;
        }
      for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) {
          double s = 0.0;
          for (int c = 0; c < 9; ++c) s = fma(T[i * 9 + c], h9[k * 9 + c], s);
          Mo[i * 3 + k] = s;
        }
      double res[3], Sg[9], Si[9];
      for (int e = 0; e < 3; ++e) res[e] = -G[e] + po[1 + e];  // multi_slam_update.cpp:131
      for (int e = 0; e < 9; ++e) Sg[e] = Mo[e] + po[4 + e] + ((e % 4 == 0) ? var_lm : 0.0);
      xb_inv33(Sg, Si);
      double gamma = 0.0;
      for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) gamma += res[i] * Si[i * 3 + k] * res[k];
      o[1] = gamma;
      if (gamma < chi2_90_3) {  // chi2(0.9, 3), multi_slam_update.cpp:216-220
        const double w_res = 1.0 / (1.0 - w_other);  // ci.cpp:94-127 (fixed weight)
        double Sj[9];
        for (int e = 0; e < 9; ++e) Sj[e] = w_res * Mo[e] + (1.0 / w_other) * po[4 + e] + ((e % 4 == 0) ? var_lm : 0.0);
        xb_inv33(Sj, Si);
        o[0] = 1.0;
        o[2] = w_res;
        for (int e = 0; e < 9; ++e) o[3 + e] = Si[e];
        for (int e = 0; e < 3; ++e) o[12 + e] = res[e];
        for (int e = 0; e < 27; ++e) o[15 + e] = h9[e];
        for (int e = 0; e < 9; ++e) o[42 + e] = (double)c9[e];
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *last_inlier = -1;
}
__global__ void k_ci_last_inlier(const double* __restrict__ rec, int n_matches, int* __restrict__ last_inlier) {
  int last = -1;
  for (int j = 0; j < n_matches; ++j)
    if (rec[64 * (size_t)j] == 1.0) last = j;
  *last_inlier = last;
}

// K_j = P_j h_j^T S_j^-1 (N x 3), delta_j = K_j res_j.  P_j = P with the three diagonal 3x3 blocks scaled by w.
__global__ void k_ci_gain(const double* __restrict__ P, int N, const double* __restrict__ rec, int n_matches,
                          double* __restrict__ Kall /*[n][N][3]*/, double* __restrict__ delta /*[n][N]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= N || j >= n_matches) return;
  const double* o = rec + 64 * (size_t)j;
  double* dj = delta + (size_t)j * N;
  if (o[0] != 1.0) { dj[i] = 0.0; return; }
  const double w = o[2];
  double a[3] = {0, 0, 0};
  for (int c = 0; c < 9; ++c) {
    const int col = (int)o[42 + c];
    double p = P[(size_t)i * N + col];
    // P_j: only the diagonal 3x3 blocks at the three column groups are scaled (multi_slam_update.cpp:229-239)
    const int g0 = (int)o[42 + (c / 3) * 3];
    if (i >= g0 && i < g0 + 3) p *= w;
    for (int r = 0; r < 3; ++r) a[r] = fma(p, o[15 + r * 9 + c], a[r]);
  }
  double k3[3];
  for (int r = 0; r < 3; ++r) k3[r] = a[0] * o[3 + 0 * 3 + r] + a[1] * o[3 + 1 * 3 + r] + a[2] * o[3 + 2 * 3 + r];
  for (int r = 0; r < 3; ++r) Kall[((size_t)j * N + i) * 3 + r] = k3[r];
  dj[i] = k3[0] * o[12] + k3[1] * o[13] + k3[2] * o[14];
}

// Sequential State::correct for every inlier match in list order (updater.cpp:31-34 -> applyCI -> state.correct).
__global__ void k_ci_correct_seq(int M, int F, int N, const double* __restrict__ rec, int n_matches,
                                 const double* __restrict__ delta, double* __restrict__ xv) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int j = 0; j < n_matches; ++j) {
    if (rec[64 * (size_t)j] != 1.0) continue;
    const double* d = delta + (size_t)j * N;
    if (t < 3) {
      xv[XV_P + t] += d[t];
      xv[XV_V + t] += d[3 + t];
      xv[XV_BW + t] += d[9 + t];
      xv[XV_BA + t] += d[12 + t];
    }
    if (t < 3 * M) xv[XV_ARR + t] += d[XB_CORE + t];
    if (t < 3 * F) xv[XV_ARR + 7 * M + t] += d[XB_CORE + 6 * M + t];
    if (t <= M) {
      double* q = (t == M) ? xv + XV_Q : xv + XV_ARR + 3 * M + 4 * t;
      const double* dd = (t == M) ? d + 6 : d + XB_CORE + 3 * M + 3 * t;
      double dq[4], qo[4];
      xb_small_angle_quat(dd, dq);
      xb_qmul(q, dq, qo);
      const double n = sqrt(qo[0] * qo[0] + qo[1] * qo[1] + qo[2] * qo[2] + qo[3] * qo[3]);
      if (n > 0.0) { qo[0] /= n; qo[1] /= n; qo[2] /= n; qo[3] /= n; }
      q[0] = qo[0]; q[1] = qo[1]; q[2] = qo[2]; q[3] = qo[3];
    }
  }
}

// covariance of the LAST inlier match ("last match wins", updater.cpp:31-34,155): P <- sym((I - K H) P_j)
//   step 1: HP = h (P_j rows)  (3 x N);  step 2: P_ij <- ((P_j - K HP)_ij + (P_j - K HP)_ji) / 2
__global__ void k_ci_hp(const double* __restrict__ P, int N, const double* __restrict__ rec, const int* __restrict__ last_inlier,
                        double* __restrict__ HP) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int jl = *last_inlier;
  if (c >= N || jl < 0) return;
  const double* o = rec + 64 * (size_t)jl;
  const double w = o[2];
  double a[3] = {0, 0, 0};
  for (int e = 0; e < 9; ++e) {
    const int row = (int)o[42 + e];
    double p = P[(size_t)row * N + c];
    const int g0 = (int)o[42 + (e / 3) * 3];
    if (c >= g0 && c < g0 + 3) p *= w;
    for (int r = 0; r < 3; ++r) a[r] = fma(o[15 + r * 9 + e], p, a[r]);
  }
  for (int r = 0; r < 3; ++r) HP[(size_t)r * N + c] = a[r];
}
__global__ void __launch_bounds__(256) k_ci_cov_last(double* __restrict__ P, int N, const double* __restrict__ rec,
                                                     const int* __restrict__ last_inlier, const double* __restrict__ Kall,
                                                     const double* __restrict__ HP) {
  const int jl = *last_inlier;
  if (jl < 0) return;
  const int j = blockIdx.x * 16 + (threadIdx.x & 15), i = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (i >= N || j >= N || i > j) return;
  const double* o = rec + 64 * (size_t)jl;
  const double w = o[2];
  const double* K = Kall + (size_t)jl * N * 3;
  double pij = P[(size_t)i * N + j], pji = P[(size_t)j * N + i];
  for (int g = 0; g < 3; ++g) {  // P_j: scaled diagonal blocks
    const int g0 = (int)o[42 + 3 * g];
    if (i >= g0 && i < g0 + 3 && j >= g0 && j < g0 + 3) { pij *= w; pji *= w; }
  }
  double a = 0.0, b = 0.0;
  for (int k = 0; k < 3; ++k) {
    a = fma(K[(size_t)i * 3 + k], HP[(size_t)k * N + j], a);
    b = fma(K[(size_t)j * 3 + k], HP[(size_t)k * N + i], b);
  }
  const double v = 0.5 * ((pij - a) + (pji - b));
  P[(size_t)i * N + j] = v;
  P[(size_t)j * N + i] = v;
}

void launch_ci_slam(cudaStream_t s, double* xv, double* P, int N, int M, int F, int n_poses, int n_features,
                    const int* anchor, const double* gathered, int payload_len, const int* matches, int n_matches,
                    double var_lm, double w_other, double chi2_90_3, double* rec, int* last_inlier, double* Kall,
                    double* delta, double* HP) {
  if (n_matches <= 0) return;
  k_ci_slam_construct<<<(n_matches + 63) / 64, 64, 0, s>>>(xv, P, N, M, n_poses, n_features, anchor, gathered, payload_len,
                                                           matches, n_matches, var_lm, w_other, chi2_90_3, rec, last_inlier);
  count_launch();
  k_ci_last_inlier<<<1, 1, 0, s>>>(rec, n_matches, last_inlier);
  count_launch();
  dim3 g((N + 127) / 128, n_matches);
  k_ci_gain<<<g, 128, 0, s>>>(P, N, rec, n_matches, Kall, delta);
  count_launch();
  k_ci_hp<<<(N + 127) / 128, 128, 0, s>>>(P, N, rec, last_inlier, HP);
  count_launch();
  dim3 gc((N + 15) / 16, (N + 15) / 16);
  k_ci_cov_last<<<gc, 256, 0, s>>>(P, N, rec, last_inlier, Kall, HP);
  count_launch();
  k_ci_correct_seq<<<(N + 127) / 128, 128, 0, s>>>(M, F, N, rec, n_matches, delta, xv);
  count_launch();
}

}  // namespace xb
