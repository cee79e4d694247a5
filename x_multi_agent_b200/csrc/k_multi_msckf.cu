// Multi-agent MSCKF-MSCKF block of the MULTI_UAV build: an own MSCKF track matched with other agents' tracks of
// the same landmark is triangulated jointly, the feature-dependent 3-row parts of every agent's measurement are
// stacked, projected on the left nullspace of the stacked feature Jacobian, gated, fused by covariance
// intersection and applied through Updater::applyCI.
//
// reference: src/x/vio/msckf_update.cpp:65-281 (preProcessOneTrack), :306-492 (processOneTrack, is_multi_msckf),
//            :494-501 (nullSpaceProjection), src/x/ekf/ci.cpp:49-92 (k-agent fuseCI, fixed weights),
//            src/x/ekf/updater.cpp:84-92,144-161 (applyCI over the list), src/x/ekf/simple_state.cpp:34-65.
//
// Formulation.  For agent i (0 = own) let U_i = orth(range(Hf_i)), B_i = U_i^T J_i (3 x 6M_i), F_i = U_i^T Hf_i,
// b_i = U_i^T r_i -- the reference's A_up^T products (:439-445); any orthonormal U_i gives the same update.  With
// A (3(k+1) x 3k) a basis of the left nullspace of [F_0; ..; F_k] and A_i its 3 rows of agent i:
//   h_j = A_0^T B_0,   res = A^T b,   S = sum_i c_i A_i^T (B_i P_i B_i^T) A_i + s^2 I
// (c_i = 1 for the gate, 1/w_i for the CI-fused S).  A peer enters only through M_i = B_i P_i B_i^T (3x3), which
// needs its pose window and the 6M x 6M pose block of its covariance: that is the "pose payload" agents exchange.
// h_j has rank 3, so applyCI reduces to 3-row quantities:  K res = P_j B_0^T a,  a = A_0 S^-1 res;
//   (I - K h_j) P_j = P_j - (P_j B_0^T) C3 (B_0 P_j),  C3 = A_0 S^-1 A_0^T.
// P_j = prior P with the diagonal 3x3 pose blocks of the track's window slots scaled by 1/w_0 (:258-267); every
// list entry restarts from the prior, so the LAST inlier entry defines the covariance while all state corrections
// accumulate in list order (updater.cpp:88-92,155).
#include "xb_kernels.h"
#include "xb_svd4.cuh"

namespace xb {

struct MmObs { const double* q; const double* p; double z0, z1; };

// joint observation t of a group: the peers' tracks first (msckf_update.cpp:96-139), the own track last (:141-145).
// A peer's pose list is its whole window (all M slots, simple_state.cpp:34-65); its track uses the last n_obs.
__device__ __forceinline__ MmObs mm_obs(const MmParams& mp, const int* gd, int t) {
  MmObs s;
  const int k = gd[1], e0 = gd[2];
  for (int e = 0; e < k; ++e) {
    const int* en = mp.ent + 3 * (e0 + e);
    const int Lp = en[2];
    if (t < Lp) {
      const double* pp = mp.gathered + (size_t)en[0] * mp.pp_len;
      const int slot = mp.M - Lp + t;
      s.p = pp + 8 + 3 * slot;
      s.q = pp + 8 + 3 * mp.M + 4 * slot;
      s.z0 = mp.pobs[2 * (size_t)(en[1] + t)];
      s.z1 = mp.pobs[2 * (size_t)(en[1] + t) + 1];
      return s;
    }
    t -= Lp;
  }
  const int trk = gd[0];
  const int o0 = mp.off[trk], L = mp.off[trk + 1] - o0;
  const int slot = mp.n_poses - L + t;
  s.p = mp.xv + XV_ARR + 3 * slot;
  s.q = mp.xv + XV_ARR + 3 * mp.M + 4 * slot;
  s.z0 = mp.obs[2 * (size_t)(o0 + t)];
  s.z1 = mp.obs[2 * (size_t)(o0 + t) + 1];
  return s;
}

// ---- joint triangulation (triangulation.cpp:102-206 over the concatenated pose / observation lists) ------------
__global__ void __launch_bounds__(128) k_mm_triangulate(MmParams mp) {
  const int lane = threadIdx.x & 31, g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= mp.n_groups) return;
  const int* gd = mp.grp + 4 * g;
  const int n_tot = gd[3];
  const MmObs ol = mm_obs(mp, gd, n_tot - 1);  // anchor: the own newest pose of the track
  double Rl[9];
  xb_rot(ol.q, Rl);
  const double pl[3] = {ol.p[0], ol.p[1], ol.p[2]};
  double alpha = 0.0, beta = 0.0, rho = 1.0;
  if (lane == 0) {
    const MmObs o1 = mm_obs(mp, gd, 0);
    double R1[9], A[16], P1[12], P2[12];
    xb_rot(o1.q, R1);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) { P1[r * 4 + c] = R1[c * 3 + r]; P2[r * 4 + c] = Rl[c * 3 + r]; }
      P1[r * 4 + 3] = -(R1[0 * 3 + r] * o1.p[0] + R1[1 * 3 + r] * o1.p[1] + R1[2 * 3 + r] * o1.p[2]);
      P2[r * 4 + 3] = -(Rl[0 * 3 + r] * pl[0] + Rl[1 * 3 + r] * pl[1] + Rl[2 * 3 + r] * pl[2]);
    }
    for (int c = 0; c < 4; ++c) {
      A[0 + c] = o1.z0 * P1[8 + c] - P1[0 + c];
      A[4 + c] = o1.z1 * P1[8 + c] - P1[4 + c];
      A[8 + c] = ol.z0 * P2[8 + c] - P2[0 + c];
      A[12 + c] = ol.z1 * P2[8 + c] - P2[4 + c];
    }
    double vh[4];
    smallest_right_singular_vector4(A, vh);
    const double x = vh[0] / vh[3], y = vh[1] / vh[3], z = vh[2] / vh[3];
    double c2[3];
    for (int r = 0; r < 3; ++r) c2[r] = P2[r * 4] * x + P2[r * 4 + 1] * y + P2[r * 4 + 2] * z + P2[r * 4 + 3];
    alpha = c2[0] / c2[2];
    beta = c2[1] / c2[2];
    rho = 1.0 / c2[2];
  }
  alpha = __shfl_sync(0xffffffffu, alpha, 0);
  beta = __shfl_sync(0xffffffffu, beta, 0);
  rho = __shfl_sync(0xffffffffu, rho, 0);
  double r_norm_last = 1000.0, r_norm = 100.0;
  int iter = 0;
  while (r_norm_last - r_norm > mp.gn_term) {
    ++iter;
    if (iter > mp.gn_max_iter) break;
    double jtj[6] = {0, 0, 0, 0, 0, 0}, jtr[3] = {0, 0, 0}, rr = 0.0;
    for (int t = lane; t < n_tot; t += 32) {
      const MmObs o = mm_obs(mp, gd, t);
      double Ri[9], dR[9], dp[3];
      xb_rot(o.q, Ri);
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
          dR[r * 3 + c] = Ri[0 * 3 + r] * Rl[0 * 3 + c] + Ri[1 * 3 + r] * Rl[1 * 3 + c] + Ri[2 * 3 + r] * Rl[2 * 3 + c];
        const double a = Ri[0 * 3 + r] * pl[0] + Ri[1 * 3 + r] * pl[1] + Ri[2 * 3 + r] * pl[2];
        const double b = Ri[0 * 3 + r] * o.p[0] + Ri[1 * 3 + r] * o.p[1] + Ri[2 * 3 + r] * o.p[2];
        dp[r] = a - b;
      }
      double h[3];
      for (int r = 0; r < 3; ++r) h[r] = dR[r * 3] * alpha + dR[r * 3 + 1] * beta + dR[r * 3 + 2] + rho * dp[r];
      const double r0 = o.z0 - h[0] / h[2], r1 = o.z1 - h[1] / h[2];
      const double j1a = -1.0 / h[2], j1b = h[0] / (h[2] * h[2]), j1c = h[1] / (h[2] * h[2]);
      double J0[3], J1[3];
      J0[0] = j1a * dR[0] + j1b * dR[6]; J0[1] = j1a * dR[1] + j1b * dR[7]; J0[2] = j1a * dp[0] + j1b * dp[2];
      J1[0] = j1a * dR[3] + j1c * dR[6]; J1[1] = j1a * dR[4] + j1c * dR[7]; J1[2] = j1a * dp[1] + j1c * dp[2];
      jtj[0] += J0[0] * J0[0] + J1[0] * J1[0];
      jtj[1] += J0[0] * J0[1] + J1[0] * J1[1];
      jtj[2] += J0[0] * J0[2] + J1[0] * J1[2];
      jtj[3] += J0[1] * J0[1] + J1[1] * J1[1];
      jtj[4] += J0[1] * J0[2] + J1[1] * J1[2];
      jtj[5] += J0[2] * J0[2] + J1[2] * J1[2];
      jtr[0] += J0[0] * r0 + J1[0] * r1;
      jtr[1] += J0[1] * r0 + J1[1] * r1;
      jtr[2] += J0[2] * r0 + J1[2] * r1;
      rr += r0 * r0 + r1 * r1;
    }
#pragma unroll
    for (int e = 0; e < 6; ++e) jtj[e] = xb_warp_sum(jtj[e]);
#pragma unroll
    for (int e = 0; e < 3; ++e) jtr[e] = xb_warp_sum(jtr[e]);
    rr = xb_warp_sum(rr);
    const double Am[9] = {jtj[0], jtj[1], jtj[2], jtj[1], jtj[3], jtj[4], jtj[2], jtj[4], jtj[5]};
    double Ai[9], d[3];
    xb_inv33(Am, Ai);
    xb_mv33(Ai, jtr, d);
    alpha -= d[0];
    beta -= d[1];
    rho -= d[2];
    r_norm_last = r_norm;
    r_norm = sqrt(rr);
  }
  if (lane == 0) { mp.ivd[3 * g] = alpha; mp.ivd[3 * g + 1] = beta; mp.ivd[3 * g + 2] = rho; }
}

// ---- small dense helpers in shared memory --------------------------------------------------------------------------
#define MM_LDG 44  // row stride of the Gauss-Jordan work matrix: up to 21 rows x (21 + 21 + 1) columns
#define MM_LDQ 24

__device__ __forceinline__ double mm_wdot(const double* a, int sa, const double* b, int sb, int n, int lane) {
  double s = 0.0;
  for (int i = lane; i < n; i += 32) s = fma(a[i * sa], b[i * sb], s);
  return xb_warp_sum(s);
}

// in-place Gauss-Jordan with partial pivoting on [S | rhs...] (m rows, w columns); lanes own rows.
// Eigen's dynamic-size inverse() is PartialPivLU (msckf_update.cpp:241); the gate only needs S^-1 applied.
__device__ bool mm_gj(double* Sg, int m, int w, int lane) {
  for (int c = 0; c < m; ++c) {
    int best = c;
    double bv = fabs(Sg[c * MM_LDG + c]);
    for (int r = c + 1; r < m; ++r) {
      const double x = fabs(Sg[r * MM_LDG + c]);
      if (x > bv) { bv = x; best = r; }
    }
    if (!(bv > 0.0)) return false;
    __syncwarp();
    if (best != c)
      for (int col = lane; col < w; col += 32) {
        const double tmp = Sg[c * MM_LDG + col];
        Sg[c * MM_LDG + col] = Sg[best * MM_LDG + col];
        Sg[best * MM_LDG + col] = tmp;
      }
    __syncwarp();
    const double piv = Sg[c * MM_LDG + c];
    const double f = (lane < m) ? Sg[lane * MM_LDG + c] : 0.0;
    __syncwarp();
    for (int col = lane; col < w; col += 32) Sg[c * MM_LDG + col] /= piv;
    __syncwarp();
    if (lane < m && lane != c)
      for (int col = 0; col < w; ++col) Sg[lane * MM_LDG + col] = fma(-f, Sg[c * MM_LDG + col], Sg[lane * MM_LDG + col]);
    __syncwarp();
  }
  return true;
}

static __host__ __device__ size_t mm_smem_doubles(int Lm) {
  return (size_t)62 * Lm + 72 + 24 + 72 + 72 + 4 + MM_LDQ * MM_LDQ + 21 * MM_LDG + 24 + 24;
}

// ---- per-group construction: peers' 3-row blocks, nullspace, gate, CI --------------------------------------------------
// MM_NW warps per group: warp 0 runs the (short, serial) geometry; the two covariance products M = B P B^T -- 6M columns x
// 6M rows of L2-resident covariance each, 95 us apiece when one warp walks them -- are spread over all warps.
#define MM_NW 6
__global__ void __launch_bounds__(32 * MM_NW) k_mm_construct(MmParams mp) {
  extern __shared__ double sm[];
  __shared__ double Mpart[MM_NW][9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = blockIdx.x;
  const bool is_w0 = warp == 0;
  const int M = mp.M, N = mp.N, np = mp.n_poses, W = 6 * M + 1;
  const int* gd = mp.grp + 4 * g;
  const int trk = gd[0], k = gd[1], e0 = gd[2], n_tot = gd[3];
  double* rec = mp.rec + (size_t)XB_MM_REC * g;
  const int o0 = mp.off[trk], L = mp.off[trk + 1] - o0, i1 = np - L;
  if (is_w0 && lane < XB_MM_REC) rec[lane] = 0.0;
  __syncwarp();
  if (is_w0 && lane == 0) { rec[16] = trk; rec[17] = i1; rec[18] = L; rec[19] = k; rec[1] = NAN; rec[2] = mp.chi2[g]; }
  // proceed_with_multi_: the own track must have passed its own gate (msckf_update.cpp:476-481)
  if (!mp.inlier[trk] || k < 1 || k > XB_MM_KMAX) return;

  const int Lm = M;
  double* Jp = sm;
  double* Ja = Jp + 6 * Lm;
  double* Hf = Ja + 6 * Lm;
  double* U = Hf + 6 * Lm;
  double* res = U + 6 * Lm;
  double* Bc = res + 2 * Lm;       // 3 x 6Lp compact: [pos 3Lp | att 3Lp]
  double* Tc = Bc + 18 * Lm;
  double* Fs = Tc + 18 * Lm;       // 3(k+1) x 3
  double* bs = Fs + 72;
  double* Mi = bs + 24;            // (k+1) x 9
  double* Vh = Mi + 72;            // 3 reflectors x 24
  double* tau = Vh + 72;
  double* Q = tau + 4;             // n x n, row stride MM_LDQ
  double* Sg = Q + MM_LDQ * MM_LDQ;
  double* rs = Sg + 21 * MM_LDG;   // res_pf after the projection
  double* ys = rs + 24;

  // global feature position from the joint estimate (msckf_update.cpp:164-166, 283-304)
  double Gf[3];
  {
    const double* ql = mp.xv + XV_ARR + 3 * M + 4 * (np - 1);
    const double* pl = mp.xv + XV_ARR + 3 * (np - 1);
    double Rl[9], t3[3];
    xb_rot(ql, Rl);
    const double ab1[3] = {mp.ivd[3 * g], mp.ivd[3 * g + 1], 1.0};
    xb_mv33(Rl, ab1, t3);
    for (int e = 0; e < 3; ++e) Gf[e] = 1.0 / mp.ivd[3 * g + 2] * t3[e] + pl[e];
  }

  // ---- own agent: M_0 = B_0 P B_0^T, F_0, b_0 (B_0 = A_up^T jac_j was written by k_tracks)
  const double* B0 = mp.B + (size_t)trk * 3 * W;
  {
    double m0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = threadIdx.x; c < 6 * M; c += 32 * MM_NW) {
      double t0 = 0.0, t1 = 0.0, t2 = 0.0;
      for (int i = 0; i < L; ++i)
        for (int blk = 0; blk < 2; ++blk)
          for (int a = 0; a < 3; ++a) {
            const int r = blk * 3 * M + 3 * (i1 + i) + a;
            const double p = mp.P[(size_t)(XB_CORE + r) * N + XB_CORE + c];
            t0 = fma(B0[r], p, t0);
            t1 = fma(B0[W + r], p, t1);
            t2 = fma(B0[2 * W + r], p, t2);
          }
      for (int v = 0; v < 3; ++v) {
        const double bv = B0[v * W + c];
        m0[v] = fma(t0, bv, m0[v]);
        m0[3 + v] = fma(t1, bv, m0[3 + v]);
        m0[6 + v] = fma(t2, bv, m0[6 + v]);
      }
    }
    for (int e = 0; e < 9; ++e) m0[e] = xb_warp_sum(m0[e]);
    if (lane == 0)
      for (int e = 0; e < 9; ++e) Mpart[warp][e] = m0[e];
    __syncthreads();
    if (is_w0 && lane == 0) {
      for (int e = 0; e < 9; ++e) {
        double v = 0.0;
        for (int w = 0; w < MM_NW; ++w) v += Mpart[w][e];   // fixed order
        Mi[e] = v;
        Fs[e] = mp.F0[9 * (size_t)g + e];
      }
      for (int u = 0; u < 3; ++u) bs[u] = B0[u * W + 6 * M];
    }
  }
  __syncthreads();

  // ---- peers: processOneTrack(..., is_multi_msckf = true) on the peer's poses with the joint feature position
  for (int e = 0; e < k; ++e) {
    const int* en = mp.ent + 3 * (e0 + e);
    const int Lp = en[2], s0 = M - Lp, R2 = 2 * Lp, C6 = 6 * Lp;
    const double* pp = mp.gathered + (size_t)en[0] * mp.pp_len;
    const double* ppos = pp + 8;
    const double* pq = pp + 8 + 3 * M;
    const double* pcov = pp + 8 + 7 * M;  // 6M x 6M row-major
    const double* z = mp.pobs + 2 * (size_t)en[1];
    double ur[3] = {0, 0, 0};
    double fi[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (is_w0) {
    for (int i = lane; i < Lp; i += 32) {
      double R[9];
      xb_rot(pq + 4 * (s0 + i), R);
      const double* pc = ppos + 3 * (s0 + i);
      const double dG[3] = {Gf[0] - pc[0], Gf[1] - pc[1], Gf[2] - pc[2]};
      double cp[3];
      xb_mtv33(R, dG, cp);
      res[2 * i] = z[2 * i] - cp[0] / cp[2];
      res[2 * i + 1] = z[2 * i + 1] - cp[1] / cp[2];
      const double Ji[6] = {1.0 / cp[2], 0.0, -cp[0] / (cp[2] * cp[2]), 0.0, 1.0 / cp[2], -cp[1] / (cp[2] * cp[2])};
      double Rt[9];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Rt[r * 3 + c] = R[c * 3 + r];
      double Jpos[6], Jatt[6], sk[9];
      xb_mm23(Ji, Rt, Jpos);
      for (int q = 0; q < 6; ++q) Jpos[q] = -Jpos[q];
      xb_skew(cp, sk);
      xb_mm23(Ji, sk, Jatt);
      // observability-constrained projection (msckf_update.cpp:393-406), g hard-coded
      if (mp.oc) {
        const double gv[3] = {0.0, 0.0, -9.81};
        double u[3], t2[2];
        xb_mv33(R, gv, u);
        double uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
        for (int r = 0; r < 2; ++r) t2[r] = (Jpos[r * 3] * u[0] + Jpos[r * 3 + 1] * u[1] + Jpos[r * 3 + 2] * u[2]) * (1.0 / uu);
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 3; ++c) Jpos[r * 3 + c] -= t2[r] * u[c];
        double skd[9];
        xb_skew(dG, skd);
        xb_mv33(skd, gv, u);
        uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
        for (int r = 0; r < 2; ++r) t2[r] = (Jatt[r * 3] * u[0] + Jatt[r * 3 + 1] * u[1] + Jatt[r * 3 + 2] * u[2]) * (1.0 / uu);
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 3; ++c) Jatt[r * 3 + c] -= t2[r] * u[c];
      }
      for (int q = 0; q < 6; ++q) { Jp[6 * i + q] = Jpos[q]; Ja[6 * i + q] = Jatt[q]; Hf[6 * i + q] = -Jpos[q]; U[6 * i + q] = -Jpos[q]; }
    }
    __syncwarp();
    // U = orth(Hf), modified Gram-Schmidt with re-orthogonalisation (any basis of range(Hf) is equivalent)
    for (int c = 0; c < 3; ++c) {
      for (int pass = 0; pass < 2; ++pass)
        for (int p = 0; p < c; ++p) {
          const double d = mm_wdot(U + p, 3, U + c, 3, R2, lane);
          for (int i = lane; i < R2; i += 32) U[i * 3 + c] -= d * U[i * 3 + p];
          __syncwarp();
        }
      const double n2 = mm_wdot(U + c, 3, U + c, 3, R2, lane);
      const double inv = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
      for (int i = lane; i < R2; i += 32) U[i * 3 + c] *= inv;
      __syncwarp();
    }
    // B_i (compact columns), b_i, F_i
    for (int i = lane; i < Lp; i += 32) {
      const double* U0 = U + 6 * i;
      for (int u = 0; u < 3; ++u) {
        for (int c = 0; c < 3; ++c) {
          Bc[u * C6 + 3 * i + c] = U0[u] * Jp[6 * i + c] + U0[3 + u] * Jp[6 * i + 3 + c];
          Bc[u * C6 + 3 * Lp + 3 * i + c] = U0[u] * Ja[6 * i + c] + U0[3 + u] * Ja[6 * i + 3 + c];
        }
        ur[u] += U0[u] * res[2 * i] + U0[3 + u] * res[2 * i + 1];
      }
    }
    for (int u = 0; u < 3; ++u) ur[u] = xb_warp_sum(ur[u]);
    for (int u = 0; u < 3; ++u)
      for (int c = 0; c < 3; ++c) fi[u * 3 + c] = mm_wdot(U + u, 3, Hf + c, 3, R2, lane);
    }  // warp 0
    __syncthreads();
    // M_i = B_i P_i B_i^T on the peer's pose covariance block (all warps)
    double mi[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int cc = threadIdx.x; cc < C6; cc += 32 * MM_NW) {
      const int gc = (cc < 3 * Lp) ? 3 * s0 + cc : 3 * M + 3 * s0 + (cc - 3 * Lp);
      double t0 = 0.0, t1 = 0.0, t2 = 0.0;
      for (int rr = 0; rr < C6; ++rr) {
        const int gr = (rr < 3 * Lp) ? 3 * s0 + rr : 3 * M + 3 * s0 + (rr - 3 * Lp);
        const double p = pcov[(size_t)gr * 6 * M + gc];
        t0 = fma(Bc[rr], p, t0);
        t1 = fma(Bc[C6 + rr], p, t1);
        t2 = fma(Bc[2 * C6 + rr], p, t2);
      }
      for (int v = 0; v < 3; ++v) {
        const double bv = Bc[v * C6 + cc];
        mi[v] = fma(t0, bv, mi[v]);
        mi[3 + v] = fma(t1, bv, mi[3 + v]);
        mi[6 + v] = fma(t2, bv, mi[6 + v]);
      }
    }
    for (int q = 0; q < 9; ++q) mi[q] = xb_warp_sum(mi[q]);
    if (lane == 0)
      for (int q = 0; q < 9; ++q) Mpart[warp][q] = mi[q];
    __syncthreads();
    if (is_w0 && lane == 0) {
      for (int q = 0; q < 9; ++q) {
        double v = 0.0;
        for (int w = 0; w < MM_NW; ++w) v += Mpart[w][q];   // fixed order
        Mi[9 * (e + 1) + q] = v;
        Fs[9 * (e + 1) + q] = fi[q];
      }
      for (int u = 0; u < 3; ++u) bs[3 * (e + 1) + u] = ur[u];
    }
    __syncthreads();
  }
  if (!is_w0) return;   // nullspace, gate and CI fusion are 3 (k + 1) x 3 problems: one warp
  (void)Tc;

  // ---- nullSpaceProjection (msckf_update.cpp:494-501): full Householder Q of the stacked 3(k+1) x 3 feature Jacobian
  const int n = 3 * (k + 1), m = 3 * k;
  if (lane == 0) {
    for (int j = 0; j < 3; ++j) {
      double nx = 0.0;
      for (int r = j; r < n; ++r) nx += Fs[r * 3 + j] * Fs[r * 3 + j];
      nx = sqrt(nx);
      for (int r = 0; r < 24; ++r) Vh[j * 24 + r] = 0.0;
      if (nx == 0.0) { tau[j] = 0.0; continue; }
      const double al = Fs[j * 3 + j];
      const double be = -copysign(nx, al);
      const double v0 = al - be;
      Vh[j * 24 + j] = 1.0;
      for (int r = j + 1; r < n; ++r) Vh[j * 24 + r] = Fs[r * 3 + j] / v0;
      tau[j] = (be - al) / be;
      for (int c = j + 1; c < 3; ++c) {
        double d = 0.0;
        for (int r = j; r < n; ++r) d += Vh[j * 24 + r] * Fs[r * 3 + c];
        d *= tau[j];
        for (int r = j; r < n; ++r) Fs[r * 3 + c] -= d * Vh[j * 24 + r];
      }
    }
  }
  __syncwarp();
  if (lane < n) {  // column `lane` of Q = H_0 H_1 H_2
    double q[24];
#pragma unroll
    for (int r = 0; r < 24; ++r) q[r] = (r == lane) ? 1.0 : 0.0;
    for (int j = 2; j >= 0; --j) {
      double d = 0.0;
#pragma unroll
      for (int r = 0; r < 24; ++r) d = fma(Vh[j * 24 + r], q[r], d);
      d *= tau[j];
#pragma unroll
      for (int r = 0; r < 24; ++r) q[r] = fma(-d, Vh[j * 24 + r], q[r]);
    }
#pragma unroll
    for (int r = 0; r < 24; ++r) Q[r * MM_LDQ + lane] = q[r];
  }
  __syncwarp();
  // res_pf = A^T b, A = Q[:, 3:]
  if (lane < m) {
    double s = 0.0;
    for (int r = 0; r < n; ++r) s = fma(Q[r * MM_LDQ + 3 + lane], bs[r], s);
    rs[lane] = s;
  }
  __syncwarp();
  const double w0 = 1.0 - (double)k * mp.w_other;  // ci.cpp:64-74 (fixed weights)
  const double w_result = 1.0 / w0;
  // S(a,b) = sum_i c_i A_i^T M_i A_i + var I
  auto build_s = [&](double c_own, double c_peer, int w) {
    for (int ab = lane; ab < m * m; ab += 32) {
      const int a = ab / m, b = ab % m;
      double s = (a == b) ? mp.var_img : 0.0;
      for (int i = 0; i <= k; ++i) {
        const double ci = (i == 0) ? c_own : c_peer;
        double acc = 0.0;
        for (int u = 0; u < 3; ++u)
          for (int v = 0; v < 3; ++v) acc = fma(Q[(3 * i + u) * MM_LDQ + 3 + a] * Mi[9 * i + u * 3 + v], Q[(3 * i + v) * MM_LDQ + 3 + b], acc);
        s = fma(ci, acc, s);
      }
      Sg[a * MM_LDG + b] = s;
    }
    for (int a = lane; a < m; a += 32) {
      Sg[a * MM_LDG + w - 1] = rs[a];
      if (w > m + 1)
        for (int b = 0; b < m; ++b) Sg[a * MM_LDG + m + b] = (a == b) ? 1.0 : 0.0;
    }
    __syncwarp();
  };
  // gate: gamma = res^T S^-1 res < chi2_0.95(2 n_obs,all - 3)   (msckf_update.cpp:240-247)
  build_s(1.0, 1.0, m + 1);
  const bool ok = mm_gj(Sg, m, m + 1, lane);
  double gamma = 0.0;
  for (int a = 0; a < m; ++a) gamma = fma(rs[a], Sg[a * MM_LDG + m], gamma);
  if (!ok) gamma = NAN;
  if (lane == 0) rec[1] = gamma;
  (void)n_tot;
  if (!(gamma < mp.chi2[g])) return;
  // CI-fused S (ci.cpp:76-84) + the second noise term (msckf_update.cpp:255)
  __syncwarp();
  build_s(w_result, 1.0 / mp.w_other, 2 * m + 1);
  if (!mm_gj(Sg, m, 2 * m + 1, lane)) return;
  if (lane < m) ys[lane] = Sg[lane * MM_LDG + 2 * m];
  __syncwarp();
  if (lane < 3) {  // a = A_0 S^-1 res
    double s = 0.0;
    for (int b = 0; b < m; ++b) s = fma(Q[lane * MM_LDQ + 3 + b], ys[b], s);
    rec[4 + lane] = s;
  }
  if (lane < 9) {  // C3 = A_0 S^-1 A_0^T
    const int u = lane / 3, v = lane % 3;
    double s = 0.0;
    for (int a = 0; a < m; ++a) {
      double t = 0.0;
      for (int b = 0; b < m; ++b) t = fma(Sg[a * MM_LDG + m + b], Q[v * MM_LDQ + 3 + b], t);
      s = fma(Q[u * MM_LDQ + 3 + a], t, s);
    }
    rec[7 + lane] = s;
  }
  if (lane == 0) {
    rec[3] = w_result;
    rec[0] = 1.0;
    atomicMax(mp.last, g);
  }
}

void launch_mm_triangulate(cudaStream_t s, const MmParams& mp) {
  if (mp.n_groups <= 0) return;
  k_mm_triangulate<<<(mp.n_groups + 3) / 4, 128, 0, s>>>(mp);
  count_launch();
}
int launch_mm_construct(cudaStream_t s, const MmParams& mp) {
  if (mp.n_groups <= 0) return 0;
  const size_t bytes = sizeof(double) * mm_smem_doubles(mp.M);
  if (bytes > 200 * 1024) return -1;
  cudaMemsetAsync(mp.last, 0xFF, sizeof(int), s);
  cudaFuncSetAttribute(k_mm_construct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  k_mm_construct<<<mp.n_groups, 32 * MM_NW, bytes, s>>>(mp);
  count_launch();
  return 0;
}

// ---- applyCI over the list ----------------------------------------------------------------------------------------------
// is P[i][j] inside a diagonal 3x3 pose block that P_j scales (msckf_update.cpp:258-267)?
__device__ __forceinline__ bool mm_scaled(int i, int j, int M, int i1, int np) {
  const int a = i - XB_CORE, b = j - XB_CORE;
  if (a < 0 || b < 0 || a >= 6 * M || b >= 6 * M) return false;
  if (a / 3 != b / 3) return false;
  const int slot = (a / 3) % M;
  return slot >= i1 && slot < np;
}

// V[:, g] = B_0^T a   (the direction h_j^T S^-1 res of entry g; zero column for rejected entries)
__global__ void k_mm_v(MmParams mp, double* __restrict__ V, int ldv) {
  const int g = blockIdx.x;
  const double* rec = mp.rec + (size_t)XB_MM_REC * g;
  const int W = 6 * mp.M + 1;
  const bool inl = rec[0] == 1.0;
  const double* B0 = mp.B + (size_t)((int)rec[16]) * 3 * W;
  for (int c = threadIdx.x; c < 6 * mp.M; c += blockDim.x)
    V[(size_t)c * ldv + g] = inl ? B0[c] * rec[4] + B0[W + c] * rec[5] + B0[2 * W + c] * rec[6] : 0.0;
}
// D[:, g] += (w - 1) * (diagonal-block part of P) v_g on the scaled rows: delta_g = P_j h_j^T S^-1 res
__global__ void k_mm_fix(MmParams mp, const double* __restrict__ P, const double* __restrict__ V, int ldv, double* __restrict__ D) {
  const int g = blockIdx.x;
  const double* rec = mp.rec + (size_t)XB_MM_REC * g;
  if (rec[0] != 1.0) return;
  const int M = mp.M, N = mp.N, i1 = (int)rec[17], L = (int)rec[18];
  const double w = rec[3];
  for (int t = threadIdx.x; t < 6 * L; t += blockDim.x) {
    const int blk = t / (3 * L), rem = t % (3 * L);
    const int c0 = blk * 3 * M + 3 * (i1 + rem / 3);  // first pose column of the 3x3 block
    const int i = XB_CORE + c0 + rem % 3;
    double s = 0.0;
    for (int c = 0; c < 3; ++c) s = fma(P[(size_t)i * N + XB_CORE + c0 + c], V[(size_t)(c0 + c) * ldv + g], s);
    D[(size_t)i * ldv + g] += (w - 1.0) * s;
  }
}
// sequential State::correct for every inlier entry in list order (updater.cpp:88-92 -> applyCI -> state.correct)
__global__ void k_mm_correct_seq(int M, int F, int N, const double* __restrict__ rec, int n_groups,
                                 const double* __restrict__ D, int ldv, double* __restrict__ xv) {
  // State::correct once per inlier entry, in list order (updater.cpp:144-161).  Every thread keeps ITS state entries in
  // registers over the whole list; the inlier flags are fetched once into shared memory and the corrections of four entries
  // are loaded at a time, unconditionally -- a read-modify-write through L2 per entry behind a dependent branch was two
  // round trips (1.4 us) per entry: 0.3 ms of the 8-agent update (224 entries).
  __shared__ unsigned char inl[1024];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int g = threadIdx.x; g < n_groups && g < 1024; g += blockDim.x) inl[g] = rec[(size_t)XB_MM_REC * g] == 1.0;
  __syncthreads();
  double c4[4] = {0, 0, 0, 0}, pa = 0.0, fa = 0.0, q[4] = {0, 0, 0, 1};
  double* qp = (t == M) ? xv + XV_Q : xv + XV_ARR + 3 * M + 4 * t;
  const int qb = (t == M) ? 6 : XB_CORE + 3 * M + 3 * t;
  const bool hc = t < 3, hp = t < 3 * M, hf = t < 3 * F, hq = t <= M;
  if (hc) { c4[0] = xv[XV_P + t]; c4[1] = xv[XV_V + t]; c4[2] = xv[XV_BW + t]; c4[3] = xv[XV_BA + t]; }
  if (hp) pa = xv[XV_ARR + t];
  if (hf) fa = xv[XV_ARR + 7 * M + t];
  if (hq) { q[0] = qp[0]; q[1] = qp[1]; q[2] = qp[2]; q[3] = qp[3]; }
  for (int g0 = 0; g0 < n_groups; g0 += 4) {
    double dc[4][4], dp[4], df[4], dq3[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int g = min(g0 + u, n_groups - 1);
      auto d = [&](int i) { return D[(size_t)i * ldv + g]; };
      if (hc) { dc[u][0] = d(t); dc[u][1] = d(3 + t); dc[u][2] = d(9 + t); dc[u][3] = d(12 + t); }
      if (hp) dp[u] = d(XB_CORE + t);
      if (hf) df[u] = d(XB_CORE + 6 * M + t);
      if (hq) { dq3[u][0] = d(qb); dq3[u][1] = d(qb + 1); dq3[u][2] = d(qb + 2); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int g = g0 + u;
      if (g >= n_groups || (g < 1024 ? !inl[g] : rec[(size_t)XB_MM_REC * g] != 1.0)) continue;
      if (hc) { c4[0] += dc[u][0]; c4[1] += dc[u][1]; c4[2] += dc[u][2]; c4[3] += dc[u][3]; }
      if (hp) pa += dp[u];
      if (hf) fa += df[u];
      if (hq) {
        double dq[4], qo[4];
        xb_small_angle_quat(dq3[u], dq);
        xb_qmul(q, dq, qo);
        const double nq = sqrt(qo[0] * qo[0] + qo[1] * qo[1] + qo[2] * qo[2] + qo[3] * qo[3]);
        if (nq > 0.0) { qo[0] /= nq; qo[1] /= nq; qo[2] /= nq; qo[3] /= nq; }
        q[0] = qo[0]; q[1] = qo[1]; q[2] = qo[2]; q[3] = qo[3];
      }
    }
  }
  if (hc) { xv[XV_P + t] = c4[0]; xv[XV_V + t] = c4[1]; xv[XV_BW + t] = c4[2]; xv[XV_BA + t] = c4[3]; }
  if (hp) xv[XV_ARR + t] = pa;
  if (hf) xv[XV_ARR + 7 * M + t] = fa;
  if (hq) { qp[0] = q[0]; qp[1] = q[1]; qp[2] = q[2]; qp[3] = q[3]; }
}
// K3 = (P_j B_0^T) C3 for the last inlier entry: one warp per row
__global__ void __launch_bounds__(128) k_mm_k3(MmParams mp, const double* __restrict__ P, double* __restrict__ K3) {
  const int gl = *mp.last;
  if (gl < 0) return;
  const int lane = threadIdx.x & 31, i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int M = mp.M, N = mp.N, W = 6 * M + 1;
  if (i >= N) return;
  const double* rec = mp.rec + (size_t)XB_MM_REC * gl;
  const int i1 = (int)rec[17], np = i1 + (int)rec[18];
  const double w = rec[3];
  const double* B0 = mp.B + (size_t)((int)rec[16]) * 3 * W;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int c = lane; c < 6 * M; c += 32) {
    double p = P[(size_t)i * N + XB_CORE + c];
    if (mm_scaled(i, XB_CORE + c, M, i1, np)) p *= w;
    a0 = fma(p, B0[c], a0);
    a1 = fma(p, B0[W + c], a1);
    a2 = fma(p, B0[2 * W + c], a2);
  }
  a0 = xb_warp_sum(a0); a1 = xb_warp_sum(a1); a2 = xb_warp_sum(a2);
  if (lane < 3) K3[(size_t)i * 3 + lane] = a0 * rec[7 + lane] + a1 * rec[10 + lane] + a2 * rec[13 + lane];
}
// HP3 = B_0 P_j (3 x N): one thread per column
__global__ void k_mm_hp3(MmParams mp, const double* __restrict__ P, double* __restrict__ HP3) {
  const int gl = *mp.last;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int M = mp.M, N = mp.N, W = 6 * M + 1;
  if (gl < 0 || j >= N) return;
  const double* rec = mp.rec + (size_t)XB_MM_REC * gl;
  const int i1 = (int)rec[17], L = (int)rec[18], np = i1 + L;
  const double w = rec[3];
  const double* B0 = mp.B + (size_t)((int)rec[16]) * 3 * W;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int i = 0; i < L; ++i)
    for (int blk = 0; blk < 2; ++blk)
      for (int a = 0; a < 3; ++a) {
        const int r = blk * 3 * M + 3 * (i1 + i) + a;
        double p = P[(size_t)(XB_CORE + r) * N + j];
        if (mm_scaled(XB_CORE + r, j, M, i1, np)) p *= w;
        a0 = fma(B0[r], p, a0);
        a1 = fma(B0[W + r], p, a1);
        a2 = fma(B0[2 * W + r], p, a2);
      }
  HP3[j] = a0;
  HP3[(size_t)N + j] = a1;
  HP3[(size_t)2 * N + j] = a2;
}
// P <- sym((I - K h_j) P_j) of the last inlier entry (updater.cpp:155-157)
__global__ void __launch_bounds__(256) k_mm_cov_last(MmParams mp, double* __restrict__ P, const double* __restrict__ K3,
                                                     const double* __restrict__ HP3) {
  const int gl = *mp.last;
  if (gl < 0) return;
  const int N = mp.N;
  const int j = blockIdx.x * 16 + (threadIdx.x & 15), i = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (i >= N || j >= N || i > j) return;
  const double* rec = mp.rec + (size_t)XB_MM_REC * gl;
  const int i1 = (int)rec[17], np = i1 + (int)rec[18];
  double pij = P[(size_t)i * N + j], pji = P[(size_t)j * N + i];
  if (mm_scaled(i, j, mp.M, i1, np)) { pij *= rec[3]; pji *= rec[3]; }
  double a = 0.0, b = 0.0;
  for (int u = 0; u < 3; ++u) {
    a = fma(K3[(size_t)i * 3 + u], HP3[(size_t)u * N + j], a);
    b = fma(K3[(size_t)j * 3 + u], HP3[(size_t)u * N + i], b);
  }
  const double v = 0.5 * ((pij - a) + (pji - b));
  P[(size_t)i * N + j] = v;
  P[(size_t)j * N + i] = v;
}

void launch_mm_apply(cudaStream_t s, const MmParams& mp, double* P, double* xv, int F, double* V, int ldv, double* D,
                     double* K3, double* HP3) {
  const int G = mp.n_groups, N = mp.N, M = mp.M;
  if (G <= 0) return;
  k_mm_v<<<G, 64, 0, s>>>(mp, V, ldv);
  count_launch();
  gemm_nn(s, N, G, 6 * M, 1.0, P + XB_CORE, N, V, ldv, 0.0, D, ldv);
  k_mm_fix<<<G, 64, 0, s>>>(mp, P, V, ldv, D);
  count_launch();
  k_mm_k3<<<(N + 3) / 4, 128, 0, s>>>(mp, P, K3);
  count_launch();
  k_mm_hp3<<<(N + 127) / 128, 128, 0, s>>>(mp, P, HP3);
  count_launch();
  dim3 gc((N + 15) / 16, (N + 15) / 16);
  k_mm_cov_last<<<gc, 256, 0, s>>>(mp, P, K3, HP3);
  count_launch();
  k_mm_correct_seq<<<(N + 127) / 128, 128, 0, s>>>(M, F, N, mp.rec, G, D, ldv, xv);
  count_launch();
}

// ---- pose payload: what an agent publishes for MSCKF-MSCKF matches ---------------------------------------------------
// [0]=1 (valid) [1]=time [2]=M | camera positions 3M | camera attitudes 4M | P[pose, pose] 6M x 6M row-major
__global__ void k_pack_poses(const double* __restrict__ xv, const double* __restrict__ P, int N, int M, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n6 = 6 * M;
  if (t < 8) out[t] = (t == 0) ? 1.0 : (t == 1) ? xv[XV_TIME] : (t == 2) ? (double)M : 0.0;
  if (t < 7 * M) out[8 + t] = xv[XV_ARR + t];
  if (t < n6 * n6) {
    const int r = t / n6, c = t % n6;
    out[8 + 7 * M + t] = P[(size_t)(XB_CORE + r) * N + XB_CORE + c];
  }
}
void launch_pack_poses(cudaStream_t s, const double* xv, const double* P, int N, int M, double* out) {
  const int n = std::max(36 * M * M, 7 * M + 8);
  k_pack_poses<<<(n + 255) / 256, 256, 0, s>>>(xv, P, N, M, out);
  count_launch();
}

}  // namespace xb
