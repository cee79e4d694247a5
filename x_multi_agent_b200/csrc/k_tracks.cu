// Per-track MSCKF / MSCKF-SLAM measurement construction: triangulation, reprojection Jacobians,
// left-nullspace projection and chi-square gating -- one warp per track.
//
// reference: src/x/vision/triangulation.cpp:48-206 (DLT + Gauss-Newton),
//            src/x/vio/msckf_update.cpp:283-492 (Jacobians, OC projection, nullspace, gate),
//            src/x/vio/msckf_slam_update.cpp:64-267 (Li 2012 promotion: H1, H2, r1),
//            src/x/vio/slam_update.cpp:49-214 (inverse-depth SLAM rows).
//
// Formulation.  With U an orthonormal basis of range(Hf_j) (2L x 3) and A its orthogonal complement,
// the reference stacks jac0 = A^T J, res0 = A^T r.  Everything downstream depends on A only through
// the projector Pi = A A^T = I - U U^T:
//   gate   gamma = (Pi r)^T (Pi J P J^T Pi + s^2 I)^-1 (Pi r)       (2L x 2L, identical value)
//   update jac0^T jac0 = J^T J - B^T B,  jac0^T res0 = J^T r - B^T (U^T r),   B = U^T J  (3 x 6M)
// so the 2L-3 dense rows are never formed: a track emits its sparse J blocks and the 3 dense rows B.
#include <algorithm>
#include <cstdlib>

#include "xb_kernels.h"
#include "xb_svd4.cuh"

namespace xb {

__device__ __forceinline__ int tri_idx(int r, int c) { return r * (r + 1) / 2 + c; }  // r >= c

// Per-warp shared-memory carve-up (doubles), Lm = max track length handled by the launch.
// Y overlays Hf (Hf is dead once U and H2 exist); the anchor blocks exist only in MSCKF-SLAM mode.
struct WarpSmem {
  double *Jp, *Ja, *Jap, *Jaa, *Hf, *U, *Y, *V, *res, *X, *scr;
  __device__ WarpSmem(double* base, int Lm, int mode) {
    scr = base; base += 128;  // small-matrix scratch (6x6 Woodbury solve)
    Jp = base; base += 6 * Lm;
    Ja = base; base += 6 * Lm;
    Hf = base; Y = base; base += 6 * Lm;
    U = base; base += 6 * Lm;
    V = base; base += 6 * Lm;
    res = base; base += 2 * Lm;
    Jap = base; Jaa = base;
    if (mode == 1) { Jaa = base + 6 * Lm; base += 12 * Lm; }
    X = base;  // (2Lm+7)(2Lm+8)/2 packed lower triangle incl. 7 augmented rows (Pi r and the 6 clone columns)
  }
  static __host__ __device__ size_t doubles(int Lm, int mode) {
    return 128 + (size_t)(mode == 1 ? 44 : 32) * Lm + (size_t)(2 * Lm + 7) * (2 * Lm + 8) / 2;
  }
};

// dot of two length-n smem vectors with stride, over the warp
__device__ __forceinline__ double wdot(const double* a, int sa, const double* b, int sb, int n, int lane) {
  double s = 0.0;
  for (int i = lane; i < n; i += 32) s = fma(a[i * sa], b[i * sb], s);
  return xb_warp_sum(s);
}

// advance a packed-lower index (r, c) by n elements (row-major order over the lower triangle)
__device__ __forceinline__ void tri_advance(int& r, int& c, int n) {
  c += n;
  while (c > r) { c -= r + 1; ++r; }
}

// ------------------------------------------------------------------------------------------------
// Cholesky of the packed lower triangle X: R2 factor rows followed by 7 right-hand-side rows R2..R2+6 that are
// solved along (they become y = L^-1 (Pi r) and Vt = L^-1 V); row q starts at tri_idx(q, 0).  One warp, LEFT-looking.
// Measured per cfg-2 track (R2 = 60): right-looking row-per-lane 182k cycles, right-looking column-per-lane 193k (every
// iteration waits for its own store), fully unrolled register tiles 128k (instruction-fetch bound: the kernel runs once
// per warp), left-looking column by column with shrinking row groups 84-99k, blocked by 4 columns (below) 66k.
// ------------------------------------------------------------------------------------------------
// Blocked by 4 columns (round 2): per block step every lane forms the four dot products of each of its rows against the
// finished rows c0..c0+3 (the own-row entries are loaded once for the four columns), the owners of the rows c0..c0+3 put
// the updated 4x4 diagonal block into shared memory, every lane factors it redundantly in registers (4 dependent rsqrt:
// the serial part) and solves the four entries of its rows: one shuffle-free hand-over, two __syncwarp and one rsqrt chain
// per FOUR pivots instead of per pivot.  Row r belongs to lane r % 32.
template <int NR>  // rows per lane: ceil((R2 + 7) / 32) <= NR
__device__ __forceinline__ bool gate_chol_blocked(double* __restrict__ X, int R2, double* __restrict__ blk, int lane) {
  const int Rend = R2 + 6;
  for (int c0 = 0; c0 < R2; c0 += 4) {
    const int nb = min(4, R2 - c0);  // R2 = 2 L is even: the last block may have 2 columns
    double v[NR][4];
    const double* L0 = X + tri_idx(c0, 0);
    const double* L1 = X + tri_idx(c0 + 1, 0);
    const double* L2 = nb > 2 ? X + tri_idx(c0 + 2, 0) : L0;
    const double* L3 = nb > 2 ? X + tri_idx(c0 + 3, 0) : L0;
#pragma unroll
    for (int u = 0; u < NR; ++u) {
      const int r = lane + 32 * u;
      const bool act = r >= c0 && r <= Rend;
      v[u][0] = 0.0; v[u][1] = 0.0; v[u][2] = 0.0; v[u][3] = 0.0;
      if (act) {
        const double* row = X + tri_idx(r, 0);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll 2
        for (int k = 0; k + 1 < c0; k += 2) {
          const double a0 = row[k], a1 = row[k + 1];
          s0 = fma(a0, L0[k], s0); s1 = fma(a0, L1[k], s1); s2 = fma(a0, L2[k], s2); s3 = fma(a0, L3[k], s3);
          q0 = fma(a1, L0[k + 1], q0); q1 = fma(a1, L1[k + 1], q1); q2 = fma(a1, L2[k + 1], q2); q3 = fma(a1, L3[k + 1], q3);
        }
        // entries (r, c0 + j) exist only for c0 + j <= r (packed lower triangle)
        v[u][0] = row[c0] - (s0 + q0);
        if (r >= c0 + 1) v[u][1] = row[c0 + 1] - (s1 + q1);
        if (nb > 2 && r >= c0 + 2) v[u][2] = row[c0 + 2] - (s2 + q2);
        if (nb > 2 && r >= c0 + 3) v[u][3] = row[c0 + 3] - (s3 + q3);
        if (r < c0 + nb) {
#pragma unroll
          for (int j = 0; j < 4; ++j) blk[(r - c0) * 4 + j] = v[u][j];
        }
      }
    }
    __syncwarp();
    const double a00 = blk[0], a10 = blk[4], a11 = blk[5];
    double a20 = 0.0, a21 = 0.0, a22 = 1.0, a30 = 0.0, a31 = 0.0, a32 = 0.0, a33 = 1.0;
    if (nb > 2) { a20 = blk[8]; a21 = blk[9]; a22 = blk[10]; a30 = blk[12]; a31 = blk[13]; a32 = blk[14]; a33 = blk[15]; }
    const double r0 = rsqrt(a00);
    const double l10 = a10 * r0, l20 = a20 * r0, l30 = a30 * r0;
    const double p1 = fma(-l10, l10, a11);
    const double r1 = rsqrt(p1);
    const double l21 = fma(-l20, l10, a21) * r1, l31 = fma(-l30, l10, a31) * r1;
    const double p2 = fma(-l21, l21, fma(-l20, l20, a22));
    const double r2 = rsqrt(p2);
    const double l32 = fma(-l31, l21, fma(-l30, l20, a32)) * r2;
    const double p3 = fma(-l32, l32, fma(-l31, l31, fma(-l30, l30, a33)));
    const double r3 = rsqrt(p3);
    if (!(a00 > 0.0) || !(p1 > 0.0) || !(p2 > 0.0) || !(p3 > 0.0)) return false;  // same values in every lane
#pragma unroll
    for (int u = 0; u < NR; ++u) {
      const int r = lane + 32 * u;
      if (r >= c0 && r <= Rend) {
        double* row = X + tri_idx(r, 0);
        // for the rows of the block itself the same formulas give l_ij (j < i) and l_ii = p_i r_i
        const double x0 = v[u][0] * r0;
        const double x1 = fma(-x0, l10, v[u][1]) * r1;
        const double x2 = fma(-x1, l21, fma(-x0, l20, v[u][2])) * r2;
        const double x3 = fma(-x2, l32, fma(-x1, l31, fma(-x0, l30, v[u][3]))) * r3;
        row[c0] = x0;
        if (r >= c0 + 1) row[c0 + 1] = x1;
        if (nb > 2 && r >= c0 + 2) row[c0 + 2] = x2;
        if (nb > 2 && r >= c0 + 3) row[c0 + 3] = x3;
      }
    }
    __syncwarp();
  }
  return true;
}

template <int OPL>  // observations per lane (track length <= 32*OPL)
__global__ void __launch_bounds__(128, 3) k_tracks(TrackParams tp) {
  XB_PDL_LONG();  // <= 168 registers: leaves the register file room for the side-stream kernels
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int M = tp.M, np = tp.n_poses;
  double* Rk = smem;            // [M][9]  rot(q_k) (body -> global), normalised
  double* pk = Rk + 9 * M;      // [M][3]
  double* wbase = pk + 3 * M + warp * WarpSmem::doubles(tp.Lmax, tp.mode);
  WarpSmem ws(wbase, tp.Lmax, tp.mode);

  const double* parr = tp.xv + XV_ARR;
  const double* qarr = tp.xv + XV_ARR + 3 * M;
  for (int k = threadIdx.x; k < np; k += blockDim.x) {
    xb_rot(qarr + 4 * k, Rk + 9 * k);
    pk[3 * k] = parr[3 * k]; pk[3 * k + 1] = parr[3 * k + 1]; pk[3 * k + 2] = parr[3 * k + 2];
  }
  __syncthreads();

  const int trk = blockIdx.x * nwarp + warp;
  if (trk >= tp.n_tracks) return;
  long long* pf = tp.prof ? tp.prof + 12 * (size_t)trk : nullptr;
#define TPROF(k) do { if (pf && lane == 0) pf[k] = clock64(); } while (0)
  TPROF(0);
  const int o0 = tp.off[trk], L = tp.off[trk + 1] - o0;
  const int W = 6 * M + 1;  // width of a B row: pose columns + residual column
  double* Bt = tp.B + (size_t)trk * 3 * W;
  const int i1 = np - L;  // window slot of the first observation (msckf_update.cpp:329-331)
  const bool slam_mode = tp.mode == 1;
  bool bad = (L < 2) || (i1 < 0) || (L > 32 * OPL);

  // ---------------------------------------------------------------- triangulation (triangulation.cpp:102-206)
  const double* Rl = Rk + 9 * (np - 1);  // last pose = inverse-depth anchor
  const double* pl = pk + 3 * (np - 1);
  double alpha = 0.0, beta = 0.0, rho = 1.0;
  // MULTI_UAV: a track matched with other agents' tracks is triangulated jointly with their observations
  // (msckf_update.cpp:96-166) by k_mm_triangulate; its inverse-depth estimate arrives through mm_ivd.
  const int mm_g = tp.mm_grp ? tp.mm_grp[trk] : -1;
  if (!bad && mm_g >= 0) {
    alpha = tp.mm_ivd[3 * mm_g]; beta = tp.mm_ivd[3 * mm_g + 1]; rho = tp.mm_ivd[3 * mm_g + 2];
  } else if (!bad) {
    if (lane == 0) {
      // projection matrices [R^T | -R^T p] of the first and last pose (triangulation.cpp:208-216)
      const double* z1 = tp.obs + 2 * (size_t)o0;
      const double* z2 = tp.obs + 2 * (size_t)(o0 + L - 1);
      const double* R1 = Rk + 9 * i1;
      const double* p1 = pk + 3 * i1;
      double A[16], P1[12], P2[12];
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) { P1[r * 4 + c] = R1[c * 3 + r]; P2[r * 4 + c] = Rl[c * 3 + r]; }
        P1[r * 4 + 3] = -(R1[0 * 3 + r] * p1[0] + R1[1 * 3 + r] * p1[1] + R1[2 * 3 + r] * p1[2]);
        P2[r * 4 + 3] = -(Rl[0 * 3 + r] * pl[0] + Rl[1 * 3 + r] * pl[1] + Rl[2 * 3 + r] * pl[2]);
      }
      for (int c = 0; c < 4; ++c) {
        A[0 + c] = z1[0] * P1[8 + c] - P1[0 + c];
        A[4 + c] = z1[1] * P1[8 + c] - P1[4 + c];
        A[8 + c] = z2[0] * P2[8 + c] - P2[0 + c];
        A[12 + c] = z2[1] * P2[8 + c] - P2[4 + c];
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
    TPROF(1);

    // iteration-invariant relative poses of this lane's observations
    double dR[OPL][9], dp[OPL][3], zz[OPL][2];
#pragma unroll
    for (int o = 0; o < OPL; ++o) {
      const int i = lane + 32 * o;
      if (i < L) {
        const double* Ri = Rk + 9 * (i1 + i);
        const double* pi = pk + 3 * (i1 + i);
        // rot = R_i^T ; delta_rot = rot * rot_a^T = R_i^T R_l ; delta_pos = rot p_a - rot p_i
        for (int r = 0; r < 3; ++r) {
          for (int c = 0; c < 3; ++c)
            dR[o][r * 3 + c] = Ri[0 * 3 + r] * Rl[0 * 3 + c] + Ri[1 * 3 + r] * Rl[1 * 3 + c] + Ri[2 * 3 + r] * Rl[2 * 3 + c];
          const double a = Ri[0 * 3 + r] * pl[0] + Ri[1 * 3 + r] * pl[1] + Ri[2 * 3 + r] * pl[2];
          const double b = Ri[0 * 3 + r] * pi[0] + Ri[1 * 3 + r] * pi[1] + Ri[2 * 3 + r] * pi[2];
          dp[o][r] = a - b;
        }
        zz[o][0] = tp.obs[2 * (size_t)(o0 + i)];
        zz[o][1] = tp.obs[2 * (size_t)(o0 + i) + 1];
      }
    }
    double r_norm_last = 1000.0, r_norm = 100.0;
    int iter = 0;
    while (r_norm_last - r_norm > tp.gn_term) {
      ++iter;
      if (iter > tp.gn_max_iter) break;
      double jtj[6] = {0, 0, 0, 0, 0, 0}, jtr[3] = {0, 0, 0}, rr = 0.0;
#pragma unroll
      for (int o = 0; o < OPL; ++o) {
        const int i = lane + 32 * o;
        if (i < L) {
          double h[3];
          for (int r = 0; r < 3; ++r)
            h[r] = dR[o][r * 3] * alpha + dR[o][r * 3 + 1] * beta + dR[o][r * 3 + 2] + rho * dp[o][r];
          const double r0 = zz[o][0] - h[0] / h[2], r1 = zz[o][1] - h[1] / h[2];
          const double j1a = -1.0 / h[2], j1b = h[0] / (h[2] * h[2]), j1c = h[1] / (h[2] * h[2]);
          // J = j1 * j0, j0 = [dR(:,0) dR(:,1) dp]
          double J0[3], J1[3];
          const double c0[3] = {dR[o][0], dR[o][3], dR[o][6]}, c1[3] = {dR[o][1], dR[o][4], dR[o][7]};
          J0[0] = j1a * c0[0] + j1b * c0[2]; J0[1] = j1a * c1[0] + j1b * c1[2]; J0[2] = j1a * dp[o][0] + j1b * dp[o][2];
          J1[0] = j1a * c0[1] + j1c * c0[2]; J1[1] = j1a * c1[1] + j1c * c1[2]; J1[2] = j1a * dp[o][1] + j1c * dp[o][2];
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
      }
#pragma unroll
      for (int e = 0; e < 6; ++e) jtj[e] = xb_warp_sum(jtj[e]);
#pragma unroll
      for (int e = 0; e < 3; ++e) jtr[e] = xb_warp_sum(jtr[e]);
      rr = xb_warp_sum(rr);
      const double Am[9] = {jtj[0], jtj[1], jtj[2], jtj[1], jtj[3], jtj[4], jtj[2], jtj[4], jtj[5]};
      double Ai[9];
      xb_inv33(Am, Ai);
      double d[3];
      xb_mv33(Ai, jtr, d);
      alpha -= d[0];
      beta -= d[1];
      rho -= d[2];
      r_norm_last = r_norm;
      r_norm = sqrt(rr);
    }
  }
  TPROF(2);
  if (lane == 0) { tp.ivd[3 * trk] = alpha; tp.ivd[3 * trk + 1] = beta; tp.ivd[3 * trk + 2] = rho; }

  // ---------------------------------------------------------------- Jacobians
  // global feature position (msckf_update.cpp:283-304): 1/rho * R_l (alpha,beta,1) + p_l
  double Gf[3];
  {
    const double ab1[3] = {alpha, beta, 1.0};
    double t3[3];
    xb_mv33(Rl, ab1, t3);
    for (int e = 0; e < 3; ++e) Gf[e] = 1.0 / rho * t3[e] + pl[e];
  }
  int nan_flag = 0;
  for (int i = lane; i < L && !bad; i += 32) {
    const int pos = i1 + i;
    const double* R = Rk + 9 * pos;
    const double* pc = pk + 3 * pos;
    const double dG[3] = {Gf[0] - pc[0], Gf[1] - pc[1], Gf[2] - pc[2]};
    double cp[3];
    xb_mtv33(R, dG, cp);  // R^T (G_p_f - p_c)
    if (!(cp[0] == cp[0] && cp[1] == cp[1] && cp[2] == cp[2])) nan_flag = 1;
    const double z0 = tp.obs[2 * (size_t)(o0 + i)], z1 = tp.obs[2 * (size_t)(o0 + i) + 1];
    ws.res[2 * i] = z0 - cp[0] / cp[2];
    ws.res[2 * i + 1] = z1 - cp[1] / cp[2];
    const double Ji[6] = {1.0 / cp[2], 0.0, -cp[0] / (cp[2] * cp[2]), 0.0, 1.0 / cp[2], -cp[1] / (cp[2] * cp[2])};
    double Rt[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Rt[r * 3 + c] = R[c * 3 + r];
    double Jpos[6], Jatt[6], sk[9];
    xb_mm23(Ji, Rt, Jpos);
    for (int e = 0; e < 6; ++e) Jpos[e] = -Jpos[e];
    xb_skew(cp, sk);
    xb_mm23(Ji, sk, Jatt);
    double* Jp = ws.Jp + 6 * i;
    double* Ja = ws.Ja + 6 * i;
    double* Jap = ws.Jap + 6 * i;
    double* Jaa = ws.Jaa + 6 * i;
    double* Hf = ws.Hf + 6 * i;
    if (!slam_mode && !tp.oc) {
      for (int e = 0; e < 6; ++e) { Jp[e] = Jpos[e]; Ja[e] = Jatt[e]; Hf[e] = -Jpos[e]; }
    } else if (!slam_mode) {
      // observability-constrained projection (msckf_update.cpp:393-406), g hard-coded
      const double g[3] = {0.0, 0.0, -9.81};
      double u[3], t2[2];
      xb_mv33(R, g, u);
      double uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
      for (int r = 0; r < 2; ++r) t2[r] = (Jpos[r * 3] * u[0] + Jpos[r * 3 + 1] * u[1] + Jpos[r * 3 + 2] * u[2]) * (1.0 / uu);
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) Jpos[r * 3 + c] -= t2[r] * u[c];
      double skd[9];
      xb_skew(dG, skd);
      xb_mv33(skd, g, u);
      uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
      for (int r = 0; r < 2; ++r) t2[r] = (Jatt[r * 3] * u[0] + Jatt[r * 3 + 1] * u[1] + Jatt[r * 3 + 2] * u[2]) * (1.0 / uu);
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) Jatt[r * 3 + c] -= t2[r] * u[c];
      for (int e = 0; e < 6; ++e) { Jp[e] = Jpos[e]; Ja[e] = Jatt[e]; Hf[e] = -Jpos[e]; }
    } else if (i == L - 1) {  // msckf_slam_update.cpp:133-143
      for (int e = 0; e < 6; ++e) { Jp[e] = 0.0; Ja[e] = 0.0; Jap[e] = 0.0; Jaa[e] = 0.0; Hf[e] = 0.0; }
      Hf[0] = 1.0;
      Hf[4] = 1.0;
    } else {  // msckf_slam_update.cpp:144-198
      double RtRn[9], JR[6], skab[9], m3[9], tmp[6];
      xb_mm33(Rt, Rl, RtRn);
      xb_mm23(Ji, RtRn, JR);
      const double ab1[3] = {alpha, beta, 1.0};
      xb_skew(ab1, skab);
      xb_mm23(JR, skab, tmp);
      for (int e = 0; e < 6; ++e) { Jaa[e] = -1.0 / rho * tmp[e]; Jap[e] = -Jpos[e]; Jp[e] = Jpos[e]; Ja[e] = Jatt[e]; }
      xb_mat_ivd(alpha, beta, rho, m3);
      xb_mm23(JR, m3, tmp);
      for (int e = 0; e < 6; ++e) Hf[e] = 1.0 / rho * tmp[e];
    }
  }
  nan_flag = __any_sync(0xffffffffu, nan_flag);
  bad = bad || nan_flag;
  __syncwarp();
  TPROF(3);

  // ---------------------------------------------------------------- U = orth(Hf): MGS with re-orthogonalisation
  // (msckf_update.cpp:423: Hf.householderQr().householderQ(); only range(Hf) matters)
  const int R2 = 2 * L;
  if (!bad) {
    for (int i = lane; i < R2 * 3; i += 32) ws.U[i] = ws.Hf[i];
    __syncwarp();
    for (int c = 0; c < 3; ++c) {
      for (int pass = 0; pass < 2; ++pass)
        for (int p = 0; p < c; ++p) {
          const double d = wdot(ws.U + p, 3, ws.U + c, 3, R2, lane);
          for (int i = lane; i < R2; i += 32) ws.U[i * 3 + c] -= d * ws.U[i * 3 + p];
          __syncwarp();
        }
      const double n2 = wdot(ws.U + c, 3, ws.U + c, 3, R2, lane);
      const double inv = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
      for (int i = lane; i < R2; i += 32) ws.U[i * 3 + c] *= inv;
      __syncwarp();
    }
  }

  TPROF(4);
  // ---------------------------------------------------------------- B = U^T [J | r]   (3 x (6M+1))
  for (int e = lane; e < 3 * W; e += 32) Bt[e] = 0.0;
  __syncwarp();
  double ur[3] = {0, 0, 0};
  if (!bad) {
    double anc[18];
    for (int e = 0; e < 18; ++e) anc[e] = 0.0;
    for (int i = lane; i < L; i += 32) {
      const int pos = i1 + i;
      const double* U0 = ws.U + 6 * i;  // rows 2i, 2i+1 (3 each)
      const double* Jp = ws.Jp + 6 * i;
      const double* Ja = ws.Ja + 6 * i;
      for (int u = 0; u < 3; ++u) {
        for (int c = 0; c < 3; ++c) {
          Bt[u * W + 3 * pos + c] = U0[u] * Jp[c] + U0[3 + u] * Jp[3 + c];
          Bt[u * W + 3 * M + 3 * pos + c] = U0[u] * Ja[c] + U0[3 + u] * Ja[3 + c];
        }
        ur[u] += U0[u] * ws.res[2 * i] + U0[3 + u] * ws.res[2 * i + 1];
      }
      if (slam_mode) {
        const double* Jap = ws.Jap + 6 * i;
        const double* Jaa = ws.Jaa + 6 * i;
        for (int u = 0; u < 3; ++u)
          for (int c = 0; c < 3; ++c) {
            anc[u * 6 + c] += U0[u] * Jap[c] + U0[3 + u] * Jap[3 + c];
            anc[u * 6 + 3 + c] += U0[u] * Jaa[c] + U0[3 + u] * Jaa[3 + c];
          }
      }
    }
    for (int u = 0; u < 3; ++u) ur[u] = xb_warp_sum(ur[u]);
    __syncwarp();
    if (slam_mode) {
      for (int e = 0; e < 18; ++e) anc[e] = xb_warp_sum(anc[e]);
      if (lane == 0) {
        const int pos = np - 1;  // own block of the last observation is zero, so plain stores are exact
        for (int u = 0; u < 3; ++u)
          for (int c = 0; c < 3; ++c) {
            Bt[u * W + 3 * pos + c] = anc[u * 6 + c];
            Bt[u * W + 3 * M + 3 * pos + c] = anc[u * 6 + 3 + c];
          }
      }
      // H2 = U^T Hf (3x3), msckf_slam_update.cpp:225
      double h2[9];
      for (int u = 0; u < 3; ++u)
        for (int c = 0; c < 3; ++c) h2[u * 3 + c] = wdot(ws.U + u, 3, ws.Hf + c, 3, R2, lane);
      if (lane == 0)
        for (int e = 0; e < 9; ++e) tp.H2[9 * (size_t)trk + e] = h2[e];
    }
    if (lane == 0)
      for (int u = 0; u < 3; ++u) Bt[u * W + 6 * M] = ur[u];
    if (mm_g >= 0) {  // jac_pf_ block of the own agent: A_up^T Hf (msckf_update.cpp:442-443)
      double f0[9];
      for (int u = 0; u < 3; ++u)
        for (int c = 0; c < 3; ++c) f0[u * 3 + c] = wdot(ws.U + u, 3, ws.Hf + c, 3, R2, lane);
      if (lane == 0)
        for (int e = 0; e < 9; ++e) tp.mm_F0[9 * (size_t)mm_g + e] = f0[e];
    }
  }
  __syncwarp();

  TPROF(5);
  // ---------------------------------------------------------------- gate: X = J P J^T over block pairs (k <= i)
  double gamma = NAN;
  int inl = 0;
  if (!bad) {
    const double* P = tp.P;
    const int ld = tp.ldp;
    const int npairs = L * (L + 1) / 2;
    const int nslot = slam_mode ? 2 : 1;
    int i = 0, k = 0;
    tri_advance(i, k, lane);
    for (int e = lane; e < npairs; e += 32, tri_advance(i, k, 32)) {
      double x00 = 0, x01 = 0, x10 = 0, x11 = 0;
      for (int si = 0; si < nslot; ++si) {
        const int pi_ = si ? np - 1 : i1 + i;
        const double* Ji_p = (si ? ws.Jap : ws.Jp) + 6 * i;
        const double* Ji_a = (si ? ws.Jaa : ws.Ja) + 6 * i;
        for (int sk = 0; sk < nslot; ++sk) {
          const int pk_ = sk ? np - 1 : i1 + k;
          const double* Jk_p = (sk ? ws.Jap : ws.Jp) + 6 * k;
          const double* Jk_a = (sk ? ws.Jaa : ws.Ja) + 6 * k;
          // T (2x6) = [Ji_p Ji_a] * P[pose pi_, pose pk_]: the 36 entries of the 6x6 block are loaded first (one L2
          // latency per block pair instead of one per column)
          double T[12];
          const int rp = XB_CORE + 3 * pi_, ra = XB_CORE + 3 * M + 3 * pi_;
          const int cpn = XB_CORE + 3 * pk_, can = XB_CORE + 3 * M + 3 * pk_;
          double pb[6][6];  // rows: pos a, att a of pose pi_; columns: pos c, att c of pose pk_
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              pb[a][c] = P[(size_t)(rp + a) * ld + cpn + c];
              pb[a][3 + c] = P[(size_t)(rp + a) * ld + can + c];
              pb[3 + a][c] = P[(size_t)(ra + a) * ld + cpn + c];
              pb[3 + a][3 + c] = P[(size_t)(ra + a) * ld + can + c];
            }
          // blocks among the newest clones are not symmetric (one clone in the regular case, more after updates without
          // measurement rows): use sym(P) there
          const bool symblk = pi_ >= np - tp.asym_clones && pk_ >= np - tp.asym_clones;
          if (symblk) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                pb[a][c] = 0.5 * (pb[a][c] + P[(size_t)(cpn + c) * ld + rp + a]);
                pb[a][3 + c] = 0.5 * (pb[a][3 + c] + P[(size_t)(can + c) * ld + rp + a]);
                pb[3 + a][c] = 0.5 * (pb[3 + a][c] + P[(size_t)(cpn + c) * ld + ra + a]);
                pb[3 + a][3 + c] = 0.5 * (pb[3 + a][3 + c] + P[(size_t)(can + c) * ld + ra + a]);
              }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            double tp0 = 0, tp1 = 0, ta0 = 0, ta1 = 0;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const double ppp = pb[a][c], pap = pb[3 + a][c], ppa = pb[a][3 + c], paa = pb[3 + a][3 + c];
              tp0 = fma(Ji_p[a], ppp, tp0); tp0 = fma(Ji_a[a], pap, tp0);
              tp1 = fma(Ji_p[3 + a], ppp, tp1); tp1 = fma(Ji_a[3 + a], pap, tp1);
              ta0 = fma(Ji_p[a], ppa, ta0); ta0 = fma(Ji_a[a], paa, ta0);
              ta1 = fma(Ji_p[3 + a], ppa, ta1); ta1 = fma(Ji_a[3 + a], paa, ta1);
            }
            T[c] = tp0; T[6 + c] = tp1; T[3 + c] = ta0; T[9 + c] = ta1;
          }
          for (int c = 0; c < 3; ++c) {
            x00 = fma(T[c], Jk_p[c], x00); x00 = fma(T[3 + c], Jk_a[c], x00);
            x01 = fma(T[c], Jk_p[3 + c], x01); x01 = fma(T[3 + c], Jk_a[3 + c], x01);
            x10 = fma(T[6 + c], Jk_p[c], x10); x10 = fma(T[9 + c], Jk_a[c], x10);
            x11 = fma(T[6 + c], Jk_p[3 + c], x11); x11 = fma(T[9 + c], Jk_a[3 + c], x11);
          }
        }
      }
      ws.X[tri_idx(2 * i, 2 * k)] = x00;
      ws.X[tri_idx(2 * i + 1, 2 * k)] = x10;
      ws.X[tri_idx(2 * i + 1, 2 * k + 1)] = x11;
      if (k < i) ws.X[tri_idx(2 * i, 2 * k + 1)] = x01;
    }
    __syncwarp();
    TPROF(6);
    // Y = X U  (2L x 3)
    for (int r = lane; r < R2; r += 32) {
      double y0 = 0, y1 = 0, y2 = 0;
      const double* xr = ws.X + tri_idx(r, 0);
#pragma unroll 4
      for (int c = 0; c <= r; ++c) {  // row part (stored), then the column below the diagonal (symmetric half)
        const double x = xr[c];
        y0 = fma(x, ws.U[c * 3], y0);
        y1 = fma(x, ws.U[c * 3 + 1], y1);
        y2 = fma(x, ws.U[c * 3 + 2], y2);
      }
      int idx = tri_idx(r + 1, r);
#pragma unroll 4
      for (int c = r + 1; c < R2; ++c) {
        const double x = ws.X[idx];
        idx += c + 1;
        y0 = fma(x, ws.U[c * 3], y0);
        y1 = fma(x, ws.U[c * 3 + 1], y1);
        y2 = fma(x, ws.U[c * 3 + 2], y2);
      }
      ws.Y[r * 3] = y0; ws.Y[r * 3 + 1] = y1; ws.Y[r * 3 + 2] = y2;
    }
    __syncwarp();
    double Z[9];  // U^T X U
    for (int u = 0; u < 3; ++u)
      for (int v = 0; v < 3; ++v) Z[u * 3 + v] = wdot(ws.U + u, 3, ws.Y + v, 3, R2, lane);
    // S = Pi X Pi + var I = X - U W^T - W U^T + var I  with  W = Y - U Z / 2  (Z = U^T X U symmetric)
    for (int r = lane; r < R2; r += 32)
      for (int v = 0; v < 3; ++v)
        ws.V[r * 3 + v] = ws.Y[r * 3 + v] - 0.5 * (ws.U[r * 3] * Z[v] + ws.U[r * 3 + 1] * Z[3 + v] + ws.U[r * 3 + 2] * Z[6 + v]);
    __syncwarp();
    // (in place, packed), augmented row R2 = (Pi r)^T.  Row-wise: a lane takes the rows r and R2-1-r (R2 + 1 elements
    // together, so the lanes are balanced), keeps U_r, V_r in registers and walks the columns; all lanes read U_c, V_c of
    // the same column at the same time (broadcasts).
    for (int h = lane; h < (R2 + 1) / 2; h += 32) {
#pragma unroll 1
      for (int side = 0; side < 2; ++side) {
        const int r = side ? R2 - 1 - h : h;
        if (side && r == h) break;
        const double u0 = ws.U[r * 3], u1 = ws.U[r * 3 + 1], u2 = ws.U[r * 3 + 2];
        const double v0 = ws.V[r * 3], v1 = ws.V[r * 3 + 1], v2 = ws.V[r * 3 + 2];
        double* xr = ws.X + tri_idx(r, 0);
#pragma unroll 4
        for (int c = 0; c <= r; ++c) {
          double sv = xr[c];
          sv -= u0 * ws.V[c * 3] + v0 * ws.U[c * 3];
          sv -= u1 * ws.V[c * 3 + 1] + v1 * ws.U[c * 3 + 1];
          sv -= u2 * ws.V[c * 3 + 2] + v2 * ws.U[c * 3 + 2];
          if (c == r) sv += tp.var_img;
          xr[c] = sv;
        }
      }
    }
    double* aug = ws.X + tri_idx(R2, 0);
    for (int r = lane; r < R2; r += 32)
      aug[r] = ws.res[r] - (ws.U[r * 3] * ur[0] + ws.U[r * 3 + 1] * ur[1] + ws.U[r * 3 + 2] * ur[2]);
    // rows R2+1+k: V^T, V = Pi J_c with J_c = J[:, 6 columns of the newest clone]  (antisymmetric part of P lives there)
    const int ccol[6] = {3 * (np - 1), 3 * (np - 1) + 1, 3 * (np - 1) + 2, 3 * M + 3 * (np - 1), 3 * M + 3 * (np - 1) + 1,
                         3 * M + 3 * (np - 1) + 2};
    for (int k = 0; k < 6; ++k) {
      double* vr = ws.X + tri_idx(R2 + 1 + k, 0);
      const double b0 = Bt[ccol[k]], b1 = Bt[W + ccol[k]], b2 = Bt[2 * W + ccol[k]];
      for (int r = lane; r < R2; r += 32) {
        const int i = r >> 1, h = r & 1;
        double jc = 0.0;
        if (i1 + i == np - 1) jc += (k < 3 ? ws.Jp : ws.Ja)[6 * i + 3 * h + (k % 3)];
        if (slam_mode) jc += (k < 3 ? ws.Jap : ws.Jaa)[6 * i + 3 * h + (k % 3)];
        vr[r] = jc - (ws.U[r * 3] * b0 + ws.U[r * 3 + 1] * b1 + ws.U[r * 3 + 2] * b2);
      }
    }
    __syncwarp();
    // Cholesky of the augmented lower triangle (rows R2..R2+6 are right-hand sides): the last rows become
    // y = L^-1 (Pi r) and Vt = L^-1 V.  Lane-owned rows, 4-wide batches so that loads overlap the FMAs.
    TPROF(7);
    const bool spd = gate_chol_blocked<(64 * OPL + 7 + 31) / 32>(ws.X, R2, ws.scr, lane);
    __syncwarp();
    TPROF(8);
    if (spd) {
      // gamma = r^T S^-1 r with S = S_s + V E V^T (E = antisymmetric part of the clone block of P):
      //   y^T y - (Vt^T y)^T E (I + G E)^-1 (Vt^T y),  Vt = L^-1 V,  G = Vt^T Vt      (Woodbury)
      gamma = wdot(aug, 1, aug, 1, R2, lane);
      double Em[36], emax = 0.0;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) {
          Em[a * 6 + b] = 0.5 * (P[(size_t)(XB_CORE + ccol[a]) * ld + XB_CORE + ccol[b]] -
                                 P[(size_t)(XB_CORE + ccol[b]) * ld + XB_CORE + ccol[a]]);
          emax = fmax(emax, fabs(Em[a * 6 + b]));
        }
      if (emax > 0.0) {
        // small dense algebra cooperatively in shared memory: Am = [I + G E | I] (6 x 12), lane b owns column b
        double* Gs = ws.scr;            // 36
        double* gvs = ws.scr + 36;      // 6
        double* Am = ws.scr + 42;       // 6 x 13
        // the 21 + 6 dot products G = Vt^T Vt, gv = Vt^T y: one per lane (rows 0..6 of the right-hand-side block: y, Vt_0..5)
        if (lane < 27) {
          int a = 0, b = lane;                     // lane < 6: (y, Vt_lane); else the pair (a, b <= a) of G
          if (lane >= 6) { int e = lane - 6; a = 0; while (e > a) { e -= a + 1; ++a; } b = e; }
          const double* ra = lane < 6 ? aug : ws.X + tri_idx(R2 + 1 + a, 0);
          const double* rb = ws.X + tri_idx(R2 + 1 + b, 0);
          double d0 = 0.0, d1 = 0.0;
          int i = 0;
          for (; i + 1 < R2; i += 2) { d0 = fma(ra[i], rb[i], d0); d1 = fma(ra[i + 1], rb[i + 1], d1); }
          if (i < R2) d0 = fma(ra[i], rb[i], d0);
          d0 += d1;
          if (lane < 6) gvs[lane] = d0;
          else { Gs[a * 6 + b] = d0; Gs[b * 6 + a] = d0; }
        }
        __syncwarp();
        if (lane < 12) {
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            double v;
            if (lane < 6) {
              v = (a == lane) ? 1.0 : 0.0;
#pragma unroll
              for (int x = 0; x < 6; ++x) v = fma(Gs[a * 6 + x], Em[x * 6 + lane], v);
            } else {
              v = (lane - 6 == a) ? 1.0 : 0.0;
            }
            Am[a * 13 + lane] = v;
          }
        }
        __syncwarp();
#pragma unroll 1
        for (int c = 0; c < 6; ++c) {  // Gauss-Jordan with partial pivoting
          int best = c;
          double bv = fabs(Am[c * 13 + c]);
          for (int r = c + 1; r < 6; ++r) { const double x = fabs(Am[r * 13 + c]); if (x > bv) { bv = x; best = r; } }
          __syncwarp();
          if (lane < 12 && best != c) { const double tmp = Am[c * 13 + lane]; Am[c * 13 + lane] = Am[best * 13 + lane]; Am[best * 13 + lane] = tmp; }
          __syncwarp();
          double f6[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) f6[r] = Am[r * 13 + c];
          __syncwarp();
          if (lane < 12) {
            const double pr = Am[c * 13 + lane] / f6[c];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
              if (r == c) Am[r * 13 + lane] = pr;
              else Am[r * 13 + lane] = fma(-f6[r], pr, Am[r * 13 + lane]);
            }
          }
          __syncwarp();
        }
        double t6[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          t6[a] = 0.0;
#pragma unroll
          for (int b = 0; b < 6; ++b) t6[a] = fma(Am[a * 13 + 6 + b], gvs[b], t6[a]);  // (I+GE)^-1 g
        }
        double corr = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = 0; b < 6; ++b) corr = fma(gvs[a] * Em[a * 6 + b], t6[b], corr);
        gamma -= corr;
      }
      const double chi = tp.chi2_95[2 * L - 3];
      inl = gamma < chi;
    }
  }
  TPROF(9);
  if (lane == 0) {
    tp.gamma[trk] = gamma;
    tp.inlier[trk] = inl;
  }
  // emit the sparse J blocks + residuals of the track (used by the Gram stage)
  for (int i = lane; i < L; i += 32) {
    double* o = tp.Jout + 14 * (size_t)(o0 + i);
    for (int e = 0; e < 6; ++e) { o[e] = bad ? 0.0 : ws.Jp[6 * i + e]; o[6 + e] = bad ? 0.0 : ws.Ja[6 * i + e]; }
    o[12] = bad ? 0.0 : ws.res[2 * i];
    o[13] = bad ? 0.0 : ws.res[2 * i + 1];
  }
  if (slam_mode && (bad || !inl)) {
    double* D = tp.D + (size_t)2 * o0 * W;
    for (int e = lane; e < 2 * L * W; e += 32) D[e] = 0.0;
  } else if (slam_mode) {
    // dense rows D = Pi J (2L x W) for the Gram stage: D = J - U B   (mode-1 tracks are few)
    double* D = tp.D + (size_t)2 * o0 * W;
    __syncwarp();
    for (int e = lane; e < R2 * W; e += 32) {
      const int r = e / W, c = e % W;
      double v = -(ws.U[r * 3] * Bt[c] + ws.U[r * 3 + 1] * Bt[W + c] + ws.U[r * 3 + 2] * Bt[2 * W + c]);
      const int i = r >> 1, h = r & 1;
      const int pos = i1 + i;
      if (c == 6 * M) v += ws.res[r];
      else {
        const bool att = c >= 3 * M;
        const int cc = att ? c - 3 * M : c;
        const int cpz = cc / 3, a = cc % 3;
        if (cpz == pos) v += (att ? ws.Ja : ws.Jp)[6 * i + 3 * h + a];
        if (cpz == np - 1) v += (att ? ws.Jaa : ws.Jap)[6 * i + 3 * h + a];
      }
      D[e] = v;
    }
  }
  // outliers contribute nothing downstream: B is kept only for inliers in MSCKF mode.  In MSCKF-SLAM
  // mode B (= H1) and U^T r (= r1) also feed the feature initialisation of *every* new track
  // (vio_updater.cpp:430-437), so they are copied out before B is masked.
  if (slam_mode) {
    double* h1 = tp.H1 + (size_t)trk * 3 * W;
    for (int e = lane; e < 3 * W; e += 32) h1[e] = Bt[e];
  }
  __syncwarp();
  if (!inl)
    for (int e = lane; e < 3 * W; e += 32) Bt[e] = 0.0;
  TPROF(10);
#undef TPROF
}

size_t tracks_smem_bytes(int M, int Lmax, int warps, int mode) {
  return sizeof(double) * ((size_t)12 * M + (size_t)warps * WarpSmem::doubles(Lmax, mode));
}

int launch_tracks(cudaStream_t s, const TrackParams& tp) {
  if (tp.n_tracks <= 0) return 0;
  // two warps (tracks) per CTA: the CTA's shared memory (<= 55 KB) leaves room on every SM for the dataflow Cholesky CTAs
  // that xb_api.cu runs concurrently on the side stream
  static int warps_cfg = 0;
  if (!warps_cfg) { const char* e = getenv("XB_TRACK_WARPS"); warps_cfg = e ? std::max(1, std::min(4, atoi(e))) : 2; }
  int warps = warps_cfg;
  const size_t cap = 110 * 1024;  // two CTAs per SM at four warps
  while (warps > 1 && tracks_smem_bytes(tp.M, tp.Lmax, warps, tp.mode) > cap) --warps;
  const size_t bytes = tracks_smem_bytes(tp.M, tp.Lmax, warps, tp.mode);
  if (bytes > cap) return -1;
  const int grid = (tp.n_tracks + warps - 1) / warps;
  if (tp.Lmax <= 32) {
    cudaFuncSetAttribute(k_tracks<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    XB_LAUNCH((k_tracks<1>), grid, warps * 32, bytes, s, tp);
  } else if (tp.Lmax <= 64) {
    cudaFuncSetAttribute(k_tracks<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    XB_LAUNCH((k_tracks<2>), grid, warps * 32, bytes, s, tp);
  } else {
    return -1;
  }
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// SLAM rows (slam_update.cpp:49-214): one warp per SLAM feature; emits 2 sparse rows (<=15 columns).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_slam_rows(SlamParams sp) {
  XB_PDL_SHORT();
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= sp.n_tracks) return;
  const int M = sp.M, np = sp.n_poses, N = sp.N;
  const double* parr = sp.xv + XV_ARR;
  const double* qarr = sp.xv + XV_ARR + 3 * M;
  const double* farr = sp.xv + XV_ARR + 7 * M;
  int* cols = sp.cols + 15 * j;
  double* vals = sp.vals + 30 * j;
  if (lane < 15) cols[lane] = 0;
  if (lane < 30) vals[lane] = 0.0;
  if (lane == 0) {
    sp.res[2 * j] = 0.0;
    sp.res[2 * j + 1] = 0.0;
    sp.inlier[j] = 0;
    sp.gamma[j] = NAN;
  }
  __syncwarp();
  const int L = sp.off[j + 1] - sp.off[j];
  if (L < 1) return;
  const double a = farr[3 * j], b = farr[3 * j + 1], r = farr[3 * j + 2];
  const int anchor = sp.anchor[j];
  if (anchor < 0 || anchor >= np) return;
  double Ra[9], Rn[9];
  xb_rot(qarr + 4 * anchor, Ra);
  xb_rot(qarr + 4 * (np - 1), Rn);
  const double ab1[3] = {a, b, 1.0};
  double t3[3], Gf[3], dG[3], cp[3];
  xb_mv33(Ra, ab1, t3);
  for (int e = 0; e < 3; ++e) { Gf[e] = 1.0 / r * t3[e] + parr[3 * anchor + e]; dG[e] = Gf[e] - parr[3 * (np - 1) + e]; }
  xb_mtv33(Rn, dG, cp);
  const double* z = sp.obs + 2 * (size_t)(sp.off[j + 1] - 1);
  const double r0 = z[0] - cp[0] / cp[2], r1 = z[1] - cp[1] / cp[2];
  const int pos = np - 1;
  const int fcol = XB_CORE + (2 * M + j) * 3;
  int nc;
  double h[30];
  int lc[15];
  for (int e = 0; e < 30; ++e) h[e] = 0.0;
  for (int e = 0; e < 15; ++e) lc[e] = 0;
  if (anchor == pos) {  // slam_update.cpp:120-130
    nc = 3;
    for (int c = 0; c < 3; ++c) lc[c] = fcol + c;
    h[0] = 1.0;
    h[15 + 1] = 1.0;
  } else {
    nc = 15;
    const double Ji[6] = {1.0 / cp[2], 0.0, -cp[0] / (cp[2] * cp[2]), 0.0, 1.0 / cp[2], -cp[1] / (cp[2] * cp[2])};
    double Rt[9], sk[9], Jpos[6], Jatt[6], RtRa[9], JR[6], skab[9], Jaa[6], m3[9], Hf[6];
    for (int rr = 0; rr < 3; ++rr)
      for (int c = 0; c < 3; ++c) Rt[rr * 3 + c] = Rn[c * 3 + rr];
    xb_skew(cp, sk);
    xb_mm23(Ji, sk, Jatt);
    xb_mm23(Ji, Rt, Jpos);
    for (int e = 0; e < 6; ++e) Jpos[e] = -Jpos[e];
    xb_mm33(Rt, Ra, RtRa);
    xb_mm23(Ji, RtRa, JR);
    xb_skew(ab1, skab);
    xb_mm23(JR, skab, Jaa);
    xb_mat_ivd(a, b, r, m3);
    xb_mm23(JR, m3, Hf);
    for (int c = 0; c < 3; ++c) {
      lc[c] = XB_CORE + 3 * pos + c;
      lc[3 + c] = XB_CORE + 3 * M + 3 * pos + c;
      lc[6 + c] = XB_CORE + 3 * anchor + c;
      lc[9 + c] = XB_CORE + 3 * M + 3 * anchor + c;
      lc[12 + c] = fcol + c;
      for (int rr = 0; rr < 2; ++rr) {
        h[rr * 15 + c] = Jpos[rr * 3 + c];
        h[rr * 15 + 3 + c] = Jatt[rr * 3 + c];
        h[rr * 15 + 6 + c] = -Jpos[rr * 3 + c];
        h[rr * 15 + 9 + c] = -1.0 / r * Jaa[rr * 3 + c];
        h[rr * 15 + 12 + c] = 1.0 / r * Hf[rr * 3 + c];
      }
    }
  }
  // gate: S = h P h^T + var I (2x2), chi2(0.9, 2*track_size)  (slam_update.cpp:192-199)
  // S = h P h^T + var I is evaluated with the (possibly non-symmetric) P exactly as the reference does
  double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
  for (int e = lane; e < nc * nc; e += 32) {  // the (d, c) pairs of the 15x15 gather are spread over the warp
    const int d = e / nc, c = e % nc;
    double hd0 = 0, hd1 = 0, hc0 = 0, hc1 = 0;
    int cd = 0, cc = 0;
#pragma unroll
    for (int q = 0; q < 15; ++q) {
      if (q == d) { hd0 = h[q]; hd1 = h[15 + q]; cd = lc[q]; }
      if (q == c) { hc0 = h[q]; hc1 = h[15 + q]; cc = lc[q]; }
    }
    const double p = sp.P[(size_t)cd * N + cc];
    s00 = fma(hd0 * p, hc0, s00);
    s01 = fma(hd0 * p, hc1, s01);
    s10 = fma(hd1 * p, hc0, s10);
    s11 = fma(hd1 * p, hc1, s11);
  }
  s00 = xb_warp_sum(s00); s01 = xb_warp_sum(s01); s10 = xb_warp_sum(s10); s11 = xb_warp_sum(s11);
  s00 += sp.var_img;
  s11 += sp.var_img;
  const double det = s00 * s11 - s01 * s10;
  const double gamma = (r0 * (s11 * r0 - s01 * r1) + r1 * (s00 * r1 - s10 * r0)) / det;
  const int inl = gamma < sp.chi2[j];
  if (lane == 0) {
    sp.gamma[j] = gamma;
    sp.inlier[j] = inl;
    for (int e = 0; e < 15; ++e) cols[e] = lc[e];
    if (inl) {
      for (int e = 0; e < 30; ++e) vals[e] = h[e];
      sp.res[2 * j] = r0;
      sp.res[2 * j + 1] = r1;
    }
  }
}

void launch_slam_rows(cudaStream_t s, const SlamParams& sp) {
  if (sp.n_tracks <= 0) return;
  XB_LAUNCH(k_slam_rows, (sp.n_tracks + 3) / 4, 128, 0, s, sp);
  count_launch();
}

// ------------------------------------------------------------------------------------------------
// Range and sun-sensor rows (range_update.cpp:61-265, solar_update.cpp:39-94): one warp.  Lane 0 evaluates the (tiny)
// closed forms, the warp evaluates the range gate h P h^T over the 33 x 33 gathered covariance entries.
// Row w of the output: XB_WNZ (column, value) entries (duplicates add: the anchor blocks of the range row accumulate
// exactly like range_update.cpp:217-242 when anchors coincide), unused entries are (0, 0.0).
// ------------------------------------------------------------------------------------------------
__device__ inline void xb_cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
// row vector (1x3) times 3x3 matrix
__device__ inline void xb_vtm33(const double* v, const double* A, double* y) {
  for (int c = 0; c < 3; ++c) y[c] = v[0] * A[c] + v[1] * A[3 + c] + v[2] * A[6 + c];
}
__global__ void __launch_bounds__(32) k_sensor_rows(SensorParams sp) {
  XB_PDL_LONG();
  __shared__ int lc[XB_WMAX * XB_WNZ];
  __shared__ double h[XB_WMAX * XB_WNZ];
  __shared__ double rs[XB_WMAX];
  __shared__ int ok;
  const int lane = threadIdx.x;
  const int M = sp.M, np = sp.n_poses, N = sp.N;
  const double* parr = sp.xv + XV_ARR;
  const double* qarr = sp.xv + XV_ARR + 3 * M;
  const double* farr = sp.xv + XV_ARR + 7 * M;
  for (int e = lane; e < XB_WMAX * XB_WNZ; e += 32) { lc[e] = 0; h[e] = 0.0; }
  if (lane < XB_WMAX) rs[lane] = 0.0;
  if (lane == 0) ok = 0;
  __syncwarp();
  int row = 0;
  if (sp.range_on) {
    if (lane == 0) {
      double Gf[3][3], Ra[3][9], al[3], be[3], rho[3];
      int an[3];
      bool valid = np >= 1;
      for (int j = 0; j < 3 && valid; ++j) {  // range_update.cpp:76-97
        const int id = sp.tri[j];
        al[j] = farr[3 * id]; be[j] = farr[3 * id + 1]; rho[j] = farr[3 * id + 2];
        an[j] = sp.anchor[id];
        if (an[j] < 0 || an[j] >= np) { valid = false; break; }
        xb_rot(qarr + 4 * an[j], Ra[j]);
        const double ab1[3] = {al[j], be[j], 1.0};
        double t3[3];
        xb_mv33(Ra[j], ab1, t3);
        for (int e = 0; e < 3; ++e) Gf[j][e] = 1.0 / rho[j] * t3[e] + parr[3 * an[j] + e];
      }
      if (valid) {
        double Ri[9], d01[3], d21[3], Gn[3];
        xb_rot(qarr + 4 * (np - 1), Ri);
        const double* pci = parr + 3 * (np - 1);
        for (int e = 0; e < 3; ++e) { d01[e] = Gf[0][e] - Gf[1][e]; d21[e] = Gf[2][e] - Gf[1][e]; }
        xb_cross3(d01, d21, Gn);  // :123
        const double pt[3] = {sp.pt_x, sp.pt_y, 1.0};
        double RtGn[3];
        xb_mtv33(Ri, Gn, RtGn);
        double a = 0.0, b = 0.0;
        for (int e = 0; e < 3; ++e) { a += (Gf[1][e] - pci[e]) * Gn[e]; b += pt[e] * RtGn[e]; }
        const double res = sp.range - a / b;  // :129-139
        double GnR[3], skpt[9], Jqc[3], Rpt[3], Gpr[3], bary[3];
        xb_vtm33(Gn, Ri, GnR);
        xb_skew(pt, skpt);
        xb_vtm33(GnR, skpt, Jqc);
        xb_mv33(Ri, pt, Rpt);
        for (int e = 0; e < 3; ++e) {
          Gpr[e] = a / b * Rpt[e] + pci[e];
          bary[e] = 1.0 / 3.0 * (Gf[0][e] + Gf[1][e] + Gf[2][e]);
        }
        const int pos = np - 1;
        for (int c = 0; c < 3; ++c) {  // :148-153, :209-215
          lc[c] = XB_CORE + 3 * pos + c;
          h[c] = -1.0 / b * Gn[c];
          lc[3 + c] = XB_CORE + 3 * M + 3 * pos + c;
          h[3 + c] = a / (b * b) * Jqc[c];
        }
        const int ep[3] = {2, 0, 1}, eq[3] = {1, 2, 0};  // :162, :178, :194
        for (int j = 0; j < 3; ++j) {
          double ed[3], bd[3], cr[3], Jf[3], JfR[3], skab[9], Jqa[3], m3[9], Jfi[3];
          for (int e = 0; e < 3; ++e) { ed[e] = Gf[ep[j]][e] - Gf[eq[j]][e]; bd[e] = bary[e] - Gpr[e]; }
          xb_cross3(ed, bd, cr);
          for (int e = 0; e < 3; ++e) Jf[e] = 1.0 / b * (1.0 / 3.0 * Gn[e] + cr[e]);
          xb_vtm33(Jf, Ra[j], JfR);
          const double ab1[3] = {al[j], be[j], 1.0};
          xb_skew(ab1, skab);
          xb_vtm33(JfR, skab, Jqa);
          xb_mat_ivd(al[j], be[j], rho[j], m3);
          xb_vtm33(JfR, m3, Jfi);
          const int o = 6 + 9 * j;
          for (int c = 0; c < 3; ++c) {
            lc[o + c] = XB_CORE + 3 * an[j] + c;
            h[o + c] = Jf[c];
            lc[o + 3 + c] = XB_CORE + 3 * M + 3 * an[j] + c;
            h[o + 3 + c] = -1.0 / rho[j] * Jqa[c];
            lc[o + 6 + c] = XB_CORE + (2 * M + sp.tri[j]) * 3 + c;
            h[o + 6 + c] = 1.0 / rho[j] * Jfi[c];
          }
        }
        rs[0] = res;
        ok = 1;
      }
    }
    __syncwarp();
    // gate: gamma = res^2 / (h P h^T + sigma_range^2) < chi2(0.9, 1)   (range_update.cpp:246-252)
    double s = 0.0;
    if (ok) {
      for (int e = lane; e < 33 * 33; e += 32) {
        const int d = e / 33, c = e % 33;
        s = fma(h[d] * sp.P[(size_t)lc[d] * N + lc[c]], h[c], s);
      }
    }
    s = xb_warp_sum(s) + sp.var_range;
    const double gamma = ok ? rs[0] * rs[0] / s : NAN;
    const int inl = ok && gamma < sp.chi2_1;
    if (lane == 0) { sp.gamma[0] = gamma; sp.inlier[0] = inl; }
    __syncwarp();
    for (int e = lane; e < XB_WNZ; e += 32) h[e] = inl ? h[e] * sp.w_range : 0.0;  // an outlier leaves a zero row (:254)
    if (lane == 0) rs[0] = inl ? rs[0] * sp.w_range : 0.0;
    row = 1;
  }
  __syncwarp();
  if (sp.sun_on && lane == 0) {  // solar_update.cpp:39-94 (calibration constants as hard-coded there)
    const double SqI[4] = {-0.063338979194957, 0.007502445522018, 0.930635612981541, 0.360346005598587};  // (x,y,z,w)
    double gs[3] = {-0.29385515271891938, -0.55080445540063927, 0.78119370269565391};
    const double gn = sqrt(gs[0] * gs[0] + gs[1] * gs[1] + gs[2] * gs[2]);
    for (int e = 0; e < 3; ++e) gs[e] /= gn;
    double Rs[9], Rq[9], sv[3], sh[3];
    xb_rot(SqI, Rs);
    xb_rot(sp.xv + 6, Rq);
    xb_mtv33(Rq, gs, sv);
    xb_mtv33(Rs, sv, sh);
    const double sn = sqrt(sh[0] * sh[0] + sh[1] * sh[1] + sh[2] * sh[2]);
    for (int e = 0; e < 3; ++e) sh[e] /= sn;
    const double RAD2DEG = 57.2957795130;
    const double r0 = sp.sun_x - RAD2DEG * atan2(sh[0], sh[2]), r1 = sp.sun_y - RAD2DEG * atan2(sh[1], sh[2]);
    const double d0 = sh[0] * sh[0] + sh[2] * sh[2], d1 = sh[1] * sh[1] + sh[2] * sh[2];
    const double mat[6] = {sh[2] / d0, 0.0, -sh[0] / d0, 0.0, sh[2] / d1, -sh[1] / d1};
    double Rst[9], sk[9], m1[6], J[6];
    for (int rr = 0; rr < 3; ++rr)
      for (int c = 0; c < 3; ++c) Rst[rr * 3 + c] = Rs[c * 3 + rr];
    xb_skew(sv, sk);
    xb_mm23(mat, Rst, m1);
    xb_mm23(m1, sk, J);
    for (int rr = 0; rr < 2; ++rr) {
      for (int c = 0; c < 3; ++c) {
        lc[(row + rr) * XB_WNZ + c] = 6 + c;  // kIdxQ
        h[(row + rr) * XB_WNZ + c] = RAD2DEG * J[rr * 3 + c] * sp.w_sun;
      }
    }
    rs[row] = r0 * sp.w_sun;
    rs[row + 1] = r1 * sp.w_sun;
  }
  __syncwarp();
  for (int e = lane; e < XB_WMAX * XB_WNZ; e += 32) { sp.cols[e] = lc[e]; sp.vals[e] = h[e]; }
  if (lane < XB_WMAX) sp.res[lane] = rs[lane];
}
void launch_sensor_rows(cudaStream_t s, const SensorParams& sp) {
  if (!sp.range_on && !sp.sun_on) return;
  XB_LAUNCH(k_sensor_rows, 1, 32, 0, s, sp);
  count_launch();
}

// ------------------------------------------------------------------------------------------------
// Gram stage:  G = sum_inliers [J|r]^T [J|r] - B_all^T B_all + D^T D   ((6M+1) x (6M+1))
//   k_gram_partial: split-K  A^T A  of a tall row-major matrix A [rows x W] -> partial[z][W x W]
//   k_gram_reduce : fixed-order sum of the partials (deterministic), block-diagonal J^T J terms and
//                   scatter into the tall Cholesky buffer [G ; g^T].
// ------------------------------------------------------------------------------------------------
static const bool g_gram_mma = [] { const char* e = getenv("XB_GEMM"); return !(e && e[0] == 'd'); }();  // see k_linalg.cu
__global__ void __launch_bounds__(256) k_gram_partial(const double* __restrict__ A, int rows, int W, int chunk,
                                                      double* __restrict__ part) {
  XB_PDL_SHORT();
  __shared__ double As[16][64 + 4];
  __shared__ double Bs[16][64 + 4];
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64, z = blockIdx.z;
  const int kb = z * chunk, ke = min(rows, kb + chunk);
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = kb; k0 < ke; k0 += 16) {
    const int kk = t >> 4, c = (t & 15) * 4;
    const int gk = k0 + kk;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int ga = m0 + c + u, gb = n0 + c + u;
      As[kk][c + u] = (gk < ke && ga < W) ? A[(size_t)gk * W + ga] : 0.0;
      Bs[kk][c + u] = (gk < ke && gb < W) ? A[(size_t)gk * W + gb] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k2][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k2][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  double* out = part + (size_t)z * W * W;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gr = m0 + ty * 4 + i;
    if (gr >= W) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = n0 + tx * 4 + j;
      if (gc < W) out[(size_t)gr * W + gc] = acc[i][j];
    }
  }
}

// One CTA per window pose: sum over the tracks' observations at that pose of [Jp Ja r]^T [Jp Ja r] (7x7).
// pose of observation i of track t = n_poses - L_t + i.
// 512 threads, two tracks per thread and trip with all index loads, then all Jacobian loads, in flight together (the kernel
// is a chain of three dependent L2 round trips per track: inlier/offsets -> Jacobian block -> products); the 28 sums are
// reduced by warp shuffles and one fixed-order pass over the 16 warp partials (deterministic).
#define GJ_THREADS 512
__global__ void __launch_bounds__(GJ_THREADS) k_gram_jtj(const int* __restrict__ off, const int* __restrict__ inlier, int n_tracks,
                                                         const double* __restrict__ Jout, int n_poses, double* __restrict__ blocks) {
  XB_PDL_SHORT();
  const int pose = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ double red[GJ_THREADS / 32][28];
  double acc[28];
#pragma unroll
  for (int e = 0; e < 28; ++e) acc[e] = 0.0;
  for (int t0 = 0; t0 < n_tracks; t0 += 2 * GJ_THREADS) {
    const double* src[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int t = t0 + u * GJ_THREADS + threadIdx.x;
      src[u] = nullptr;
      if (t < n_tracks) {
        const int o0 = off[t], o1 = off[t + 1], inl = inlier[t];
        const int i = pose - (n_poses - (o1 - o0));
        if (inl && i >= 0 && i < o1 - o0) src[u] = Jout + 14 * (size_t)(o0 + i);
      }
    }
    double o[2][14];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int e = 0; e < 14; ++e) o[u][e] = src[u] ? src[u][e] : 0.0;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      // 7 "columns": Jp(:,0..2), Ja(:,0..2), r ; rows 0/1
      double c0[7], c1[7];
#pragma unroll
      for (int e = 0; e < 3; ++e) { c0[e] = o[u][e]; c1[e] = o[u][3 + e]; c0[3 + e] = o[u][6 + e]; c1[3 + e] = o[u][9 + e]; }
      c0[6] = o[u][12];
      c1[6] = o[u][13];
      int q = 0;
#pragma unroll
      for (int a = 0; a < 7; ++a)
#pragma unroll
        for (int b = a; b < 7; ++b) { acc[q] += c0[a] * c0[b] + c1[a] * c1[b]; ++q; }
    }
  }
#pragma unroll
  for (int e = 0; e < 28; ++e) acc[e] = xb_warp_sum(acc[e]);
  if (lane == 0)
    for (int e = 0; e < 28; ++e) red[warp][e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 28) {
    double v = 0.0;
    for (int w = 0; w < GJ_THREADS / 32; ++w) v += red[w][threadIdx.x];
    blocks[28 * pose + threadIdx.x] = v;
  }
}

__global__ void k_gram_reduce(const double* __restrict__ partB, int nzB, const double* __restrict__ partD, int nzD,
                              const double* __restrict__ blocks, int M, int n_poses, double* __restrict__ T, int ld,
                              int rows_pad, int cols_pad, double* __restrict__ diag0) {
  XB_PDL_SHORT();
  // T is the tall buffer [cols_pad (G) + 32 (row 0 = g^T)] x ld ; everything outside G/g is identity/zero padding.
  const int W = 6 * M + 1;
  const int r = blockIdx.y * 16 + threadIdx.y, c = blockIdx.x * 16 + threadIdx.x;
  if (r >= rows_pad || c >= cols_pad) return;
  const int n = 6 * M;
  double v = 0.0;
  int gr = -1;
  if (r < n) gr = r;
  else if (r == cols_pad) gr = n;  // augmented row: g^T
  if (gr >= 0 && c < n) {
    // fixed-order sums of the split-K partials, eight loads in flight at a time (one L2 round trip per eight partials
    // instead of one per partial)
    const size_t e0 = (size_t)gr * W + c, zs = (size_t)W * W;
    for (int z0 = 0; z0 < nzB; z0 += 8) {
      double p[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) p[u] = z0 + u < nzB ? partB[(size_t)(z0 + u) * zs + e0] : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) v -= p[u];
    }
    for (int z0 = 0; z0 < nzD; z0 += 8) {
      double p[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) p[u] = z0 + u < nzD ? partD[(size_t)(z0 + u) * zs + e0] : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) v += p[u];
    }
    // block-diagonal J^T J: element (gr, c) is non-zero when both belong to the same pose (or gr is the r column)
    auto pose_of = [&](int x, int& k) { const bool att = x >= 3 * M; const int xx = att ? x - 3 * M : x; k = (att ? 3 : 0) + xx % 3; return xx / 3; };
    int kc, pc = pose_of(c, kc);
    int kr = 6, pr = pc;
    if (gr < n) pr = pose_of(gr, kr);
    if (pr == pc && pc < n_poses) {
      const int a = min(kr, kc), b = max(kr, kc);
      const int q = a * 7 - a * (a - 1) / 2 + (b - a);
      v += blocks[28 * pc + q];
    }
  } else if (r == c && r >= n && r < cols_pad) {
    v = 1.0;  // identity padding keeps the factor well defined
  }
  if (r == c && r < cols_pad) diag0[r] = v;
  if (r < cols_pad && (c >> 5) > (r >> 5)) v = 0.0;  // tiles above the diagonal: the factor is lower triangular
  T[(size_t)r * ld + c] = v;
}

void launch_gram(cudaStream_t s, const GramParams& gp, cudaStream_t s_jtj, cudaEvent_t ev_fork, cudaEvent_t ev_join) {
  const int W = 6 * gp.M + 1;
  // the block-diagonal J^T J terms (one CTA per pose) are independent of the B^T B partials: optional second stream
  const bool fork = s_jtj != nullptr && s_jtj != s;
  if (fork) {
    cudaEventRecord(ev_fork, s);
    cudaStreamWaitEvent(s_jtj, ev_fork, 0);
    XB_LAUNCH(k_gram_jtj, gp.M, GJ_THREADS, 0, s_jtj, gp.off, gp.inlier, gp.n_tracks_msckf, gp.Jout, gp.n_poses, gp.blocks);
    count_launch();
    cudaEventRecord(ev_join, s_jtj);
  }
  const int tiles = (W + 63) / 64;
  int nzB = 0, nzD = 0;
  if (gp.rowsB > 0) {
    nzB = gp.nzB;
    const int chunk = ((gp.rowsB + nzB - 1) / nzB + 15) / 16 * 16;
    if (g_gram_mma) {
      gemm_tn_splitk(s, W, W, gp.rowsB, gp.B, W, gp.B, W, gp.partB, W, (size_t)W * W, nzB);
    } else {
      dim3 g(tiles, tiles, nzB);
      XB_LAUNCH(k_gram_partial, g, 256, 0, s, gp.B, gp.rowsB, W, chunk, gp.partB);
      count_launch();
    }
  }
  if (gp.rowsD > 0) {
    nzD = gp.nzD;
    const int chunk = ((gp.rowsD + nzD - 1) / nzD + 15) / 16 * 16;
    if (g_gram_mma) {
      gemm_tn_splitk(s, W, W, gp.rowsD, gp.D, W, gp.D, W, gp.partD, W, (size_t)W * W, nzD);
    } else {
      dim3 g(tiles, tiles, nzD);
      XB_LAUNCH(k_gram_partial, g, 256, 0, s, gp.D, gp.rowsD, W, chunk, gp.partD);
      count_launch();
    }
  }
  if (fork) {
    cudaStreamWaitEvent(s, ev_join, 0);
  } else {
    XB_LAUNCH(k_gram_jtj, gp.M, GJ_THREADS, 0, s, gp.off, gp.inlier, gp.n_tracks_msckf, gp.Jout, gp.n_poses, gp.blocks);
    count_launch();
  }
  dim3 b(16, 16), g((gp.cols_pad + 15) / 16, (gp.rows_pad + 15) / 16);
  XB_LAUNCH(k_gram_reduce, g, b, 0, s, gp.partB, nzB, gp.partD, nzD, gp.blocks, gp.M, gp.n_poses, gp.T, gp.ld, gp.rows_pad,
                                gp.cols_pad, gp.diag0);
  count_launch();
}

}  // namespace xb
