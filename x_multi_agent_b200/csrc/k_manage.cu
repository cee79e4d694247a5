// Sliding-window state management on the device-resident state and covariance.
// reference: src/x/vio/state_manager.cpp:31-149 (manage), :273-349 (augmentCovariance),
//            :351-482 (reparametrizeFeatures), :484-537 (slideWindow), :151-227 (feature init).
//
// The reference evaluates feature removal, anchor re-parametrisation (J P J^T), the window slide
// (L P R, 0/1 matrices) and clone augmentation (J P J^T) as dense N^3 products.  Their composition is
//     P' = A P A^T,   A = J_aug * Z * L_slide * J_reparam * C_removal
// where every row of A is zero, a unit vector e_src, or one of a few <=15-entry rows (3 per re-anchored
// feature, 6 for the new clone).  So P' is a gather of P plus a thin correction: one HBM pass.
#include "xb_kernels.h"

namespace xb {

// rowmap encoding: >=0 unit row (source index), -1 zero row, <=-2 computed row with index (-2 - v)
struct ManageDev {
  int M, F, N, n_poses, n_features, slide, n_reanch;
  const int* feat_src;   // [F]
  const int* reanch;     // [n_reanch] new feature slots whose anchor moves from pose 0 to pose M-1
  double* combo_vals;    // [n_comp][15]
  double* scratch;       // >= 7M + 3F doubles
};

__global__ void __launch_bounds__(128) k_manage_prep(ManageDev md, double* __restrict__ xv) {
  XB_PDL_SHORT();
  const int t = threadIdx.x, nt = blockDim.x;
  const int M = md.M, F = md.F;
  double* parr = xv + XV_ARR;
  double* qarr = xv + XV_ARR + 3 * M;
  double* farr = xv + XV_ARR + 7 * M;
  double* sc_f = md.scratch;            // 3F
  double* sc_p = md.scratch + 3 * F;    // 3M
  double* sc_q = sc_p + 3 * M;          // 4M
  __shared__ double cam[7];             // new camera pose (q x,y,z,w ; p)
  __shared__ double Jth[9], Rci[9];

  if (t == 0) {
    // State::computeCameraOrientation / computeCameraPosition (state.cpp:185-195)
    double qn[4] = {xv[XV_Q], xv[XV_Q + 1], xv[XV_Q + 2], xv[XV_Q + 3]};
    double qi[4] = {xv[XV_QIC], xv[XV_QIC + 1], xv[XV_QIC + 2], xv[XV_QIC + 3]};
    xb_qnormalize(qn);
    xb_qnormalize(qi);
    xb_qmul(qn, qi, cam);
    double R[9], rp[3], sk[9], m[9];
    xb_rot_raw(qn, R);
    xb_mv33(R, &xv[XV_PIC], rp);
    for (int e = 0; e < 3; ++e) cam[4 + e] = xv[XV_P + e] + rp[e];
    // augmentation Jacobians (state_manager.cpp:309-324)
    xb_skew(&xv[XV_PIC], sk);
    xb_mm33(R, sk, m);
    for (int e = 0; e < 9; ++e) Jth[e] = -m[e];
    const double qc[4] = {-qi[0], -qi[1], -qi[2], qi[3]};
    xb_rot_raw(qc, Rci);
  }
  // feature compaction (state_manager.cpp:62-70)
  for (int e = t; e < 3 * F; e += nt) {
    const int k = e / 3, src = md.feat_src[k];
    sc_f[e] = src >= 0 ? farr[3 * src + e % 3] : 0.0;
  }
  for (int e = t; e < 3 * M; e += nt) sc_p[e] = parr[e];
  for (int e = t; e < 4 * M; e += nt) sc_q[e] = qarr[e];
  __syncthreads();
  for (int e = t; e < 3 * F; e += nt) farr[e] = sc_f[e];
  __syncthreads();

  if (md.slide) {
    // anchor re-parametrisation, eq. 38 of Li 2012 (state_manager.cpp:351-482); old window still in sc_p / sc_q
    for (int i = t; i < md.n_reanch; i += nt) {
      const int j = md.reanch[i];
      const int idx1 = M - 1;
      double Rn[9], Ro[9], RnT[9];
      xb_rot(sc_q + 4 * idx1, Rn);
      xb_rot(sc_q, Ro);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) RnT[r * 3 + c] = Rn[c * 3 + r];
      const double ao = sc_f[3 * j], bo = sc_f[3 * j + 1], ro = sc_f[3 * j + 2];
      const double ab1[3] = {ao, bo, 1.0};
      double t3[3], vv[3], np_[3];
      xb_mv33(Ro, ab1, t3);
      for (int e = 0; e < 3; ++e) vv[e] = -sc_p[3 * idx1 + e] + sc_p[e] + 1.0 / ro * t3[e];
      xb_mv33(RnT, vv, np_);
      const double rn = 1.0 / np_[2], an = np_[0] * rn, bn = np_[1] * rn;
      farr[3 * j] = an; farr[3 * j + 1] = bn; farr[3 * j + 2] = rn;
      double RR[9], sk[9], Jao[9], Jan[9], m3[9], Jfo[9], tmp[9];
      xb_mm33(RnT, Ro, RR);
      xb_skew(ab1, sk);
      xb_mm33(RR, sk, tmp);
      for (int e = 0; e < 9; ++e) Jao[e] = -1.0 / ro * tmp[e];
      xb_skew(np_, Jan);
      xb_mat_ivd(ao, bo, ro, m3);
      xb_mm33(RR, m3, tmp);
      for (int e = 0; e < 9; ++e) Jfo[e] = 1.0 / ro * tmp[e];
      // A_j (3x15 over [pos_new att_new pos_old att_old feat_old]) then rho_new * mat * A_j
      double Aj[45];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          Aj[r * 15 + c] = -RnT[r * 3 + c];
          Aj[r * 15 + 3 + c] = Jan[r * 3 + c];
          Aj[r * 15 + 6 + c] = RnT[r * 3 + c];
          Aj[r * 15 + 9 + c] = Jao[r * 3 + c];
          Aj[r * 15 + 12 + c] = Jfo[r * 3 + c];
        }
      double mat[9] = {1.0, 0.0, -an, 0.0, 1.0, -bn, 0.0, 0.0, -rn};
      double* out = md.combo_vals + (size_t)(6 + 3 * i) * 15;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 15; ++c)
          out[r * 15 + c] = rn * (mat[r * 3] * Aj[c] + mat[r * 3 + 1] * Aj[15 + c] + mat[r * 3 + 2] * Aj[30 + c]);
    }
    // slide the pose window (state_manager.cpp:486-493)
    for (int e = t; e < 3 * M; e += nt) parr[e] = (e < 3 * (M - 1)) ? sc_p[e + 3] : 0.0;
    for (int e = t; e < 4 * M; e += nt) qarr[e] = (e < 4 * (M - 1)) ? sc_q[e + 4] : 0.0;
  }
  __syncthreads();
  const int pos = md.slide ? M - 1 : md.n_poses;
  if (t < 4) qarr[4 * pos + t] = cam[t];
  if (t < 3) parr[3 * pos + t] = cam[4 + t];
  // combo rows 0..2: new camera position (cols 0,1,2,6,7,8); rows 3..5: new camera attitude (cols 6,7,8)
  for (int e = t; e < 6 * 15; e += nt) {
    const int r = e / 15, c = e % 15;
    double v = 0.0;
    if (r < 3) {
      if (c < 3) v = (c == r) ? 1.0 : 0.0;
      else if (c < 6) v = Jth[r * 3 + (c - 3)];
    } else if (c < 3) {
      v = Rci[(r - 3) * 3 + c];
    }
    md.combo_vals[e] = v;
  }
}

// Source covariance of manage: either a materialised N x N matrix, or the "virtual" form a ring slot is stored in -- the
// 15 x N strip [P_ii | P_iv] of the slot plus the covariance generation that holds P_vv (P_vi = P_iv^T).  Reading the
// virtual form directly saves the assemble pass (one 5 MB read + write at cfg-2) at the head of every update.
struct PSrc {
  const double* P;      // materialised matrix, or nullptr
  const double* strip;  // 15 x N: [P_ii | P_iv]
  const double* strip2; // 15 x N: P_vi^T; == strip for a slot whose P_vi equals P_iv^T (every slot written by an update
                        // that symmetrised); differs after updates without measurement rows (propagator.cpp:197-203)
  const double* gen;    // N x N (only rows/cols >= 15 are read)
  int N;
  __device__ __forceinline__ double at(int r, int c) const {
    if (P) return P[(size_t)r * N + c];
    if (r < XB_CORE) return strip[(size_t)r * N + c];
    if (c < XB_CORE) return strip2[(size_t)c * N + r];
    return gen[(size_t)r * N + c];
  }
};

// T[ci][b] = sum_e val_e * P[col_e][b]   and   T2[ci][r] = sum_e P[r][col_e] * val_e
// The column-side products T2 are needed wherever P[r][col] != P[col][r]: for the core rows (r < 15) always, for every
// row when the source covariance is unsymmetric beyond the core block (`general`: the previous updates applied no
// measurement, state_manager.cpp:273-349 then copies unsymmetric blocks from clone to clone).
__global__ void k_manage_T(int N, int n_comp, const int* __restrict__ ccols, const double* __restrict__ cvals, PSrc P,
                           double* __restrict__ T, double* __restrict__ T2, int general) {
  XB_PDL_SHORT();
  const int ci = blockIdx.y;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (ci >= n_comp || b >= N) return;
  const bool col_side = general || b < XB_CORE;
  double s = 0.0, s2 = 0.0;
  for (int e = 0; e < 15; ++e) {
    const int col = ccols[ci * 15 + e];
    const double v = cvals[ci * 15 + e];
    s = fma(v, P.at(col, b), s);
    if (col_side) s2 = fma(P.at(b, col), v, s2);
  }
  T[(size_t)ci * N + b] = s;
  if (col_side) T2[(size_t)ci * N + b] = s2;
}

__global__ void __launch_bounds__(256) k_manage_apply(int N, const int* __restrict__ rowmap, const int* __restrict__ ccols,
                                                       const double* __restrict__ cvals, PSrc P, const double* __restrict__ T,
                                                       const double* __restrict__ T2, double* __restrict__ Pn, int general) {
  XB_PDL_SHORT();
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int i0 = blockIdx.y * 32 + (threadIdx.x >> 5) * 4;
  if (j >= N) return;
  const int mj = rowmap[j];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + u;
    if (i >= N) break;
    const int mi = rowmap[i];
    double v = 0.0;
    if (mi == -1 || mj == -1) {
      v = 0.0;
    } else if (mi >= 0 && mj >= 0) {
      v = P.at(mi, mj);
    } else if (mi <= -2 && mj >= 0) {
      v = T[(size_t)(-2 - mi) * N + mj];
    } else if (mi >= 0 && mj <= -2) {
      v = (general || mi < XB_CORE) ? T2[(size_t)(-2 - mj) * N + mi] : T[(size_t)(-2 - mj) * N + mi];
    } else {
      const int ci = -2 - mi, cj = -2 - mj;
      for (int e = 0; e < 15; ++e) v = fma(T[(size_t)ci * N + ccols[cj * 15 + e]], cvals[cj * 15 + e], v);
    }
    Pn[(size_t)i * N + j] = v;
  }
}

void launch_manage_dev(cudaStream_t s, int M, int F, int N, int n_poses, int n_features, int slide, int n_reanch,
                       const int* d_feat_src, const int* d_reanch, const int* d_rowmap, const int* d_ccols,
                       double* d_cvals, double* d_scratch, double* xv, const double* Pold, double* Pnew, double* Tm,
                       double* T2, const double* strip, const double* gen, const double* strip2, int general) {
  ManageDev md{M, F, N, n_poses, n_features, slide, n_reanch, d_feat_src, d_reanch, d_cvals, d_scratch};
  XB_LAUNCH(k_manage_prep, 1, 128, 0, s, md, xv);
  count_launch();
  const PSrc src{Pold, strip, strip2 ? strip2 : strip, gen, N};  // Pold == nullptr: read the slot's strips + generation directly
  const int n_comp = 6 + 3 * n_reanch;
  dim3 gt((N + 127) / 128, n_comp);
  XB_LAUNCH(k_manage_T, gt, 128, 0, s, N, n_comp, d_ccols, d_cvals, src, Tm, T2, general);
  count_launch();
  dim3 ga((N + 31) / 32, (N + 31) / 32);
  XB_LAUNCH(k_manage_apply, ga, 256, 0, s, N, d_rowmap, d_ccols, d_cvals, src, Tm, T2, Pnew, general);
  count_launch();
}

// ---- assemble / extract: strip <-> full covariance ------------------------------------------------
__global__ void k_assemble(int N, const double* __restrict__ strip, const double* __restrict__ strip2,
                           const double* __restrict__ Pg, double* __restrict__ Pw) {
  XB_PDL_SHORT();
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= N) return;
  double v;
  if (i < XB_CORE) v = strip[(size_t)i * N + j];
  else if (j < XB_CORE) v = strip2[(size_t)j * N + i];  // P_vi (strip2 == strip when P_vi = P_iv^T)
  else v = Pg[(size_t)i * N + j];
  Pw[(size_t)i * N + j] = v;
}
void launch_assemble(cudaStream_t s, int N, const double* strip, const double* Pgen, double* Pwork, const double* strip2) {
  dim3 g((N + 255) / 256, N);
  XB_LAUNCH(k_assemble, g, 256, 0, s, N, strip, strip2 ? strip2 : strip, Pgen, Pwork);
  count_launch();
}
// strip2[r][j] = Pwork[j][r]: the first 15 columns of the work covariance, stored like the row strip
__global__ void k_extract_cols(int N, const double* __restrict__ Pw, double* __restrict__ strip2) {
  XB_PDL_SHORT();
  const int j = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (j < N) strip2[(size_t)r * N + j] = Pw[(size_t)j * N + r];
}
void launch_extract_strip2(cudaStream_t s, int N, const double* Pwork, double* strip2) {
  dim3 g((N + 255) / 256, XB_CORE);
  XB_LAUNCH(k_extract_cols, g, 256, 0, s, N, Pwork, strip2);
  count_launch();
}
void launch_extract_strip(cudaStream_t s, int N, const double* Pwork, double* strip) {
  cudaMemcpyAsync(strip, Pwork, sizeof(double) * 15 * (size_t)N, cudaMemcpyDeviceToDevice, s);
}

// ---- feature initialisation ------------------------------------------------------------------------
// E[3t+r][b] = (H2_t^-1 H1_t)[r][b]  (3n x 6M);  f_new = f - E corr_pose + H2^-1 r1   (state_manager.cpp:151-174)
__global__ void k_featinit_E(FeatInitParams fp, double* __restrict__ E, double* __restrict__ H2inv, double* __restrict__ xv) {
  XB_PDL_SHORT();
  const int t = blockIdx.x;
  const int W = 6 * fp.M + 1;
  __shared__ double Hi[9];
  __shared__ double red[3][128];
  if (threadIdx.x == 0) xb_inv33(fp.H2 + 9 * (size_t)t, Hi);
  __syncthreads();
  const double* H1 = fp.H1 + (size_t)t * 3 * W;
  double acc[3] = {0, 0, 0};
  for (int b = threadIdx.x; b < 6 * fp.M; b += blockDim.x) {
    for (int r = 0; r < 3; ++r) {
      const double v = Hi[r * 3] * H1[b] + Hi[r * 3 + 1] * H1[W + b] + Hi[r * 3 + 2] * H1[2 * W + b];
      E[(size_t)(3 * t + r) * 6 * fp.M + b] = v;
      acc[r] = fma(v, fp.corr[XB_CORE + b], acc[r]);
    }
  }
  for (int r = 0; r < 3; ++r) red[r][threadIdx.x] = acc[r];
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int r = 0; r < 3; ++r) red[r][threadIdx.x] += red[r][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 9) H2inv[9 * (size_t)t + threadIdx.x] = Hi[threadIdx.x];
  if (threadIdx.x < 3) {
    const int r = threadIdx.x;
    const double r1[3] = {H1[6 * fp.M], H1[W + 6 * fp.M], H1[2 * W + 6 * fp.M]};
    const double f = fp.ivd[3 * t + r] - red[r][0] + (Hi[r * 3] * r1[0] + Hi[r * 3 + 1] * r1[1] + Hi[r * 3 + 2] * r1[2]);
    xv[XV_ARR + 7 * fp.M + 3 * (fp.n_features + t) + r] = f;
  }
}
// write cross and diagonal blocks (state_manager.cpp:216-219): rows/cols ns..ns+3n
__global__ void k_featinit_write(int N, int ns, int n3, const double* __restrict__ C, const double* __restrict__ Pdd,
                                 const double* __restrict__ H2inv, double var, double* __restrict__ P) {
  XB_PDL_SHORT();
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= N || r >= n3) return;
  double v;
  if (c >= ns && c < ns + n3) {
    const int cc = c - ns;
    v = Pdd[r * n3 + cc];
    if (r / 3 == cc / 3) {  // var * H2inv H2inv^T on the diagonal 3x3 block
      const double* Hi = H2inv + 9 * (r / 3);
      const int a = r % 3, b = cc % 3;
      v += var * (Hi[a * 3] * Hi[b * 3] + Hi[a * 3 + 1] * Hi[b * 3 + 1] + Hi[a * 3 + 2] * Hi[b * 3 + 2]);
    }
    P[(size_t)(ns + r) * N + c] = v;
  } else {
    v = -C[(size_t)r * N + c];
    P[(size_t)(ns + r) * N + c] = v;
    P[(size_t)c * N + ns + r] = v;
  }
}

void launch_init_msckf_slam(cudaStream_t s, const FeatInitParams& fp, double* xv, double* P, double* scratch) {
  const int n3 = 3 * fp.n_new, K6 = 6 * fp.M, N = fp.N;
  double* E = scratch;                       // n3 x 6M
  double* H2inv = E + (size_t)n3 * K6;       // n_new x 9
  double* C = H2inv + 9 * (size_t)fp.n_new;  // n3 x N
  double* Pdd = C + (size_t)n3 * N;          // n3 x n3
  XB_LAUNCH(k_featinit_E, fp.n_new, 128, 0, s, fp, E, H2inv, xv);
  count_launch();
  gemm_nn(s, n3, N, K6, 1.0, E, K6, P + (size_t)XB_CORE * N, N, 0.0, C, N);
  gemm_nt(s, n3, n3, K6, 1.0, C + XB_CORE, N, E, K6, 0.0, Pdd, n3);
  const int ns = XB_CORE + 6 * fp.M + 3 * fp.n_features;
  dim3 g((N + 127) / 128, n3);
  XB_LAUNCH(k_featinit_write, g, 128, 0, s, N, ns, n3, C, Pdd, H2inv, fp.var_img, P);
  count_launch();
}

// state_manager.cpp:176-198 + slam_update.cpp:216-242
__global__ void k_init_std(int M, int N, int n_features, int n_new, const int* __restrict__ off, const double* __restrict__ obs,
                           double rho0, double var_img, double var_rho0, double* __restrict__ xv, double* __restrict__ P) {
  XB_PDL_SHORT();
  const int ns = XB_CORE + 6 * M + 3 * n_features;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n3 = 3 * n_new;
  if (idx < n3 * N) {
    const int r = idx / N, c = idx % N;
    double v = 0.0;
    if (c == ns + r) v = (r % 3 == 2) ? var_rho0 : var_img;
    P[(size_t)(ns + r) * N + c] = v;
    if (c < ns || c >= ns + n3) P[(size_t)c * N + ns + r] = 0.0;
  }
  if (idx < n_new) {
    const double* z = obs + 2 * (size_t)(off[idx + 1] - 1);
    double* f = xv + XV_ARR + 7 * M + 3 * (n_features + idx);
    f[0] = z[0]; f[1] = z[1]; f[2] = rho0;
  }
}
void launch_init_std_slam(cudaStream_t s, int M, int F, int N, int n_features, int n_new, const int* off,
                          const double* obs, double rho0, double var_img, double var_rho0, double* xv, double* P) {
  (void)F;
  const int tot = 3 * n_new * N;
  XB_LAUNCH(k_init_std, (tot + 255) / 256, 256, 0, s, M, N, n_features, n_new, off, obs, rho0, var_img, var_rho0, xv, P);
  count_launch();
}

// P_j: scale listed 3x3 diagonal blocks by w (msckf_update.cpp:258-267, multi_slam_update.cpp:229-239)
__global__ void k_scale_blocks(double* P, int N, const int* cols, int n_blocks, double w) {
  XB_PDL_SHORT();
  const int b = blockIdx.x, e = threadIdx.x;
  if (b >= n_blocks || e >= 9) return;
  const int c = cols[b];
  P[(size_t)(c + e / 3) * N + c + e % 3] *= w;
}
void launch_scale_blocks(cudaStream_t s, double* P, int N, const int* cols, int n_blocks, double w) {
  if (n_blocks <= 0) return;
  XB_LAUNCH(k_scale_blocks, n_blocks, 32, 0, s, P, N, cols, n_blocks, w);
  count_launch();
}

__global__ void k_add_diag(double* A, int ld, int n, double v) {
  XB_PDL_SHORT();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[(size_t)i * ld + i] += v;
}
void launch_add_diag(cudaStream_t s, double* A, int ld, int n, double v) {
  XB_LAUNCH(k_add_diag, (n + 255) / 256, 256, 0, s, A, ld, n, v);
  count_launch();
}

}  // namespace xb
