// Drives the C++ x::VIO facade (include/x/vio/vio.h) the way a caller of the reference does -- loadParamsFromYaml, setUp,
// initAtTime, processImu, setLastRangeMeasurement / setLastSunAngleMeasurement, processMatchesMeasurement on 10-double
// match vectors -- over an event stream written by tests/test_gpu_vio.py, and dumps the updated states so that pytest can
// compare them with the Python facade on the same stream.  Build: g++ -std=c++17 -I include -I <Eigen> ... -lxb200
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "x/vio/vio.h"

static std::vector<double> read_all(const char* path) {
  FILE* fp = fopen(path, "rb");
  if (!fp) { perror(path); exit(2); }
  fseek(fp, 0, SEEK_END);
  long n = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  std::vector<double> d(n / 8);
  if (fread(d.data(), 8, d.size(), fp) != d.size()) exit(2);
  fclose(fp);
  return d;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  x::VIO vio;
  x::fsm::path yaml(argv[1]);
  const x::Params params = vio.loadParamsFromYaml(yaml);
  if (argc > 4) {   // parameter dump only (CPU test of the loader)
    // Camera::undistort + Camera::normalize of the binding (camera.cpp:69-87, 122-135) against the camera model of the
    // library (xb_tm_normalize_point), without and with FOV distortion
    for (const double s : {0.0, 0.9}) {
      const x::Camera cam(0.46, 0.61, 0.5, 0.5, s, 640, 480);
      for (int k = 0; k < 50; ++k) {
        x::Feature f;
        f.setXDist(13.0 * k + 0.25);
        f.setYDist(479.0 - 9.5 * k);
        const x::Feature lib = cam.undistortAndNormalize(f);
        cam.undistort(f);
        const x::Feature own = cam.normalize(f);
        if (std::fabs(own.getX() - lib.getX()) > 1e-15 || std::fabs(own.getY() - lib.getY()) > 1e-15) {
          fprintf(stderr, "camera model mismatch at s = %g, k = %d: %.17g %.17g vs %.17g %.17g\n", s, k, own.getX(), own.getY(),
                  lib.getX(), lib.getY());
          return 10;
        }
      }
    }
    printf("%d %d %d %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %d %d %s\n", params.n_poses_max,
           params.n_slam_features_max, params.min_track_length, params.state_buffer_size, params.n_tiles_h, params.cam_fx,
           params.sigma_img, params.q.w(), params.q.x(), params.q_ic.z(), params.g(2), params.sigma_dtheta(1), params.msckf_baseline,
           params.img_width, params.non_max_supp ? 1 : 0, params.vocabulary_path.c_str());
    return 0;
  }
  const std::vector<double> ev = read_all(argv[2]);
  vio.setUp(params, 256);
  if (vio.isInitialized()) return 3;
  vio.initAtTime(0.0);
  if (!vio.isInitialized()) return 3;
  const int M = params.n_poses_max, F = params.n_slam_features_max, LX = XB_XVEC_LEN(M, F);
  std::vector<double> out;
  size_t p = 0;
  int n_updates = 0;
  while (p < ev.size()) {
    const int type = (int)ev[p++];
    if (type == 1) {          // IMU sample
      const double t = ev[p]; const unsigned seq = (unsigned)ev[p + 1];
      const x::Vector3 w(ev[p + 2], ev[p + 3], ev[p + 4]), a(ev[p + 5], ev[p + 6], ev[p + 7]);
      p += 8;
      vio.processImu(t, seq, w, a);
    } else if (type == 3) {   // laser range
      x::RangeMeasurement r;
      r.timestamp = ev[p]; r.range = ev[p + 1];
      p += 2;
      vio.setLastRangeMeasurement(r);
    } else if (type == 4) {   // sun angles
      x::SunAngleMeasurement s;
      s.timestamp = ev[p]; s.x_angle = ev[p + 1]; s.y_angle = ev[p + 2];
      p += 3;
      vio.setLastSunAngleMeasurement(s);
    } else {                  // matches
      const double t = ev[p]; const unsigned seq = (unsigned)ev[p + 1]; const size_t n = (size_t)ev[p + 2];
      p += 3;
      const std::vector<double> mv(ev.begin() + p, ev.begin() + p + 10 * n);
      p += 10 * n;
      const auto updated = vio.processMatchesMeasurement(t, seq, mv);
      out.push_back(updated.has_value() ? 1.0 : 0.0);
      if (updated.has_value()) {
        if (updated->getTime() != t) return 4;   // the state carries the image timestamp (vio.cpp:314-316)
        const std::vector<double> x = updated->xvec();
        out.insert(out.end(), x.begin(), x.end());
        ++n_updates;
      } else {
        out.insert(out.end(), (size_t)LX, 0.0);
      }
    }
  }
  // SLAM features of the newest state in world coordinates (computeSLAMCartesianFeaturesForState)
  std::vector<double> newest(LX);
  xb_ekf_get_state(vio.ekf().handle(), -1, newest.data());
  x::State s(M, F);
  s.setFromXvec(newest.data(), M, F);
  const std::vector<x::Vector3> xyz = vio.computeSLAMCartesianFeaturesForState(s);
  out.push_back((double)xyz.size());
  for (const auto& v : xyz) { out.push_back(v(0)); out.push_back(v(1)); out.push_back(v(2)); }
  FILE* fo = fopen(argv[3], "wb");
  fwrite(out.data(), 8, out.size(), fo);
  fclose(fo);
  printf("ok %d updates\n", n_updates);
#ifdef MULTI_UAV
  // -DMULTI_UAV build: a second agent (a second filter on the same GPU) replays the same stream, sends its data
  // (VIO::getDataToSend, vio.cpp:440-452) and this agent fuses it (VIO::processOtherMeasurements, vio.cpp:498-574) with
  // every SLAM feature matched to the peer's feature of the same index.  The peer's estimates equal this agent's, so
  // the landmark residuals are zero: covariance intersection must leave the estimates where they are and return a
  // finite, symmetric covariance that differs from the prior one.
  {
    x::VIO peer;
    peer.setUp(params, 256);
    peer.initAtTime(0.0);
    size_t q = 0;
    double t_last = 0.0;
    while (q < ev.size()) {
      const int type = (int)ev[q++];
      if (type == 1) {
        peer.processImu(ev[q], (unsigned)ev[q + 1], x::Vector3(ev[q + 2], ev[q + 3], ev[q + 4]), x::Vector3(ev[q + 5], ev[q + 6], ev[q + 7]));
        q += 8;
      } else if (type == 3) {
        x::RangeMeasurement r; r.timestamp = ev[q]; r.range = ev[q + 1]; q += 2;
        peer.setLastRangeMeasurement(r);
      } else if (type == 4) {
        x::SunAngleMeasurement sa; sa.timestamp = ev[q]; sa.x_angle = ev[q + 1]; sa.y_angle = ev[q + 2]; q += 3;
        peer.setLastSunAngleMeasurement(sa);
      } else {
        const size_t n = (size_t)ev[q + 2];
        t_last = ev[q];
        peer.processMatchesMeasurement(ev[q], (unsigned)ev[q + 1], std::vector<double>(ev.begin() + q + 3, ev.begin() + q + 3 + 10 * n));
        q += 3 + 10 * n;
      }
    }
    // the peer's newest state with its covariance
    std::vector<double> px(LX);
    const int N = XB_NERR(M, F);
    std::vector<double> pc((size_t)N * N);
    xb_ekf_get_state(peer.ekf().handle(), -1, px.data());
    xb_ekf_get_covariance(peer.ekf().handle(), -1, pc.data(), XB_COL_MAJOR);
    x::State ps(M, F);
    ps.setFromXvec(px.data(), M, F);
    x::Matrix pcov(N, N);
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) pcov(i, j) = pc[(size_t)j * N + i];
    ps.setCovariance(pcov);
    std::shared_ptr<x::SimpleState> sent;
    x::TrackList msckf_tracks, slam_tracks, opp_tracks;
    std::vector<int> anchors;
    peer.getDataToSend(sent, ps, msckf_tracks, slam_tracks, anchors, opp_tracks);
    if (!sent || (int)anchors.size() != F || sent->getErrorStateSize() != N) { fprintf(stderr, "getDataToSend\n"); return 6; }
    const int nf = (int)xyz.size();
    std::vector<x::VIO::Correspondence> slam_corr, msckf_corr;
    for (int i = 0; i < nf; ++i) slam_corr.push_back({(unsigned long long)i, i});
    x::TrackListPtr received_msckf;
    for (const auto& t : msckf_tracks) received_msckf.push_back(std::make_shared<x::Track>(t));
    std::vector<double> before(LX), pb((size_t)N * N);
    xb_ekf_get_state(vio.ekf().handle(), -1, before.data());
    xb_ekf_get_covariance(vio.ekf().handle(), -1, pb.data(), XB_COL_MAJOR);
    // no correspondence: nothing happens (vio.cpp:548-550)
    if (vio.processOtherMeasurements(t_last, 1, sent->getDynamicState(), sent->getPositionState(), sent->getOrientationState(),
                                     sent->getFeatureState(), sent->getCovariance(), received_msckf, anchors, {}, {}).has_value()) return 7;
    const auto fused = vio.processOtherMeasurements(t_last, 1, sent->getDynamicState(), sent->getPositionState(),
                                                    sent->getOrientationState(), sent->getFeatureState(), sent->getCovariance(),
                                                    received_msckf, anchors, slam_corr, msckf_corr);
    if (!fused.has_value()) { fprintf(stderr, "processOtherMeasurements returned nullopt\n"); return 7; }
    std::vector<double> after(LX), pa((size_t)N * N);
    xb_ekf_get_state(vio.ekf().handle(), -1, after.data());
    xb_ekf_get_covariance(vio.ekf().handle(), -1, pa.data(), XB_COL_MAJOR);
    double dx = 0.0, dp = 0.0, asym = 0.0;
    for (int i = 0; i < LX; ++i) if (i != 31) dx = std::max(dx, std::fabs(after[i] - before[i]));
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) {
        if (!std::isfinite(pa[(size_t)j * N + i])) { fprintf(stderr, "covariance not finite\n"); return 8; }
        dp = std::max(dp, std::fabs(pa[(size_t)j * N + i] - pb[(size_t)j * N + i]));
        asym = std::max(asym, std::fabs(pa[(size_t)j * N + i] - pa[(size_t)i * N + j]));
      }
    printf("multi: %d SLAM correspondences, max |dx| = %.3e, max |dP| = %.3e, asym = %.3e\n", nf, dx, dp, asym);
    if (dx > 1e-9 || dp == 0.0) return 9;
    printf("multi ok\n");
  }
#endif
  return 0;
}
