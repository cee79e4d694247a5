// Drives the C++ x::VIO facade (include/x/vio/vio.h) the way a caller of the reference does -- loadParamsFromYaml, setUp,
// initAtTime, processImu, setLastRangeMeasurement / setLastSunAngleMeasurement, processMatchesMeasurement on 10-double
// match vectors -- over an event stream written by tests/test_gpu_vio.py, and dumps the updated states so that pytest can
// compare them with the Python facade on the same stream.  Build: g++ -std=c++17 -I include -I <Eigen> ... -lxb200
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
  return 0;
}
