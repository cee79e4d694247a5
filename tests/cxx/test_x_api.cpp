// Replays a recorded event stream (written by tests/test_gpu_cxx.py) through the C++ x:: API binding and dumps the
// resulting states, so that pytest can compare them with the ctypes path.  Build: g++ -std=c++17 ... -lxb200
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "x/xb200_binding.hpp"

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
  if (argc < 3) return 2;
  const std::vector<double> ev = read_all(argv[1]);
  size_t p = 0;
  const int M = (int)ev[p++], F = (int)ev[p++], max_tracks = (int)ev[p++];
  const double sigma_img = ev[p++];
  x::VioUpdater updater(sigma_img, 0.05, 0.5, 0.25, 10);
  x::Ekf ekf(updater);
  x::State def(M, F);
  ekf.set(updater, x::Vector3(0, 0, -9.81), x::ImuNoise(), 64, def, 50.0, 1, 0.005, max_tracks);
  const int N = XB_NERR(M, F), LX = XB_XVEC_LEN(M, F);
  std::vector<double> out;
  while (p < ev.size()) {
    const int type = (int)ev[p++];
    if (type == 0) {  // init: xvec + covariance (row-major in the file)
      x::State s(M, F);
      for (int i = 0; i < LX; ++i) s.xvec()[i] = ev[p++];
      x::Matrix c(N, N);
      for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) c(i, j) = ev[p++];
      s.setCovariance(c);
      ekf.initializeFromState(s);
    } else if (type == 1) {
      const double t = ev[p++]; const unsigned seq = (unsigned)ev[p++];
      x::Vector3 w(ev[p], ev[p + 1], ev[p + 2]), a(ev[p + 3], ev[p + 4], ev[p + 5]);
      p += 6;
      ekf.processImu(t, seq, w, a);
    } else {
      x::VioMeasurement m;
      m.timestamp = ev[p++];
      x::TrackList* lists[5] = {&m.slam_trks, &m.msckf_trks, &m.msckf_short_trks, &m.new_slam_std_trks, &m.new_msckf_slam_trks};
      for (auto* tl : lists) {
        const int nt = (int)ev[p++];
        for (int t = 0; t < nt; ++t) {
          const int L = (int)ev[p++];
          x::Track trk;
          for (int i = 0; i < L; ++i) { trk.emplace_back(ev[p], ev[p + 1]); p += 2; }
          tl->push_back(trk);
        }
      }
      const int nl = (int)ev[p++];
      for (int i = 0; i < nl; ++i) m.lost_slam_trk_idxs.push_back((unsigned)ev[p++]);
      updater.setMeasurement(m);
      auto st = ekf.processUpdateMeasurement();
      if (!st) { fprintf(stderr, "update returned nullopt\n"); return 3; }
      out.insert(out.end(), st->xvec().begin(), st->xvec().end());
    }
  }
  // newest (re-propagated) state
  x::State newest(M, F);
  xb_ekf_get_state(ekf.handle(), -1, newest.xvec().data());
  out.insert(out.end(), newest.xvec().begin(), newest.xvec().end());
  FILE* fo = fopen(argv[2], "wb");
  fwrite(out.data(), 8, out.size(), fo);
  fclose(fo);
  printf("ok %zu doubles\n", out.size());
  return 0;
}
