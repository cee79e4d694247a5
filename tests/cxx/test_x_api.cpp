// Replays a recorded event stream (written by tests/test_gpu_cxx.py) through the C++ x:: API binding -- the
// reference's call sequence: VIO::setUp (vio.cpp:201-214), Ekf::initializeFromState, Ekf::processImu,
// VioUpdater::setMeasurement + Ekf::processUpdateMeasurement -- and dumps the resulting states, so that pytest can
// compare them with the ctypes path.  Build: g++ -std=c++17 -I include -I <Eigen> ... -lxb200
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "x/ekf/ekf.h"
#include "x/vio/vio_updater.h"

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
  x::Tracker tracker;
  x::StateManager state_manager(M, F);
  x::TrackManager track_manager;
  x::VioUpdater vio_updater;
  x::Ekf ekf(vio_updater);
  vio_updater = x::VioUpdater(tracker, state_manager, track_manager, sigma_img, 0.05, 0.5, 0.25, 10);
  const x::State default_state = x::State(M, F);
  ekf.set(vio_updater, x::Vector3(0, 0, -9.81), x::ImuNoise(), 64, default_state, 50.0, 1, 0.005, max_tracks);
  const int N = XB_NERR(M, F), LX = XB_XVEC_LEN(M, F);
  std::vector<double> out;
  bool cov_checked = false;
  while (p < ev.size()) {
    const int type = (int)ev[p++];
    if (type == 0) {  // init: xvec + covariance (row-major in the file)
      x::State s(M, F);
      s.setFromXvec(&ev[p], M, F);
      p += LX;
      x::Matrix c(N, N);
      for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) c(i, j) = ev[p++];
      s.setCovariance(c);
      vio_updater.stateManager().clear();   // VIO::initAtTime (vio.cpp:58-59)
      ekf.initializeFromState(s);
    } else if (type == 1) {
      const double t = ev[p++]; const unsigned seq = (unsigned)ev[p++];
      x::Vector3 w(ev[p], ev[p + 1], ev[p + 2]), a(ev[p + 3], ev[p + 4], ev[p + 5]);
      p += 6;
      ekf.processImu(t, seq, w, a);
    } else {
      x::VioMeasurement m;
      m.timestamp = ev[p++];
      x::TrackList lists[5];
      for (auto& tl : lists) {
        const int nt = (int)ev[p++];
        for (int t = 0; t < nt; ++t) {
          const int L = (int)ev[p++];
          x::Track trk;
          for (int i = 0; i < L; ++i) { trk.push_back(x::Feature(m.timestamp, ev[p], ev[p + 1])); p += 2; }
          tl.push_back(trk);
        }
      }
      std::vector<unsigned int> lost;
      const int nl = (int)ev[p++];
      for (int i = 0; i < nl; ++i) lost.push_back((unsigned)ev[p++]);
      vio_updater.trackManager().setTracks(lists[0], lists[1], lists[2], lists[3], lists[4], lost);
      // range / sun-angle members of the VioMeasurement (vio/types.h:300-305) + the facet the track manager reports
      std::vector<int> facet;
      if (ev[p++] != 0.0) {
        m.range.timestamp = ev[p]; m.range.range = ev[p + 1];
        m.range.img_pt_n.setX(ev[p + 2]); m.range.img_pt_n.setY(ev[p + 3]);
        facet = {(int)ev[p + 4], (int)ev[p + 5], (int)ev[p + 6]};
        p += 7;
      }
      vio_updater.trackManager().setFacet(facet);
      if (ev[p++] != 0.0) {
        m.sun_angle.timestamp = ev[p]; m.sun_angle.x_angle = ev[p + 1]; m.sun_angle.y_angle = ev[p + 2];
        p += 3;
      }
      vio_updater.setMeasurement(m);
      auto updated_state = ekf.processUpdateMeasurement();
      if (!updated_state.has_value()) { fprintf(stderr, "update returned nullopt\n"); return 3; }
      const std::vector<double> x = updated_state->xvec();
      out.insert(out.end(), x.begin(), x.end());
      if (!cov_checked && out.size() > 4 * (size_t)LX) {
        // the covariance of a returned State is fetched lazily from the ring slot it was written to: symmetric after an
        // update with rows, and the same matrix the C ABI hands out for that slot
        const x::Matrix P = updated_state->getCovariance();
        std::vector<double> Pc((size_t)N * N);
        xb_ekf_get_covariance(ekf.handle(), xb_ekf_last_update_slot(ekf.handle()), Pc.data(), XB_COL_MAJOR);
        for (int i = 0; i < N; ++i)
          for (int j = 0; j < N; ++j)
            if (P(i, j) != Pc[(size_t)j * N + i] || P(i, j) != P(j, i)) { fprintf(stderr, "lazy covariance mismatch\n"); return 4; }
        if (updated_state->getPoseCovariance().rows() != 6) return 4;
        cov_checked = true;
      }
    }
  }
  // newest (re-propagated) state
  std::vector<double> newest(LX);
  xb_ekf_get_state(ekf.handle(), -1, newest.data());
  out.insert(out.end(), newest.begin(), newest.end());
  FILE* fo = fopen(argv[2], "wb");
  fwrite(out.data(), 8, out.size(), fo);
  fclose(fo);
  printf("ok %zu doubles\n", out.size());
  return cov_checked ? 0 : 5;
}
