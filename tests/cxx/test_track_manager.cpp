// x::TrackManager::manageTracks through the source-compatible C++ API (include/x/vio/track_manager.h) against the raw
// C ABI (xb_tm_*) on the same inputs: checks the marshalling of matches / attitudes / lists and the copy semantics of
// the manager (VioUpdater keeps a copy of the TrackManager it is constructed with).  Host code only: runs without a GPU.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "x/vio/track_manager.h"
#include "x/vision/camera.h"
#include "x/vision/tiled_image.h"

using namespace x;

static unsigned long long lcg = 12345;
static double rnd() { lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL; return (double)((lcg >> 11) & 0xFFFFFFFFFFFFFULL) / (double)0x10000000000000ULL; }

int main() {
  const Camera cam(0.46, 0.61, 0.5, 0.5, 0.95, 640, 480);
  TrackManager tm(cam, 0.02, 0.02);
  TiledImage img(3, 4);
  xb_tm_config c{};
  c.fx = 0.46; c.fy = 0.61; c.cx = 0.5; c.cy = 0.5; c.s = 0.95; c.img_width = 640; c.img_height = 480;
  c.min_baseline_x_n = 0.02; c.min_baseline_y_n = 0.02; c.n_tiles_h = 3; c.n_tiles_w = 4;
  xb_track_manager* raw = xb_tm_create(&c);
  const int NP = 80, n_poses_max = 6, n_slam_max = 8, min_len = 3;
  std::vector<double> px(NP), py(NP), vx(NP), vy(NP);
  for (int i = 0; i < NP; ++i) { px[i] = 20 + 600 * rnd(); py[i] = 20 + 440 * rnd(); vx[i] = 3 + 4 * (rnd() - 0.5); vy[i] = 1 + 4 * (rnd() - 0.5); }
  AttitudeList rots;
  int checked = 0;
  TrackManager copy = tm;   // shares nothing yet
  for (int k = 0; k < 30; ++k) {
    const double ang = 0.012 * (k + 1);
    rots.emplace_back(std::sin(ang / 2) * 0.6, std::sin(ang / 2) * 0.8, 0.0, std::cos(ang / 2));
    if (rots.size() > (size_t)n_poses_max + 1) rots.erase(rots.begin());
    MatchList matches;
    std::vector<double> mv;
    for (int i = 0; i < NP; ++i) {
      const double nx = px[i] + vx[i], ny = py[i] + vy[i];
      const bool ok = nx > 5 && nx < 635 && ny > 5 && ny < 475 && rnd() > 0.05;
      if (ok && k > 0) {
        Match m;
        m.previous = Feature(0.1 * k, k, 0.0, 0.0, px[i], py[i]);
        m.current = Feature(0.1 * (k + 1), k + 1, 0.0, 0.0, nx, ny);
        matches.push_back(m);
        const double row[10] = {0, 0.1 * k, px[i], py[i], 0.1 * (k + 1), nx, ny, 0, 0, 0};
        mv.insert(mv.end(), row, row + 10);
      }
      if (ok) { px[i] = nx; py[i] = ny; } else { px[i] = 20 + 600 * rnd(); py[i] = 20 + 440 * rnd(); }
    }
    std::vector<double> r4;
    for (const Attitude& a : rots) { r4.push_back(a.ax); r4.push_back(a.ay); r4.push_back(a.az); r4.push_back(a.aw); }
    tm.manageTracks(matches, rots, n_poses_max, n_slam_max, min_len, img);
    if (xb_tm_manage_tracks(raw, mv.data(), (int)mv.size() / 10, r4.data(), (int)rots.size(), n_poses_max, n_slam_max, min_len) != XB_OK) return 2;
    if (k == 10) copy = tm;   // a copy made mid-sequence keeps following the same track state
    const TrackManager& use = k >= 10 ? copy : tm;
    const TrackList lists[5] = {use.getMsckfTracks(), use.getShortMsckfTracks(), use.getNewSlamStdTracks(), use.getNewSlamMsckfTracks(),
                                use.normalizeSlamTracks(n_poses_max)};
    for (int w = 0; w < 5; ++w) {
      int nt = 0, no = 0;
      xb_tm_list_size(raw, w, w == 4 ? n_poses_max : 0, &nt, &no);
      std::vector<int> off(nt + 1);
      std::vector<double> xy(2 * (no + 1));
      xb_tm_get_list(raw, w, w == 4 ? n_poses_max : 0, off.data(), xy.data(), nullptr);
      if ((int)lists[w].size() != nt) { std::printf("frame %d list %d: %zu vs %d tracks\n", k, w, lists[w].size(), nt); return 1; }
      for (int i = 0; i < nt; ++i) {
        if ((int)lists[w][i].size() != off[i + 1] - off[i]) return 1;
        for (int j = off[i]; j < off[i + 1]; ++j) {
          if (lists[w][i][j - off[i]].getX() != xy[2 * j] || lists[w][i][j - off[i]].getY() != xy[2 * j + 1]) return 1;
          ++checked;
        }
        if (w == 4 && (int)lists[w][i].size() > n_poses_max) return 1;
      }
    }
    std::vector<int> lost(64);
    const int nl = xb_tm_lost_slam_idxs(raw, lost.data(), 64);
    const std::vector<unsigned int> l2 = use.getLostSlamTrackIndexes();
    if ((int)l2.size() != nl) return 1;
    for (int i = 0; i < nl; ++i) if ((int)l2[i] != lost[i]) return 1;
  }
  xb_tm_destroy(raw);
  std::printf("x::TrackManager == xb_tm_* over 30 frames, %d observations compared\n", checked);
  return checked > 500 ? 0 : 3;
}
