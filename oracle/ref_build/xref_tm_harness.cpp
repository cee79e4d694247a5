// Test infrastructure (oracle/): drives the reference's own TrackManager (src/x/vio/track_manager.cpp, compiled unmodified
// where it lies, with src/x/vision/{camera,feature,track,tiled_image}.cpp) through the path VIO::importMatches ->
// TrackManager::manageTracks -> getters, so that oracle/track_manager.py and the product (xb_tm_*, libxb200.so) can be
// pinned against it (single-agent build flavour; the -DMULTI_UAV short-track rule needs the tracker's opportunistic
// ids and is not pinned).  No reference source is copied; the OpenCV types come from shim/opencv2/xref_cv.hpp.
#include <cstring>
#include <vector>

#include "x/vio/track_manager.h"
#include "x/vision/camera.h"
#include "x/vision/tiled_image.h"
#include "x/vision/types.h"

using namespace x;

struct TmHandle {
  Camera cam;
  TrackManager tm;
  unsigned n_tiles_h, n_tiles_w, w, h;
};

static const TrackList get_list(TmHandle* t, int which, int size_out) {
  switch (which) {
    case 0: return t->tm.getMsckfTracks();
    case 1: return t->tm.getShortMsckfTracks();
    case 2: return t->tm.getNewSlamStdTracks();
    case 3: return t->tm.getNewSlamMsckfTracks();
    case 4: return t->tm.normalizeSlamTracks(size_out);
    default: return t->tm.getOppTracks();
  }
}

extern "C" {
void* xref_tm_create(double fx, double fy, double cx, double cy, double s, unsigned w, unsigned h, double bx, double by,
                     unsigned n_tiles_h, unsigned n_tiles_w) {
  TmHandle* t = new TmHandle();
  t->cam = Camera(fx, fy, cx, cy, s, w, h);
  t->tm = TrackManager(t->cam, bx, by);
  t->n_tiles_h = n_tiles_h; t->n_tiles_w = n_tiles_w; t->w = w; t->h = h;
  return t;
}
void xref_tm_destroy(void* p) { delete (TmHandle*)p; }
// VIO::importMatches (vio.cpp:372-434) restated for the harness (VIO itself needs the whole front end), then the
// reference's manageTracks
int xref_tm_manage(void* p, const double* mv, int n_matches, unsigned seq, const double* rots, int n_rots, int n_poses_max,
                   int n_slam_max, int min_track_length) {
  TmHandle* t = (TmHandle*)p;
  MatchList matches(n_matches);
  for (int i = 0; i < n_matches; ++i) {
    Feature prev(mv[10 * i + 1], seq - 1, 0.0, 0.0, mv[10 * i + 2], mv[10 * i + 3], -1.0);
    t->cam.undistort(prev);
    Feature cur(mv[10 * i + 4], seq, 0.0, 0.0, mv[10 * i + 5], mv[10 * i + 6], -1.0);
    t->cam.undistort(cur);
    matches[i].previous = prev;
    matches[i].current = cur;
  }
  AttitudeList cam_rots;
  for (int i = 0; i < n_rots; ++i) cam_rots.push_back(Attitude(rots[4 * i], rots[4 * i + 1], rots[4 * i + 2], rots[4 * i + 3]));
  cv::Mat base((int)t->h, (int)t->w, CV_8UC1);
  TiledImage img(base, 0.0, seq, t->n_tiles_h, t->n_tiles_w, 40);
  t->tm.manageTracks(matches, cam_rots, n_poses_max, n_slam_max, min_track_length, img);
  return 0;
}
int xref_tm_list_size(void* p, int which, int size_out, int* n_tracks, int* n_obs) {
  const TrackList l = get_list((TmHandle*)p, which, size_out);
  int n = 0;
  for (const Track& tr : l) n += (int)tr.size();
  *n_tracks = (int)l.size();
  *n_obs = n;
  return 0;
}
int xref_tm_get_list(void* p, int which, int size_out, int* offsets, double* xy) {
  const TrackList l = get_list((TmHandle*)p, which, size_out);
  int o = 0;
  offsets[0] = 0;
  for (size_t i = 0; i < l.size(); ++i) {
    for (const Feature& f : l[i]) { xy[2 * o] = f.getX(); xy[2 * o + 1] = f.getY(); ++o; }
    offsets[i + 1] = o;
  }
  return (int)l.size();
}
int xref_tm_lost(void* p, int* idx, int cap) {
  const std::vector<unsigned int> l = ((TmHandle*)p)->tm.getLostSlamTrackIndexes();
  for (size_t i = 0; i < l.size() && (int)i < cap; ++i) idx[i] = (int)l[i];
  return (int)l.size();
}
int xref_tm_remove_persistent(void* p, unsigned idx) { ((TmHandle*)p)->tm.removePersistentTracksAtIndex(idx); return 0; }
int xref_tm_remove_new_persistent(void* p, const unsigned* idx, int n) {
  ((TmHandle*)p)->tm.removeNewPersistentTracksAtIndexes(std::vector<unsigned int>(idx, idx + n));
  return 0;
}
}
