// Track management: turns the front end's feature matches into the five track lists the filter update consumes.
// Host code (branchy list bookkeeping on <= ~1k items per frame); lives in libxb200.so behind the C ABI of
// include/xb200.h (xb_tm_*), no device work.
//
// reference: src/x/vio/track_manager.cpp:115-436 (TrackManager::manageTracks), :576-636 (checkBaseline), :36-113 (getters,
//            removal), src/x/vio/vio.cpp:372-434 (VIO::importMatches: the 10-double match vector), src/x/vision/camera.cpp:
//            69-160 (FOV undistortion, normalisation), src/x/vision/tiled_image.cpp:139-158 (tile of a feature),
//            src/x/vision/feature.cpp:47-67 (feature equality).
//
// Behaviour that is reproduced on purpose (tests/test_track_manager.py pins it against the reference's own
// track_manager.cpp compiled in place):
//   * a persistent (SLAM) track is continued by the FIRST match whose previous feature equals its last feature
//     (relative-epsilon comparison of the undistorted pixel coordinates), and that match is consumed;
//   * lost persistent tracks are reported with the index they had BEFORE the removals of this call;
//   * opportunistic tracks are ordered by std::sort on the length (not stable: the order among equal lengths is the one
//     libstdc++'s introsort produces, and it decides which tracks get the free SLAM slots);
//   * the tile-balancing rule may evict the youngest persistent track of the fullest tile for a long-enough
//     opportunistic track of a tile that has at least two features less.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/xb200.h"

namespace {

struct Feat {
  double t = 0.0;
  double x = 0.0, y = 0.0;    // undistorted pixel coordinates
  double xd = 0.0, yd = 0.0;  // distorted (measured) pixel coordinates
};
struct Trk {
  std::vector<Feat> f;
  unsigned long long id = 0;
};
typedef std::vector<Trk> TrkList;

// Feature::operator== (feature.cpp:47-67)
bool nearly_equal(double a, double b) {
  const double aa = std::fabs(a), ab = std::fabs(b), diff = std::fabs(a - b);
  if (a == b) return true;
  if (a == 0 || b == 0 || (aa + ab < std::numeric_limits<double>::min()))
    return diff < (std::numeric_limits<double>::epsilon() * std::numeric_limits<double>::min());
  return diff / std::min(aa + ab, std::numeric_limits<double>::max()) < std::numeric_limits<double>::epsilon();
}
bool same_feature(const Feat& a, const Feat& b) { return nearly_equal(a.x, b.x) && nearly_equal(a.y, b.y); }

struct Quat {
  double w, x, y, z;
};
Quat qmul(const Quat& a, const Quat& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
Quat qconj(const Quat& a) { return {a.w, -a.x, -a.y, -a.z}; }
Quat qnormalized(const Quat& a) {
  const double n = std::sqrt(a.w * a.w + a.x * a.x + a.y * a.y + a.z * a.z);
  return {a.w / n, a.x / n, a.y / n, a.z / n};
}

}  // namespace

struct xb_track_manager {
  xb_tm_config cfg;
  // Camera (camera.cpp:27-48): focal lengths / principal point are given as fractions of the image size
  double fx, fy, cx, cy, inv_fx, inv_fy, cx_n, cy_n, s_term;
  double tile_w, tile_h;
  TrkList slam, new_slam, opp;
  TrkList msckf_n, msckf_short_n, new_std_n, new_msckf_n;  // normalised coordinates
  std::vector<unsigned> lost;
  std::vector<unsigned long long> opp_ids;  // MULTI_UAV: opportunistic tracks matched with another agent's
  unsigned long long next_id = 0;

  void undistort(Feat& f) const {  // Camera::undistort (camera.cpp:69-87)
    const double dx = f.xd * inv_fx - cx_n, dy = f.yd * inv_fy - cy_n;
    const double r = std::sqrt(dx * dx + dy * dy);
    double k = 1.0;
    if (r > 0.01) k = (cfg.s == 0.0 ? r : std::tan(r * cfg.s) * s_term) / r;
    f.x = k * dx * fx + cx;
    f.y = k * dy * fy + cy;
  }
  Trk normalize(const Trk& t, size_t max_size) const {  // Camera::normalize (camera.cpp:103-137): crops from the end
    const size_t n = t.f.size(), n_out = max_size ? std::min(max_size, n) : n;
    Trk o;
    o.id = t.id;
    o.f.resize(n_out);
    for (size_t j = n - n_out; j < n; ++j) {
      Feat g = t.f[j];
      g.x = t.f[j].x * inv_fx - cx_n;
      g.y = t.f[j].y * inv_fy - cy_n;
      g.xd = t.f[j].xd * inv_fx - cx_n;
      g.yd = t.f[j].yd * inv_fy - cy_n;
      o.f[j - (n - n_out)] = g;
    }
    return o;
  }
  int tile_of(const Feat& f) const {  // TiledImage::setTileForFeature (tiled_image.cpp:139-158)
    double c = f.xd - tile_w - 0.5;
    int col = 0;
    while (c > 0) { col += 1; c -= tile_w; }
    double r = (double)cfg.img_height - f.yd - 0.5;
    int row = (int)cfg.n_tiles_h - 1;
    while (r > tile_h) { row -= 1; r -= tile_h; }
    return row * (int)cfg.n_tiles_w + col;
  }
  // TrackManager::checkBaseline (track_manager.cpp:576-636): the observations are rotated into the last camera frame
  // and the spread of the normalised coordinates is compared with the thresholds
  bool baseline_ok(const Trk& t, const double* rots, int n_rots) const {
    const int n_obs = (int)t.f.size();
    if (n_obs < 2 || n_rots < n_obs) return false;
    const int i_last = n_rots - 1, i_first = i_last - n_obs + 1;
    double min_x = t.f[n_obs - 1].x, max_x = min_x, min_y = t.f[n_obs - 1].y, max_y = min_y;
    const Quat qn = qnormalized({rots[4 * i_last + 3], rots[4 * i_last], rots[4 * i_last + 1], rots[4 * i_last + 2]});
    for (int i = i_first; i <= i_last; ++i) {
      const Quat qi = qnormalized({rots[4 * i + 3], rots[4 * i], rots[4 * i + 1], rots[4 * i + 2]});
      const Quat rel = qmul(qconj(qi), qn);  // Cn_q_Ci
      const Quat ray = {0.0, t.f[i - i_first].x, t.f[i - i_first].y, 1.0};
      const Quat rn = qmul(qmul(qconj(rel), ray), rel);
      const double fx_ = rn.x / rn.z, fy_ = rn.y / rn.z;
      if (fx_ < min_x) min_x = fx_; else if (fx_ > max_x) max_x = fx_;
      if (fy_ < min_y) min_y = fy_; else if (fy_ > max_y) max_y = fy_;
    }
    return (max_x - min_x) > cfg.min_baseline_x_n || (max_y - min_y) > cfg.min_baseline_y_n;
  }
  const TrkList* list(int which, int size_out, TrkList& tmp) const {
    switch (which) {
      case XB_TM_MSCKF: return &msckf_n;
      case XB_TM_MSCKF_SHORT: return &msckf_short_n;
      case XB_TM_NEW_SLAM_STD: return &new_std_n;
      case XB_TM_NEW_SLAM_MSCKF: return &new_msckf_n;
      case XB_TM_SLAM:  // TrackManager::normalizeSlamTracks (track_manager.cpp:36-38)
        tmp.clear();
        for (const Trk& t : slam) tmp.push_back(normalize(t, (size_t)std::max(0, size_out)));
        return &tmp;
      case XB_TM_OPP:   // TrackManager::getOppTracks (track_manager.cpp:54-61): cropped to the number of tracks, as written
        tmp.clear();
        for (const Trk& t : opp) tmp.push_back(normalize(t, opp.size()));
        return &tmp;
      default: return nullptr;
    }
  }
};

extern "C" {

XB_API xb_track_manager* xb_tm_create(const xb_tm_config* cfg) {
  if (!cfg || cfg->img_width == 0 || cfg->img_height == 0 || cfg->n_tiles_h == 0 || cfg->n_tiles_w == 0) return nullptr;
  xb_track_manager* tm = new xb_track_manager();
  tm->cfg = *cfg;
  tm->fx = cfg->img_width * cfg->fx;
  tm->fy = cfg->img_height * cfg->fy;
  tm->cx = cfg->img_width * cfg->cx;
  tm->cy = cfg->img_height * cfg->cy;
  tm->inv_fx = 1.0 / tm->fx;
  tm->inv_fy = 1.0 / tm->fy;
  tm->cx_n = tm->cx * tm->inv_fx;
  tm->cy_n = tm->cy * tm->inv_fy;
  tm->s_term = 1.0 / (2.0 * std::tan(cfg->s / 2.0));
  tm->tile_h = (double)cfg->img_height / cfg->n_tiles_h;  // tiled_image.cpp:47-50
  tm->tile_w = (double)cfg->img_width / cfg->n_tiles_w;
  return tm;
}
XB_API void xb_tm_destroy(xb_track_manager* tm) { delete tm; }
XB_API void xb_tm_clear(xb_track_manager* tm) {  // TrackManager::clear (track_manager.cpp:74-81)
  if (!tm) return;
  tm->slam.clear(); tm->new_slam.clear(); tm->lost.clear(); tm->opp.clear(); tm->msckf_n.clear(); tm->msckf_short_n.clear();
}
XB_API int xb_tm_set_opp_ids(xb_track_manager* tm, const unsigned long long* ids, int n) {
  if (!tm || n < 0) return XB_E_INVALID;
  tm->opp_ids.assign(ids, ids + n);
  return XB_OK;
}

XB_API int xb_tm_manage_tracks(xb_track_manager* tm, const double* match_vector, int n_matches, const double* cam_rots,
                               int n_rots, int n_poses_max, int n_slam_features_max, int min_track_length) {
  if (!tm || n_matches < 0 || (n_matches > 0 && !match_vector) || n_rots < 1 || !cam_rots) return XB_E_INVALID;
  // ---- VIO::importMatches (vio.cpp:372-434): [cam_id, t_prev, x_prev, y_prev, t_cur, x_cur, y_cur, landmark xyz]
  struct Mt { Feat prev, cur; };
  std::vector<Mt> matches((size_t)n_matches);
  for (int i = 0; i < n_matches; ++i) {
    const double* v = match_vector + 10 * (size_t)i;
    matches[i].prev.t = v[1]; matches[i].prev.xd = v[2]; matches[i].prev.yd = v[3];
    matches[i].cur.t = v[4]; matches[i].cur.xd = v[5]; matches[i].cur.yd = v[6];
    tm->undistort(matches[i].prev);
    tm->undistort(matches[i].cur);
  }
  const int n_bins = (int)(tm->cfg.n_tiles_h * tm->cfg.n_tiles_w);
  // ---- the persistent tracks announced last frame join the persistent list (track_manager.cpp:120-124)
  tm->slam.insert(tm->slam.end(), tm->new_slam.begin(), tm->new_slam.end());
  tm->new_slam.clear();
  // per tile: indexes into (slam ++ new_slam) as they are NOW, and the indexes the persistent tracks had at entry
  std::vector<std::vector<unsigned>> bin_now((size_t)n_bins), bin_entry((size_t)n_bins);
  unsigned fullest = 0;
  // ---- 1. continue or lose the persistent tracks (track_manager.cpp:139-187)
  tm->lost.clear();
  unsigned n_lost = 0;
  for (unsigned t = 0; t < tm->slam.size();) {
    size_t hit = matches.size();
    for (size_t m = 0; m < matches.size(); ++m)
      if (same_feature(tm->slam[t].f.back(), matches[m].prev)) { hit = m; break; }
    if (hit == matches.size()) {
      tm->lost.push_back(t + n_lost);
      tm->slam.erase(tm->slam.begin() + t);
      ++n_lost;
      continue;
    }
    const int bin = tm->tile_of(matches[hit].cur);
    if (bin < 0 || bin >= n_bins) return XB_E_INVALID;  // feature outside the image
    bin_now[bin].push_back(t);
    bin_entry[bin].push_back(t + n_lost);
    if (bin_now[bin].size() > bin_now[fullest].size()) fullest = (unsigned)bin;
    tm->slam[t].f.push_back(matches[hit].cur);
    matches.erase(matches.begin() + hit);
    ++t;
  }
  // ---- 2. remaining matches extend opportunistic tracks or start new ones (track_manager.cpp:189-232)
  tm->msckf_n.clear();
  tm->msckf_short_n.clear();
  TrkList prev_opp;
  prev_opp.swap(tm->opp);
  for (const Mt& m : matches) {
    size_t hit = prev_opp.size();
    for (size_t t = 0; t < prev_opp.size(); ++t)
      if (same_feature(prev_opp[t].f.back(), m.prev)) { hit = t; break; }
    if (hit < prev_opp.size()) {
      prev_opp[hit].f.push_back(m.cur);
      tm->opp.push_back(prev_opp[hit]);
      prev_opp.erase(prev_opp.begin() + hit);
    } else {
      Trk nt;
      nt.id = ++tm->next_id;
      nt.f.push_back(m.prev);
      nt.f.push_back(m.cur);
      tm->opp.push_back(nt);
    }
  }
  // ---- 3. opportunistic tracks that just ended: short MSCKF tracks.  They belong to the previous frame, hence the
  //         attitude list without its last entry (track_manager.cpp:234-273)
  if (tm->cfg.multi_uav) {
    for (const Trk& dead : prev_opp) {
      if (dead.f.size() < 3) continue;
      for (size_t o = 0; o < tm->opp_ids.size(); ++o)
        if (dead.id == tm->opp_ids[o]) {
          tm->msckf_short_n.push_back(tm->normalize(dead, (size_t)n_rots));
          tm->opp_ids.erase(tm->opp_ids.begin() + o);
          break;
        }
    }
  } else {
    for (const Trk& dead : prev_opp) {
      if (dead.f.size() < 2) continue;
      const Trk nt = tm->normalize(dead, (size_t)(n_rots - 1));
      if (tm->baseline_ok(nt, cam_rots, n_rots - 1)) tm->msckf_short_n.push_back(nt);
    }
  }
  // ---- 4. longest first (std::sort, as the reference: see the header) and distribute (track_manager.cpp:274-398)
  std::sort(tm->opp.begin(), tm->opp.end(), [](const Trk& a, const Trk& b) { return a.f.size() > b.f.size(); });
  for (size_t t = 0; t < tm->opp.size();) {
    const int bin = tm->tile_of(tm->opp[t].f.back());
    if (bin < 0 || bin >= n_bins) return XB_E_INVALID;
    if (!(tm->opp[t].f.size() > (size_t)min_track_length - 1)) { ++t; continue; }
    if (tm->slam.size() + tm->new_slam.size() < (size_t)n_slam_features_max) {
      // a free SLAM slot
      tm->new_slam.push_back(tm->opp[t]);
      tm->opp.erase(tm->opp.begin() + t);
      bin_now[bin].push_back((unsigned)(tm->slam.size() + tm->new_slam.size() - 1));
      if (bin_now[bin].size() > bin_now[fullest].size()) fullest = (unsigned)bin;
    } else if (bin_now[fullest].size() > bin_now[bin].size() + 1) {
      // spread the SLAM features over the tiles: the youngest track of the fullest tile makes room
      const unsigned victim = bin_now[fullest].back();
      if (victim >= tm->slam.size()) {
        tm->new_slam.erase(tm->new_slam.begin() + (victim - tm->slam.size()));
      } else {
        tm->lost.push_back(bin_entry[fullest].back());
        bin_entry[fullest].pop_back();
        tm->slam.erase(tm->slam.begin() + victim);
      }
      for (auto& b : bin_now)
        for (unsigned& i : b)
          if (i > victim) --i;
      bin_now[fullest].pop_back();
      tm->new_slam.push_back(tm->opp[t]);
      tm->opp.erase(tm->opp.begin() + t);
      bin_now[bin].push_back((unsigned)(tm->slam.size() + tm->new_slam.size() - 1));
      for (int i = 0; i < n_bins; ++i)
        if (bin_now[i].size() > bin_now[fullest].size()) fullest = (unsigned)i;
    } else if (tm->opp[t].f.size() > (size_t)n_poses_max - 1) {
      // as long as the window: an MSCKF measurement if the baseline is sufficient; dropped either way
      const Trk nt = tm->normalize(tm->opp[t], (size_t)n_rots);
      if (tm->baseline_ok(nt, cam_rots, n_rots)) tm->msckf_n.push_back(nt);
      tm->opp.erase(tm->opp.begin() + t);
    } else {
      ++t;
    }
  }
  if (tm->cfg.multi_uav) tm->opp_ids.clear();
  // ---- 5. new SLAM tracks: MSCKF-SLAM initialisation if the baseline allows, standard otherwise; the MSCKF-SLAM ones
  //         first, which is the order in which the update inserts the features (track_manager.cpp:403-432)
  tm->new_std_n.clear();
  tm->new_msckf_n.clear();
  TrkList with_baseline, without;
  for (const Trk& t : tm->new_slam) {
    const Trk nt = tm->normalize(t, (size_t)n_rots);
    if (tm->baseline_ok(nt, cam_rots, n_rots)) { tm->new_msckf_n.push_back(nt); with_baseline.push_back(t); }
    else { tm->new_std_n.push_back(nt); without.push_back(t); }
  }
  tm->new_slam = with_baseline;
  tm->new_slam.insert(tm->new_slam.end(), without.begin(), without.end());
  return XB_OK;
}

XB_API int xb_tm_list_size(const xb_track_manager* tm, int which, int size_out, int* n_tracks, int* n_obs) {
  if (!tm) return XB_E_INVALID;
  TrkList tmp;
  const TrkList* l = tm->list(which, size_out, tmp);
  if (!l) return XB_E_INVALID;
  size_t n = 0;
  for (const Trk& t : *l) n += t.f.size();
  if (n_tracks) *n_tracks = (int)l->size();
  if (n_obs) *n_obs = (int)n;
  return XB_OK;
}
XB_API int xb_tm_get_list(const xb_track_manager* tm, int which, int size_out, int* offsets, double* xy,
                          unsigned long long* ids) {
  if (!tm || !offsets) return XB_E_INVALID;
  TrkList tmp;
  const TrkList* l = tm->list(which, size_out, tmp);
  if (!l) return XB_E_INVALID;
  int o = 0;
  offsets[0] = 0;
  for (size_t i = 0; i < l->size(); ++i) {
    for (const Feat& f : (*l)[i].f) {
      if (xy) { xy[2 * o] = f.x; xy[2 * o + 1] = f.y; }
      ++o;
    }
    offsets[i + 1] = o;
    if (ids) ids[i] = (*l)[i].id;
  }
  return (int)l->size();
}
XB_API int xb_tm_lost_slam_idxs(const xb_track_manager* tm, int* idxs, int cap) {
  if (!tm) return XB_E_INVALID;
  for (size_t i = 0; i < tm->lost.size() && (int)i < cap; ++i) idxs[i] = (int)tm->lost[i];
  return (int)tm->lost.size();
}
XB_API int xb_tm_remove_persistent_track(xb_track_manager* tm, unsigned idx) {  // track_manager.cpp:83-85
  if (!tm || idx >= tm->slam.size()) return XB_E_INVALID;
  tm->slam.erase(tm->slam.begin() + idx);
  return XB_OK;
}
XB_API int xb_tm_remove_new_persistent_tracks(xb_track_manager* tm, const unsigned* idxs, int n) {  // :87-97, increasing order
  if (!tm || n < 0) return XB_E_INVALID;
  for (int i = n; i > 0; --i) {
    if (idxs[i - 1] >= tm->new_slam.size()) return XB_E_INVALID;
    tm->new_slam.erase(tm->new_slam.begin() + idxs[i - 1]);
  }
  return XB_OK;
}
XB_API int xb_tm_counts(const xb_track_manager* tm, int* n_slam, int* n_new_slam, int* n_opp) {
  if (!tm) return XB_E_INVALID;
  if (n_slam) *n_slam = (int)tm->slam.size();
  if (n_new_slam) *n_new_slam = (int)tm->new_slam.size();
  if (n_opp) *n_opp = (int)tm->opp.size();
  return XB_OK;
}

// The Delaunay facet of a point set that contains the query point (TrackManager::featureTriangleAtPoint,
// track_manager.cpp:443-560).  The reference inserts the points into a cv::Subdiv2D over the rectangle
// (-1, -1, w + 2, h + 2) and locates the query; Subdiv2D stores single-precision points and bounds the plane with three
// virtual vertices at (rx + B, ry), (rx, ry + B), (rx - B, ry - B), B = 3 max(w + 2, h + 2): a facet that touches one of
// them is not a facet of SLAM features (the reference then reports none).  The Delaunay triangulation of points in
// general position is unique, so the same point set is triangulated here by Bowyer-Watson insertion; a query exactly on
// an edge or a vertex takes the first facet found (the reference takes the one its edge walk ends in).
// xy: n x 2 (distorted pixel coordinates).  Returns 3 and the vertex indexes, or 0 when there is no such facet.
XB_API int xb_tm_delaunay_facet(const double* xy, int n, int img_width, int img_height, double qx, double qy, int* ids) {
  if (!xy || !ids || n < 3) return 0;
  struct P2 { double x, y; };
  std::vector<P2> p(3 + (size_t)n);
  const double rx = -1.0, ry = -1.0, B = 3.0 * std::max(img_width + 2, img_height + 2);
  p[0] = {rx + B, ry};
  p[1] = {rx, ry + B};
  p[2] = {rx - B, ry - B};
  for (int i = 0; i < n; ++i) p[3 + i] = {(double)(float)xy[2 * i], (double)(float)xy[2 * i + 1]};
  struct Tri { int a, b, c; };
  auto orient = [&](const P2& a, const P2& b, const P2& c) { return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x); };
  std::vector<Tri> tris;
  tris.push_back(orient(p[0], p[1], p[2]) > 0 ? Tri{0, 1, 2} : Tri{0, 2, 1});  // counter-clockwise
  auto in_circle = [&](const Tri& t, const P2& d) {
    const double ax = p[t.a].x - d.x, ay = p[t.a].y - d.y, bx = p[t.b].x - d.x, by = p[t.b].y - d.y;
    const double cx = p[t.c].x - d.x, cy = p[t.c].y - d.y;
    const double det = (ax * ax + ay * ay) * (bx * cy - cx * by) - (bx * bx + by * by) * (ax * cy - cx * ay) +
                       (cx * cx + cy * cy) * (ax * by - bx * ay);
    return det > 0.0;
  };
  std::vector<Tri> keep;
  std::vector<std::pair<int, int>> edges;
  for (int i = 3; i < 3 + n; ++i) {
    bool dup = false;
    for (int j = 0; j < i && !dup; ++j) dup = p[j].x == p[i].x && p[j].y == p[i].y;
    if (dup) continue;  // Subdiv2D::insert returns the existing vertex
    keep.clear();
    edges.clear();
    for (const Tri& t : tris) {
      if (!in_circle(t, p[i])) { keep.push_back(t); continue; }
      const int e[3][2] = {{t.a, t.b}, {t.b, t.c}, {t.c, t.a}};
      for (const auto& ed : e) {
        bool shared = false;
        for (size_t k = 0; k < edges.size(); ++k)
          if (edges[k].first == ed[1] && edges[k].second == ed[0]) { edges.erase(edges.begin() + k); shared = true; break; }
        if (!shared) edges.emplace_back(ed[0], ed[1]);
      }
    }
    for (const auto& ed : edges) keep.push_back({ed.first, ed.second, i});
    tris.swap(keep);
  }
  const P2 q = {(double)(float)qx, (double)(float)qy};
  for (const Tri& t : tris) {
    if (orient(p[t.a], p[t.b], q) < 0 || orient(p[t.b], p[t.c], q) < 0 || orient(p[t.c], p[t.a], q) < 0) continue;
    if (t.a < 3 || t.b < 3 || t.c < 3) return 0;  // bounded by a virtual vertex: not a facet of features
    ids[0] = t.a - 3; ids[1] = t.b - 3; ids[2] = t.c - 3;
    return 3;
  }
  return 0;
}

// Camera::undistort + Camera::normalize of one image point (camera.cpp:69-87, 95-101): how VIO::processMatchesMeasurement
// turns the LRF image point into RangeMeasurement::img_pt_n (vio.cpp:288-294).  out = normalised (x, y).
XB_API int xb_tm_normalize_point(const xb_track_manager* tm, double x_dist, double y_dist, double* out) {
  if (!tm || !out) return XB_E_INVALID;
  Feat f;
  f.xd = x_dist;
  f.yd = y_dist;
  tm->undistort(f);
  out[0] = f.x * tm->inv_fx - tm->cx_n;
  out[1] = f.y * tm->inv_fy - tm->cy_n;
  return XB_OK;
}

// TrackManager::featureTriangleAtPoint (track_manager.cpp:443-560) on the last observations of the persistent (SLAM)
// tracks: the ids of the three SLAM features around the LRF image point (distorted pixel coordinates), or 0.
XB_API int xb_tm_feature_triangle_at_point(const xb_track_manager* tm, double x_dist, double y_dist, int* ids) {
  if (!tm || !ids) return XB_E_INVALID;
  std::vector<double> xy;
  xy.reserve(2 * tm->slam.size());
  for (const Trk& t : tm->slam) {
    if (t.f.empty()) return 0;
    xy.push_back(t.f.back().xd);
    xy.push_back(t.f.back().yd);
  }
  return xb_tm_delaunay_facet(xy.data(), (int)tm->slam.size(), (int)tm->cfg.img_width, (int)tm->cfg.img_height, x_dist, y_dist, ids);
}

}  // extern "C"
