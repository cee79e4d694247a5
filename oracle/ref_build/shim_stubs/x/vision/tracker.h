// Test infrastructure (oracle/): stands in for the reference's image front end (include/x/vision/tracker.h,
// FAST + KLT + RANSAC on OpenCV -- out of scope) when its filter back end is compiled in place.
// Found first on the include path, it gives VioUpdater a Tracker that tracks nothing and hands out the
// inter-agent match lists the harness loaded (the seam of vio_updater.cpp:172-186).
#ifndef JPL_VPP_TRACKER_H_
#define JPL_VPP_TRACKER_H_
#include <memory>

#include "x/vision/types.h"

namespace x {
class Tracker {
 public:
  struct Shared {
    MsckfMatches msckf_matches;
    SlamMatches slam_matches;
  };
  std::shared_ptr<Shared> sh = std::make_shared<Shared>();  // shared by the copy VioUpdater keeps

  void track(TiledImage&, const double&, unsigned int) {}
  bool checkMatches() { return false; }
  [[nodiscard]] MatchList getMatches() const { return {}; }
  void cleanSlamMatches() { sh->slam_matches.clear(); }
  MsckfMatches& getMsckfMatches() { return sh->msckf_matches; }
  SlamMatches& getSlamMatches() { return sh->slam_matches; }
  void updateOppMatches(const TrackList&, const TrackList&, const TrackList&) {}
  OppIDListPtr getOppIds() { return std::make_shared<std::vector<unsigned long long>>(); }
};
}  // namespace x
#endif
