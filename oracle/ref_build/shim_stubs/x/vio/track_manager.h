// Test infrastructure (oracle/): stands in for include/x/vio/track_manager.h when the reference's filter back end
// is compiled in place.  manageTracks() does nothing; the getters return the five track lists + lost-feature
// indexes the harness loaded -- exactly the seam VioUpdater::preProcess reads (vio_updater.cpp:172-179), which is
// where SURVEY.md 8(b) injects synthetic tracks.
#ifndef TRACK_MANAGER_H_
#define TRACK_MANAGER_H_
#include <memory>
#include <vector>

#include <x/vio/types.h>
#include <x/vision/types.h>

namespace x {
class TrackManager {
 public:
  struct Lists {
    TrackList slam, msckf, msckf_short, new_slam_std, new_slam_msckf, opp;
    std::vector<unsigned int> lost;
    std::vector<int> tri;  // the facet featureTriangleAtPoint() reports (3 SLAM feature ids, or empty)
  };
  std::shared_ptr<Lists> lists = std::make_shared<Lists>();  // shared by the copy VioUpdater keeps

  TrackList getMsckfTracks() const { return lists->msckf; }
  TrackList getShortMsckfTracks() const { return lists->msckf_short; }
  TrackList getNewSlamStdTracks() const { return lists->new_slam_std; }
  TrackList getNewSlamMsckfTracks() const { return lists->new_slam_msckf; }
  TrackList getOppTracks() { return lists->opp; }
  void setOppUpgradesMSCKF(OppIDListPtr&) {}
  TrackList normalizeSlamTracks(const int) const { return lists->slam; }
  void clear() { *lists = Lists(); }
  std::vector<unsigned int> getLostSlamTrackIndexes() const { return lists->lost; }
  void manageTracks(MatchList&, const AttitudeList, const size_t, const size_t, const size_t, TiledImage&) {}
  std::vector<int> featureTriangleAtPoint(const Feature&, TiledImage&) const { return lists->tri; }
};
}  // namespace x
#endif
