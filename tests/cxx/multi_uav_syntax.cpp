// Syntax check of the operator API in both build flavours, incl. a user-defined Updater with host matrices
// (include/x/ekf/updater.h:106-231: the pure virtuals differ between the single-agent and the -DMULTI_UAV build).
#include "x/ekf/ekf.h"
#include "x/ekf/simple_state.h"
#include "x/vio/vio_updater.h"

class RangeOnlyUpdater : public x::Updater {   // what a maintainer's own sensor updater looks like
 public:
  double getTime() const override { return t_; }
  double t_ = 0.0;

 protected:
  using L = std::vector<std::shared_ptr<x::Matrix>>;
  void preProcess(const x::State&) override {}
  bool preUpdate(x::State&) override { return true; }
  bool preUpdateShortMsckf() override { return false; }
#ifdef MULTI_UAV
  bool preUpdateCI() override { return false; }
  void constructSlamCIUpdate(const x::State&, L&, L&, L&, L&) override {}
  void constructUpdate(const x::State& s, x::Matrix& h, x::Matrix& res, x::Matrix& r, L&, L&, L&, L&) override { build(s, h, res, r); }
  void constructShortMsckfUpdate(const x::State&, x::Matrix&, x::Matrix&, x::Matrix&, L&, L&, L&, L&) override {}
#else
  void constructUpdate(const x::State& s, x::Matrix& h, x::Matrix& res, x::Matrix& r) override { build(s, h, res, r); }
  void constructShortMsckfUpdate(const x::State&, x::Matrix&, x::Matrix&, x::Matrix&) override {}
#endif
  void postUpdate(x::State&, const x::Matrix&) override {}
  void build(const x::State& s, x::Matrix& h, x::Matrix& res, x::Matrix& r) {
    h = x::Matrix::Zero(1, s.nErrorStates());
    h(0, 2) = 1.0;
    res = x::Matrix::Constant(1, 1, 0.1);
    r = x::Matrix::Identity(1, 1) * 0.01;
  }
};

int main() {
  x::Tracker tracker;
  x::StateManager sm(4, 2);
  x::TrackManager tm;
  x::VioUpdater u(tracker, sm, tm, 1.0 / 320, 0.1, 0.5, 0.25, 2, 0.3, 0.1, 0.1);
  x::Ekf ekf(u);
  x::Ekf copy(ekf);
  RangeOnlyUpdater r;
  x::Ekf ekf2(r);
  x::State s(4, 2);
  s.getCovarianceRef()(0, 0) = 2.0;
  s.correct(Eigen::VectorXd::Zero(s.nErrorStates()));
  (void)s.computeCameraAttitude();
  (void)s.computeCameraPosition();
  (void)s.computeCameraOrientation();
  s.reset();
#ifdef MULTI_UAV
  x::MsckfMatches mm;
  x::SlamMatches smm;
  u.setMsckfMatches(mm);
  u.setSlamMatches(smm);
  (void)&x::Ekf::processOthersMeasurement;
  (void)&x::Updater::collaborativeUpdate;
#endif
  return 0;
}
