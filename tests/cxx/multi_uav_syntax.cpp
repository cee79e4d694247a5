#include "x/xb200_binding.hpp"
int main() {
  x::VioUpdater u(1.0 / 320, 0.1, 0.5, 0.25, 2, 0.3, 0.1, 0.1);
  x::Ekf ekf(u);
  x::MsckfMatches mm; x::SlamMatches sm;
  u.setMsckfMatches(mm); u.setSlamMatches(sm);
  (void)&x::Ekf::processOthersMeasurement;
  return 0;
}
