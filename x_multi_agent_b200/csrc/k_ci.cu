// Multi-agent covariance-intersection entry points (filled in by the CI milestone).
#include "../../include/xb200.h"
#include "xb_kernels.h"
extern "C" int xb_ci_payload_len(const xb_filter*) { return XB_E_UNSUPPORTED; }
extern "C" int xb_ci_pack(xb_filter*, int, double*) { return XB_E_UNSUPPORTED; }
extern "C" int xb_ekf_process_others(xb_filter*, double, const xb_peer_state*, int, const xb_slam_match*, int, double*) {
  return XB_E_UNSUPPORTED;
}
extern "C" int xb_vio_set_msckf_matches(xb_filter*, const xb_peer_state*, int, const xb_msckf_match*, int) {
  return XB_E_UNSUPPORTED;
}
