// the usage example of the reference's README (README.md:196-252), against include/x/
#include <x/vio/vio.h>
#include <ctime>
using namespace x;
int main(int argc, char**) {
  VIO vio_;
  if (argc > 99) {
    const auto params = vio_.loadParamsFromYaml("PATH_TO_A_YAML_FILE");
    vio_.setUp(params);
    time_t now = time(0);
    vio_.initAtTime((double)now);
    SunAngleMeasurement angle;
    angle.timestamp = 1.0; angle.x_angle = 0.1; angle.y_angle = 0.2;
    vio_.setLastSunAngleMeasurement(angle);
    RangeMeasurement range;
    range.timestamp = 1.0; range.range = 2.0;
    vio_.setLastRangeMeasurement(range);
    double timestamp = 1.0; int seq = 3;
    Vector3 w_m(0, 0, 0); Vector3 a_m(0, 0, 9.81);
    const auto propagated_state = vio_.processImu(timestamp, seq, w_m, a_m);
    (void)propagated_state;
  }
  return 0;
}
