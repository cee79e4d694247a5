// Test infrastructure (oracle/): OpenCV stand-in, see opencv2/xref_cv.hpp.
#include "xref_cv.hpp"
