// Forwards to the B200 back end's implementation of this reference header (jpl-x/x_multi_agent include/x/vision/camera.h):
// all classes of the hot-path operator API live in include/x/xb200_binding.hpp.
#pragma once
#include "../xb200_binding.hpp"
