// Test-infrastructure stand-in for an OpenCV 3.2 header (see ../cvshim.hpp).
#include "../cvshim.hpp"
