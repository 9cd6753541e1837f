// Forwarding header of the test-only OpenCV shim (oracle/ref_shim/cvshim.hpp): lets the reference sources keep their #include lines.
#pragma once
#include "../cvshim.hpp"
