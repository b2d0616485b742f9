// oracle/ref_shim: every OpenCV header the reference includes resolves to the one stand-in header (test infrastructure only)
#include "../core/core.hpp"
