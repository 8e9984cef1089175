// stand-in, see opencv2/core/core.hpp in this tree (test infrastructure, NOT OpenCV)
#include "../core/core.hpp"
