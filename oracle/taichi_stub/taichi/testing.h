// TEST INFRASTRUCTURE — see taichi/util.h.  src/async/async_mpm.cpp:9 includes this path.
#pragma once
#include <taichi/common/testing.h>
