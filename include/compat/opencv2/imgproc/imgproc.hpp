/* the sobfu application includes this header but uses nothing from it (src/apps/demo.cpp:9) */
#pragma once
#include <opencv2/core/core.hpp>
