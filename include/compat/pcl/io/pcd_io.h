/* included by the sobfu application (src/apps/demo.cpp:12), nothing of it is used */
#pragma once
#include <pcl/PCLPointCloud2.h>
