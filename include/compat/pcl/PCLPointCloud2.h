#pragma once
#include <cstdint>
#include <vector>
namespace pcl {
struct PCLPointCloud2 { unsigned width = 0, height = 0, point_step = 0, row_step = 0; std::vector<uint8_t> data; };
}
