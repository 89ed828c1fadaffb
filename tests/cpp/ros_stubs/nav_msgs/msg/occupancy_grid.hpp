// TEST STUB of nav_msgs/msg/OccupancyGrid
#pragma once
#include <cstdint>
#include <vector>

#include "geometry_msgs/msg/pose_stamped.hpp"
namespace nav_msgs { namespace msg {
struct MapMetaData { float resolution = 0; uint32_t width = 0, height = 0; geometry_msgs::msg::Pose origin; };
struct OccupancyGrid { std_msgs::msg::Header header; MapMetaData info; std::vector<int8_t> data; };
}}
