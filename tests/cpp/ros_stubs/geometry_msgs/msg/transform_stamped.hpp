// TEST STUB of geometry_msgs/msg/TransformStamped
#pragma once
#include "geometry_msgs/msg/pose_stamped.hpp"
namespace geometry_msgs { namespace msg {
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::msg::Header header; std::string child_frame_id; Transform transform; };
}}
