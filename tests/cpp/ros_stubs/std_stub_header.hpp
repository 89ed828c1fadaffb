// TEST STUB: std_msgs/Header as the slam_viz messages use it
#pragma once
#include <string>

#include "rclcpp/rclcpp.hpp"
namespace std_msgs { namespace msg {
struct Header {
    rclcpp::Time stamp;
    std::string frame_id;
};
}}
