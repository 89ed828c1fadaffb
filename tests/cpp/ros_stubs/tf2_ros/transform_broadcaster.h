// TEST STUB of tf2_ros::TransformBroadcaster
#pragma once
#include "geometry_msgs/msg/transform_stamped.hpp"
#include "rclcpp/rclcpp.hpp"
namespace tf2_ros {
class TransformBroadcaster {
public:
    explicit TransformBroadcaster(rclcpp::Node&) {}
    void sendTransform(const geometry_msgs::msg::TransformStamped&) { ++count; }
    size_t count = 0;
};
}
