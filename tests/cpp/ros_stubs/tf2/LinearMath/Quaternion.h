// TEST STUB (slam_node.cpp includes it, uses Eigen::Quaterniond instead)
#pragma once
