#!/bin/sh
# Builds the C++17 host-API test against the in-tree libraries (no GPU needed to compile).
set -e
cd "$(dirname "$0")"
ROOT=../..
g++ -std=c++17 -O2 -Wall -Wno-unused-function -I$ROOT/include -I$ROOT/lidar-slam-from-scratch_b200/host \
    host_api_test.cpp -o host_api_test \
    -L$ROOT/lidar-slam-from-scratch_b200 -lslam_b200 -L$ROOT/oracle -loracle -L$ROOT/synth -lsynth \
    -Wl,-rpath,'$ORIGIN/../../lidar-slam-from-scratch_b200' -Wl,-rpath,'$ORIGIN/../../oracle' -Wl,-rpath,'$ORIGIN/../../synth'
# the same test through the headers' Eigen branch (`__has_include(<Eigen/Dense>)`): Eigen itself is not in this image, the
# Eigen-API stand-in of the oracle (oracle/eigen_standin) takes its place, so PointCloud::Matrix etc. are Eigen-style types
g++ -std=c++17 -O2 -Wall -Wno-unused-function -I$ROOT/oracle/eigen_standin -I$ROOT/include \
    -I$ROOT/lidar-slam-from-scratch_b200/host host_api_test.cpp -o host_api_test_eigenapi \
    -L$ROOT/lidar-slam-from-scratch_b200 -lslam_b200 -L$ROOT/oracle -loracle -L$ROOT/synth -lsynth \
    -Wl,-rpath,'$ORIGIN/../../lidar-slam-from-scratch_b200' -Wl,-rpath,'$ORIGIN/../../oracle' -Wl,-rpath,'$ORIGIN/../../synth'
echo "built tests/cpp/host_api_test and tests/cpp/host_api_test_eigenapi"
